# one gpurun job (1 GPU): last verification of the host-side touches (launch decisions, consumer threads)
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_pytest_misc.log 2>&1; tail -4 gpurun_out/r2_pytest_misc.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
