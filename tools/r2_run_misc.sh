# one gpurun job (1 GPU): the GPU suite and smoke() on the final library
mkdir -p gpurun_out
timeout 140 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
