# one gpurun job (1 GPU): K=240 real fast form, the consumer thread's fallback/helper paths
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_air.py tests/test_gpu_parity.py -m gpu -q -x -k "fast_form or consumer or full_path" > gpurun_out/r2_pytest_misc.log 2>&1; tail -4 gpurun_out/r2_pytest_misc.log
python tools/bench_air.py 3000000 246 8 fast | tail -1
