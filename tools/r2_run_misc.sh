# one gpurun job (1 GPU): drop-in tests incl. the UDP sinks
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compat.py -m gpu -q > gpurun_out/r2_pytest_misc.log 2>&1; tail -6 gpurun_out/r2_pytest_misc.log
