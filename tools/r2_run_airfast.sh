# one gpurun job (1 GPU): real-input fast form (k_channelize_rdft): tests, kernel numbers, ncu capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_air.py tests/test_gpu_cs16.py tests/test_compat.py -m gpu -q -x > gpurun_out/r2_pytest_air.log 2>&1; tail -6 gpurun_out/r2_pytest_air.log
: > gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 2500000 296 8 | tail -1 >> gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 2500000 296 8 fast | tail -1 >> gpurun_out/r2_airfast.jsonl
ACB_FAST_WARPS=1 python tools/bench_air.py 2500000 296 8 fast | tail -1 >> gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 10000000 74 8 fast | tail -1 >> gpurun_out/r2_airfast.jsonl
cat gpurun_out/r2_airfast.jsonl
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_channelize_rdft -s 2 -c 1 -f -o gpurun_out/r2_k1_air_fast python tools/bench_air.py 2500000 296 8 fast > gpurun_out/ncu_k1_air_fast.log 2>&1
tail -2 gpurun_out/ncu_k1_air_fast.log
