# one gpurun job (1 GPU): real-input fast form (k_channelize_rdft, LPR lanes per row): tests, kernel numbers, ncu capture;
# demod register-cap / F2F experiments at one lane per channel in the pipeline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_air.py tests/test_gpu_cs16.py -m gpu -q -x > gpurun_out/r2_pytest_air.log 2>&1; tail -4 gpurun_out/r2_pytest_air.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lane_width" > gpurun_out/r2_pytest_lanes.log 2>&1; tail -3 gpurun_out/r2_pytest_lanes.log
: > gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 2500000 296 8 | tail -1 >> gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 2500000 296 8 fast | tail -1 >> gpurun_out/r2_airfast.jsonl
ACB_FAST_WARPS=2 python tools/bench_air.py 2500000 296 8 fast | tail -1 >> gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 10000000 74 8 fast | tail -1 >> gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 10000000 74 8 | tail -1 >> gpurun_out/r2_airfast.jsonl
python tools/bench_air.py 6000000 123 8 fast | tail -1 >> gpurun_out/r2_airfast.jsonl
cat gpurun_out/r2_airfast.jsonl
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_channelize_rdft -s 2 -c 1 -f -o gpurun_out/r2_k1_air_fast python tools/bench_air.py 2500000 296 8 fast > gpurun_out/ncu_k1_air_fast.log 2>&1
tail -2 gpurun_out/ncu_k1_air_fast.log
timeout 400 python tools/ab_demod.py 4736 1,49,50,17 fast > gpurun_out/r2_ab8.jsonl 2>/dev/null; cat gpurun_out/r2_ab8.jsonl
