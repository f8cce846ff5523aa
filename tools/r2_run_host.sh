# one gpurun job (1 GPU): consumer thread's bucket ordering: order-sensitive tests, fast headline with the host-side timings
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_compat.py tests/test_gpu_air.py tests/test_gpu_cs16.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_pytest_host.log 2>&1; tail -4 gpurun_out/r2_pytest_host.log
python bench.py --steps 20 --warmup 5 --channelizer fast --config none --no-cpu-baseline | tail -1 > gpurun_out/bench_n1_fast.json
python - <<'PY'
import json
f=json.load(open('gpurun_out/bench_n1_fast.json')); print('fast headline', round(f['value']), f['ms_per_step'], f['roofline']['frac'], f['roofline']['isolated']['frac'], f['kernels'], f['checked'])
PY
