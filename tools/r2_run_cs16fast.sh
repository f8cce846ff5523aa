# one gpurun job (1 GPU): CS16 fast form + multi-threaded consumer: full GPU suite, bench lines, CS16 kernel numbers, one ncu capture
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
python bench.py --steps 20 --warmup 5 --channelizer fast --config none --no-cpu-baseline | tail -1 > gpurun_out/bench_n1_fast.json
python - <<'PY'
import json
f=json.load(open('gpurun_out/bench_n1_fast.json')); print('fast headline', round(f['value']), f['ms_per_step'], f['roofline']['frac'], f['roofline']['isolated']['frac'], f['kernels'], f['checked']['bit_exact'])
PY
: > gpurun_out/r2_cs16fast.jsonl
python tools/bench_cs16.py 0 296 8 | tail -1 >> gpurun_out/r2_cs16fast.jsonl
python tools/bench_cs16.py 0 296 8 fast | tail -1 >> gpurun_out/r2_cs16fast.jsonl
ACB_FAST_WARPS=1 python tools/bench_cs16.py 0 296 8 fast | tail -1 >> gpurun_out/r2_cs16fast.jsonl
python tools/bench_cs16.py 1 296 8 fast | tail -1 >> gpurun_out/r2_cs16fast.jsonl
cat gpurun_out/r2_cs16fast.jsonl
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_channelize_dft1 -s 2 -c 1 -f -o gpurun_out/r2_k1_cs16_fast python tools/bench_cs16.py 0 296 8 fast > gpurun_out/ncu_k1_cs16_fast.log 2>&1
tail -2 gpurun_out/ncu_k1_cs16_fast.log
