# one gpurun job (1 GPU): the one-row fast channelizer after the issue-slot work (uniform bulk-copy issue, slots sorted by
# residue, twiddles one slot ahead): parity tests, kernel alone, in the pipeline, one ncu capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_pytest_dft1b.log 2>&1; tail -5 gpurun_out/r2_pytest_dft1b.log
: > gpurun_out/r2_dft1b.jsonl
python tools/bench_k1.py fast | tail -1 >> gpurun_out/r2_dft1b.jsonl
ACB_FAST_PF=0 python tools/bench_k1.py fast | tail -1 >> gpurun_out/r2_dft1b.jsonl
ACB_FAST_WARPS=4 python tools/bench_k1.py fast | tail -1 >> gpurun_out/r2_dft1b.jsonl
ACB_FAST_ROWS=2 python tools/bench_k1.py fast | tail -1 >> gpurun_out/r2_dft1b.jsonl
timeout 300 python tools/ab_demod.py 592,4736 4,1 fast >> gpurun_out/r2_dft1b.jsonl 2>/dev/null
ACB_FAST_PF=0 timeout 300 python tools/ab_demod.py 4736 1 fast >> gpurun_out/r2_dft1b.jsonl 2>/dev/null
cat gpurun_out/r2_dft1b.jsonl
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_channelize_dft -s 3 -c 1 -f -o gpurun_out/r2_k1_fast python tools/bench_k1.py fast > gpurun_out/ncu_k1_fast.log 2>&1
python bench.py --steps 20 --warmup 5 --channelizer fast --config none | tail -1 > gpurun_out/bench_n1_fast.json
python - <<'PY'
import json
f=json.load(open('gpurun_out/bench_n1_fast.json')); print('fast headline', round(f['value']), f['config']['streams_per_gpu'], f['roofline']['frac'], f['roofline']['isolated']['frac'], f['kernels'], f.get('sweep'))
PY
