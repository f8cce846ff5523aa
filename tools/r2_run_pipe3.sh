# one gpurun job (1 GPU): pipeline depth 3 + block FEC on its own stream: full GPU suite, bench lines, pipeline sweep
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --channelizer fast --config none | tail -1 > gpurun_out/bench_n1_fast.json
timeout 300 python tools/ab_demod.py 592,4736 4,1 fast > gpurun_out/r2_ab7.jsonl 2>/dev/null; cat gpurun_out/r2_ab7.jsonl
python bench.py --steps 20 --warmup 5 | tail -1 > gpurun_out/bench_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json')); a=d['alt_channelizer']
print('exact', round(d['value']), d['config']['streams_per_gpu'], d['roofline']['frac'], d['roofline']['isolated']['frac'], 'e2e', round(d['e2e']['value'],2), d['checked'], d.get('sweep'))
print('fast(alt)', round(a['value']), a['roofline']['frac'], a['roofline']['isolated']['frac'])
print('configs', json.dumps(d.get('configs'))[:1500])
f=json.load(open('gpurun_out/bench_n1_fast.json')); print('fast headline', round(f['value']), f['config']['streams_per_gpu'], f['roofline']['frac'], f['roofline']['isolated']['frac'], f['kernels'], f.get('sweep'), f['checked'])
PY
