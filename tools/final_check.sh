# the round-end sequence on one GPU: GPU tests + smoke() + both bench arms
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref_n1.json
python bench.py --steps 20 --warmup 5 | tail -1 > gpurun_out/bench_n1.json
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json')); a=d['alt_channelizer']
print(round(d['value']), d['config']['streams_per_gpu'], round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['isolated']['frac'], round(a['value']), a['roofline']['frac'], a['roofline']['isolated']['frac'], d['checked'], d['clocks'])
r=json.load(open('gpurun_out/bench_ref_n1.json')); print(r['value'])"
