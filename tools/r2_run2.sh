# round-2 GPU job 2: full GPU suite, demod lanes x streams sweep with the pipelined loads, first full bench line, ncu of K2
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; tail -8 gpurun_out/r2_pytest2.log
timeout 600 python tools/ab_demod.py 592,2368,4736 8,4,2,1 fast > gpurun_out/r2_ab2_fast.jsonl 2> gpurun_out/r2_ab2_fast.err
timeout 300 python tools/ab_demod.py 592,2368 4,2 exact > gpurun_out/r2_ab2_exact.jsonl 2> gpurun_out/r2_ab2_exact.err
cat gpurun_out/r2_ab2_fast.jsonl gpurun_out/r2_ab2_exact.jsonl
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err ) 2>&1 | tail -3
tail -c 1500 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_a.json'))
    print('value',round(d['value']),'S',d['config']['streams_per_gpu'],'B',d['config']['blocks_per_step'],'e2e',round(d['e2e']['value']) if d.get('e2e') else None)
    print('sweep',d['streams_sweep']); print('roofline',d['roofline']['frac'],d['roofline'].get('isolated')); print('kernels',d['kernels'])
    a=d['alt_channelizer']; print('alt',round(a['value']),a['roofline']['frac'],a['roofline'].get('isolated'),a['checked'])
    print('checked',d['checked']); print('configs',json.dumps(d['configs'])[:3000])
except Exception as e: print('bench parse failed',e)
PY
ACB_DEMOD_LANES=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_demod2 -s 2 -c 1 -f -o gpurun_out/r2_k2_l4b python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k2_l4b.log 2>&1
