# one gpurun job (1 GPU): GPU tests, smoke, the ncu --set full captures of every kernel, the launch list, the bench lines;
# outputs land in gpurun_out/ (tools/summarize_ncu.py r2 turns them into profiles/r2_*)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_channelize -s 2 -c 1 -f -o gpurun_out/r2_k1 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k1.log 2>&1
timeout 300 $NCU -k regex:k_channelize_dft -s 3 -c 1 -f -o gpurun_out/r2_k1_fast python tools/bench_k1.py fast > gpurun_out/ncu_k1_fast.log 2>&1
timeout 300 $NCU -k regex:k_demod2 -s 2 -c 1 -f -o gpurun_out/r2_k2 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k2.log 2>&1
ACB_DEMOD_LANES=4 timeout 400 $NCU -k regex:k_demod2 -s 1 -c 1 -f -o gpurun_out/r2_k2_sat python tools/prof_run.py 2368 16 3 > gpurun_out/ncu_k2_sat.log 2>&1
timeout 300 $NCU -k regex:k_block_fec -s 2 -c 1 -f -o gpurun_out/r2_k3 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k3.log 2>&1
ACB_DEMOD_LANES=1 timeout 600 $NCU -k regex:k_demod2 -s 1 -c 1 -f -o gpurun_out/r2_k2_l1 python tools/prof_run.py 4736 8 3 > gpurun_out/ncu_k2_l1.log 2>&1
timeout 300 $NCU -k regex:k_channelize_dft1 -s 2 -c 1 -f -o gpurun_out/r2_k1_cs16_fast python tools/bench_cs16.py 0 296 8 fast > gpurun_out/ncu_k1_cs16_fast.log 2>&1
timeout 300 $NCU -k regex:k_channelize_rdft -s 2 -c 1 -f -o gpurun_out/r2_k1_air_fast python tools/bench_air.py 2500000 296 8 fast > gpurun_out/ncu_k1_air_fast.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --config none --no-check > gpurun_out/r2_launches_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref_n1.json
python bench.py --steps 20 --warmup 5 | tail -1 > gpurun_out/bench_n1.json
python bench.py --steps 20 --warmup 5 --channelizer fast --config none | tail -1 > gpurun_out/bench_n1_fast.json
python tools/bench_k1.py fast | tail -1 > gpurun_out/bench_k1_fast.json
python tools/bench_k1.py exact | tail -1 > gpurun_out/bench_k1_exact.json
python tools/bench_air.py 2500000 296 8 | tail -1 > gpurun_out/bench_air_c8.json
python tools/bench_cs16.py 0 | tail -1 > gpurun_out/bench_cs16_soapy.json
python tools/bench_cs16.py 1 | tail -1 > gpurun_out/bench_cs16_sdrplay.json
: > gpurun_out/r2_air_fast.jsonl; : > gpurun_out/r2_cs16_fast.jsonl
for a in "2500000 296" "6000000 123" "10000000 74"; do python tools/bench_air.py $a 8 fast | tail -1 >> gpurun_out/r2_air_fast.jsonl; done
for v in 0 1; do python tools/bench_cs16.py $v 296 8 fast | tail -1 >> gpurun_out/r2_cs16_fast.jsonl; done
cat gpurun_out/bench_air_c8.json gpurun_out/bench_cs16_soapy.json gpurun_out/bench_cs16_sdrplay.json gpurun_out/r2_air_fast.jsonl gpurun_out/r2_cs16_fast.jsonl
timeout 300 python tools/ab_demod.py 592,4736 8,4,1 fast > gpurun_out/r2_ab6.jsonl 2>/dev/null; cat gpurun_out/r2_ab6.jsonl
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json')); a=d['alt_channelizer']
print('exact', round(d['value']), d['config']['streams_per_gpu'], d['roofline']['frac'], d['roofline']['isolated']['frac'], 'e2e', round(d['e2e']['value']), d['checked'])
print('fast(alt)', round(a['value']), a['roofline']['frac'], a['roofline']['isolated']['frac'])
f=json.load(open('gpurun_out/bench_n1_fast.json')); print('fast headline', round(f['value']), f['config']['streams_per_gpu'], f['roofline']['frac'], f['roofline']['isolated']['frac'], f['kernels'])
print(open('gpurun_out/bench_k1_fast.json').read())
PY
ls -la gpurun_out | tail -30
