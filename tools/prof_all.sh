set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_channelize -s 2 -c 1 -f -o gpurun_out/r1_k1 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k1.log 2>&1
timeout 300 $NCU -k regex:k_channelize_dft -s 3 -c 1 -f -o gpurun_out/r1_k1_fast python tools/bench_k1.py fast > gpurun_out/ncu_k1_fast.log 2>&1
timeout 300 $NCU -k regex:k_demod -s 2 -c 1 -f -o gpurun_out/r1_k2 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k2.log 2>&1
timeout 300 $NCU -k regex:k_block_fec -s 2 -c 1 -f -o gpurun_out/r1_k3 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k3.log 2>&1
timeout 300 $NCU -k regex:k_channelize -s 3 -c 1 -f -o gpurun_out/r1_k1_real python tools/bench_air.py 2500000 296 8 > gpurun_out/ncu_k1_real.log 2>&1
timeout 300 $NCU -k regex:k_channelize -s 3 -c 1 -f -o gpurun_out/r1_k1_cs16 python tools/bench_cs16.py 0 296 8 > gpurun_out/ncu_k1_cs16.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref_n1.json
python bench.py --steps 20 --warmup 3 | tail -1 > gpurun_out/bench_n1.json
python tools/bench_k1.py fast | tail -1 > gpurun_out/bench_k1_fast.json
python tools/bench_k1.py exact | tail -1 > gpurun_out/bench_k1_exact.json
python bench.py --steps 20 --warmup 3 --channelizer fast --streams 2368 --no-e2e --no-alt --no-cpu-baseline | tail -1 > gpurun_out/bench_n1_fast_s2368.json
python tools/bench_cs16.py 0 | tail -1 > gpurun_out/bench_cs16_soapy.json
python tools/bench_cs16.py 1 | tail -1 > gpurun_out/bench_cs16_sdrplay.json
python tools/bench_air.py 2500000 296 8 | tail -1 > gpurun_out/bench_air_c8.json
cat gpurun_out/bench_k1_fast.json gpurun_out/bench_n1_fast_s2368.json
ls -la gpurun_out
