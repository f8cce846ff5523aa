# one gpurun job (1 GPU): BASELINE configs[4]'s capture — the generalised-FIR channelizer at K = 1600, 256 channels, taps 65..513:
# kernel times (CUDA events) and one ncu --set full capture at 513 taps
set -x
mkdir -p gpurun_out
python tools/bench_fir.py | tail -1 > gpurun_out/bench_fir.json; cat gpurun_out/bench_fir.json | cut -c1-1500
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_channelize -s 17 -c 1 -f -o gpurun_out/r2_k1_fir python tools/bench_fir.py > gpurun_out/ncu_k1_fir.log 2>&1; tail -2 gpurun_out/ncu_k1_fir.log
