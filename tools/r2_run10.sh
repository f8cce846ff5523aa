# 1-GPU job: one-row-per-lane fast channelizer — parity, kernel timing (2 and 4 warps per CTA), pipelined step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_air.py -q > gpurun_out/r2_pytest10.log 2>&1; tail -5 gpurun_out/r2_pytest10.log
for cfg in "2 0" "1 4" "1 2"; do set -- $cfg
  ACB_FAST_ROWS=$1 ACB_FAST_WARPS=$2 python tools/bench_k1.py fast 592 16 | tail -1
  ACB_FAST_ROWS=$1 ACB_FAST_WARPS=$2 python tools/bench_k1.py fast 4736 8 | tail -1
  ACB_FAST_ROWS=$1 ACB_FAST_WARPS=$2 timeout 300 python tools/ab_demod.py 592,4736 4,1 fast 2>/dev/null
done > gpurun_out/r2_dft1.jsonl
cat gpurun_out/r2_dft1.jsonl
NCU="ncu --set full --clock-control none --import-source on"
ACB_FAST_ROWS=1 ACB_FAST_WARPS=4 timeout 300 $NCU -k regex:k_channelize_dft1 -s 3 -c 1 -f -o gpurun_out/r2_k1_fast1 python tools/bench_k1.py fast > gpurun_out/ncu_k1_fast1.log 2>&1
