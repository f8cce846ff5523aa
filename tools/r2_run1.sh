# round-2 GPU job 1: correctness of the rewritten demod kernel at every lane width, lanes x streams sweep, one ncu capture
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; tail -5 gpurun_out/r2_pytest1.log
timeout 900 python tools/ab_demod.py 592,1184,2368,4736 -4,-8,8,4,2,1,24,20,18,17 fast > gpurun_out/r2_ab_demod_fast.jsonl 2> gpurun_out/r2_ab_demod_fast.err
timeout 300 python tools/ab_demod.py 592,2368 -4,4,2,1 exact > gpurun_out/r2_ab_demod_exact.jsonl 2> gpurun_out/r2_ab_demod_exact.err
ACB_DEMOD_LANES=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_demod2 -s 2 -c 1 -f -o gpurun_out/r2_k2_l4 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k2_l4.log 2>&1
cat gpurun_out/r2_ab_demod_fast.jsonl | head -60
tail -3 gpurun_out/r2_ab_demod_fast.err gpurun_out/r2_ab_demod_exact.err
