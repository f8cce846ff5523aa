# one gpurun job (1 GPU): the demod's two register budgets: parity over every lane code, ncu at 592 streams (wide schedule)
# and at 2368 streams x 4 lanes (lean schedule, 16 warps per SM)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -q -x > gpurun_out/r2_pytest_lanes.log 2>&1; tail -3 gpurun_out/r2_pytest_lanes.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_demod2 -s 2 -c 1 -f -o gpurun_out/r2_k2 python tools/prof_run.py 592 16 4 > gpurun_out/ncu_k2.log 2>&1
ACB_DEMOD_LANES=4 timeout 400 $NCU -k regex:k_demod2 -s 1 -c 1 -f -o gpurun_out/r2_k2_sat python tools/prof_run.py 2368 16 3 > gpurun_out/ncu_k2_sat.log 2>&1
timeout 300 python tools/ab_demod.py 592,2368 4,1 fast > gpurun_out/r2_ab9.jsonl 2>/dev/null; cat gpurun_out/r2_ab9.jsonl
