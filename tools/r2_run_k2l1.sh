# one gpurun job: ncu capture of the demod kernel at one lane per channel, 4736 streams x 8 blocks (the bench's saturating point)
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
ACB_DEMOD_LANES=1 timeout 600 $NCU -k regex:k_demod2 -s 1 -c 1 -f -o gpurun_out/r2_k2_l1 python tools/prof_run.py 4736 8 3 > gpurun_out/ncu_k2_l1.log 2>&1
tail -3 gpurun_out/ncu_k2_l1.log
