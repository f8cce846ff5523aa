# 1-GPU job: full GPU suite on the latest kernels, lanes/pinning sweep, compute-sanitizer memcheck + racecheck
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest5.log 2>&1; tail -6 gpurun_out/r2_pytest5.log
timeout 600 python tools/ab_demod.py 592 8,4,40,36 fast > gpurun_out/r2_ab5.jsonl 2> gpurun_out/r2_ab5.err
timeout 600 python tools/ab_demod.py 2368,4736 1,33,2,34,4,36 fast >> gpurun_out/r2_ab5.jsonl 2>> gpurun_out/r2_ab5.err
timeout 300 python tools/ab_demod.py 4736 1,33 exact >> gpurun_out/r2_ab5.jsonl 2>> gpurun_out/r2_ab5.err
cat gpurun_out/r2_ab5.jsonl
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
ACB_DEMOD_LANES=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_racecheck_l1.log 2>&1; tail -3 gpurun_out/r2_sanitizer_racecheck_l1.log
