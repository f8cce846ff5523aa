# one gpurun job (1 GPU): compute-sanitizer memcheck + racecheck over tools/sanitize_run.py (every kernel incl. the fast forms,
# three submits in flight), and the stress test under racecheck
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
ACB_DEMOD_LANES=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_racecheck_l1.log 2>&1; tail -3 gpurun_out/r2_sanitizer_racecheck_l1.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_stress.py -m gpu -q -x > gpurun_out/r2_sanitizer_stress_memcheck.log 2>&1; tail -5 gpurun_out/r2_sanitizer_stress_memcheck.log
