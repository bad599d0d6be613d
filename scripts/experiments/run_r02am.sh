python scripts/sanitize_smoke.py 2>&1 | tail -1
timeout 800 compute-sanitizer --tool memcheck --kernel-regex kns=small_ --error-exitcode 1 python scripts/sanitize_smoke.py > gpurun_out/r02am_memcheck.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/r02am_memcheck.log
DICP_SMALL_MID_R=4 timeout 800 compute-sanitizer --tool racecheck --kernel-regex kns=small_ --error-exitcode 1 python scripts/sanitize_smoke.py > gpurun_out/r02am_racecheck.log 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/r02am_racecheck.log
