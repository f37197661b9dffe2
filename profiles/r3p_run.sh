set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_small.py all > gpurun_out/r3p_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 5 gpurun_out/r3p_memcheck.log
MTFJSP_TRUNK_FORM=1 timeout 600 compute-sanitizer --tool memcheck python profiles/sanitize_small.py enc > gpurun_out/r3p_memcheck_form1.log 2>&1; echo "memcheck form1 rc=$?"; tail -n 3 gpurun_out/r3p_memcheck_form1.log
timeout 900 compute-sanitizer --tool racecheck python profiles/sanitize_small.py env > gpurun_out/r3p_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 3 gpurun_out/r3p_racecheck.log; grep -c "hazard" gpurun_out/r3p_racecheck.log
