set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_small.py all > gpurun_out/r2o_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 5 gpurun_out/r2o_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python profiles/sanitize_small.py env > gpurun_out/r2o_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 5 gpurun_out/r2o_racecheck.log; grep -c "hazard" gpurun_out/r2o_racecheck.log
timeout 600 compute-sanitizer --tool initcheck python profiles/sanitize_small.py env > gpurun_out/r2o_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -n 4 gpurun_out/r2o_initcheck.log
