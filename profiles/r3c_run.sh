set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3c_ro_launches.csv python profiles/prof_rollout.py 65536 2 > gpurun_out/r3c_ncu0.log 2>&1
