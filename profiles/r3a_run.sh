set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder.py -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3a_pytest.log
timeout 300 python profiles/prof_encoder.py 65536 tf32 5 > gpurun_out/r3a_enc.log 2>&1; tail -1 gpurun_out/r3a_enc.log
timeout 300 python profiles/prof_rollout.py 65536 6 > gpurun_out/r3a_ro.log 2>&1; tail -1 gpurun_out/r3a_ro.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3a_enc_launches.csv python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r3a_ncu0.log 2>&1
grep "linear_tf32_kernel<128>" gpurun_out/r3a_enc_launches.csv | tail -12 | awk -F'","' '{print $NF}'
