set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder.py -m gpu -x -q > gpurun_out/r3k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3k_pytest.log
MTFJSP_TRUNK_FORM=1 timeout 600 python -m pytest tests/test_encoder.py -m gpu -x -q -k trunk > gpurun_out/r3k_pytest1.log 2>&1; echo "pytest form1 rc=$?"; tail -1 gpurun_out/r3k_pytest1.log
timeout 300 python profiles/prof_rollout.py 65536 6 > gpurun_out/r3k_ro.log 2>&1; tail -1 gpurun_out/r3k_ro.log
MTFJSP_TRUNK_FORM=1 timeout 300 python profiles/prof_rollout.py 65536 6 > gpurun_out/r3k_ro1.log 2>&1; tail -1 gpurun_out/r3k_ro1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3k_ro_launches.csv python profiles/prof_rollout.py 65536 2 > gpurun_out/r3k_ncu0.log 2>&1
grep "gat_trunk" gpurun_out/r3k_ro_launches.csv | tail -2 | awk -F'","' '{print $NF}'
