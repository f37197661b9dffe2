set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder.py -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2x_pytest.log
timeout 300 python profiles/prof_rollout.py 65536 6 > gpurun_out/r2x_ro.log 2>&1; tail -1 gpurun_out/r2x_ro.log
MTFJSP_FUSED_TRUNK=0 timeout 300 python profiles/prof_rollout.py 65536 6 > gpurun_out/r2x_ro0.log 2>&1; tail -1 gpurun_out/r2x_ro0.log
MTFJSP_FUSED_TRUNK=0 MTFJSP_FUSED_HEAD=0 timeout 300 python profiles/prof_rollout.py 65536 6 > gpurun_out/r2x_ro00.log 2>&1; tail -1 gpurun_out/r2x_ro00.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_ro_launches.csv python profiles/prof_rollout.py 65536 2 > gpurun_out/r2x_ncu0.log 2>&1
