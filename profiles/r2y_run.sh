set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2y_pytest.log
timeout 300 python profiles/prof_rollout.py 65536 6 > gpurun_out/r2y_ro.log 2>&1; tail -1 gpurun_out/r2y_ro.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_ro_launches.csv python profiles/prof_rollout.py 65536 2 > gpurun_out/r2y_ncu0.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tf32_kernelILi128 -s 1 -c 1 -o gpurun_out/r2y_gemm128 -f python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r2y_ncu1.log 2>&1; tail -2 gpurun_out/r2y_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gat_trunk -s 1 -c 1 -o gpurun_out/r2y_trunk -f python profiles/prof_rollout.py 65536 1 > gpurun_out/r2y_ncu2.log 2>&1; tail -2 gpurun_out/r2y_ncu2.log
