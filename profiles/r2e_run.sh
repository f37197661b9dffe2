set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_single_env.py tests/test_gae_buffer.py tests/test_ppo.py tests/test_parallel_env_dropin.py -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2e_pytest.log
NCU="ncu --set full --clock-control none --import-source on -k regex:env_kernel_s"
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r2e_A_random -f python profiles/prof_step.py A 16 random > gpurun_out/r2e_ncu1.log 2>&1
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r2e_A_step -f python profiles/prof_step.py A 16 step > gpurun_out/r2e_ncu3.log 2>&1
timeout 300 $NCU -s 200 -c 1 -o gpurun_out/r2e_C_random -f python profiles/prof_step.py C 210 random > gpurun_out/r2e_ncu5.log 2>&1
ls -la gpurun_out/
