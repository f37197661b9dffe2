set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_single_env.py tests/test_ppo.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2c_pytest.log
timeout 1200 python profiles/checkpoint_parity.py > gpurun_out/r2c_ckpt.log 2>&1; echo "ckpt rc=$?"; tail -50 gpurun_out/r2c_ckpt.log
