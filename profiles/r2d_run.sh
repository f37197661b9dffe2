set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_single_env.py tests/test_ppo.py tests/test_checkpoint_rollout.py -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2d_pytest.log
export MTFJSP_LIB=$PWD/e2e-mappo-for-mt-fjsp_b200/build/libmtfjsp_b200_v2.so
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_rules.py tests/test_rollout.py tests/test_parallel_env_dropin.py -m gpu -x -q > gpurun_out/r2d_pytest_v2.log 2>&1; echo "pytest v2 rc=$?"; tail -30 gpurun_out/r2d_pytest_v2.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r2d_bench_v2.json 2> gpurun_out/r2d_bench_v2.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2d_bench_v2.err
