set -x
mkdir -p gpurun_out
export MTFJSP_LIB=$PWD/e2e-mappo-for-mt-fjsp_b200/build/libmtfjsp_b200_v7.so
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_rollout.py -m gpu -x -q -k "not replay" > gpurun_out/r2j_pytest_v7.log 2>&1; echo "pytest v7 rc=$?"; tail -n 4 gpurun_out/r2j_pytest_v7.log
for lp in 1 0; do
MTFJSP_L2_PERSIST=$lp timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r2j_bench_lp$lp.json 2> gpurun_out/r2j_bench_lp$lp.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2j_bench_lp$lp.err
done
