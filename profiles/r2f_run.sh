set -x
mkdir -p gpurun_out
for v in v3 v3b; do
export MTFJSP_LIB=$PWD/e2e-mappo-for-mt-fjsp_b200/build/libmtfjsp_b200_$v.so
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q > gpurun_out/r2f_pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -4 gpurun_out/r2f_pytest_$v.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r2f_bench_$v.json 2> gpurun_out/r2f_bench_$v.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2f_bench_$v.err
done
