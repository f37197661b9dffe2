set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
for ao in 1 0; do
MTFJSP_ALTERNATE_ORDER=$ao timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r2b_bench_ao$ao.json 2> gpurun_out/r2b_bench_ao$ao.err; echo "bench rc=$?"
done
