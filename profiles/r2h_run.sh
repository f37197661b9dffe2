set -x
mkdir -p gpurun_out
export MTFJSP_LIB=$PWD/e2e-mappo-for-mt-fjsp_b200/build/libmtfjsp_b200_v5.so
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "30 or 15 or 20 or replay" > gpurun_out/r2h_pytest_v5.log 2>&1; echo "pytest v5 rc=$?"; tail -n 4 gpurun_out/r2h_pytest_v5.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r2h_bench_v5.json 2> gpurun_out/r2h_bench_v5.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2h_bench_v5.err
NCU="ncu --set full --clock-control none --import-source on -k regex:env_kernel_s"
timeout 300 $NCU -s 200 -c 1 -o gpurun_out/r2h_C_random -f python profiles/prof_step.py C 210 random > gpurun_out/r2h_ncu5.log 2>&1
