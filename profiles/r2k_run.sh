set -x
mkdir -p gpurun_out
export MTFJSP_LIB=$PWD/e2e-mappo-for-mt-fjsp_b200/build/libmtfjsp_b200_v8.so
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
import ctypes
rt=ctypes.CDLL("libcudart.so.12")
for name,attr in (("maxPersistingL2",108),("maxAccessPolicyWindow",109)):
    v=ctypes.c_int(); rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print(name, v.value)
PY
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "not replay and not 30 and not 15 and not 20" > gpurun_out/r2k_pytest_v8.log 2>&1; echo "pytest v8 rc=$?"; tail -n 3 gpurun_out/r2k_pytest_v8.log
for lp in 2 1; do
MTFJSP_L2_PERSIST=$lp timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r2k_bench_lp$lp.json 2> gpurun_out/r2k_bench_lp$lp.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2k_bench_lp$lp.err
done
