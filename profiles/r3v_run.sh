set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cuda_parity.py -m gpu -x -q > gpurun_out/r3v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3v_pytest.log
timeout 600 python bench.py --workload B --steps 100 --warmup 10 --repeats 7 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r3v_bench_B.json 2> gpurun_out/r3v_bench_B.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r3v_bench_B.json'))
r=d['roofline']
print('B value %.3e random %.1f us step_obs %.1f step_only %.1f' % (d['value'], r['kernel_us'], r['step_obs_kernel']['kernel_us'], r['step_only']['kernel_us']))
PY
