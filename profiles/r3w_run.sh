mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "golden or replay" > gpurun_out/r3w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3w_pytest.log
timeout 600 python bench.py --steps 72 --warmup 5 --repeats 7 --no-policy --no-train --no-dropin --no-cpu-baseline --no-workloads > gpurun_out/r3w_bench.json 2> gpurun_out/r3w_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r3w_bench.json'))
r=d['roofline']
print('value %.3e e2e %.3e (%.1f us) random %.1f step_obs %.1f step_only %.1f' % (d['value'], d['e2e']['value'], d['e2e']['us_per_step'], r['kernel_us'], r['step_obs_kernel']['kernel_us'], r['step_only']['kernel_us']))
PY
