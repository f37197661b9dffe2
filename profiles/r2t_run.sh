set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cuda_parity.py tests/test_parallel_env_dropin.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest.log
timeout 900 python bench.py --steps 72 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench.json'))
print('value %.3e e2e %.3e (%.1f us) info-only %.3e (%.1f us)' % (d['value'], d['e2e']['value'], d['e2e']['us_per_step'], d['e2e']['step_info_only']['value'], d['e2e']['step_info_only']['us_per_step']))
print('random %.1f step_obs %.1f step_only %.1f' % (d['roofline']['kernel_us'], d['roofline']['step_obs_kernel']['kernel_us'], d['roofline']['step_only']['kernel_us']))
for k,w in d['workloads'].items(): print(k, 'e2e', w.get('e2e',{}).get('value'), w.get('e2e',{}).get('us_per_step'))
print('dropin', d['e2e']['dropin_parallel_env']['value'], d['e2e']['dropin_parallel_env'].get('compat_ell',{}).get('value'))
PY
