set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "host_step or packed_host" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2r_pytest.log
timeout 600 python profiles/prof_zerocopy.py A > gpurun_out/r2r_zc_A.log 2>&1; cat gpurun_out/r2r_zc_A.log | tail -8
timeout 600 python profiles/prof_zerocopy.py C > gpurun_out/r2r_zc_C.log 2>&1; cat gpurun_out/r2r_zc_C.log | tail -8
timeout 600 python bench.py --steps 72 --warmup 5 --repeats 5 --no-policy --no-train --no-dropin --no-cpu-baseline --no-workloads > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print('value %.3e e2e %.3e (%.1f us) info-only %.3e' % (d['value'], d['e2e']['value'], d['e2e']['us_per_step'], d['e2e']['step_info_only']['value']))
print('random %.1f step_obs %.1f step_only %.1f' % (d['roofline']['kernel_us'], d['roofline']['step_obs_kernel']['kernel_us'], d['roofline']['step_only']['kernel_us']))
PY
