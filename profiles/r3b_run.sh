set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_encoder.py tests/test_ppo.py tests/test_rollout.py tests/test_checkpoint_rollout.py -m gpu -x -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3b_pytest.log
timeout 900 python bench.py --steps 72 --warmup 5 --repeats 3 --no-dropin --no-cpu-baseline --no-workloads > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r3b_bench.json'))
print('value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))
print('policy %.3e (%.2f ms) train %.3e collect %.1f update %.1f fp32 %.3e' % (d['policy_rollout']['value'], d['policy_rollout']['ms_per_step'], d['train_iteration']['value'], d['train_iteration']['collect_ms'], d['train_iteration']['update_ms'], d['train_iteration']['value_library_fp32']))
PY
