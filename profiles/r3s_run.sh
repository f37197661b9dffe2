set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 4 --steps 72 --warmup 5 --repeats 5 --no-dropin --no-cpu-baseline --no-workloads > gpurun_out/r3s_bench_4gpu.json 2> gpurun_out/r3s_bench_4gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3s_bench_4gpu.json'))
print('n_gpus', d['n_gpus'], 'value %.3e e2e %.3e (%.1f us)' % (d['value'], d['e2e']['value'], d['e2e']['us_per_step']))
print('policy %.3e train %.3e collect %.1f update %.1f allreduce share %s' % (d['policy_rollout']['value'], d['train_iteration']['value'], d['train_iteration']['collect_ms'], d['train_iteration']['update_ms'], d['train_iteration']['allreduce_share']))
PY
