set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --scaling strong --steps 72 --warmup 5 --repeats 5 --no-dropin --no-cpu-baseline --no-workloads --no-train > gpurun_out/r3r_bench_8gpu_strong.json 2> gpurun_out/r3r_bench_8gpu_strong.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3r_bench_8gpu_strong.json'))
print('n_gpus', d['n_gpus'], d['scaling'], d['config'], 'value %.3e e2e %.3e (%.1f us) policy %.3e' % (d['value'], d['e2e']['value'], d['e2e']['us_per_step'], d['policy_rollout']['value']))
PY
