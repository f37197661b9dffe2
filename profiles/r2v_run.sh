set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 72 --warmup 5 --repeats 5 --no-policy --no-train --no-dropin --no-cpu-baseline --no-workloads > gpurun_out/r2v_bench_8gpu.json 2> gpurun_out/r2v_bench_8gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2v_bench_8gpu.json'))
print('n_gpus', d['n_gpus'], 'value %.3e e2e %.3e (%.1f us) info-only %.3e (%.1f us)' % (d['value'], d['e2e']['value'], d['e2e']['us_per_step'], d['e2e']['step_info_only']['value'], d['e2e']['step_info_only']['us_per_step']))
PY
