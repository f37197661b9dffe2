mkdir -p gpurun_out
for W in B C; do
timeout 600 python bench.py --workload $W --steps 100 --warmup 10 --repeats 5 --no-policy --no-train --no-dropin --no-cpu-baseline > gpurun_out/r3y_bench_$W.json 2> gpurun_out/r3y_bench_$W.err
python - <<PY
import json
d=json.load(open('gpurun_out/r3y_bench_$W.json'))
print('$W', d['config']['workload'], 'value %.3e e2e %.3e (%.1f us/step, %d B/env back)' % (d['value'], d['e2e']['value'], d['e2e']['us_per_step'], d['e2e']['d2h_bytes_per_step'] // d['config']['envs_per_gpu']))
PY
done
