mkdir -p gpurun_out
s=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r3u_bench_default.json 2> gpurun_out/r3u_bench_default.err; echo rc=$?
e=$(date +%s); echo "default bench.py wall: $((e-s)) s"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3u_bench_default.json"))
print(d["steps"], d["warmup"], "value %.3e e2e %.3e policy %.3e train %.3e update %.1f" % (d["value"], d["e2e"]["value"], d["policy_rollout"]["value"], d["train_iteration"]["value"], d["train_iteration"]["update_ms"]))
PY
