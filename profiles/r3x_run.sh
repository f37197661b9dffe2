set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3x_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r3x_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r3x_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3x_bench.json 2> gpurun_out/r3x_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r3x_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3x_ref.json 2> gpurun_out/r3x_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3x_ro_launches.csv python profiles/prof_rollout.py 65536 2 > gpurun_out/r3x_ncu0.log 2>&1
