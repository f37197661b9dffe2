set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r3g_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r3g_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r3g_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3g_ref.json 2> gpurun_out/r3g_ref.err; echo "ref rc=$?"
