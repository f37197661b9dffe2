set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2p_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2p_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_ref.json 2> gpurun_out/r2p_ref.err; echo "ref rc=$?"
NCU="ncu --set full --clock-control none --import-source on -k regex:env_kernel_s"
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r2p_A_random -f python profiles/prof_step.py A 16 random > gpurun_out/r2p_ncu1.log 2>&1
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r2p_A_replay -f python profiles/prof_step.py A 16 replay > gpurun_out/r2p_ncu2.log 2>&1
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r2p_A_step -f python profiles/prof_step.py A 16 step > gpurun_out/r2p_ncu3.log 2>&1
