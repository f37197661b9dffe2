set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_cuda_parity.py > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
NCU="ncu --set full --clock-control none --import-source on -k regex:env_kernel_s"
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r2a_A_random -f python profiles/prof_step.py A 16 random > gpurun_out/r2a_ncu1.log 2>&1
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r2a_A_step -f python profiles/prof_step.py A 16 step > gpurun_out/r2a_ncu3.log 2>&1
timeout 300 $NCU -s 200 -c 1 -o gpurun_out/r2a_C_random -f python profiles/prof_step.py C 210 random > gpurun_out/r2a_ncu5.log 2>&1
ls -la gpurun_out/
