set -x
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_instances.py -m gpu -q 2>&1 | tail -n 3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 600 gpurun_out/r2l_bench_2gpu.err; wc -c gpurun_out/r2l_bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_ref_2gpu.json 2> gpurun_out/r2l_ref_2gpu.err; echo "ref2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 profiles/prof_pcie.py > gpurun_out/r2l_pcie_2gpu.log 2>&1; tail -8 gpurun_out/r2l_pcie_2gpu.log
