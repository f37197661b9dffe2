set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 profiles/prof_pcie.py > gpurun_out/r2m_pcie_8gpu.log 2>&1; tail -8 gpurun_out/r2m_pcie_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 --no-policy --no-train > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/r2m_bench_8gpu.err; wc -c gpurun_out/r2m_bench_8gpu.json
