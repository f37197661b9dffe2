set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder.py -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2w_pytest.log
timeout 300 python profiles/prof_encoder.py 65536 tf32 5 > gpurun_out/r2w_enc.log 2>&1; tail -1 gpurun_out/r2w_enc.log
MTFJSP_FUSED_HEAD=0 timeout 300 python profiles/prof_encoder.py 65536 tf32 5 > gpurun_out/r2w_enc0.log 2>&1; tail -1 gpurun_out/r2w_enc0.log
