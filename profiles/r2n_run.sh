set -x
mkdir -p gpurun_out
export MTFJSP_LIB=$PWD/e2e-mappo-for-mt-fjsp_b200/build/libmtfjsp_b200_v11.so
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "host" > gpurun_out/r2n_pytest_v11.log 2>&1; echo "pytest v10 rc=$?"; tail -n 3 gpurun_out/r2n_pytest_v11.log
MTFJSP_HOST_PIPE=1 timeout 300 python profiles/prof_host_step.py A 2>&1 | tail -10
MTFJSP_HOST_PIPE=0 timeout 300 python profiles/prof_host_step.py A 2>&1 | head -6
