set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tf32_kernel -s 1 -c 1 -o gpurun_out/r2z_gemm128 -f python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r2z_ncu1.log 2>&1; tail -2 gpurun_out/r2z_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_tf32 -s 1 -c 1 -o gpurun_out/r2z_head -f python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r2z_ncu2.log 2>&1; tail -2 gpurun_out/r2z_ncu2.log
