set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tf32_tma_kernel -s 2 -c 1 -o gpurun_out/r3d_agg -f python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r3d_ncu1.log 2>&1; tail -2 gpurun_out/r3d_ncu1.log
