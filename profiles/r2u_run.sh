set -x
mkdir -p gpurun_out
timeout 300 python profiles/prof_encoder.py 65536 tf32 5 > gpurun_out/r2u_enc.log 2>&1; tail -2 gpurun_out/r2u_enc.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_enc_launches.csv python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r2u_ncu0.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tf32_kernel -s 6 -c 1 -o gpurun_out/r2u_gemm128 -f python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r2u_ncu1.log 2>&1; tail -2 gpurun_out/r2u_ncu1.log
