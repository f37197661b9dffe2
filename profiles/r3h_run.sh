set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:linear_tf32_tma_kernel -s 1 -c 1 -o gpurun_out/r3h_tma -f python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r3h_ncu1.log 2>&1
timeout 300 $NCU -k regex:linear_tf32_tma_kernel -s 2 -c 1 -o gpurun_out/r3h_agg -f python profiles/prof_encoder.py 65536 tf32 1 > gpurun_out/r3h_ncu2.log 2>&1
timeout 300 $NCU -k regex:gat_trunk -s 1 -c 1 -o gpurun_out/r3h_trunk -f python profiles/prof_rollout.py 65536 1 > gpurun_out/r3h_ncu3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3h_ro_launches.csv python profiles/prof_rollout.py 65536 2 > gpurun_out/r3h_ncu0.log 2>&1
