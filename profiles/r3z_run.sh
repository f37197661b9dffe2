set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:env_kernel_s"
timeout 300 $NCU -s 10 -c 1 -o gpurun_out/r3z_A_random -f python profiles/prof_step.py A 16 random > gpurun_out/r3z_ncu1.log 2>&1
timeout 300 $NCU -s 40 -c 1 -o gpurun_out/r3z_B_random -f python profiles/prof_step.py B 48 random > gpurun_out/r3z_ncu2.log 2>&1
