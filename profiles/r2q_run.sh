set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:env_kernel_s"
timeout 300 $NCU -s 40 -c 1 -o gpurun_out/r2q_B_random -f python profiles/prof_step.py B 48 random > gpurun_out/r2q_ncu4.log 2>&1
timeout 300 $NCU -s 200 -c 1 -o gpurun_out/r2q_C_random -f python profiles/prof_step.py C 210 random > gpurun_out/r2q_ncu5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches_bench.csv python bench.py --steps 72 --warmup 3 --repeats 1 --no-policy --no-train --no-dropin --no-cpu-baseline --no-workloads > gpurun_out/r2q_ncu_b.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2q_launches_bench.csv
