# ncu evidence of a round (run on the GPU box under gpurun): launch list of two bench steps + of one native
# planner run + full captures of the dominant kernels.  Summaries are extracted in the build container with
# `python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/y.txt`.
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r02b}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-plan-latency --no-cpu-baseline --no-extras --no-cuda-graph > gpurun_out/${T}_launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_plan.csv python tools/bench_c3_native.py 64 --no-python --once > gpurun_out/${T}_launches_plan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_iris_fused -s 2 -c 1 -f -o gpurun_out/${T}_prof_iris python tools/ncu_driver.py > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_iris_fused -s 4 -c 1 -f -o gpurun_out/${T}_prof_iris_sat python tools/ncu_driver.py > gpurun_out/${T}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_lp -s 1 -c 1 -f -o gpurun_out/${T}_prof_pair_lp python tools/ncu_driver.py > gpurun_out/${T}_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:k_pair_filter -s 1 -c 1 -f -o gpurun_out/${T}_prof_pair_filter python tools/ncu_driver.py > gpurun_out/${T}_ncu4.log 2>&1
ncu --set full --clock-control none -k regex:k_fk -s 1 -c 1 -f -o gpurun_out/${T}_prof_fk python tools/ncu_driver.py > gpurun_out/${T}_ncu5.log 2>&1
tail -3 gpurun_out/${T}_ncu1.log
ls -la gpurun_out/${T}_*.ncu-rep
