# ncu evidence of a round (run on the GPU box under gpurun): launch list of two bench steps + full captures of
# the dominant kernels.  Summaries are extracted here with `ncu -i ... --page raw --csv` and kept under profiles/.
cd ${GRAFT_REPO_ROOT:-.}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-plan-latency --no-cpu-baseline --no-extras --no-cuda-graph > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_iris_fused -s 2 -c 1 -f -o gpurun_out/r02_prof_iris python tools/ncu_driver.py > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_iris_fused -s 4 -c 1 -f -o gpurun_out/r02_prof_iris_sat python tools/ncu_driver.py > gpurun_out/r02_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"k_pair_lp|k_pair_filter|k_fk" -s 6 -c 3 -f -o gpurun_out/r02_prof_pair_fk python tools/ncu_driver.py > gpurun_out/r02_ncu3.log 2>&1
tail -3 gpurun_out/r02_ncu1.log
ls -la gpurun_out/*.ncu-rep
