# bench.py on N GPUs of one box the way the driver launches it: bash tools/run_n.sh N [tag]
N=${1:-8}; T=${2:-n$N}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 3 --no-cpu-baseline --no-plan-latency > gpurun_out/bench_r2_s2_$T.json 2> gpurun_out/bench_r2_s2_$T.err
tail -c 300 gpurun_out/bench_r2_s2_$T.err
python - <<PY
import json
for l in open("gpurun_out/bench_r2_s2_$T.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("adjacency_equals_single_rank"))
        print({k:(round(v["max"],4),round(v["min"],4)) for k,v in d["stages_ms"].items()})
        c3=d["c3"]; print("c3", c3["plan_queries_per_sec"], c3["plan_queries_per_sec_incl_scene_ingest"], c3["latency_ms_p50"], c3["python_driver"]["results_identical_to_native"])
        print("c4", d["c4"]["ms_per_step"], d["c4"]["stages_ms"])
PY
