timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 50 --warmup 3 --no-cpu-baseline --no-plan-latency > gpurun_out/bench_r2_s2_n8.json 2> gpurun_out/bench_r2_s2_n8.err
tail -c 300 gpurun_out/bench_r2_s2_n8.err
python - <<PY
import json
for l in open("gpurun_out/bench_r2_s2_n8.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], d["stages_ms"], d.get("adjacency_equals_single_rank")); print(json.dumps(d["c3"])[:900]); print(json.dumps(d["c4"])[:700])
PY
