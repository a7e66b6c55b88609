for T in 1 0; do
BPGEO_TAIL=$T timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --no-extras --no-cpu-baseline --no-plan-latency > gpurun_out/bench_r2_s2_n2_tail$T.json 2> gpurun_out/bench_r2_s2_n2_tail$T.err
python - <<PY
import json
for l in open("gpurun_out/bench_r2_s2_n2_tail$T.json"):
    if l.startswith("{"):
        d=json.loads(l); print("TAIL=$T", d["ms_per_step"], d["value"], d.get("stages_ms"), d.get("adjacency_equals_single_rank"))
PY
done
