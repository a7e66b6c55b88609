#!/bin/bash
# A/B of library build variants on one GPU: for every boundplanner_b200/libbpgeo_<name>.so given, the golden-graph
# parity tests and a short bench without the extras.  Usage: bash tools/ab_libs.sh base gap9 ...
for name in "$@"; do
  if [ "$name" = base ]; then lib=$PWD/boundplanner_b200/libbpgeo.so; else lib=$PWD/boundplanner_b200/libbpgeo_$name.so; fi
  echo "=== $name"
  BPGEO_LIB=$lib python -m pytest tests/test_gpu_golden_graph.py tests/test_gpu_sets.py -x -q -k "not sweep" 2>&1 | tail -3
  BPGEO_LIB=$lib python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline --no-plan-latency > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
for l in open("gpurun_out/ab_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", "ms/step", round(d["ms_per_step"], 4), "stages", {k: round(v, 4) for k, v in d["stages_ms"].items()}, "newton", d["kernels"].get("newton_iters_fixed_mid"), d["kernels"].get("k_mvie_fixed_mid_ms"))
PY
done
