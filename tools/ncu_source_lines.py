"""Stall samples of a kernel by SOURCE LINE, from an ncu report with --import-source on (run here, no GPU).
The SASS page of the report is joined with nvdisasm's line table of the same library build (order of instructions
is the key), barrier stalls -- warps parked at __syncthreads while another warp of the CTA runs a one-warp solver
-- are taken out, and the lines are ranked.  This is how the dependent-chain hot spots of the Newton solve were
found (bp_mvie.cuh back substitution / pivots, bp_mvie_warp.cuh log1p of the Armijo test, predictor divisions).

  python tools/ncu_source_lines.py gpurun_out/x.ncu-rep '_Z12k_iris_fusedILi0ELb0ELi1EEv9SceneView11FusedParams' [lib.so] [top]"""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sym = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "boundplanner_b200", "libbpgeo.so")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(f".text.{sym}:")][0]
cur, seq = None, []
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
        seq.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((int(r[ix["# Samples"]]) - int(r[ix["stall_barrier"]]), int(r[ix["Instructions Executed"]])))
    except ValueError:
        continue
if len(data) != len(seq):
    sys.exit(f"instruction counts differ (report {len(data)}, library {len(seq)}): the report is of another build")
agg, inst = collections.Counter(), collections.Counter()
for key, (smp, n) in zip(seq, data):
    agg[key or ("?", 0)] += smp
    inst[key or ("?", 0)] += n
tot = sum(agg.values())
byfile = collections.Counter()
for key, v in agg.items():
    byfile[key[0]] += v
print(f"# {os.path.basename(rep)}: {tot} stall samples outside barriers; by file:", ", ".join(f"{f} {100 * v / tot:.1f}%" for f, v in byfile.most_common(6)))
print(f"{'file:line':34s} {'samples':>8s} {'share':>7s} {'warp instr':>11s}  source")
src_cache = {}
for key, v in agg.most_common(top):
    f, ln = key
    text = ""
    for d in ("boundplanner_b200/csrc",):
        pth = os.path.join(ROOT, d, f)
        if os.path.exists(pth):
            src_cache.setdefault(pth, open(pth).read().splitlines())
            if 0 < ln <= len(src_cache[pth]):
                text = src_cache[pth][ln - 1].strip()[:90]
    print(f"{f + ':' + str(ln):34s} {v:8d} {100 * v / tot:6.1f}% {inst[key]:11d}  {text}")
