"""SASS evidence for profiles/: per kernel of libbpgeo.so the instruction mix that shows what the hot loops are made
of (DFMA / DMUL / DADD fp64 pipe, MUFU.RCP64H / RSQ64H seeds, REDUX warp reductions, SHFL, BAR, UBLKCP = TMA bulk
copies, SYNCS = mbarrier, LDS/STS, ATOMS) plus the lines around the first UBLKCP of the kernels that use TMA.

  python tools/sass_evidence.py > profiles/r02_sass_summary.txt        (no GPU needed: cuobjdump on the built .so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "boundplanner_b200", "libbpgeo.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, name = collections.OrderedDict(), None
for line in txt.splitlines():
    mt = re.search(r"Function : (\S+)", line)
    if mt:
        name = subprocess.run(["c++filt", mt.group(1)], capture_output=True, text=True).stdout.strip()
        kern[name] = []
        continue
    mt = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if mt and name:
        kern[name].append(mt.group(2))
KEYS = ["DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H", "REDUX", "SHFL", "VOTE", "BAR.SYNC", "UBLKCP",
        "SYNCS", "LDS", "STS", "ATOMS", "LDG", "STG", "ST.E", "LDL", "STL", "CALL"]
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}: instruction mix per kernel (static counts)")
for name, ins in kern.items():
    short = re.sub(r"\(.*", "", name)
    if not any(k in short for k in ("k_iris_fused<0, false, 1>", "k_fk<", "k_pair_lp", "k_mvie(", "k_poly_point<false>",
                                    "k_pair_filter", "bp_mvie_warp_fn", "k_fk_kin")):
        continue
    cnt = collections.Counter()
    for i in ins:
        op = i.split()[0] if not i.startswith("@") else i.split()[1]
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
    print(f"\n{short}: {len(ins)} instructions ({len(ins) * 16 // 1024} KB)")
    print("   " + "  ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
    for k in ("UBLKCP", "REDUX", "MUFU.RSQ64H"):
        hit = [n for n, i in enumerate(ins) if k in i]
        if hit and ("k_fk<false, false>" in short or "k_iris_fused<0, false, 1>" in short):
            n0 = hit[0]
            print(f"   first {k} (instruction {n0}):")
            for i in ins[max(0, n0 - 2): n0 + 3]:
                print("      " + i)
