"""Aggregate an ncu SASS source page by CUDA source line (development aid).

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    python tools/ncu_lines.py sass.csv gasoline_b200/lib/libgasoline_b200.so k_tree_gravityILi4 [top]

Maps SASS instruction order to source lines with `nvdisasm -g` on the cubin extracted from the library (needs
-lineinfo), then prints per-line executed warp instructions and stall samples."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

sass_csv, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
lines = []
for f in sorted(os.listdir(tmp)):
    if "sm_100a" not in f:
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in out:
        continue
    infn, cur = False, None
    for ln in out.splitlines():
        if ln.startswith(".text."):
            infn = kern in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ci = {n: h.index(n) for n in ("Source", "Instructions Executed", "Warp Stall Sampling (All Samples)",
                              "Thread Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_wait",
                              "stall_not_selected", "stall_math", "stall_branch_resolving")}
body = [r for r in rows[hdr + 1:] if len(r) > ci["Instructions Executed"]]
print(f"{len(body)} SASS rows, {len(lines)} disassembled instructions", file=sys.stderr)
agg = collections.defaultdict(lambda: collections.Counter())
for i, r in enumerate(body):
    key = lines[i] if i < len(lines) else None
    a = agg[key]
    a["inst"] += int(r[ci["Instructions Executed"]] or 0)
    a["thr"] += int(r[ci["Thread Instructions Executed"]] or 0)
    a["samp"] += int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_not_selected", "stall_math", "stall_branch_resolving"):
        a[k] += int(r[ci[k]] or 0)
    a["n"] += 1
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
print(f"total warp inst {ti:.4g}, samples {ts}")
print(f"{'line':>28s} {'sass':>5s} {'inst%':>6s} {'samp%':>6s} {'lanes':>5s}  long short wait notsel math br")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    lanes = a["thr"] / a["inst"] if a["inst"] else 0
    print(f"{str(key):>28s} {a['n']:5d} {100*a['inst']/ti:6.2f} {100*a['samp']/ts:6.2f} {lanes:5.1f}  "
          f"{a['stall_long_sb']:5d} {a['stall_short_sb']:5d} {a['stall_wait']:5d} {a['stall_not_selected']:5d} "
          f"{a['stall_math']:5d} {a['stall_branch_resolving']:5d}")

# ---- optional: totals by line ranges given as name=lo-hi[,lo-hi] arguments after `top`
if len(sys.argv) > 5:
    print("\nregion totals:")
    for spec in sys.argv[5:]:
        name, rngs = spec.split("=")
        rr = [tuple(int(v) for v in x.split("-")) for x in rngs.split(",")]
        tot = collections.Counter()
        for key, a in agg.items():
            if key and key[0].endswith(".cu") and any(lo <= key[1] <= hi for lo, hi in rr):
                tot.update(a)
        print(f"{name:>16s}: inst {100*tot['inst']/ti:6.2f}%  samples {100*tot['samp']/ts:6.2f}%  long_sb {100*tot['stall_long_sb']/ts:5.2f}% "
              f"short_sb {100*tot['stall_short_sb']/ts:5.2f}% lanes {tot['thr']/max(tot['inst'],1):4.1f}")
    tot = collections.Counter()
    for key, a in agg.items():
        if not (key and key[0].endswith(".cu")):
            tot.update(a)
    print(f"{'non-.cu (intrinsics)':>16s}: inst {100*tot['inst']/ti:6.2f}%  samples {100*tot['samp']/ts:6.2f}%")
