"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel markdown table.

    python tools/launch_table.py gpurun_out/launches.csv profiles/r01_launches_bench "title line" [exclude-from-shares ...]
"""
import csv
import re
import shutil
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"<unnamed>::|void |\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:70]


def main():
    src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    excl = set(sys.argv[4:])
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[kn])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1e3
    tot = sum(v[1] for k, v in agg.items() if k not in excl)
    with open(out + ".md", "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` -- per-launch times are "
                f"cold-cache and serialised, compare SHARES.  Raw list: `{out.split('/')[-1]}.csv`.\n\n")
        f.write("| kernel | launches | total us | avg us | share of step kernels |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            sh = "" if k in excl else f"{100 * t / tot:.1f} %"
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {sh} |\n")
    shutil.copy(src, out + ".csv")
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
