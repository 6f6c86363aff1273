"""Summarise an ncu report (--set full capture of ONE kernel launch) into a small JSON + markdown table.

    python tools/ncu_summary.py gpurun_out/prof_k_eval.ncu-rep profiles/r01_k_eval   -> .json and .md

Read here on the CPU box (ncu -i ... --page raw --csv).  The JSON is what bench.py reads for roofline.traffic."""
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/CTA"), ("launch__shared_mem_per_block_static", "static smem/CTA"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"), ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (smem)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "selected", "dispatch_stall", "math_pipe_throttle",
          "mio_throttle", "lg_throttle", "branch_resolving", "no_instruction", "barrier", "membar", "sleeping"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(hdr)}
    d = {"kernel": vals[col["Kernel Name"]], "report": rep.split("/")[-1], "metrics": {}, "stalls_per_issue": {}}
    for k, label in KEYS:
        if k in col:
            v = vals[col[k]].replace(",", "")
            try:
                v = float(v)
            except ValueError:
                pass
            d["metrics"][k] = {"label": label, "value": v, "unit": units[col[k]]}
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in col:
            d["stalls_per_issue"][s] = float(vals[col[k]])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tr = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = d["metrics"].get(k)
        if m:
            tr += m["value"] * scale.get(m["unit"], 1.0)
    d["dram_bytes_per_launch"] = tr
    json.dump(d, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full: `{d['kernel']}`\n\nsource report: `{d['report']}` (one launch, `--clock-control none`)\n\n")
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for k, m in d["metrics"].items():
            v = m["value"]
            f.write(f"| {m['label']} (`{k}`) | {v:.4g} | {m['unit']} |\n" if isinstance(v, float) else f"| {m['label']} | {v} | {m['unit']} |\n")
        f.write(f"| DRAM traffic per launch (read + write) | {tr/1e6:.1f} | MB |\n\n")
        f.write("Warp stall reasons (warps stalled per issue-active cycle, per scheduler):\n\n| reason | warps |\n|---|---|\n")
        for s, v in sorted(d["stalls_per_issue"].items(), key=lambda kv: -kv[1]):
            f.write(f"| {s} | {v:.3f} |\n")
    print("wrote", out + ".json", out + ".md")


if __name__ == "__main__":
    main()
