#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into (1) a per-launch table of the metrics quoted in DESIGN.md /
profiles/README.md and (2) profiles/ncu_traffic.json (mean DRAM bytes per launch and kernel, read by bench.py for
`roofline.traffic`).

  ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/X_raw.csv
  python tools/ncu_summary.py /tmp/X_raw.csv profiles/r01f_ncu_summary.csv [--traffic profiles/ncu_traffic.json --source "..."]
"""
import argparse
import csv
import json

KEYS = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "duration_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def to_bytes(value, unit):
    return value * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("out")
    ap.add_argument("--traffic")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        d = {}
        for k, name in KEYS:
            if k not in col:
                continue
            v = num(r[col[k]])
            if name in ("dram_read", "dram_write") and isinstance(v, float):
                v = to_bytes(v, units[col[k]])
            if name == "duration_us" and isinstance(v, float):
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[col[k]], 1.0)
            d[name] = v
        out.append(d)
    names = [n for _, n in KEYS if any(n in d for d in out)]
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(names)
        for d in out:
            w.writerow([("%.6g" % d[n]) if isinstance(d.get(n), float) else d.get(n, "") for n in names])
    if a.traffic:
        cls = {"trace": [], "shadow": [], "shade": [], "accumulate": []}
        for d in out:
            k = d["kernel"]
            closest = any(s in k for s in ("k_trace<0>", "k_trace<(int)0>", "k_trace<0,", "k_trace<(int)0,"))
            c = "shade" if "k_shade" in k else ("accumulate" if "k_accumulate" in k else ("trace" if closest else ("shadow" if "k_trace" in k else None)))
            if c:
                cls[c].append(d["dram_read"] + d["dram_write"])
        t = {c: (sum(v) / len(v) if v else None) for c, v in cls.items()}
        t.update({"unit": "bytes per launch", "launches_averaged": {c: len(v) for c, v in cls.items()}, "source": a.source})
        json.dump(t, open(a.traffic, "w"), indent=1)
    for d in out:
        print("%-44s %8.1f us  dram %6.1f MB  thr/inst %5.2f  issue %5.1f%%  warps %5.1f%%" % (
            str(d["kernel"])[:44], d["duration_us"], (d["dram_read"] + d["dram_write"]) * 1e-6, d["threads_per_inst"], d["issue_active_pct"], d["warps_active_pct"]))


if __name__ == "__main__":
    main()
