#!/usr/bin/env python
"""Summarises an ncu --set full report as a markdown table (one column per captured launch).

  python tools/ncu_summary.py gpurun_out/r2_trace_c2.ncu-rep [more.ncu-rep ...] > profiles/<name>.md

Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed); picks the metrics the design argues with: duration,
registers, occupancy, issue utilisation, active threads per instruction, pipe utilisation, cache hit rates, DRAM bytes and
throughput, the leading stall reasons."""
import csv
import io
import subprocess
import sys

ROWS = [  # label, list of candidate metric-name suffixes (first match wins), scale
    ("duration [us]", ["gpu__time_duration.sum"], 1e-3),
    ("registers / thread", ["launch__registers_per_thread"], 1),
    ("grid size", ["launch__grid_size"], 1),
    ("achieved occupancy [%]", ["sm__warps_active.avg.pct_of_peak_sustained_active"], 1),
    ("SM issue utilisation [%]", ["sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active"], 1),
    ("active threads / warp instruction", ["smsp__thread_inst_executed_per_inst_executed.ratio"], 1),
    ("warp instructions", ["smsp__inst_executed.sum", "sm__inst_executed.sum"], 1),
    ("ALU pipe [%]", ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"], 1),
    ("FMA pipe [%]", ["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"], 1),
    ("XU pipe [%]", ["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"], 1),
    ("LSU pipe [%]", ["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"], 1),
    ("L1 hit rate [%]", ["l1tex__t_sector_hit_rate.pct"], 1),
    ("L2 hit rate [%]", ["lts__t_sector_hit_rate.pct"], 1),
    ("DRAM read [MB]", ["dram__bytes_read.sum"], 1e-6),
    ("DRAM write [MB]", ["dram__bytes_write.sum"], 1e-6),
    ("L2 throughput [% of peak]", ["lts__throughput.avg.pct_of_peak_sustained_elapsed"], 1),
    ("DRAM throughput [% of peak]", ["dram__throughput.avg.pct_of_peak_sustained_elapsed"], 1),
    ("stall long_scoreboard [warps/issue]", ["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"], 1),
    ("stall wait", ["smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"], 1),
    ("stall not_selected", ["smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"], 1),
    ("stall math_pipe_throttle", ["smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"], 1),
    ("stall branch_resolving", ["smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"], 1),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return v * mult


def main():
    for rep in sys.argv[1:]:
        hdr, units, data = load(rep)
        name_i = hdr.index("Kernel Name")
        cols = [r[name_i].replace("brt::", "").split("(")[0] + f" #{k}" for k, r in enumerate(data)]
        print(f"### {rep}\n")
        print("| metric | " + " | ".join(cols) + " |")
        print("|---|" + "---|" * len(cols))
        for label, cands, scale in ROWS:
            idx = None
            for c in cands:  # several sections can carry a metric of the same name: take the first column that has values
                m = [i for i, h in enumerate(hdr) if (h == c or h.endswith("." + c)) and any(r[i].strip() for r in data)]
                if m:
                    idx = m[0]
                    break
            if idx is None:
                continue
            vals = []
            for r in data:
                try:
                    v = float(r[idx].replace(",", ""))
                except ValueError:
                    vals.append(r[idx])
                    continue
                u = units[idx]
                if "byte" in u:
                    v = to_bytes(v, u) * 1e-6
                elif label.startswith("duration"):
                    v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                else:
                    v = v * scale
                vals.append(f"{v:.3f}".rstrip("0").rstrip(".") if abs(v) < 1e6 else f"{v:.0f}")
            print(f"| {label} | " + " | ".join(vals) + " |")
        print()


if __name__ == "__main__":
    main()
