#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt --json profiles/x_metrics.json --updates N

--json also writes the handful of numbers bench.py quotes next to its live measurements (DRAM bytes of one launch of
the dominant kernel, warp instructions per 32 spin-updates, issue-slot utilisation), so that they are read from the
committed capture instead of being literals in bench.py.  --updates = spin-updates of the captured launch
(chains x sweeps x spins of the command that was profiled).
"""
import csv
import io
import re
import subprocess
import sys

KEEP = [
    r"^gpu__time_duration\.sum$", r"^launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit.*)$",
    r"^sm__cycles_elapsed\.avg(\.per_second)?$", r"^sm__inst_executed\.avg\.per_cycle_elapsed$", r"^smsp__inst_executed\.sum$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|lsu|xu|uniform|tensor.*)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_(alu|fma|fmaheavy|tensor.*)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^sm__pipe_tensor.*", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__t_bytes\.sum$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    json_out = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    updates = float(sys.argv[sys.argv.index("--updates") + 1]) if "--updates" in sys.argv else None
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# {rep}: ncu --set full --clock-control none, raw page, selected metrics\n")
        for r in rows[2:]:
            f.write(f"\n## kernel: {r[name_col][:160]}\n")
            for i, h in enumerate(hdr):
                if any(re.search(p, h) for p in KEEP):
                    f.write(f"{h:95s} {units[i]:12s} {r[i]}\n")
    print("wrote", out)
    if json_out:
        import json
        r = rows[2]
        get = lambda name: float(r[hdr.index(name)].replace(",", "")) if name in hdr else None
        unit = lambda name: units[hdr.index(name)] if name in hdr else ""
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = get("dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1.0)
        wr = get("dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1.0)
        inst = get("smsp__inst_executed.sum")
        d = {"kernel": r[name_col][:120], "capture": rep, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "dram_bytes_per_launch": rd + wr, "warp_instructions": inst,
             "issue_active_frac": (get("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0) / 100.0,
             "gpu_time_us": get("gpu__time_duration.sum"), "updates_in_capture": updates,
             "warp_instr_per_32_updates": (inst * 32.0 / updates) if (inst and updates) else None}
        json.dump(d, open(json_out, "w"), indent=1)
        print("wrote", json_out)


if __name__ == "__main__":
    main()
