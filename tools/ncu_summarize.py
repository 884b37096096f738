#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into a small JSON under profiles/.

    python tools/ncu_summarize.py gpurun_out/r01_k_nn_tc.ncu-rep profiles/r01_k_nn_tc_full.json [--note "..."]

Keeps the metrics the roofline discussion in DESIGN.md cites: duration, DRAM bytes, tensor-pipe / L2 / SM
throughput percentages, registers, grid, and the top stall reasons of the sampled warps.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--note" else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP:
                d[h] = f"{r[i]} {units[i]}".strip()
            elif h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued"):
                try:
                    d.setdefault("_stalls", {})[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i])
                except ValueError:
                    pass
        st = d.pop("_stalls", {})
        tot = sum(st.values()) or 1.0
        d["top_stalls_share_of_samples"] = {k: round(v / tot, 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6]}
        launches.append(d)
    js = {"report": rep, "note": note, "launches": launches}
    with open(out, "w") as f:
        json.dump(js, f, indent=1)
    print(json.dumps(js, indent=1)[:3000])


if __name__ == "__main__":
    main()
