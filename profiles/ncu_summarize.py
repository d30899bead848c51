#!/usr/bin/env python
"""Turn an `ncu --set full` report into the markdown tables committed under profiles/ (run where ncu is installed; no GPU needed).
    python profiles/ncu_summarize.py gpurun_out/r02_prof_cfg4.ncu-rep > profiles/r02_ncu_cfg4_tables.md"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        print(f"## `{r[idx['Kernel Name']]}`\n")
        print("| metric | value |\n|---|---|")
        for w in WANT:
            if w in idx:
                print(f"| `{w}` | {r[idx[w]]} {units[idx[w]]} |")
        top = sorted(((float(r[idx[h]].replace(',', '')), h) for h in stalls), reverse=True)[:6]
        names = ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in top)
        print(f"| warp stall reasons (warps per issue-active cycle, top 6) | {names} |\n")


if __name__ == "__main__":
    main()
