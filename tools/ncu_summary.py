"""Selected metrics of one `ncu --set full` report as JSON (profiles/*.json): duration, DRAM bytes per launch,
issue / pipe utilisation, shared-memory wavefronts and the stall breakdown.  Usage:
    python tools/ncu_summary.py report.ncu-rep "capture description" > profiles/name.json"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tma.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "memory_l1_wavefronts_shared_ideal", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rep, what = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    out = []
    for r in rows[hi + 2:]:
        d = dict(zip(hdr, r))
        m = {k: {"value": d[k], "unit": units[hdr.index(k)]} for k in KEYS if k in d}
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k] or 0) > 0.1:
                m[k] = {"value": d[k], "unit": "warps per issue"}
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        dram = sum(float(d[k]) * unit[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.append({"kernel": d["Kernel Name"], "capture": what, "dram_bytes_per_launch": dram, "metrics": m})
    json.dump(out[0] if len(out) == 1 else out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
