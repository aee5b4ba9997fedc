"""Summarise an .ncu-rep (ncu --set full) into a small CSV + markdown table for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in raw.splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx["Kernel Name"]
    table = []
    for r in data:
        rec = {"kernel": r[name_i]}
        for k in KEYS:
            if k in idx:
                rec[k] = f"{r[idx[k]]} {units[idx[k]]}".strip()
        table.append(rec)
    with open(out + ".json", "w") as f:
        json.dump(table, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}` (per launch; cold-cache, serialised replays)\n\n")
        for rec in table:
            f.write(f"## {rec['kernel']}\n\n| metric | value |\n|---|---|\n")
            for k in KEYS:
                if k in rec:
                    f.write(f"| {k} | {rec[k]} |\n")
            f.write("\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
