"""Summarise .ncu-rep captures (ncu --set full) into a small CSV table:
per kernel launch: duration, DRAM bytes, achieved DRAM GB/s, registers,
achieved occupancy, active threads per instruction, issue utilisation."""
import csv
import io
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "ns",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "thr_per_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "l1tex__t_sector_hit_rate.pct": "l1_hit",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
}


def scale(v, unit, table):
    return float(v.replace(",", "")) * table.get(unit, 1)


def main():
    rows_out = []
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"file": rep.split("/")[-1], "kernel": r[hdr.index("Kernel Name")][:60]}
            for m, short in METRICS.items():
                if m in hdr:
                    i = hdr.index(m)
                    if short in ("dram_rd", "dram_wr"):
                        d[short] = scale(r[i], units[i], {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9})
                    elif short == "ns":
                        d[short] = scale(r[i], units[i], {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9})
                    else:
                        d[short] = r[i][:8]
            if "ns" in d and "dram_rd" in d:
                d["dram_gbs"] = (d["dram_rd"] + d["dram_wr"]) / d["ns"]
            rows_out.append(d)
    cols = ["file", "kernel", "ns", "dram_rd", "dram_wr", "dram_gbs", "dram_pct", "regs", "grid", "block",
            "occ_pct", "thr_per_inst", "issue_pct", "warp_inst", "l1_hit", "l2_hit"]
    w = csv.DictWriter(sys.stdout, fieldnames=cols, extrasaction="ignore")
    w.writeheader()
    for d in rows_out:
        for k in ("ns", "dram_rd", "dram_wr"):
            if k in d:
                d[k] = f"{d[k]:.0f}"
        if "dram_gbs" in d:
            d["dram_gbs"] = f"{d['dram_gbs']:.1f}"
        w.writerow(d)


if __name__ == "__main__":
    main()
