"""Per kernel of `ncu --set full` captures: the top warp-stall reasons (share of pc samples),
issue-slot utilisation, achieved occupancy, SM / memory throughput -- what bounds a kernel
that is not on a bandwidth roofline.
usage: python scripts/ncu_stalls.py a.ncu-rep [b.ncu-rep ...] > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

EXTRA = {
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_thr%",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed": "mem_thr%",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "thr/inst",
}


def main():
    print("# top warp-stall reasons per kernel (share of ncu pc samples, mean over the captured launches)")
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        h = rows[0]
        stall_cols = [(i, n.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, n in enumerate(h)
                      if "pcsamp_warps_issue_stalled" in n and not n.endswith("_not_issued")]
        per = collections.OrderedDict()
        for r in rows[2:]:
            k = r[h.index("Kernel Name")][:56]
            e = per.setdefault(k, {"n": 0, "stall": collections.Counter(), "extra": collections.Counter(), "ns": 0.0})
            e["n"] += 1
            for i, n in stall_cols:
                try:
                    e["stall"][n] += float(r[i].replace(",", ""))
                except ValueError:
                    pass
            for m, short in EXTRA.items():
                if m in h:
                    try:
                        e["extra"][short] += float(r[h.index(m)].replace(",", ""))
                    except ValueError:
                        pass
        print(f"## {rep.split('/')[-1]}")
        for k, e in per.items():
            tot = sum(e["stall"].values()) or 1.0
            top = ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in e["stall"].most_common(5))
            ex = "  ".join(f"{s} {e['extra'][s] / e['n']:.1f}" for s in EXTRA.values() if s in e["extra"])
            print(f"{k:56s} x{e['n']:<2d} {ex}\n    stalls: {top}")


if __name__ == "__main__":
    main()
