"""Turn the raw evidence of a GPU batch (gpurun_out/) into the small text tables committed
under profiles/:

    python scripts/profile_tables.py launches gpurun_out/r2_launches_room.csv  > profiles/r2_launches_room_summary.txt
    python scripts/profile_tables.py ncu gpurun_out/r2_final_phys.ncu-rep ...  > profiles/r2_phys_summary.txt

`launches`: per-kernel launch count / total / share / mean from an
`ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised: compare shares).
`ncu`: per-kernel means of the scripts/ncu_summary.py columns of `ncu --set full` captures.
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        a = agg.setdefault(r[ki][:110], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us in kernels "
          f"(ncu replay: cold caches, serialised -- compare SHARES, not absolute times)")
    print(f"{'launches':>8s} {'total_us':>9s} {'share':>6s} {'mean_us':>8s}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[0]:8d} {a[1]:9.1f} {100 * a[1] / tot:5.1f}% {a[1] / a[0]:8.1f}  {k}")


def ncu(paths):
    here = __file__.rsplit("/", 1)[0]
    raw = subprocess.run([sys.executable, here + "/ncu_summary.py"] + paths, capture_output=True, text=True).stdout
    rows = list(csv.DictReader(io.StringIO(raw)))
    seen = collections.OrderedDict()
    for r in rows:
        seen.setdefault((r["file"], r["kernel"][:52]), []).append(r)
    print("# ncu --set full --clock-control none (cold caches); means over the captured launches of each kernel")
    print(f"{'kernel':52s} {'n':>2s} {'us':>8s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'GB/s':>6s} {'regs':>4s} "
          f"{'grid':>6s} {'blk':>4s} {'occ%':>5s} {'thr/inst':>8s} {'issue%':>6s} {'l1hit%':>6s} {'l2hit%':>6s}")

    def mean(v, c):
        vals = [float(x[c]) for x in v if x.get(c) not in (None, "")]
        return sum(vals) / len(vals) if vals else float("nan")

    for (f, k), v in seen.items():
        print(f"{k:52s} {len(v):2d} {mean(v, 'ns') / 1e3:8.1f} {mean(v, 'dram_rd') / 1e6:10.1f} "
              f"{mean(v, 'dram_wr') / 1e6:10.1f} {mean(v, 'dram_gbs'):6.0f} {v[0]['regs']:>4s} {v[0]['grid']:>6s} "
              f"{v[0]['block']:>4s} {mean(v, 'occ_pct'):5.1f} {mean(v, 'thr_per_inst'):8.1f} "
              f"{mean(v, 'issue_pct'):6.1f} {mean(v, 'l1_hit'):6.1f} {mean(v, 'l2_hit'):6.1f}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        ncu(sys.argv[2:])
