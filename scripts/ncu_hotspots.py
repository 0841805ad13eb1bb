"""Per-source-line hot spots of one kernel launch in an .ncu-rep captured with
--import-source on (code built with -lineinfo).
usage: ncu_hotspots.py <rep> <launch-id> [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
    cur = None
    hdr = None
    data = []
    for r in csv.reader(io.StringIO(raw)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) > 8 and r[0].isdigit() and r[2] == "-":
            inst = int(r[hdr.index("Instructions Executed")] or 0)
            thr = int(r[hdr.index("Thread Instructions Executed")] or 0)
            stall = int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0)
            if inst or stall:
                data.append((cur, int(r[0]), r[1].strip()[:90], inst, thr, stall))
    ti = sum(d[3] for d in data)
    ts = sum(d[5] for d in data)
    print(f"total warp-inst {ti}  stall samples {ts}")
    for d in sorted(data, key=lambda d: -d[5])[:top]:
        print(f"{100*d[5]/max(ts,1):5.1f}% stall {100*d[3]/max(ti,1):5.1f}% inst thr/inst {d[4]/max(d[3],1):4.1f} | {d[0]}:{d[1]} {d[2]}")


if __name__ == "__main__":
    main()
