"""Static SASS evidence per kernel of libmadrona_b200.so (cuobjdump -sass): instruction count
and the mnemonics that prove which hardware paths a kernel uses -- UBLKCP (TMA bulk copy),
SYNCS (mbarrier), LDG.E.128 / STG.E.128 (16-byte global accesses), MATCH / VOTE, ATOM / RED.
usage: python scripts/sass_summary.py madrona_b200/libmadrona_b200.so > profiles/r2_sass_functions.txt"""
import collections
import re
import subprocess
import sys

KEYS = ["UBLKCP", "SYNCS", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS.128", "MATCH", "VOTE", "ATOM", "RED",
        "SHFL", "BAR.SYNC", "LDL", "STL", "MUFU", "FFMA", "DFMA"]


def main():
    lib = sys.argv[1]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    cur, counts, total = None, {}, collections.Counter()
    for l in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", l)
        if cur and m:
            op = m.group(1)
            total[cur] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    print(f"# {lib}: SASS per kernel (sm_100a); columns = static instruction counts")
    print(f"{'kernel':70s} {'inst':>6s} " + " ".join(f"{k:>9s}" for k in KEYS))
    for fn in sorted(counts, key=lambda f: -total[f]):
        name = re.sub(r"\(mb2::.*", "", demangle(fn)).replace("void ", "")[:70]
        print(f"{name:70s} {total[fn]:6d} " + " ".join(f"{counts[fn][k]:9d}" for k in KEYS))


if __name__ == "__main__":
    main()
