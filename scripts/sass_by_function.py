"""Static SASS instruction count of one kernel, attributed to the source
function each instruction's line info points at (needs -lineinfo).
usage: sass_by_function.py <cubin> <kernel-substring> <source.cu>"""
import collections
import re
import subprocess
import sys


def main():
    cubin, kern, srcfile = sys.argv[1:4]
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    src = open(srcfile).read().splitlines()
    starts = []
    for i, l in enumerate(src, 1):
        m = re.match(r'^(?:static )?(?:__device__|__global__).*?(\w+)\(', l)
        if m:
            starts.append((i, m.group(1)))

    def fn(line):
        name = "?"
        for s, n in starts:
            if s <= line:
                name = n
            else:
                break
        return name

    base = srcfile.split("/")[-1]
    active = False
    cur = ("?", 0)
    cnt = collections.Counter()
    lines = collections.Counter()
    for l in out.splitlines():
        if l.startswith(".text."):
            active = kern in l
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', l):
            key = fn(cur[1]) if cur[0] == base else cur[0]
            cnt[key] += 1
            lines[cur] += 1
    print(sum(cnt.values()))
    for k, v in cnt.most_common(25):
        print(v, k)
    print()
    for k, v in lines.most_common(25):
        print(v, k, src[k[1] - 1].strip()[:80] if k[0] == base and k[1] <= len(src) else "")


if __name__ == "__main__":
    main()
