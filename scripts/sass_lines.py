"""Static SASS-per-source-line histogram for one kernel out of `nvdisasm -g <cubin>` output (run on the CPU box before
spending GPU time).  Usage: python scripts/sass_lines.py <disasm.txt> <kernel-substring> [top]"""
import collections
import re
import sys


def main():
    path, key = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    inside = False
    cur = None
    cnt = collections.Counter()
    ops = collections.Counter()
    for ln in open(path):
        if ln.startswith(".text."):
            inside = key in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            cnt[cur] += 1
            ops[m.group(2).split(".")[0]] += 1
    tot = sum(cnt.values())
    print("total static instructions", tot)
    byfile = collections.Counter()
    for (f, _), c in cnt.items():
        byfile[f] += c
    print(dict(byfile))
    print("ops", ops.most_common(25))
    for k, c in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
        print(f"{k[0]}:{k[1]:<5d} {c}")


main()
