"""Fold tools/hostprof/sigprof.c samples into functions: python tools/hostprof/report.py sigprof.txt [top]"""
import bisect
import collections
import subprocess
import sys


def symbols(path):
    syms = set()
    for extra in ([], ["-D"]):                      # static symbols where the object has them, the dynamic table otherwise
        out = subprocess.run(["nm", "-C", "--defined-only", "-n"] + extra + [path], capture_output=True, text=True).stdout
        for line in out.splitlines():
            parts = line.split(None, 2)
            if len(parts) == 3 and parts[1] in "tTwWiI":
                syms.add((int(parts[0], 16), parts[2]))
    return sorted(syms)


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    per_obj = collections.defaultdict(list)
    for line in open(path):
        obj, off, cnt = line.rsplit(None, 2)
        per_obj[obj].append((int(off, 16), int(cnt)))
    total = sum(c for v in per_obj.values() for _, c in v)
    fn = collections.Counter()
    objs = collections.Counter()
    for obj, samples in per_obj.items():
        syms = symbols(obj) if obj != "?" else []
        addrs = [a for a, _ in syms]
        is_exe = bool(syms) and addrs[0] > 0x100000 and not obj.endswith(".so")
        for off, cnt in samples:
            objs[obj] += cnt
            i = bisect.bisect_right(addrs, off) - 1
            name = syms[i][1] if i >= 0 else "?"
            fn[(obj.rsplit("/", 1)[-1], name)] += cnt
    print(f"{total} samples")
    for obj, c in objs.most_common():
        print(f"  {100 * c / total:5.1f}%  {obj}")
    print()
    for (obj, name), c in fn.most_common(top):
        print(f"  {100 * c / total:5.1f}%  {obj:24s} {name}")


main()
