#!/usr/bin/env python
"""Per-source-line view of an ncu capture of one kernel: joins the SASS page of the
report (instructions executed, stall samples per instruction) with the line table
of the cubin (nvdisasm -gi), attributing inlined code to the line of the kernel
that inlined it.

    python tools/ncu_lines.py REPORT.ncu-rep build/xa_kernels.o 'xa_walk_kernelILi8ELi1E' [turns]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kern = sys.argv[1:4]
    per = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, data = rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(h)}
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-c", "-gi", os.path.join(tmp, cubin)], capture_output=True,
                          text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith("_Z") and kern in l and l.endswith(":"))
    pat_file = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    pat_inst = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
    cur, insts = None, []
    for l in sass[start:]:
        if l.startswith("//-----") and insts:
            break
        m = pat_file.search(l)
        if m:
            f, ln, f2, ln2 = m.groups()
            cur = (f2.split("/")[-1], int(ln2)) if f2 else (f.split("/")[-1], int(ln))
        if pat_inst.match(l):
            insts.append(cur)
    assert len(insts) == len(data), (len(insts), len(data))
    by = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = ts = 0
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    for cur, r in zip(insts, data):
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        by[cur][0] += n
        by[cur][1] += s
        for k in stalls:
            by[cur][2][k[6:]] += int(r[ix[k]] or 0)
        tot += n
        ts += s
    src = {}
    print(f"total warp instructions {tot}, samples {ts}" + (f", {tot / per:.0f} per turn" if per else ""))
    for key in sorted(by, key=lambda k: k or ("", 0)):
        n, s, st = by[key]
        if n * 250 > tot or s * 250 > ts:
            f, ln = key if key else ("?", 0)
            if f not in src:
                p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bjxa_b200", "csrc", f)
                src[f] = open(p).read().split("\n") if os.path.exists(p) else []
            text = src[f][ln - 1].strip()[:64] if ln - 1 < len(src[f]) else ""
            top = ",".join(f"{k}:{v}" for k, v in st.most_common(2) if v)
            extra = f" ({n / per:6.1f}/turn)" if per else ""
            print(f"{f[:14]:14s} {ln:5d} inst {100 * n / tot:5.1f}%{extra} samp {100 * s / ts:5.1f}%  [{top}]  {text}")


if __name__ == "__main__":
    main()
