"""Aggregate an exported ncu SASS source page (ncu -i rep --page source --csv --print-source sass | gzip) by device function, using the
symbol table of the library the capture ran (the .ncu-rep files themselves are too large to bring back from the GPU box).
usage: python tools/ncu_sass_by_function.py <sass.csv.gz> <libb2s.so> <kernel symbol substring, e.g. phase1_kernelIf>"""
import collections
import csv
import gzip
import io
import os
import subprocess
import sys
import tempfile


def funcs_for(so, kname):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    syms = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout.splitlines()
    fs = []
    for l in syms:
        f = l.split()
        if len(f) >= 8 and f[3] == "FUNC" and kname in f[-1]:
            fs.append((int(f[1], 16), int(f[2], 0), f[-1]))
    fs.sort()
    return fs


def short(n):
    if "$" not in n:
        return "<kernel body>"
    out = subprocess.run(["c++filt", n.split("$")[-1]], capture_output=True, text=True).stdout.strip()
    return out.split("(")[0][-64:]


def main():
    path, so, kn = sys.argv[1], sys.argv[2], sys.argv[3]
    fs = funcs_for(so, kn)
    rows = list(csv.reader(io.TextIOWrapper(gzip.open(path))))
    h, data = rows[1], rows[2:]
    ia, isamp, iex = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    sidx = {s: h.index(s) for s in stalls}
    base = int(data[0][ia], 16)
    kern = [f for f in fs if "$" not in f[2]][0]
    agg = collections.defaultdict(collections.Counter)
    for r in data:
        try:
            off = int(r[ia], 16) - base  # the page lists the whole .text section of the kernel
        except Exception:
            continue
        name = "?"
        for a, sz, n in fs:  # the kernel symbol spans the whole section: the smallest enclosing symbol wins
            if a <= off < a + sz and (name == "?" or "$" in n):
                name = n
                if "$" in n:
                    break
        c = agg[name]
        c["samples"] += int(r[isamp] or 0); c["inst"] += int(r[iex] or 0); c["sass"] += 1
        for s in stalls:
            c[s] += int(r[sidx[s]] or 0)
    tot = sum(c["samples"] for c in agg.values())
    toti = sum(c["inst"] for c in agg.values())
    print("total samples %d, warp instructions %d, SASS lines %d" % (tot, toti, len(data)))
    for n, c in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:int(os.environ.get("TOP", "16"))]:
        top = sorted(((s, c[s]) for s in stalls), key=lambda x: -x[1])[:3]
        print("  %5.1f%% samples %5.1f%% inst %6d sass  %-64s %s" % (100 * c["samples"] / tot, 100 * c["inst"] / max(toti, 1), c["sass"], short(n),
              " ".join("%s %.0f%%" % (s[6:], 100 * v / max(c["samples"], 1)) for s, v in top)))


if __name__ == "__main__":
    main()
