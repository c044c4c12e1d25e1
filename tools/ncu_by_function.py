"""Aggregate an ncu source-page (SASS) dump by device function using the cubin's symbol table.
usage: python tools/ncu_by_function.py <report.ncu-rep> <libb2s.so> [kernel-substring]"""
import csv
import io
import os
import subprocess
import sys
import tempfile


def select_section(rows):
    """a report with several kernels prints one (Kernel Name, header, rows...) section per kernel: keep the first one
    whose demangled name contains $NCU_KERNEL_MATCH (default: the first section)"""
    want = os.environ.get("NCU_KERNEL_MATCH", "")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        return rows
    starts.append(len(rows))
    for a, b in zip(starts[:-1], starts[1:]):
        if want in rows[a][1]:
            return rows[a:b]
    raise SystemExit("no kernel section matches " + want)


def main():
    rep, so = sys.argv[1], sys.argv[2]
    kname = sys.argv[3] if len(sys.argv) > 3 else "step_kernelIf"
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    syms = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout.splitlines()
    funcs = []
    for l in syms:
        f = l.split()
        if len(f) >= 8 and f[3] == "FUNC" and kname in f[7]:
            funcs.append((int(f[1], 16), int(f[2], 0), f[7].split("$")[-1] if "$" in f[7] else "<kernel body>"))
    funcs.sort()
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = select_section(list(csv.reader(io.StringIO(out))))
    hdr = rows[1]
    ia, ii, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ino = hdr.index("stall_no_inst")
    base = None
    agg = {}
    tot = [0, 0, 0]
    for r in rows[2:]:
        if len(r) <= ino:
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        off = addr - base
        name = "<kernel body>"
        for v, sz, n in funcs:
            if v <= off < v + sz and n != "<kernel body>":
                name = n
        a = agg.setdefault(name, [0, 0, 0])
        for k, idx in enumerate((ii, ist, ino)):
            try:
                val = int(r[idx])
            except ValueError:
                val = 0
            a[k] += val
            tot[k] += val
    print(f"{'function':60s} {'inst%':>7s} {'samples%':>9s} {'no_inst% of fn samples':>22s}")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:60]:60s} {100*a[0]/max(tot[0],1):7.2f} {100*a[1]/max(tot[1],1):9.2f} {100*a[2]/max(a[1],1):22.1f}")
    print("total warp-instructions", tot[0], "samples", tot[1], "no_inst share %.1f%%" % (100 * tot[2] / max(tot[1], 1)))


if __name__ == "__main__":
    main()
