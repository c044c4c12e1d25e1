"""Per-source-line executed-instruction histogram for one device function of one kernel.
usage: python tools/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substr> <function-substr> [topN]"""
import csv, io, os, re, subprocess, sys, tempfile, collections


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


rep, so, kname, fname = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the section of the kernel: track current function label and current source line
in_kernel = False
cur_fn, cur_line = "<kernel body>", None
addr2 = {}
for l in dis:
    if l.startswith(".section") or l.lstrip().startswith(".section"):
        in_kernel = (".text." in l and kname in l)
        cur_fn = "<kernel body>"
        continue
    if not in_kernel:
        continue
    m = re.match(r"^(\$[^:]+):", l)
    if m:
        cur_fn = m.group(1).split("$")[-1]
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        addr2[int(m.group(1), 16)] = (cur_fn, cur_line)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = select_section(list(csv.reader(io.StringIO(out))))
hdr = rows[1]
ia, ii = hdr.index("Address"), hdr.index("Instructions Executed")
base = None
hist = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ii:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    fn, line = addr2.get(a - base, ("?", None))
    try:
        n = int(r[ii])
    except ValueError:
        n = 0
    if fname in fn:
        hist[line] += n
        tot += n
print("function", fname, "total inst", tot)
for line, n in hist.most_common(topn):
    print("%6.2f%%  %s" % (100 * n / max(tot, 1), line))
