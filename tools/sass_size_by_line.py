"""Static SASS size (instruction count) per source line of one device function of one kernel.
usage: python tools/sass_size_by_line.py <lib.so> <kernel-substr> <function-substr> [topN]"""
import collections, os, re, subprocess, sys, tempfile
so, kname, fname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
in_kernel, cur_fn, cur_line = False, "<kernel body>", None
hist, fsz = collections.Counter(), collections.Counter()
for l in dis:
    if l.lstrip().startswith(".section"):
        in_kernel = (".text." in l and kname in l); cur_fn = "<kernel body>"; continue
    if not in_kernel: continue
    m = re.match(r"^(\$[^:]+):", l)
    if m: cur_fn = m.group(1).split("$")[-1]; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        fsz[cur_fn] += 1
        if fname in cur_fn: hist[cur_line] += 1
print("functions (instructions):", ", ".join("%s=%d" % (k[:40], v) for k, v in fsz.most_common(12)))
tot = sum(hist.values())
print("function", fname, "instructions", tot)
for line, n in hist.most_common(topn):
    print("%6d %5.1f%%  %s" % (n, 100 * n / max(tot, 1), line))
