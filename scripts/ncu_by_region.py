"""Per-source-line executed warp instructions of one kernel, in SASS address order, with instructions of
inlined header code (shuffles, reductions) charged to the last line of the kernel's own file seen before
them.  usage: ncu_by_region.py <report.ncu-rep> <cubin> <kernel-substring> <file.cu> <units>
Prints instructions per unit (e.g. per solve) for every line of <file.cu> and cumulative totals."""
import collections, csv, io, re, subprocess, sys
rep, cubin, kern, fname, units = sys.argv[1:6]
units = float(units)
sass = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout
fn = None; own = None; amap = {}
for l in sass.splitlines():
    s = l.strip()
    a = re.match(r"\.text\.(\S+):", s)
    if a: fn = a.group(1); own = None; continue
    a = re.match(r'//## File "([^"]+)", line (\d+)', s)
    if a:
        if a.group(1).endswith(fname): own = int(a.group(2))
        continue
    a = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
    if a and fn and kern in fn: amap[int(a.group(1), 16)] = (own, a.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kstart = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern.split("_kernel")[0] in r[1])
start = next(i for i in range(kstart, len(rows)) if rows[i] and rows[i][0] == "Address")
h = rows[start]; ix = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0]); base = None
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    addr = int(r[0], 16)
    if base is None: base = addr
    key = amap.get(addr - base, (None, "?"))[0]
    agg[key][0] += int(r[ix["Instructions Executed"]] or 0); agg[key][1] += int(r[ix["# Samples"]] or 0)
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total {tot / units:.0f} instructions per unit, {ts} samples")
cum = 0
for k in sorted(agg, key=lambda x: (x is None, x)):
    cum += agg[k][0]
    if agg[k][0] / units >= 5:
        print(f"line {k}: {agg[k][0] / units:7.1f} inst/unit  {100 * agg[k][1] / ts:5.1f}% samples   cum {cum / units:7.0f}")
