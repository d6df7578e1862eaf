"""Aggregate an ncu SASS source page by CUDA source line.
usage: ncu_by_line.py <report.ncu-rep> <cubin> <kernel-mangled-substring> [top]
Joins `ncu --page source --csv` (per-SASS-address samples / instructions executed) with
`nvdisasm -g` line markers of the same cubin."""
import collections, csv, io, re, subprocess, sys

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout
fn = None; line = None; amap = {}
for l in sass.splitlines():
    s = l.strip()
    a = re.match(r"\.text\.(\S+):", s)
    if a:
        fn = a.group(1); line = None; continue
    a = re.match(r'//## File "([^"]+)", line (\d+)', s)
    if a:
        line = (a.group(1).split("/")[-1], int(a.group(2))); continue
    a = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
    if a and fn and kern in fn:
        amap[int(a.group(1), 16)] = (line, a.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the block of the kernel whose name contains the first word of <kernel-mangled-substring> (else the first block)
key = re.split(r"_kernel|IL", kern)[0]
kstart = next((i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and key in r[1]), 0)
start = next(i for i in range(kstart, len(rows)) if rows[i] and rows[i][0] == "Address")
h = rows[start]; ix = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0, 0]); base = None
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
stalls = collections.defaultdict(lambda: collections.Counter())
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = amap.get(addr - base, (None, "?"))[0]
    samp = int(r[ix["# Samples"]] or 0); inst = int(r[ix["Instructions Executed"]] or 0)
    agg[key][0] += samp; agg[key][1] += inst; agg[key][2] += 1
    for c in stall_cols:
        v = r[ix[c]]
        if v and v != "0":
            stalls[key][c] += int(v)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instructions {ti}, sass instrs {sum(v[2] for v in agg.values())}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{c[6:]}:{n}" for c, n in stalls[k].most_common(3))
    print(f"{100 * v[0] / ts:5.1f}% samp {100 * v[1] / ti:5.1f}% inst {v[2]:5d} sass  {k}  [{st}]")
