"""Attribute the PC samples of one kernel in an .ncu-rep to source lines (needs -lineinfo and the cubin of the
SAME build): ncu_lines.py report.ncu-rep kernel_regex object.o mangled_function source.cu"""
import collections, csv, re, subprocess, sys, tempfile, os

rep, kre, obj, fn, src_path = sys.argv[1:6]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
in_fn, cur, off2line = False, None, {}
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        in_fn = m.group(1) == fn
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
h = rows[hi]; ai, ns, ie = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
data, base = [], None
for r in rows[hi + 1:]:
    try:
        a = int(r[ai], 16)
        base = a if base is None else base
        data.append((a - base, int(r[ns]), int(r[ie])))
    except ValueError:
        pass
ts, ti = sum(d[1] for d in data), sum(d[2] for d in data)
by = collections.defaultdict(lambda: [0, 0])
for off, s, i in data:
    k = off2line.get(off, ("?", 0)); by[k][0] += s; by[k][1] += i
srcs = {}
print(f"{len(data)} instructions, {ts} samples, {ti} warp-instructions executed")
for (f, l), v in sorted(by.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[6]) if len(sys.argv) > 6 else 40]:
    if f not in srcs:
        cand = [os.path.join(os.path.dirname(src_path), f), src_path]
        srcs[f] = next((open(c).read().splitlines() for c in cand if os.path.basename(c) == f and os.path.exists(c)), [])
    txt = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
    print(f"{f}:{l:4d} smp {100 * v[0] / ts:5.1f}% exec {100 * v[1] / ti:5.1f}%  {txt}")
