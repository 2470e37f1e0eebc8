"""Join an ncu SASS source page (per-address stall samples) with nvdisasm line info and print
the hottest source lines of one kernel.
usage: python profiles/hotlines.py <report.ncu-rep> <kernel-regex> <lib.so> [top]"""
import csv, re, subprocess, sys, tempfile, os, glob, collections

rep, kre, lib = sys.argv[1], sys.argv[2], os.path.abspath(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout.splitlines()
# first kernel only
hdr_i = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
rows = []
for l in out[hdr_i:]:
    if l.startswith('"Kernel Name"') and rows:
        break
    rows.append(l)
rd = list(csv.DictReader(rows))
kname = out[hdr_i - 1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
line_of = {}
mangled = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    cur_fn, cur_line = None, None
    for l in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
        if m:
            cur_fn = m.group(1)
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur_fn and re.search(kre, cur_fn):
            line_of.setdefault(cur_fn, {})[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
fn = max(line_of, key=lambda k: len(line_of[k]))
base = min(int(r["Address"], 16) for r in rd)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in rd:
    off = int(r["Address"], 16) - base
    ln, sass = line_of[fn].get(off, (None, r["Source"]))
    s = int(r["# Samples"] or 0)
    ie = int(r["Instructions Executed"] or 0)
    tot += s
    a = agg[ln]
    a[0] += s
    a[1] += ie
    for k in ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_math",
              "stall_not_selected", "stall_branch_resolving", "stall_no_inst"):
        a[2][k] += int(r.get(k) or 0)
print(kname[:120], " total samples", tot)
for ln, (s, ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3) if v)
    print(f"{str(ln):32s} samples={s:6d} ({100.0 * s / max(1, tot):5.1f}%) inst={ie:8d}  {tops}")
