"""Attribute ncu warp-stall samples to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all <obj or .so>;  nvdisasm -g <cubin> > all.dis
    python tools/ncu_hot_lines.py sass.csv all.dis <mangled-kernel-substring> [section index] [top N]

The SASS page lists the instructions of the kernel in order; nvdisasm -g annotates the same instructions
with `//## File "...", line N`, so the two are aligned by instruction index."""
import collections
import csv
import re
import sys

sass_csv, dis, key = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else -1
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

# --- per-instruction (file, line) from nvdisasm
lines_of = []
cur = ("?", 0)
infn = False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        infn = key in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines_of.append(cur)

# --- per-instruction samples from ncu
sections, sec = [], None
for row in csv.reader(open(sass_csv, newline="")):
    if not row:
        continue
    if row[0] == "Kernel Name":
        sec = {"name": row[1], "hdr": None, "rows": []}
        sections.append(sec)
    elif sec is not None and sec["hdr"] is None:
        sec["hdr"] = row
    elif sec is not None:
        sec["rows"].append(row)
sec = sections[which]
h = sec["hdr"]
i_smp, i_ins, i_src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall_cols = [(i, c[6:]) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
print(sec["name"][:120])
print("instructions: ncu %d, nvdisasm %d" % (len(sec["rows"]), len(lines_of)))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for idx, r in enumerate(sec["rows"]):
    n = int(r[i_smp] or 0)
    ins = int(r[i_ins] or 0)
    loc = lines_of[idx] if idx < len(lines_of) else ("?", 0)
    a = agg[loc]
    a[0] += n
    a[1] += ins
    for i, c in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[2][c] += v
    tot += n
print("total samples", tot)
byfile = collections.Counter()
for (f, l), a in agg.items():
    byfile[f] += a[0]
print("by file:", ", ".join("%s %.1f%%" % (f, 100.0 * n / tot) for f, n in byfile.most_common(6)))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6d %5.1f%% inst %10d  %s:%d  %s" % (a[0], 100.0 * a[0] / tot, a[1], f, l,
                                               " ".join("%s:%d" % kv for kv in a[2].most_common(3))))
