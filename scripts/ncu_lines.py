"""Per source line sample totals from an ncu report (needs -lineinfo): python scripts/ncu_lines.py <rep> <file-substr> [kernel-idx]"""
import csv, subprocess, sys, collections
rep, fsub = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# structure: repeated sections: "File Name",path / header / rows (Line No, Source, Address, SASS, metrics...)
sections = []
cur = None
for r in rows:
    if r and r[0] in ("File Name", "File Path"):
        cur = {"file": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
secs = [s for s in sections if fsub in s["file"] and "hdr" in s]
print("sections matching:", len(secs))
s = secs[which]
h = s["hdr"]
i_s = h.index("# Samples")
agg = collections.OrderedDict()
line, src = None, None
for r in s["rows"]:
    if not r[0] or not r[0].isdigit():
        continue
    line, src = int(r[0]), r[1]
    try:
        n = int(r[i_s] or 0)
    except (ValueError, IndexError):
        n = 0
    a = agg.setdefault(line, [0, src])
    a[0] += n
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for ln, (n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{n:6d} {100*n/max(tot,1):5.1f}%  L{ln}: {src.strip()[:120]}")
