"""Hot SASS instructions of a kernel from an ncu report: python scripts/ncu_hot.py <rep> <kernel-regex> [N]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may be concatenated; take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[hdr_i + 1:]):
    if not r or r[0] in ("Address", "Kernel Name") or len(r) <= i_s:
        if r and r[0] == "Kernel Name": break
        continue
    try:
        s = int(r[i_s] or 0)
    except ValueError:
        continue
    st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    data.append((s, int(r[i_ex] or 0), idx, r[i_src], st))
tot = sum(d[0] for d in data)
print("kernel:", rows[0][1][:100] if rows[0] else "", "total samples", tot, "instructions", len(data))
for s, e, idx, src, st in sorted(data, reverse=True)[:N]:
    print(f"{s:6d} {100*s/max(tot,1):5.1f}% exec={e:9d} #{idx:5d} {src.strip()[:70]:70s} {st}")
