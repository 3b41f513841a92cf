"""Top stalled SASS instructions of one kernel from an ncu report: python tools/ncu_hot.py rep kernel [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
si, ai = h.index("# Samples"), h.index("Source")
ie = h.index("Instructions Executed")
body = [r for r in rows[hdr + 1:] if len(r) > si and r[0].startswith("0x")]
tot = sum(int(r[si]) for r in body)
print("total samples", tot, "instructions", len(body))
idx = {id(r): i for i, r in enumerate(body)}
for r in sorted(body, key=lambda r: -int(r[si]))[:topn]:
    i = idx[id(r)]
    prev = body[i - 1][ai].strip()[:50] if i else ""
    print(f"{int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  #{i:5d} exec={r[ie]:>8s}  {r[ai].strip()[:70]:70s} | prev: {prev}")
