"""Top SASS lines by stall samples from `ncu -i rep --page source --csv` (SASS view)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr) or not r[ix["# Samples"]].strip().isdigit():
        continue
    data.append((int(r[ix["# Samples"]] or 0), n, r))
total = sum(d[0] for d in data)
agg = {h: sum(int(d[2][ix[h]] or 0) for d in data) for h in stall_cols}
print("total samples", total, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
print("instructions executed", sum(int(d[2][ix["Instructions Executed"]] or 0) for d in data))
for s, n, r in sorted(data, reverse=True)[:top]:
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{s:7d} {100*s/total:5.1f}%  line {n:5d}  exec {r[ix['Instructions Executed']]:>10}  {r[ix['Source']].strip():60s} {st}  wf={r[ix['L1 Wavefronts Shared']]}/{r[ix['L1 Wavefronts Shared Ideal']]}")
