"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`."""
import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]; si = H.index("Warp Stall Sampling (All Samples)"); ei = H.index("Instructions Executed")
data = []
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) <= si or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name": break
        continue
    data.append((idx, r[1].strip(), float(r[si]), float(r[ei])))
tot = sum(d[2] for d in data)
print("total samples", tot, "instructions", len(data))
for idx, ins, s, e in sorted(data, key=lambda d: -d[2])[:n]:
    print(f"{s/tot:6.3f}  #{idx:5d}  exec={int(e):8d}  {ins[:100]}")
