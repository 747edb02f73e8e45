"""Lists the hottest SASS regions of an .ncu-rep by executed instructions:
python tools/ncu_hot_sass.py rep [min_fraction]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ci["Instructions Executed"]]) for r in body)
totS = sum(int(r[ci["# Samples"]]) for r in body)
print(f"total warp instructions {tot:.4e}, samples {totS}, SASS lines {len(body)}")
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
for k, r in enumerate(body):
    ie = int(r[ci["Instructions Executed"]])
    if ie / tot >= thr:
        print(f"{k:5d} {100*ie/tot:5.2f}% smp {100*int(r[ci['# Samples']])/totS:5.2f}% thr {r[ci['Avg. Threads Executed']]:>4} cfl {r[ci['L1 Conflicts Shared N-Way']]:>3} {r[ci['Source']].strip()}")
