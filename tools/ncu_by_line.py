"""development probe: executed warp instructions of an .ncu-rep per CUDA source line.
usage: ncu_by_line.py <rep> <kernel mangled-name substring> [libKernel.so] [min share %]
Joins ncu's SASS page (per-instruction 'Instructions Executed') with nvdisasm -g line info of the same
function by instruction address."""
import csv, os, re, subprocess, sys, tempfile
rep, fn = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "metropolis-hastings-gpgpu_b200", "libKernel.so")
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("mh_kernels")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
OUTER = os.environ.get("BY_LINE_OUTER")      # e.g. "mh_delta.cuh:200" = attribute to the innermost frame in that file at/after that line
line_of, cur, inside, frames = {}, None, False, []
for l in dis:
    if l.startswith("//---") and ".text." in l:
        inside = fn in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        frames.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*)", l)
    if m:
        if frames:
            cur = frames[0]                                  # innermost
            if OUTER:
                of, ol = OUTER.split(":")
                pick = [f for f in frames if f[0] == of and f[1] >= int(ol)]
                cur = pick[0] if pick else frames[-1]
            frames = []
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
agg, tot = {}, 0
base = None
for r in rows[2:]:
    addr = int(r[ci["Address"]], 16) if "Address" in ci else None
    if base is None:
        base = addr
    ie = int(r[ci["Instructions Executed"]])
    smp = int(r[ci["# Samples"]])
    key = line_of.get(addr - base)
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += ie; a[1] += smp; a[2] += 1
    tot += ie
totS = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {tot:.4e}; share of instructions / samples / SASS lines per source line (>= {thr} %)")
src_cache = {}
def text(key):
    if not key: return ""
    f = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", key[0])
    if f not in src_cache:
        try: src_cache[f] = open(f).read().splitlines()
        except OSError: src_cache[f] = []
    ls = src_cache[f]
    return ls[key[1] - 1].strip()[:90] if 0 < key[1] <= len(ls) else ""
for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if 100.0 * a[0] / tot >= thr:
        print(f"{(key[0] + ':' + str(key[1])) if key else '?':22s} {100*a[0]/tot:6.2f}% {100*a[1]/totS:6.2f}% {a[2]:5d}  {text(key)}")
