"""Stall samples of one kernel in an .ncu-rep, bucketed by SASS address range and top stall reasons, with the
hottest instructions.  usage: python tools/ncu_roles.py report.ncu-rep kernel-index [bucket=100] [top=25]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2])
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, out, k = None, [], -1
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1
        if k == kidx:
            print(r[1][:110])
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if k == kidx and hdr and len(r) == len(hdr):
        out.append(r)
ia = hdr.index('Warp Stall Sampling (All Samples)')
isrc = hdr.index('Source')
iex = hdr.index('Instructions Executed')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ia]) for r in out)
print('total samples', tot, 'instructions', len(out))
for b in range(0, len(out), bucket):
    blk = out[b:b + bucket]
    s = sum(int(r[ia]) for r in blk)
    if s == 0:
        continue
    ex = sum(int(r[iex]) for r in blk)
    agg = {st: sum(int(r[hdr.index(st)]) for r in blk) for st in stalls}
    top = sorted(agg.items(), key=lambda x: -x[1])[:3]
    print(f"  [{b:5d}] samples {s:6d} ({100 * s / tot:4.1f}%)  executed {ex:10d}  " + " ".join(f"{n[6:]}={v}" for n, v in top))
top = sorted(range(len(out)), key=lambda i: -int(out[i][ia]))[:topn]
for i in sorted(top):
    r = out[i]
    st = {s: int(r[hdr.index(s)]) for s in stalls if int(r[hdr.index(s)]) > 0}
    print(i, r[ia], r[iex], r[isrc].strip()[:70], sorted(st.items(), key=lambda x: -x[1])[:2])
