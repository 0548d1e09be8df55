"""Per-instruction stall summary of one kernel in an .ncu-rep (needs -lineinfo / --import-source on).
usage: python tools/ncu_stalls.py report.ncu-rep [kernel-index] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, out, k = None, [], -1
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1
        if k == kidx:
            print(r[1][:100])
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
agg = {s: sum(int(r[hdr.index(s)]) for r in out) for s in stalls}
print('total samples', tot, 'instructions', len(out))
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for b in range(0, len(out), 100):
    s = sum(int(r[ia]) for r in out[b:b + 100])
    ex = sum(int(r[iex]) for r in out[b:b + 100])
    print(f"  [{b:5d}] samples {s:6d}  executed {ex}")
top = sorted(range(len(out)), key=lambda i: -int(out[i][ia]))[:topn]
for i in sorted(top):
    r = out[i]
    st = {s: int(r[hdr.index(s)]) for s in stalls if int(r[hdr.index(s)]) > 0}
    print(i, r[ia], r[iex], r[isrc].strip()[:72], sorted(st.items(), key=lambda x: -x[1])[:2])
