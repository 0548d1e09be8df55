"""Summarise tools/prof_gclk.py output: per launch of the LAST recorded step, the stamps of all CTAs
relative to the step's first layer-kernel entry (us): entry, setup, W staged, previous kernel complete
(griddepcontrol.wait returned), first accumulator ready, tiles done, exit; min / median / max over CTAs.
usage: python tools/gclk_summary.py gclk.txt [launches_per_step=8]"""
import statistics
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 8
launches = []
for line in open(path):
    w = line.split()
    if w[0] == "launch":
        launches.append({"meta": " ".join(w[2:]), "cta": []})
    elif w[0] == "cta":
        launches[-1]["cta"].append([int(x) for x in w[2:]])
step = launches[-per_step:]
t0 = min(c[0] for c in step[0]["cta"])
names = ["entry", "setup", "Wstaged", "prev_done", "raw0_landed", "x_full0", "mma0", "acc0", "tiles_done", "exit"]
slots = [0, 1, 2, 3, 13, 10, 11, 4, 5, 6]
prev_exit = None
for L in step:
    cs = L["cta"]
    print(f"== {L['meta']}  ({len(cs)} CTAs)" + (f"   previous kernel's last exit {prev_exit:8.2f}" if prev_exit is not None else ""))
    for e, nm in zip(slots, names):
        v = sorted((c[e] - t0) / 1e3 for c in cs if len(c) > e and c[e])
        if v:
            print(f"   {nm:11s} min {v[0]:8.2f}  med {statistics.median(v):8.2f}  max {v[-1]:8.2f}")
    fin = [c for c in cs if c[7] and c[7] > c[0]]   # (graph replays keep stale slots of other CTAs: take the live one)
    fin = [c for c in fin if c[7] > c[5]]
    if fin:
        c = fin[0]
        if len(c) > 14 and c[12]:
            print(f"   last CTA reduction: thread 0 done {(c[12] - t0) / 1e3:8.2f}, all threads done {(c[14] - t0) / 1e3:8.2f}")
        red = f", reduced {(c[8] - t0) / 1e3:8.2f}, scale/shift written {(c[9] - t0) / 1e3:8.2f}" if len(c) > 9 and c[8] else ""
        print(f"   last CTA: tiles done {(c[5] - t0) / 1e3:8.2f}, finalisation starts {(c[7] - t0) / 1e3:8.2f}{red}, exits {(c[6] - t0) / 1e3:8.2f}")
    prev_exit = max((c[6] - t0) / 1e3 for c in cs)
