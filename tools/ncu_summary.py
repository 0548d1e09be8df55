"""Summarise an `ncu --set full` report: one line per captured launch with duration, DRAM traffic,
tensor-pipe activity, issue utilisation.  usage: python tools/ncu_summary.py report.ncu-rep [--json]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default


def scale(name, v):
    u = units[col[name]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3,
                "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)


out = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    d = {
        "kernel": name.replace("void ", "").replace("papc::", "")[:70],
        "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
        "us": scale("gpu__time_duration.sum", g(r, "gpu__time_duration.sum")),
        "dram_read_MB": scale("dram__bytes_read.sum", g(r, "dram__bytes_read.sum")) / 1e6,
        "dram_write_MB": scale("dram__bytes_write.sum", g(r, "dram__bytes_write.sum")) / 1e6,
        "tensor_pipe_active_pct": g(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "regs": g(r, "launch__registers_per_thread"),
        "warps_active_pct": g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": g(r, "smsp__inst_executed.sum"),
        "sm_cycles": g(r, "sm__cycles_elapsed.max"),
    }
    out.append(d)
if "--json" in sys.argv:
    print(json.dumps(out, indent=1))
else:
    for d in out:
        print(f"{d['kernel'][:58]:58s} grid {d['grid']:>12s} {d['us']:8.1f} us  dram R {d['dram_read_MB']:7.1f} W {d['dram_write_MB']:7.1f} MB  "
              f"tensor {d['tensor_pipe_active_pct']:5.1f}%  sm {d['sm_throughput_pct']:5.1f}%  dram {d['dram_throughput_pct']:5.1f}%  regs {d['regs']:.0f}")
