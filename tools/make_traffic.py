"""profiles/traffic.json from an `ncu --set full` capture of ONE warm step of tools/prof_step.py: the kernels of the
SSG stack in launch order are matched with the bench's (name, M, cin, cout) labels.
usage: python tools/make_traffic.py gpurun_out/r02_step_full.ncu-rep profiles/r02_ncu_full_step.json
(the per-call zero_words_kernel launches are not captured)"""
import json
import subprocess
import sys

rep, summary_out = sys.argv[1], sys.argv[2]
summ = json.loads(subprocess.run([sys.executable, "tools/ncu_summary.py", rep, "--json"], capture_output=True, text=True).stdout)
json.dump(summ, open(summary_out, "w"), indent=1)
# launch order of one eager step (fused sampling on the first layer, side-stream sampling for the second)
LABELS = [("sample_group", "sample_group", 32768, 512, 32), ("point_moments", "point_moments_finish", 524288, 3, 64),
          ("mlp_layer_tt_kernel<2, 1, 0, 0, 0>", "mlp_tt<pointmlp,f16x3,Wtmem>", 524288, 64, 64),
          ("mlp_layer_tt_kernel<0, 1, 0, 1, 0>", "mlp_tt<plain,f16x3,Wtmem,pool>", 524288, 64, 128),
          ("pool_finish", "pool_finish", 16384, 0, 128),
          ("fps_reg", "fps_reg", 16384, 128, 0), ("ball_query", "ball_query", 4096, 512, 64),
          ("mlp_layer_tt_kernel<1, 0, 0, 0, 0>", "mlp_tt<gather,tf32x3,Wtmem>", 262144, 128, 128),
          ("mlp_layer_tt_kernel<0, 1, 0, 0, 0>", "mlp_tt<plain,f16x3,Wtmem>", 262144, 128, 128),
          ("mlp_layer_tt_kernel<0, 1, 0, 1, 1>", "mlp_tt<plain,f16x3,Wtmem,pool,pair>", 262144, 128, 256),
          ("pool_finish", "pool_finish", 4096, 0, 256),
          ("prep_wimg", None, 0, 0, 0),
          ("mlp_layer_tt_kernel<1, 0, 1, 0, 0>", "mlp_tt<gather,tf32x3,Wstream>", 4096, 256, 256),
          ("prep_ximg", None, 0, 0, 0),
          ("mlp_layer_tt_kernel<0, 1, 0, 0, 0>", "mlp_tt<plain,f16x3,Wtmem>", 4096, 256, 512),
          ("prep_ximg", None, 0, 0, 0),
          ("prep_wimg", None, 0, 0, 0),
          ("mlp_layer_tt_kernel<0, 1, 1, 1, 0>", "mlp_tt<plain,f16x3,Wstream,pool>", 4096, 512, 1024),
          ("pool_finish", "pool_finish", 32, 0, 1024)]
assert len(summ) == len(LABELS), (len(summ), len(LABELS))
ks, total = [], 0.0
for s, (key, name, M, cin, cout) in zip(summ, LABELS):
    assert key in s["kernel"], (key, s["kernel"])
    by = (s["dram_read_MB"] + s["dram_write_MB"]) * 1e6
    total += by
    if name is None:
        continue
    ks.append({"name": name, "M": M, "cin": cin, "cout": cout, "dram_bytes": int(by),
               "dram_read_bytes": int(s["dram_read_MB"] * 1e6), "dram_write_bytes": int(s["dram_write_MB"] * 1e6),
               "ncu_us": round(s["us"], 1), "tensor_pipe_active_pct": round(s["tensor_pipe_active_pct"], 1),
               "sass_kernel": s["kernel"]})
out = {"source": f"{summary_out}: ncu --set full --clock-control none, the second (warm) pass of tools/prof_step.py "
                 "(B=32 x 1024 points), one launch each; dram__bytes_read.sum + dram__bytes_write.sum per launch",
       "dram_bytes_per_step": int(total), "kernels": ks}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print("DRAM bytes per step: %.1f MB over %d launches" % (total / 1e6, len(summ)))
