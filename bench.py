#!/usr/bin/env python
"""bench.py -- PointNet++ SSG SetAbstraction forward points/sec on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path -- sa1 -> sa2 -> sa3 of PointNet2_SSG_Clas
(PAPC/models/classify/pointnet2/pointnet2.py:11-16, 33-35) -- over one batch of synthetic
1024-point clouds.  At N=1 the workload is BASELINE.json configs[1] (B=32, N=1024).  For N>1 the
batch is sharded over the ranks (weak scaling: 32 clouds per GPU, so N=8 is configs[4], B=256),
BatchNorm statistics are per shard, and the step ends with the ONE all-gather of the per-shard
l3 features.

JSON keys: value = whole-job points/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = the same metric through the public layer API from pinned HOST buffers (H2D of the clouds and
D2H of the l3 features inside the timed region); roofline = the dominant kernel (grouped-MLP layer
GEMM) timed live with CUDA events on its stream; cpu_baseline = the oracle (NumPy restatement of
the reference path) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_POINTS = 1024
B_PER_GPU = 32
SA_CFG = [  # (npoint, radius, nsample, in_channel, mlp, group_all) -- classify/pointnet2/pointnet2.py:11-16
    (512, 0.2, 32, 3, [64, 64, 128], False),
    (128, 0.4, 64, 131, [128, 128, 256], False),
    (None, None, None, 259, [256, 512, 1024], True),
]
METRIC = "PointNet++SSG SetAbstraction fwd points/sec"
UNIT = "points/s"


def flops_per_cloud(n=N_POINTS):
    """Algorithmic FLOPs of the grouped MLPs (2*rows*cin*cout), SURVEY.md 8(d): 53.6 G per 32 clouds."""
    total, layers = 0, []
    npts = n
    for (S, _, K, cin, mlp, ga) in SA_CFG:
        rows = npts if ga else S * K
        c = cin
        for co in mlp:
            layers.append((rows, c, co))
            total += 2 * rows * c * co
            c = co
        npts = 1 if ga else S
    return total, layers


def algorithmic_bytes_per_cloud(n=N_POINTS):
    """Compulsory API-boundary traffic per cloud, SURVEY.md 8(d) (~800 B/point)."""
    total, npts, d = 0, n, 0
    for (S, _, K, cin, mlp, ga) in SA_CFG:
        s = 1 if ga else S
        total += 12 * npts + 4 * npts * d + 12 * s + 4 * mlp[-1] * s
        npts, d = s, mlp[-1]
    return total


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_step(xyz, starts, params, acc):
    """One pass of the oracle (NumPy restatement of layers.py:179-221, statement by statement,
    including the [B,S,N] distance matrix + sort of query_ball_point) over a batch."""
    from oracle import layers_np
    x, p = xyz, None
    for i, cfg in enumerate(SA_CFG):
        sa = layers_np.PointNetSetAbstraction(*cfg)
        sa.acc = acc
        for l, prm in enumerate(params[i]):
            sa.mlp_convs[l].weight = prm["weight"].reshape(*prm["weight"].shape, 1, 1)
            sa.mlp_convs[l].bias = prm["bias"]
        x, p = sa(x, p, start_idx=starts[i])
    return p


def time_cpu_baseline(steps, warmup, batch):
    """Oracle on the host cores: returns (points/s, ms/step, cores, sample description)."""
    from papc_b200 import synth
    cores = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    try:  # torchrun exports OMP_NUM_THREADS=1: give NumPy's BLAS all the host cores back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    xyz = synth.clouds(batch, N_POINTS, seed=0)
    starts = [synth.fps_start(batch, N_POINTS, seed=1), np.zeros(batch, np.int64), None]
    params = [synth.mlp_params(c[3], c[4], seed=2 + i) for i, c in enumerate(SA_CFG)]
    for _ in range(warmup):
        cpu_reference_step(xyz, starts, params, np.float32)
    t = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(xyz, starts, params, np.float32)
        t.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(t))
    sample = (f"{steps} timed passes (after {warmup} warm-up) of the full sa1+sa2+sa3 oracle over "
              f"{batch} clouds x {N_POINTS} points, fp32 NumPy/BLAS with {cores} threads")
    return batch * N_POINTS / (ms / 1e3), ms, cores, sample


WORKLOAD = ("PointNet++SSG classify SetAbstraction stack sa1+sa2+sa3, 1024-pt clouds, "
            "B=32 per GPU (BASELINE configs[1]; N=8 is configs[4], B=256)")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  PaddlePaddle cannot
    be installed here (no wheel, no network), so this is the oracle port -- the line-by-line NumPy
    restatement of pointnet2_basic_layers.py -- on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 8  # bounded sample: 8 of the 32 clouds per pass keeps K+W passes within a few minutes
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    value, ms, cores, sample = time_cpu_baseline(steps, warmup, batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU * max(1, args.gpus), "n_points": N_POINTS,
                   "bn": "train-mode batch statistics, as the reference's unregistered SA layers run",
                   "sample": f"{batch} of the {B_PER_GPU} clouds per step (points/s is per point, so the "
                             "rate is that of the full batch on the same cores)",
                   "note": "Paddle unavailable: CPU restatement of the reference path (oracle port), rank 0 only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from papc_b200 import _lib, layers, sa_stack, synth
    from papc_b200 import dist as pdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: papc_b200 has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    B = B_PER_GPU
    Bg = B * world
    xyz_all = synth.clouds(Bg, N_POINTS, seed=0)
    st1_all = synth.fps_start(Bg, N_POINTS, seed=1)
    lo, hi = pdist.shard_range(Bg, rank, world)
    xyz_h = torch.from_numpy(xyz_all[lo:hi]).pin_memory()
    st1 = torch.from_numpy(st1_all[lo:hi]).to(dev)
    st2 = torch.zeros(B, dtype=torch.int64, device=dev)
    model = sa_stack.SSGSetAbstractionStack().to(dev)
    for i, sa in enumerate(model.layers_()):
        sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(SA_CFG[i][3], SA_CFG[i][4], seed=2 + i))
    xyz_d = xyz_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out_h = torch.empty((B, 1024), dtype=torch.float32).pin_memory()

    def step_eager():
        _, l3 = model(xyz_d, None, start_idx=(st1, st2))
        return l3.reshape(B, 1024)

    # The step as the library's public graph API runs it: the whole sa1 -> sa2 -> sa3 forward captured
    # once (papc_b200.sa_stack.GraphedForward) and replayed; --no-graph times the eager calls instead.
    graphed = None
    if not args.no_graph:
        graphed = sa_stack.GraphedForward(lambda x: model(x, None, start_idx=(st1, st2)), xyz_d)

    def forward(x=None):
        if graphed is None:
            _, l3 = model(xyz_d if x is None else x, None, start_idx=(st1, st2))
        else:
            _, l3 = graphed.replay() if x is None else graphed(x)
        return l3.reshape(B, 1024)

    def step_device():
        feats = forward()
        if world > 1:
            feats = pdist.all_gather_features(feats)
        return feats

    def step_e2e():
        if graphed is None:
            feats = forward(xyz_h.to(dev, non_blocking=True))
        else:
            feats = forward(xyz_h)          # H2D straight into the graph's static input buffer
        if world > 1:
            g = pdist.all_gather_features(feats)
            feats = g[lo:hi]
        out_h.copy_(feats, non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        l0 = lib.papc_launch_count()
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (outside the timed interval)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        launches = lib.papc_launch_count() - l0
        if graphed is not None:  # replays launch the captured kernels without passing the host counter
            launches += graphed.kernels_per_replay * steps
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    total_ms, launches = timed(step_device, args.steps, args.warmup)
    e2e_ms, _ = timed(step_e2e, args.steps, args.warmup)
    if sampler:
        sampler.stop()

    points_per_step = Bg * N_POINTS
    value = points_per_step * args.steps / (total_ms / 1e3)
    e2e_value = points_per_step * args.steps / (e2e_ms / 1e3)

    # ---- roofline of the dominant kernel: every kernel of the step timed live with CUDA events on
    #      its launching stream by the library's launch profiler (papc_prof_*), same step function
    kernels = profile_kernels(torch, lib, _lib, step_eager, flush, min(args.steps, 10))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    # kernels timed inside a long step -> the sustained figure (B200_PROFILING.md); fallback 1.59 PF
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    peak_src = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if "bf16_tflops_sustained" in peaks
                else "fallback 1.59 PFLOP/s (B200_PROFILING.md)")
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    flops_cloud, _ = flops_per_cloud()
    step_ms = total_ms / args.steps
    # the dominant KERNEL = the kernel function with the largest share of the step, all its launches
    # together (mlp_layer_tt_kernel runs 8 times per step on different layer shapes)
    fams = {}
    for k in kernels:
        f = fams.setdefault(k["name"].split("<")[0], {"ms": 0.0, "launches": 0.0, "flops": 0.0, "bytes": 0.0,
                                                      "traffic": 0.0, "traffic_ok": True, "members": []})
        f["ms"] += k["ms_per_step"]
        f["launches"] += k["per_step"]
        f["flops"] += k["flops"] * k["per_step"]
        f["bytes"] += k["bytes"] * k["per_step"]
        t = lookup_traffic(k)
        if t is None:
            f["traffic_ok"] = False
        else:
            f["traffic"] += t * k["per_step"]
        f["members"].append(k)
    fam_name, fam = max(fams.items(), key=lambda kv: kv[1]["ms"])
    dom = max(fam["members"], key=lambda k: k["ms_per_step"])
    avg_ms = fam["ms"] / fam["launches"]
    traffic = fam["traffic"] / fam["launches"] if fam["traffic_ok"] else None
    sass = {"mlp_tt": "papc::tt::mlp_layer_tt_kernel (tcgen05 grouped-MLP layer: producers -> UMMA -> fused BN-stat / "
                      "max-pool epilogue)", "fps_reg": "papc::fps_reg_kernel", "ball_query": "papc::ball_query_kernel"}
    if fam["flops"] > 0:
        ach = fam["flops"] / (fam["ms"] / 1e3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "traffic": traffic, "peak_source": peak_src,
                    "note": "achieved = algorithmic 2*M*cin*cout of the kernel's launches / their live CUDA-event time "
                            "(average launch); the fp32-parity operand split issues 3 tensor-core products per "
                            "algorithmic product, so frac <= 1/3 of the measured bf16 peak"}
    else:
        ach = fam["bytes"] / (fam["ms"] / 1e3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": traffic, "peak_source": hbm_src}
    roofline.update({
        "kernel": sass.get(fam_name, fam_name), "avg_launch_ms": avg_ms, "launches_per_step": fam["launches"],
        "share_of_step": fam["ms"] / step_ms,
        "algorithmic_flops_per_launch": fam["flops"] / fam["launches"],
        "algorithmic_bytes_per_launch": fam["bytes"] / fam["launches"],
        "hbm_gbs_algorithmic": fam["bytes"] / (fam["ms"] / 1e3) / 1e9, "hbm_peak_gbs": hbm,
        "largest_launch": {"name": dom["name"], "M": dom["M"], "cin": dom["cin"], "cout": dom["cout"],
                           "avg_ms": dom["ms"], "tflops": dom["flops"] / (dom["ms"] / 1e3) / 1e12 if dom["flops"] else None,
                           "traffic": lookup_traffic(dom)},
        "fps": next(({"avg_ms": k["ms"], "ns_per_iteration": k["ms"] * 1e6 / max(k["cin"], 1),
                      "note": "latency bound: npoint dependent argmax steps per cloud, one CTA per cloud"}
                     for k in sorted(kernels, key=lambda k: -k["ms"]) if k["name"] == "fps_reg"), None),
        "kernels": [{"name": k["name"], "M": k["M"], "cin": k["cin"], "cout": k["cout"],
                     "launches_per_step": k["per_step"], "avg_ms": round(k["ms"], 5),
                     "share_of_step": round(k["ms_per_step"] / step_ms, 4),
                     "tflops": round(k["flops"] / (k["ms"] / 1e3) / 1e12, 2) if k["flops"] > 0 else None,
                     "gbs": round(k["bytes"] / (k["ms"] / 1e3) / 1e9, 1) if k["bytes"] > 0 else None}
                    for k in sorted(kernels, key=lambda k: -k["ms_per_step"])],
        "kernel_time_share_of_step": sum(k["ms_per_step"] for k in kernels) / step_ms,
        "whole_step": {"achieved_tflops": flops_cloud * Bg * args.steps / (total_ms / 1e3) / 1e12,
                       "achieved_hbm_gbs_algorithmic": algorithmic_bytes_per_cloud() * Bg * args.steps / (total_ms / 1e3) / 1e9},
    })

    cpu_v, cpu_ms, cores, sample = time_cpu_baseline(steps=2, warmup=1, batch=8)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": Bg, "n_points": N_POINTS, "parallelism": f"batch-shard x{world}",
                   "bn": "train-mode batch statistics (per shard), as the reference's unregistered SA layers run",
                   "collective": "one all-gather of l3 features" if world > 1 else "none",
                   "launch": "eager calls" if graphed is None else "CUDA-graph replay of the captured forward "
                             "(sa_stack.GraphedForward); per-kernel times in roofline.kernels come from eager passes",
                   "l2": "256 MiB buffer rewritten between timed iterations (outside the timed interval)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(xyz_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "ms_per_step": cpu_ms},
        "clocks": sampler.summary() if sampler else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def profile_kernels(torch, lib, L, step_fn, flush, steps):
    """Run `steps` more passes of the step with the library's launch profiler on: every kernel is
    bracketed by CUDA events on its launching stream.  -> per distinct (kernel, shape): average
    launch ms, launches per step, ms per step, algorithmic FLOPs / bytes per launch."""
    step_fn()
    torch.cuda.synchronize()
    L.check(lib.papc_prof_reset(), "prof_reset")
    L.check(lib.papc_prof_enable(1), "prof_enable")
    for _ in range(steps):
        flush.zero_()
        step_fn()
    torch.cuda.synchronize()
    L.check(lib.papc_prof_enable(0), "prof_enable")
    recs = L.prof_records()
    L.check(lib.papc_prof_reset(), "prof_reset")
    agg = {}
    for r in recs:
        key = (r["name"], r["M"], r["cin"], r["cout"])
        a = agg.setdefault(key, {"name": r["name"], "M": r["M"], "cin": r["cin"], "cout": r["cout"],
                                 "flops": r["flops"], "bytes": r["bytes"], "n": 0, "tot": 0.0})
        a["n"] += 1
        a["tot"] += r["ms"]
    out = []
    for a in agg.values():
        a["ms"] = a["tot"] / a["n"]
        a["per_step"] = a["n"] / steps
        a["ms_per_step"] = a["tot"] / steps
        out.append(a)
    return out


def lookup_traffic(k):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel + shape from the
    committed `ncu --set full` capture (profiles/traffic.json), or None if it was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    try:
        table = json.load(open(path))
    except Exception:
        return None
    for e in table.get("kernels", []):
        if e["name"] == k["name"] and e["M"] == k["M"] and e["cin"] == k["cin"] and e["cout"] == k["cout"]:
            return e["dram_bytes"]
    return None


def layer_roofline(torch, lib, layers, synth, dev, B, only=None, reps=5, warm=3):
    """Time each grouped-MLP layer launch (the C-ABI step call papc_mlp_layer_forward_f32 launches
    exactly one GEMM kernel) with CUDA events on the launching stream, on the bench's own shapes."""
    import ctypes as C
    from papc_b200 import _lib as L
    rng = np.random.default_rng(0)
    st = L.stream_ptr(dev)
    res = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    npts, d = N_POINTS, 0
    for si, (S, r, K, cin, mlp, ga) in enumerate(SA_CFG):
        s, k = (1, npts) if ga else (S, K)
        M = B * s * k
        xyz = torch.from_numpy(np.ascontiguousarray(synth.clouds(B, npts, seed=si).transpose(0, 2, 1))).to(dev)
        feats = torch.randn((B, npts, d), device=dev) if d else None
        if ga:
            new_xyz, idx = None, None
        else:
            fidx, new_xyz = layers.farthest_point_sample_idx(xyz, s, torch.zeros(B, dtype=torch.int64, device=dev), True)
            idx = layers._ball_query(r, k, xyz, new_xyz, torch.int32)
        src = layers._make_src(xyz, new_xyz, feats, idx, B, npts, s, k, L.XYZ_FIRST)
        x, c = None, cin
        scale = shift = None
        for li, co in enumerate(mlp):
            last = li == len(mlp) - 1
            w = torch.from_numpy((rng.standard_normal((co, c)) * np.sqrt(2.0 / c)).astype(np.float32)).to(dev)
            bias = torch.zeros(co, device=dev)
            y = None if last else torch.empty((M, co), dtype=torch.float32, device=dev)
            pmax = torch.empty((B * s, co), device=dev) if last else None
            pmin = torch.empty((B * s, co), device=dev) if last else None
            partial = torch.empty((lib.papc_mlp_stats_partial_rows(M), 2, co), dtype=torch.float64, device=dev)
            lwsb = lib.papc_mlp_layer_workspace_bytes(c, co)
            lws = torch.empty(max(lwsb, 256), dtype=torch.uint8, device=dev)

            def launch():
                L.check(lib.papc_mlp_layer_forward_f32(C.byref(src) if li == 0 else None, L.ptr(x), L.ptr(scale),
                                                       L.ptr(shift), M, c, co, k, L.ptr(w), L.ptr(bias), L.ptr(y),
                                                       L.ptr(pmax), L.ptr(pmin), L.ptr(partial), L.ptr(lws), lwsb,
                                                       st), "layer")
            name = f"sa{si + 1}.l{li + 1} [{M}x{c}]x[{c}x{co}]"
            if only and not any(o in name for o in only):
                scale = torch.ones(co, device=dev)
                shift = torch.zeros(co, device=dev)
                if y is not None:
                    y.normal_()
                x, c = y, co
                continue
            for _ in range(warm):
                launch()
            torch.cuda.synchronize()
            ms = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                launch()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = float(np.mean(ms))
            fl = 2.0 * M * c * co
            res.append({"name": name, "ms": t, "flops": fl,
                        "tflops": fl / (t / 1e3) / 1e12})
            scale = torch.ones(co, device=dev)
            shift = torch.zeros(co, device=dev)
            x, c = y, co
        npts, d = s, mlp[-1]
    dom = max(res, key=lambda r: r["ms"])
    return {"layers": res, "dominant": dom, "mlp_ms_total": sum(r["ms"] for r in res)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time eager calls instead of CUDA-graph replays")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
