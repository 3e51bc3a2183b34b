#!/usr/bin/env python
"""bench.py - frames/s (fwd+bwd) of the MaGGIe hot path on N B200s (BASELINE.json metric, config C2/C3).

  python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A step = model(batch) + (loss * scale).backward() (+ one flat gradient all-reduce when N>1) on a synthetic
batch of 8 frames x 512x512 x 3 instances per GPU (weak scaling), training mode, iter=1.
`value`  : inputs already resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same step through the public API with the batch in pinned HOST memory: H2D copies and the D2H
           read of the loss are inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU, H, W, N_INST, EDGE_PX = 8, 512, 512, 3, 6.0
LOSS_SCALE = 128.0
F_DENSE_PER_FRAME = 69.23e9 * (H * W) / (512 * 512)


def f_sparse(counts):
    n1, n2, n4, n8 = counts
    return 2.0 * (23072 * n1 + 7680 * n2 + 113952 * n4 + 8192 * n8)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        z = json.load(open(p))
        return dict(hbm=z["hbm_gbs"], tf_burst=z["bf16_tflops"], tf_sustained=z["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.skip = [], None, index, 0

    def mark(self, wait_s=30.0):
        """The timed region starts here: samples taken so far (warm-up) are dropped.  The sampler is started long before
        (nvidia-smi's start-up - NVML initialisation over all GPUs of the box - stalls CUDA submissions for seconds), and
        the timed region does not begin until its first sample has arrived."""
        t0 = time.time()
        while self.proc is not None and not self.rows and self.proc.poll() is None and time.time() - t0 < wait_s:
            time.sleep(0.05)
        self.skip = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        if len(self.rows) > self.skip:
            self.rows = self.rows[self.skip:]
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_fn(frames=2):
    """Oracle port (CPU fp32 restatement of the reference path) fwd+bwd on a `frames`-frame sample of C2."""
    import torch

    from oracle import maggie_oracle as O
    from oracle import make_golden as G
    import synthdata as synth

    torch.set_num_threads(os.cpu_count())
    import numpy as np
    z = np.load(os.path.join(G.GOLDEN_DIR, "state_shapes.npz"))
    tmpl = {k: torch.zeros(tuple(z[k]), dtype=torch.long if k.endswith("num_batches_tracked") else torch.float32)
            for k in z.files}
    P = synth.synth_state_dict(tmpl)
    for k, v in P.items():
        if v.is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var")) \
                and "dummy_downscale" not in k:
            v.requires_grad_(True)
    batch = synth.make_batch(b=frames, n_f=1, n_i=N_INST, H=H, W=W, edge_px=EDGE_PX, train=True, it=1)
    cfg = synth.model_cfg()

    def step():
        G.seed_all()
        for v in P.values():
            v.grad = None
        _, loss = O.forward(P, batch, True, cfg)
        loss["total"].backward()
        return float(loss["total"])

    return step, frames


def cpu_c1_eval_ms():
    """Oracle port, eval forward of BASELINE config C1 (1 x 256x256, 1 instance), median of 3 after one warm-up."""
    import numpy as np
    import torch

    from oracle import maggie_oracle as O
    from oracle import make_golden as G
    import synthdata as synth

    z = np.load(os.path.join(G.GOLDEN_DIR, "state_shapes.npz"))
    tmpl = {k: torch.zeros(tuple(z[k]), dtype=torch.long if k.endswith("num_batches_tracked") else torch.float32)
            for k in z.files}
    P = synth.synth_state_dict(tmpl)
    batch = synth.make_batch(b=1, n_f=1, n_i=1, H=256, W=256, edge_px=6.0)
    cfg = synth.model_cfg()
    ts = []
    with torch.no_grad():
        for i in range(4):
            t0 = time.perf_counter()
            O.forward(P, batch, False, cfg)
            if i:
                ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[1]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, frames = cpu_step_fn(2)
    warm = min(args.warmup, 1)  # one warm-up is what a ~15 s CPU step affords; stated in `sample`
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = frames / dt
    sample = f"{frames} frames x {H}x{W} x {N_INST} inst per step (BatchNorm needs >=2), fp32, {warm} warm-up"
    c1 = cpu_c1_eval_ms()
    print(json.dumps({
        "impl": "reference", "metric": "frames_per_sec_fwd_bwd", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2: {FRAMES_PER_GPU}x{H}x{W}x{N_INST}-inst train fwd+bwd (CPU sample: {frames} frames/step)"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        # BASELINE config C1 (the reference's own CPU-runnable case), for the record: eval forward of one 256x256 image
        "c1_eval_forward": {"ms": c1, "frames_per_sec": 1e3 / c1, "sample": "1 x 256x256 x 1 inst, 1 warm-up, median of 3"},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
# (H=W of the input feature map, Cin, Cout, k, stride, dilation, transposed, count) of every dense conv of one C2 forward
C2_CONVS = [
    (512, 32, 32, 3, 2, 1, False, 1), (256, 32, 32, 3, 1, 1, False, 3), (256, 32, 64, 3, 2, 1, False, 1),
    (128, 64, 64, 3, 1, 1, False, 8), (128, 64, 128, 3, 2, 1, False, 1), (128, 64, 128, 2, 2, 1, False, 1),
    (64, 128, 128, 3, 1, 1, False, 14), (64, 128, 256, 3, 2, 1, False, 1), (64, 128, 256, 2, 2, 1, False, 1),
    (32, 256, 256, 3, 1, 1, False, 11), (32, 256, 512, 3, 2, 1, False, 1), (32, 256, 512, 2, 2, 1, False, 1),
    (16, 512, 512, 3, 1, 1, False, 3), (512, 32, 32, 3, 1, 1, False, 1), (512, 32, 32, 3, 1, 1, False, 1),
    (16, 512, 256, 1, 1, 1, False, 1), (16, 512, 256, 3, 1, 2, False, 1), (16, 512, 256, 3, 1, 4, False, 1),
    (16, 512, 256, 3, 1, 8, False, 1), (1, 512, 256, 1, 1, 1, False, 1), (16, 1280, 512, 1, 1, 1, False, 1),
    (16, 512, 512, 4, 2, 1, True, 1), (32, 512, 256, 3, 1, 1, False, 1), (16, 512, 256, 1, 1, 1, False, 1),
    (32, 256, 256, 4, 2, 1, True, 1), (64, 256, 128, 3, 1, 1, False, 1), (32, 256, 128, 1, 1, 1, False, 1),
    (64, 128, 64, 1, 1, 1, False, 1),
]


def _traffic_table():
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels below from the committed
    `ncu --set full` captures (profiles/r2_ncu_traffic.json, written by tools/ncu_kernels.py); {} when absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {}


def kernel_probes(torch, dev, peaks):
    """Live CUDA-event timing of this repo's own kernels at the C2 shapes.  Every entry is the AVERAGE launch duration over a
    timed region of back-to-back launches on rotating operand sets (more than the 126 MB L2 in total), replayed from a CUDA
    graph so that no host launch overhead sits between the events.  The conv family is timed layer by layer over the whole
    C2 layer table (forward, data gradient, weight gradient); achieved = sum of algorithmic FLOPs (or bytes) / sum of
    average launch durations x launches per step."""
    from maggie_b200 import _lib, dense, ops, sparse
    import synthdata as synth

    L = _lib.lib()
    traffic = _traffic_table()

    def burst(fn, nsets, rounds=3):
        for i in range(nsets):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(rounds):
                    for i in range(nsets):
                        fn(i)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3 / (rounds * nsets))
        return sorted(ts)[1]

    def nsets_for(nbytes):
        return int(max(2, min(12, (300 << 20) // max(int(nbytes), 1) + 1)))

    out = {}

    def hbm_entry(name, t, nbytes, note=None, launches=1):
        out[name] = dict(bound="hbm", achieved=nbytes / t / 1e9, peak=peaks["hbm"], unit="GB/s", frac=nbytes / t / 1e9 / peaks["hbm"],
                         traffic=traffic.get(name), us_per_launch=t * 1e6, algorithmic_bytes=nbytes, launches_per_step=launches)
        if note:
            out[name]["note"] = note

    # ---- bandwidth kernels
    als = [torch.stack([synth.soft_ellipse_alphas(1, 10, H, W, EDGE_PX, seed=s + 10 * k)[0] for s in range(FRAMES_PER_GPU)]).to(dev)
           for k in range(3)]
    widths = [15] * (FRAMES_PER_GPU * 10)
    hbm_entry("unknown_mask_kernel", burst(lambda i: ops.unknown_mask(als[i], widths), 3), als[0].numel() * 5.0,
              "threshold + elliptical dilation of 80 planes of 512^2: 4 B read + 1 B written per pixel", launches=5)
    imgs = [torch.randn(FRAMES_PER_GPU, 3, H, W, device=dev) for _ in range(3)]
    msk = (als[0][:, :3] > 0.5).float().contiguous()
    tab = torch.randn(11, 3, device=dev)
    with torch.no_grad():
        hbm_entry("mask_embed_fwd_kernel", burst(lambda i: ops.mask_embed(imgs[i], msk, tab, [0, 1, 2], 32), 3),
                  FRAMES_PER_GPU * H * W * (6 * 4 + 64.0), "3 image + 3 mask planes fp32 read, 32 fp16 channels written per pixel")

    # ---- sparse refinement kernels (K9b / K9c) on a C2-sized OS1 site list: 3x3 SubM 32 -> 32
    roi = (als[0][:, :N_INST] > 1.0 / 255) & (als[0][:, :N_INST] < 254.0 / 255)
    T = ops.build_sites(ops.unknown_mask(als[0][:, :N_INST].contiguous(), [15] * (FRAMES_PER_GPU * N_INST)).reshape(-1, H, W))
    n1 = T.counts[0]
    srcs = [torch.randn(n1, 32, device=dev).half() for _ in range(6)]
    gys = [torch.randn(n1, 32, device=dev).half() for _ in range(6)]
    w9 = torch.randn(32, 3, 3, 32, device=dev) / 17.0
    wp = sparse.pack_fwd(w9)
    t = burst(lambda i: sparse.sparse_conv_launch(srcs[i], wp, 9, 32, 32, table=T.nbr[0]), 6)
    rows_bytes = n1 * (32 * 2 * 2 + 9 * 4.0)          # rows read once + written once (gathers hit L2) + the rulebook
    hbm_entry("sparse_conv_persistent_kernel", t, rows_bytes,
              f"SubM 3x3 32->32 on {n1} active sites: rows in + rows out + 36 B of rulebook per site (SURVEY 8d byte model)", launches=12)
    out["sparse_conv_persistent_kernel"]["tflops"] = 2.0 * n1 * 9 * 32 * 32 / t / 1e12
    t = burst(lambda i: sparse._wgrad(gys[i], 32, srcs[i], 32, T.nbr[0], 9), 6)
    hbm_entry("sparse_wgrad_persistent_kernel", t, rows_bytes,
              f"weight gradient of the same layer: d_out rows + gathered source rows + rulebook", launches=5)
    del srcs, gys

    # ---- conv family: every layer of the C2 table, forward / data gradient / weight gradient, weight packs prepared outside
    # the timed region (in the step they come from the grouped K0 kernel).  Launches are accounted to the kernel
    # mg_conv_fprop / mg_conv_wgrad routed them to: K2 (generic), K2b / K4b (halo, <= 64 channels: HBM-bound, in bytes),
    # K2t (mid-resolution 3x3 layers, transposed form; K2h is opt-in).
    orig_pack, memo = dense.pack_weight, {}

    def cached_pack(w, ci_pad=None):
        key = (w.data_ptr(), tuple(w.shape), tuple(w.stride()), ci_pad)
        if key not in memo:
            memo[key] = (orig_pack(w, ci_pad), w)
        return memo[key][0]

    dense.pack_weight = cached_pack
    names = ["conv_tcgen05_kernel[fprop]", "conv_tcgen05_kernel[dgrad]", "conv_midt_tcgen05_kernel[fprop]",
             "conv_midt_tcgen05_kernel[dgrad]", "conv_mid_tcgen05_kernel[fprop]", "conv_mid_tcgen05_kernel[dgrad]",
             "wgrad_tcgen05_kernel", "conv_halo_tcgen05_kernel", "wgrad_halo_tcgen05_kernel"]
    tot = {k: [0.0, 0.0, 0] for k in names}
    try:
        for (hw, ci, co, k, s, d, tr, cnt) in C2_CONVS:
            N = FRAMES_PER_GPU
            w = torch.randn((ci, co, k, k) if tr else (co, ci, k, k), device=dev) / (ci * k * k) ** 0.5
            g = dense.ConvGeom("convT", 4, 2, 1, 1) if tr else dense.ConvGeom("conv", k, s, d * (k // 2) if k > 1 else 0, d)
            if not tr and k == 2:
                g = dense.ConvGeom("conv", 2, 2, 0, 1)
            x0 = torch.randn(N, hw, hw, ci, device=dev).half()
            y0 = g.fwd(x0, w)
            nbytes = 2.0 * (x0.numel() + y0.numel())
            ns = nsets_for(nbytes)
            xs = [x0] + [torch.randn_like(x0) for _ in range(ns - 1)]
            ys = [y0] + [torch.randn_like(y0) for _ in range(ns - 1)]
            flops = 2.0 * y0.shape[0] * y0.shape[1] * y0.shape[2] * co * ci * (4 if tr else k * k)
            dwp = torch.zeros_like(dense.pack_weight(w.permute(1, 0, 2, 3) if tr else w, ci), dtype=torch.float32)

            class Bank:   # accumulate into a pre-zeroed pack as the weight bank does (no fill inside the timing)
                G = dwp

            for kind, fn, nl in (("fprop", lambda i: g.fwd(xs[i], w), 1), ("dgrad", lambda i: g.dgrad(ys[i], w, x0.shape), 1),
                                 ("wgrad", lambda i: g.wgrad(ys[i], xs[i], w.shape, bank=Bank), 4 if tr else 1)):
                h0, m0, w0, t0 = (L.mg_conv_halo_launches(), L.mg_conv_mid_launches(), L.mg_wgrad_halo_launches(),
                                  L.mg_conv_midt_launches())
                fn(0)
                halo = L.mg_conv_halo_launches() > h0 or L.mg_wgrad_halo_launches() > w0
                mid = "mid" if L.mg_conv_mid_launches() > m0 else ("midt" if L.mg_conv_midt_launches() > t0 else "")
                t = burst(fn, ns)
                if kind == "wgrad":
                    key = "wgrad_halo_tcgen05_kernel" if halo else "wgrad_tcgen05_kernel"
                else:
                    key = "conv_halo_tcgen05_kernel" if halo else (f"conv_{mid}_tcgen05_kernel[{kind}]" if mid else f"conv_tcgen05_kernel[{kind}]")
                tot[key][0] += (nbytes if halo else flops) * cnt
                tot[key][1] += t * cnt
                tot[key][2] += nl * cnt
            del xs, ys
    finally:
        dense.pack_weight = orig_pack
    for key, (work, t, nl) in tot.items():
        if nl == 0:
            continue
        if key in ("conv_halo_tcgen05_kernel", "wgrad_halo_tcgen05_kernel"):
            out[key] = dict(bound="hbm", achieved=work / t / 1e9, peak=peaks["hbm"], unit="GB/s",
                            frac=work / t / 1e9 / peaks["hbm"], traffic=traffic.get(key), us_per_step=t * 1e6,
                            algorithmic_bytes_per_step=work, launches_per_step=nl,
                            note="K2b / K4b launches of the C2 layer table (stride-1 3x3 layers with <= 64 channels at >= 128-wide "
                                 "resolutions); bytes = the two activation tensors the kernel streams")
        else:
            out[key] = dict(bound="tensor", achieved=work / t / 1e12, peak=peaks["tf_sustained"], unit="TFLOP/s",
                            frac=work / t / 1e12 / peaks["tf_sustained"], traffic=traffic.get(key.split("[")[0]), us_per_step=t * 1e6,
                            algorithmic_flops_per_step=work, launches_per_step=nl,
                            note="sum over the C2 conv layers routed to this kernel")
    for k2 in ("unknown_mask_kernel", "mask_embed_fwd_kernel", "sparse_conv_persistent_kernel", "sparse_wgrad_persistent_kernel"):
        out[k2]["us_per_step"] = out[k2]["us_per_launch"] * out[k2]["launches_per_step"]
    return out


def other_configs(torch, dev, args_no_graphs=False):
    """The remaining BASELINE.json configs on one GPU, each a short measurement (CUDA events, inputs resident, 3 warm-up +
    median of >= 5): C1 = eval forward of one 256 x 256 image with one instance (both precisions); C4 = video model, one
    5-frame 480 x 832 clip with 2 instances, training step; C5 = 1024 x 1024 x 8 instances, training step at three widths
    of the alpha transition band (= three active-site fractions of the sparse refinement)."""
    import numpy as np
    import random

    from maggie_b200.config import CfgNode
    from maggie_b200.network import build_model
    import synthdata as synth

    def med(fn, n=5, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(n):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    to_dev = lambda b: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items() if k not in ("fg", "bg")}
    res = {}
    torch.manual_seed(1234)
    model, _ = build_model(CfgNode(synth.model_cfg()))
    model.to(dev)

    def train_step(m, batch):
        np.random.seed(7), random.seed(7)
        for p in m.parameters():
            p.grad = None
        _, loss = m(batch, mem_feat=None)
        (loss["total"] * LOSS_SCALE).backward()

    # C1 (evaluation forwards replay the dense stage as a CUDA graph too)
    model.eval()
    model.enable_cuda_graphs(not args_no_graphs)
    b1 = to_dev(synth.make_batch(b=1, n_f=1, n_i=1, H=256, W=256, edge_px=6.0))
    c1 = {}
    with torch.no_grad():
        for prec in ("fp16", "high"):
            model.set_precision(prec)
            ms = med(lambda: model(b1, mem_feat=None), n=9)
            c1[prec] = {"ms": ms, "frames_per_sec": 1e3 / ms}
    model.set_precision("fp16")
    res["c1_eval_256_1inst"] = dict(c1, note="eval forward incl. the one host read of the status word; 'high' = fp32-accurate mode")
    # C5 sweep (training modes replay the dense stage as CUDA graphs, as the headline run does)
    model.train()
    model.enable_cuda_graphs(not args_no_graphs)
    sweep = []
    for edge in (3.0, 8.0, 24.0):
        b5 = to_dev(synth.make_batch(b=1, n_f=1, n_i=8, H=1024, W=1024, edge_px=edge, seed=77, train=True, it=1))
        ms = med(lambda: train_step(model, b5), n=5)
        n1 = model.last_site_counts[0]
        sweep.append({"edge_px": edge, "active_fraction": n1 / (8 * 1024 * 1024), "active_sites_os1": n1, "ms_per_step": ms,
                      "frames_per_sec": 1e3 / ms})
        del b5
    res["c5_train_1024_8inst_sweep"] = sweep
    del model
    torch.cuda.empty_cache()
    # C4
    vmodel, _ = build_model(CfgNode(synth.video_cfg()))
    vmodel.to(dev).train()
    vmodel.enable_cuda_graphs(not args_no_graphs)
    b4 = to_dev(synth.make_batch(b=1, n_f=5, n_i=2, H=480, W=832, edge_px=6.0, seed=9, train=True, it=1))
    ms = med(lambda: train_step(vmodel, b4), n=5)
    res["c4_video_train_5x480x832_2inst"] = {"ms_per_clip": ms, "clips_per_sec": 1e3 / ms, "frames_per_sec": 5e3 / ms,
                                             "active_sites_os1": vmodel.last_site_counts[0]}
    vmodel.eval()
    b4e = to_dev(synth.make_batch(b=1, n_f=3, n_i=2, H=480, W=832, edge_px=6.0, seed=9))
    with torch.no_grad():
        ms = med(lambda: vmodel(b4e, mem_feat=None), n=5)
    res["c4_video_eval_window_3x480x832_2inst"] = {"ms_per_window": ms, "windows_per_sec": 1e3 / ms}
    del vmodel
    torch.cuda.empty_cache()
    return res


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from maggie_b200 import _lib
    from maggie_b200.config import CfgNode
    from maggie_b200.dp import FlatGradAllReduce
    from maggie_b200.network import build_model
    import synthdata as synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.manual_seed(1234)  # identical initial weights on every rank
    model, _ = build_model(CfgNode(synth.model_cfg()))
    model.to(dev).train()
    sync_bn = bool(args.sync_bn) and world > 1
    if sync_bn:  # SyncBatchNorm-equivalent statistics exchange (`model.sync_bn: true`); the dense stage then runs eagerly
        from maggie_b200.dp import set_sync_bn
        set_sync_bn(True)
    model.enable_cuda_graphs(not args.no_graphs)
    flat = FlatGradAllReduce(model.parameters(), bank=model.bank)

    # The synthetic batch as a loader delivers it: uint8 frames / alphas / masks / transition maps (1 byte per value over
    # PCIe instead of 4); the input stage (K16, `maggie_b200.io.prepare_batch`: ToTensor + Normalize + dataset scaling) makes
    # the float tensors `MaGGIe.forward` takes on the GPU.  Both timed loops see the SAME inputs: the resident batch is the
    # output of that input stage.
    from maggie_b200 import io as mio
    f32 = synth.make_batch(b=FRAMES_PER_GPU, n_f=1, n_i=N_INST, H=H, W=W, edge_px=EDGE_PX, seed=1234 + rank, train=True,
                           it=args.iter)
    mean = torch.tensor(mio.IMAGENET_MEAN).view(1, 1, 3, 1, 1)
    std = torch.tensor(mio.IMAGENET_STD).view(1, 1, 3, 1, 1)
    host = {
        "frames": ((f32["image"] * std + mean) * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous(),
        "alpha": (f32["alpha"] * 255.0).round().clamp(0, 255).to(torch.uint8),
        "mask": (f32["mask"] * 255.0).round().clamp(0, 255).to(torch.uint8),
        "transition": f32["transition"].to(torch.uint8),
    }
    host = {k: v.pin_memory() for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    copy_stream = torch.cuda.Stream(device=dev)
    # two persistent sets of device input buffers (no per-step allocation); a set is refilled only after the step that
    # last read it has been enqueued AND finished on the compute stream
    dev_sets = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    set_free = [None, None]
    turn = [0]

    def to_dev():
        """H2D of one step's uint8 inputs from pinned memory on a copy stream (as a prefetching loader does) and the input
        stage behind them on the same stream; the batch carries the event that marks its tensors complete (`ready_event`,
        see MaGGIe._input_stage_async: the model's own input stage waits for that event only, not for the previous
        step's backward)."""
        i = turn[0] % 2
        turn[0] += 1
        if set_free[i] is not None:
            copy_stream.wait_event(set_free[i])
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                dev_sets[i][k].copy_(v, non_blocking=True)
            u8 = dev_sets[i]
            batch = mio.prepare_batch(u8["frames"], u8["alpha"], u8["mask"], downscale_mask=False)
            batch["transition"] = u8["transition"].float()
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        for v in batch.values():      # allocated on the copy stream, consumed on the compute stream
            v.record_stream(compute)
        batch["iter"] = args.iter
        batch["ready_event"] = ev
        batch["_set"] = i
        return batch

    def release(batch):
        ev = torch.cuda.Event()
        ev.record()
        set_free[batch["_set"]] = ev

    resident = to_dev()

    import numpy as np
    import random

    # Host flow control: the host may run at most two steps ahead of the GPU (as any training loop that reads its loss
    # does).  Unbounded run-ahead lets the caching allocator pile up blocks that other streams have not released yet.
    in_flight = []
    throttle = os.environ.get("MAGGIE_B200_BENCH_NO_THROTTLE", "0") != "1"

    def step(batch):
        if throttle and len(in_flight) >= 2:
            in_flight.pop(0).synchronize()
        np.random.seed(7), random.seed(7)
        flat.zero()
        _, loss = model(batch, mem_feat=None)
        (loss["total"] * LOSS_SCALE).backward()
        flat.allreduce()
        if throttle:
            ev = torch.cuda.Event()
            ev.record()
            in_flight.append(ev)
        return loss["total"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last_median = [None]

    def timed(fn, n):
        """Mean over exactly n steps between two events (barrier + synchronize on both sides, max over ranks); the median
        of the per-step intervals (one event per step, read after the region) is kept in `last_median`."""
        # (the cyclic garbage collector is parked for the timed region, as training loops do around their hot loop: a
        #  generation-2 pass over the process's ~10^6 objects showed up as one 20-80 ms step at random positions)
        gc.collect()
        gc.disable()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        e0.record()
        for i in range(n):
            fn()
            if i + 1 < n:
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
        e1.record()
        barrier()
        gc.enable()
        pts = [e0] + marks + [e1]
        raw = [a.elapsed_time(b) for a, b in zip(pts, pts[1:])]
        per = sorted(raw)
        last_median[0] = per[len(per) // 2]
        if os.environ.get("MAGGIE_B200_BENCH_VERBOSE", "0") == "1":
            print(f"[rank {rank}] per-step ms: " + " ".join(f"{t:.2f}" for t in raw), file=sys.stderr, flush=True)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    # multi-rank runs take a few more untimed steps: the first collectives after start-up (NCCL channel set-up, allocator
    # growth on every rank) showed up as one 20-100 ms step inside the first timed steps of a 5-step warm-up
    n_warm = max(args.warmup, 3) + (8 if world > 1 else 0)
    for _ in range(n_warm):
        step(resident)
    torch.cuda.synchronize()
    sampler.mark()
    _lib.reset_launch_count()
    replayed0 = model.replayed_native_launches
    ms = timed(lambda: step(resident), args.steps)
    ms_median = last_median[0]
    launches = _lib.launch_count() + (model.replayed_native_launches - replayed0)

    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    pending = []

    def e2e_step():
        """One step through the public API with HOST inputs: H2D of this step's batch from pinned memory (copy stream),
        forward + backward (+ all-reduce), and a D2H copy of the step's loss into pinned memory.  The loss VALUE is
        consumed one step later (after the next step has been enqueued), as a pipelined training loop logs it: the
        copy itself is issued and completed inside the timed region for every step."""
        batch = to_dev()
        loss = step(batch)
        release(batch)
        slot = len(pending) % 2
        loss_host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pending.append((ev, slot))
        if len(pending) > 1:
            pev, pslot = pending[-2]
            pev.synchronize()
            return float(loss_host[pslot])
        return None

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    pending[-1][0].synchronize()
    counts_main = list(model.last_site_counts)

    # the other training regime: iter >= 9000 - the predicted OS8 alpha guides the detail stage, the uncertain
    # region is known only after the dense stage (one host read of the status word in the middle of the step)
    post = None
    if not args.no_post_warmup:
        other_iter = 100000 if args.iter < 9000 else 1
        alt = dict(resident)
        alt["iter"] = other_iter
        for _ in range(3):
            step(alt)
        n_alt = max(10, args.steps // 3)
        ms_alt = timed(lambda: step(alt), n_alt)
        post = {"iter": other_iter, "value": FRAMES_PER_GPU * world / (ms_alt * 1e-3), "unit": "frames/s", "ms_per_step": ms_alt,
                "steps": n_alt, "active_sites_os1_os2_os4_os8": list(model.last_site_counts),
                "note": "same model / batch, inputs resident; iter >= 3 * warmup_detail_iter: uncertain region from the "
                        "predicted alpha (device-side switch), one 32-byte host read per step after the dense stage.  With "
                        "RANDOM weights the predicted alpha is uncertain everywhere, so this is also the 100 %-active stress "
                        "case of the sparse stage (compare the site counts)"}
        if other_iter > 1:
            # the same control flow at EQUAL work: the OS8 head's output is replaced by the ground-truth alpha (benchmark-only
            # substitution), so the uncertain region is the warm-up regime's while the switch, the mask, the site tables
            # and the host read all happen after the dense stage as in a trained model's late iterations
            dec = model.decoder
            orig = dec._os8_alpha
            gt_dev = alt["alpha"][:, 0].float()
            dec._os8_alpha = lambda logits, masks, n_i, Hh, Ww, slots=None, status=None: (
                orig(logits, masks, n_i, Hh, Ww, slots, status) * 0.0 + gt_dev)
            try:
                for _ in range(3):
                    step(alt)
                ms_eq = timed(lambda: step(alt), n_alt)
            finally:
                dec._os8_alpha = orig
            post["equal_work"] = {"value": FRAMES_PER_GPU * world / (ms_eq * 1e-3), "ms_per_step": ms_eq,
                                  "active_sites_os1_os2_os4_os8": list(model.last_site_counts),
                                  "note": "OS8 alpha replaced by the ground truth (benchmark-only): post-warm-up control flow "
                                          "(device-side switch, mask + site tables + host read after the dense stage) at the "
                                          "warm-up regime's active fraction"}
    last_loss = float(loss_host[pending[-1][1]])
    assert last_loss == last_loss, "loss is NaN"
    clocks = sampler.stop() if rank == 0 else None
    counts = counts_main
    if sync_bn:
        from maggie_b200 import dense as _dense
        sync_bn_path = ("peer-memory kernel (K15), 2 per BatchNorm per step" if all(w is not None for w in _dense._WINDOWS.values())
                        else "one all-reduce per BatchNorm and direction")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    frames = FRAMES_PER_GPU * world
    value, e2e = frames / (ms * 1e-3), frames / (ms_e2e * 1e-3)
    f_step = 3.0 * (F_DENSE_PER_FRAME * FRAMES_PER_GPU + f_sparse(counts))  # per GPU per step
    probes = kernel_probes(torch, dev, peaks)
    top = max(probes, key=lambda k: probes[k]["us_per_step"])
    roof = dict(probes[top], kernel=top, peak_source=peaks["src"])
    line = {
        "metric": "frames_per_sec_fwd_bwd", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms, "ms_per_step_median": ms_median, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"C2: {FRAMES_PER_GPU}x{H}x{W}x{N_INST}-inst train fwd+bwd per GPU (iter={args.iter}, edge {EDGE_PX}px)",
                   "frames_per_gpu": FRAMES_PER_GPU, "active_sites_os1_os2_os4_os8": counts,
                   "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "loss_scale": LOSS_SCALE, "host_run_ahead_steps": 2 if throttle else "unbounded", "sync_bn": sync_bn, "cuda_graphs_dense_stage": bool(model._graphs),
                   **({"sync_bn_exchange": sync_bn_path} if sync_bn else {})},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e, "h2d": "uint8 frames / alphas / masks / transition maps, pinned host memory -> device on a copy stream, every step; float tensors made on the GPU by the input stage (K16)",
                "d2h": "loss copied to pinned memory every step, value consumed one step later"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernels": probes,
        "flop_roofline": {"flop_per_step_per_gpu": f_step, "achieved_tflops": f_step / (ms * 1e-3) / 1e12,
                          "peak_tflops_sustained": peaks["tf_sustained"],
                          "frac": f_step / (ms * 1e-3) / 1e12 / peaks["tf_sustained"], "peak_source": peaks["src"]},
    }
    if post is not None:
        line["other_regime"] = post
    if world == 1 and not args.no_extras:
        del model, flat
        torch.cuda.empty_cache()
        line["other_configs"] = other_configs(torch, dev, args.no_graphs)
    if world == 1 and not args.no_cpu_baseline:
        cstep, cframes = cpu_step_fn(2)
        cstep()
        t0 = time.perf_counter()
        cstep()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cframes / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"1 timed step (after 1 warm-up) of {cframes} frames x {H}x{W} x {N_INST} inst, fp32 oracle port"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="run the dense stage eagerly instead of as CUDA graphs")
    ap.add_argument("--sync-bn", action="store_true", help="N>1: exchange BatchNorm statistics across ranks (model.sync_bn true)")
    ap.add_argument("--iter", type=int, default=1, help="training iteration of the headline number (1: warm-up regime of the "
                    "first 3000 iterations; >= 9000: the predicted OS8 alpha guides the detail stage)")
    ap.add_argument("--no-post-warmup", action="store_true", help="skip the extra iter=100000 measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1 / C4 / C5 measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
