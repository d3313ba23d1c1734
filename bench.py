#!/usr/bin/env python
"""Benchmark of the B200-native TAI / bi-TAI hot path (contract: see the task statement, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Default workload = BASELINE.json configs[1]: the KTH bi-TAI TRAINING STEP (model_key TAI_gray, c_dim 1,
128x128, K = F = T = 5, batch 32 per GPU; generator + spectral-norm discriminator, losses and Adam as in
the reference's TAITrainingEnvironment), on synthetic U(-1,1) clips with xavier-normal weights.  One
"step" = forward + loss + backward + both optimiser updates.  Data-parallel over clips, weak scaling
(32 clips per GPU), gradients all-reduced over NCCL.  The line printed by rank 0 carries:

  value      frames/s (= N * B * T * K / time) with the batch already resident in HBM
  e2e        the same metric through the public API with HOST (pinned) batches: H2D copies of the three
             clip tensors and a D2H read of the losses inside the timed region, every step
  roofline   the costliest kernel of THIS library inside the timed region, timed with CUDA events on its
             own stream (tai_b200_timing_*): algorithmic flop / measured time against the FP32 FMA peak
  cpu_baseline   the CPU port of the reference step (oracle/reference_model.py) on a bounded sample

``--impl reference`` times that CPU port alone (the reference has no CPU path and cannot be imported under
Python 3 / torch 2 -- DESIGN.md), on rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bi-TAI inpainted frames/sec (KTH training step: fwd + bwd + Adam, generator and discriminator)"
UNIT = "frames/s"


def metric_name(workload):
    if workload == "slomo_train_b4":
        return "Super SloMo interpolated frames/sec (training step: fwd + bwd + Adam)"
    if WORKLOADS[workload][8]:
        return METRIC
    return "bi-TAI inpainted frames/sec (inference forward pass, %s)" % workload

WORKLOADS = {
    # name: model_key, c_dim, H, W, K, T, F, batch per GPU, training?
    "kth_train_b32": ("TAI_gray", 1, 128, 128, 5, 5, 5, 32, True),
    "kth_infer_b1": ("TAI_gray", 1, 128, 128, 5, 5, 5, 1, False),
    "ucf_infer_b8": ("TAI_color", 3, 240, 320, 5, 3, 5, 8, False),
    "slomo_infer_b8": ("SloMoFillInModel_color", 3, 256, 320, 2, 3, 2, 8, False),
    # the reference's own UCF test arguments (exp_args/default_args/UCF-101/test_3.txt:1-7): K = F = 4, batch 16
    "ucf_infer_ref443_b16": ("TAI_color", 3, 240, 320, 4, 3, 4, 16, False),
    # Super SloMo training step at the reference's training size and default batch (SuperSloMo_train.txt:3,
    # options.py:22,164-175); the perceptual loss runs on randomly initialised VGG-16 features (no network here)
    "slomo_train_b4": ("SloMoFillInModel_color", 3, 160, 192, 4, 3, 4, 4, True),
}
TRAIN_HP = dict(alpha=1.0, beta=0.02, lr=1e-4, beta1=0.5, df_dim=64, Ip=3, disc_window_size=3)  # options.py:72-99


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kth_train_b32", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--tf32", action="store_true", help="allow TF32 in the cuDNN convolutions (reported in config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="inference workloads: eager launches instead of CUDA-graph replay")
    return ap.parse_args()


SEPCONV_KERNELS = ("sepconv_fused_fwd", "sepconv_fwd", "sepconv_bwd_vh", "sepconv_bwd_i")
WARP_KERNELS = ("slomo_interp_input", "slomo_refine_blend_t", "slomo_interp_input_bwd", "slomo_refine_blend_t_bwd",
                "slomo_combine_warp", "slomo_refine_blend", "warp_fwd", "warp_bwd")
HEADLINE_KERNELS = {"kth_train_b32": SEPCONV_KERNELS, "kth_infer_b1": SEPCONV_KERNELS, "ucf_infer_b8": SEPCONV_KERNELS,
                    "ucf_infer_ref443_b16": SEPCONV_KERNELS, "slomo_infer_b8": WARP_KERNELS, "slomo_train_b4": WARP_KERNELS}


def load_traffic(workload):
    """profiles/dram_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from one
    `ncu --set full` capture per kernel, at the launch shape of the named workload."""
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if not os.path.isfile(path):
        return {}
    with open(path) as f:
        return json.load(f).get("workloads", {}).get(workload, {})


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        clocks, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                clocks.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        clocks.sort()
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": mx,
                "samples": len(clocks), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU port of the reference (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------

def workload_config(workload, B, world, tf32=False, graph=False):
    """The `config` object of the JSON line: the same dict for the B200 arm and the reference arm."""
    key, c, H, W, K, T, F_, _, training = WORKLOADS[workload]
    return {"workload": workload, "model_key": key, "clips_per_gpu": B, "global_clips": B * world,
            "frame": [c, H, W], "K": K, "T": T, "F": F_, "training": training, "ks": 51,
            "conv_math": "tf32" if tf32 else "fp32 (cudnn.allow_tf32=False)",
            "launch": "cuda-graph replay of the forward pass" if graph else "eager",
            "parallelism": "dp%d (clips sharded, NCCL all-reduce of gradients only)" % world,
            "l2": "per-step working set (activations, 4 x 107 MB kernel maps per middle frame) >> 126 MB L2"}


def cpu_reference_step_factory(workload, spatial=None, clips=1):
    import torch
    from oracle.reference_model import CpuTAITrainingStep, to_cpu_reference
    from video_frame_inpainting_b200.models.create_model import create_model
    from video_frame_inpainting_b200.util.util import weights_init
    key, c, H, W, K, T, F_, _, training = WORKLOADS[workload]
    if spatial:
        H, W = spatial
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    clip = torch.rand(clips, K + T + F_, c, H, W, generator=g) * 2 - 1
    pre, mid, fol = clip[:, :K].contiguous(), clip[:, K:K + T].contiguous(), clip[:, K + T:].contiguous()
    if training:
        st = CpuTAITrainingStep(create_model(key), (H, W), c, K, T, F_, alpha=TRAIN_HP["alpha"], beta=TRAIN_HP["beta"],
                                lr=TRAIN_HP["lr"], beta1=TRAIN_HP["beta1"], df_dim=TRAIN_HP["df_dim"],
                                Ip=TRAIN_HP["Ip"], disc_t=TRAIN_HP["disc_window_size"])
        fn = lambda: st.step(pre, fol, mid)
    else:
        model = create_model(key)
        model.apply(weights_init)
        model = to_cpu_reference(model).eval()

        def fn():
            with torch.no_grad():
                return model(T, pre, fol)
    frames = T * clips
    sample = "%s: 1 step on %d clip(s) of the %dx%d workload (%d middle frames each), FP32, torch CPU convs + C port " \
             "of the reference kernels (OpenMP), all host threads" % ("training" if training else "inference", clips,
                                                                     H, W, T)
    return fn, frames, sample


def cpu_clips_for_budget(workload, budget_s, nsteps, cap):
    """Clips per CPU step such that `nsteps` steps fit `budget_s`: one clip is timed first (after a warm-up),
    the batch is then sized from it (a larger batch only uses the host cores better)."""
    fn, _, _ = cpu_reference_step_factory(workload)
    fn()
    t0 = time.time()
    fn()
    t1 = time.time() - t0
    clips = int(budget_s / (max(1, nsteps) * max(t1, 1e-3)))
    return max(1, min(cap, clips)), t1


def run_reference(args):
    """--impl reference: the CPU port, rank 0 only (the other ranks exit 0 without work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is the CPU implementation with all host threads
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import torch
    B = args.batch or WORKLOADS[args.workload][7]
    nsteps = args.steps + max(0, args.warmup)
    # a bounded sample of the workload: as many of the B clips per step as keep the whole run near four minutes
    clips, t1 = cpu_clips_for_budget(args.workload, 240.0, nsteps, B)
    spatial = None
    if t1 * nsteps > 420.0:  # even one clip per step does not fit: crop the frames
        spatial = (64, 64)
        clips = 1
    fn, frames, sample = cpu_reference_step_factory(args.workload, spatial, clips)
    if spatial:
        sample += " [cropped to 64x64: the full-size one-clip step took %.1f s]" % t1
    for _ in range(max(1, args.warmup)):
        fn()
    t0 = time.time()
    for _ in range(args.steps):
        fn()
    dt = time.time() - t0
    value = frames * args.steps / dt
    cores = torch.get_num_threads()
    sample += "; %d of the %d clips of a step, %d warm-up + %d timed steps, %.1f s" % (clips, B, max(1, args.warmup),
                                                                                    args.steps, dt)
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, B, args.gpus, args.tf32,
                                  graph=not WORKLOADS[args.workload][8] and not args.no_graph),
        "device": "host CPU: the reference has no CPU path; this is the CPU port of its step (oracle/reference_model.py)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------

def build_workload(workload, batch=0, tf32=False, rank=0, local_rank=0, graph=True):
    """Model + environment + synthetic clips of a named workload on cuda:local_rank.
    Returns (step_resident, step_e2e, info)."""
    import torch
    from video_frame_inpainting_b200 import _lib
    from video_frame_inpainting_b200.environments.environments import (BaseVideoFillInEnvironment,
                                                                      TAITrainingEnvironment)
    from video_frame_inpainting_b200.models.create_model import create_model

    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.benchmark = True
    if os.environ.get("TAI_CUDNN_BENCHMARK_LIMIT"):           # experiment switch: 0 = let cuDNN try every algorithm
        torch.backends.cudnn.benchmark_limit = int(os.environ["TAI_CUDNN_BENCHMARK_LIMIT"])
    _lib.load()  # fail loudly here if the CUDA library is missing

    key, c, H, W, K, T, F_, B, training = WORKLOADS[workload]
    if batch:
        B = batch
    torch.manual_seed(0)  # same weights on every rank
    model = create_model(key)
    if os.environ.get("TAI_BATCH_STREAMS") in ("0", "1") and hasattr(model, "batch_streams"):
        model.batch_streams = os.environ["TAI_BATCH_STREAMS"] == "1"   # A/B switch for profiles/; default: model's own
    if os.environ.get("TAI_BATCH_HISTORY") in ("0", "1"):
        for m in model.modules():
            if hasattr(m, "batch_history"):
                m.batch_history = os.environ["TAI_BATCH_HISTORY"] == "1"
    if os.environ.get("TAI_BATCH_TIME") in ("0", "1") and hasattr(model, "batch_time"):
        model.batch_time = os.environ["TAI_BATCH_TIME"] == "1"
    if training and key.startswith("SloMo"):
        import warnings
        from video_frame_inpainting_b200.environments.environments import SloMoTrainingEnvironment
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")   # random VGG-16 features: stated in the workload's description
            env = SloMoTrainingEnvironment(model, "/tmp/tai_b200_ckpt", "bench", 1e-4, 0.5, K, T, F_, (0, 0), 0.8, 0.005,
                                           0.4, 1, 40000, 0.1)
        env.train()
    elif training:
        env = TAITrainingEnvironment(model, "/tmp/tai_b200_ckpt", "bench", (H, W), c, TRAIN_HP["alpha"],
                                     TRAIN_HP["beta"], TRAIN_HP["lr"], TRAIN_HP["beta1"], TRAIN_HP["df_dim"],
                                     TRAIN_HP["Ip"], TRAIN_HP["disc_window_size"], K, T, F_, (0, 0))
        env.train()
    else:
        env = BaseVideoFillInEnvironment(model, "/tmp/tai_b200_ckpt", "bench", (0, 0))
        env.eval()
        env.enable_cuda_graph(bool(graph))  # replay of the captured forward; the training step stays eager
    env.K, env.T, env.F = K, T, F_

    g = torch.Generator().manual_seed(1000 + rank)  # different clips per rank
    host = torch.rand(B, K + T + F_, c, H, W, generator=g) * 2 - 1
    h_pre = host[:, :K].contiguous().pin_memory()
    h_mid = host[:, K:K + T].contiguous().pin_memory()
    h_fol = host[:, K + T:].contiguous().pin_memory()
    d_pre, d_mid, d_fol = h_pre.to(dev), h_mid.to(dev), h_fol.to(dev)

    def step_resident():
        if training:
            env.preceding_frames, env.following_frames, env.gt_middle_frames = d_pre, d_fol, d_mid
            env.forward_train()
            env.optimize_parameters()
        else:
            env.preceding_frames, env.following_frames = d_pre, d_fol
            env.forward_test()

    def step_e2e():
        if training:
            env.set_train_inputs(h_pre, h_fol, h_mid)          # H2D from pinned memory
            env.forward_train()
            env.optimize_parameters()
            return env.get_current_errors()                     # D2H of the loss scalars
        env.set_test_inputs(h_pre, h_fol)
        env.forward_test()
        return float(env.gen_output['pred'].abs().mean())      # D2H of a result metric

    h2d = (h_pre.numel() + h_fol.numel() + (h_mid.numel() if training else 0)) * 4
    info = dict(key=key, c=c, H=H, W=W, K=K, T=T, F=F_, B=B, training=training, h2d=h2d,
                d2h=(10 if training else 1) * 4, env=env, graph=bool(graph) and not training)
    return step_resident, step_e2e, info


def make_step(workload, batch=0):
    """(step, info) for tools/step_profile.py and other single-GPU harnesses."""
    import torch
    torch.cuda.set_device(0)
    step, _, info = build_workload(workload, batch)
    return step, info


def run_b200(args):
    import torch
    import torch.distributed as dist
    from video_frame_inpainting_b200 import _lib
    from video_frame_inpainting_b200.parallel import init_distributed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    rank, local_rank, world = init_distributed()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    step_resident, step_e2e, info = build_workload(args.workload, args.batch, args.tf32, rank, local_rank,
                                                   graph=not args.no_graph)
    key, c, H, W, K, T, F_, B, training = (info[k] for k in ("key", "c", "H", "W", "K", "T", "F", "B", "training"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    for _ in range(max(3, args.warmup)):
        step_resident()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if info["graph"]:
        # Kernels replayed from a CUDA graph are not re-issued by the library, so its launch counter and event
        # timers see nothing: count / time them in ONE eager pass (same kernels, same order), then time replays.
        env = info["env"]
        saved, env._graphs = env._graphs, None
        launches0 = _lib.launch_count()
        _lib.timing_enable(True)
        step_resident()
        torch.cuda.synchronize()
        kernel_times = _lib.timing_report()
        _lib.timing_enable(False)
        per_step_launches = _lib.launch_count() - launches0
        for k in kernel_times:                      # scale the single pass to the timed region
            for f in ("ms", "launches", "flops", "bytes"):
                k[f] *= args.steps
        env._graphs = saved
        step_resident()                              # capture
        ms = timed(step_resident, args.steps)
        launches = per_step_launches * args.steps
    else:
        launches0 = _lib.launch_count()
        _lib.timing_enable(True)
        ms = timed(step_resident, args.steps)
        kernel_times = _lib.timing_report()
        _lib.timing_enable(False)
        launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    step_e2e()  # warm the pinned-copy path
    ms_e2e = timed(step_e2e, args.steps)

    frames = world * B * T * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    h2d, d2h = info["h2d"], info["d2h"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: the kernels BASELINE.json's metric names, inside the timed region ----
    # Headline = the costliest kernel of the workload's headline set (the separable convolutions for the TAI
    # workloads: "sepconv % of FP32 FMA peak"; the warp / blend kernels for the SloMo baseline: "% of HBM"),
    # timed by CUDA events the library records on its own stream around each launch.  The costliest kernel of
    # the whole library (an epilogue kernel in the training step) is kept beside it as `library_top`.
    peaks = measured_peaks()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fma_peak = sms * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12  # TFLOP/s, nominal at the max SM clock
    fma_src = "%d SMs x 128 lanes x 2 x %.0f MHz (sm_max_mhz, %s); not a tensor-core kernel" % (
        sms, peaks["sm_max_mhz"], peaks["source"])
    total_kernel_ms = sum(k["ms"] for k in kernel_times) or 1.0
    kernel_times.sort(key=lambda k: -k["ms"])
    traffic = load_traffic(args.workload)

    def entry_of(k):
        avg_s = k["ms"] * 1e-3 / max(1, k["launches"])
        return {"kernel": k["name"], "launches_per_step": k["launches"] / args.steps,
                "avg_us": avg_s * 1e6, "share_of_library_time": k["ms"] / total_kernel_ms,
                "share_of_step": k["ms"] / ms,
                "tflops": k["flops"] / max(1, k["launches"]) / avg_s / 1e12 if k["flops"] else 0.0,
                "gbs": k["bytes"] / max(1, k["launches"]) / avg_s / 1e9,
                "algorithmic_bytes": k["bytes"] / max(1, k["launches"]),
                "algorithmic_flop": k["flops"] / max(1, k["launches"])}

    def roofline_of(e):
        t = traffic.get(e["kernel"])
        if e["tflops"] > 0:   # the separable convolutions: FP32 CUDA-core FMA roof (HBM figures beside it)
            r = {"kernel": e["kernel"], "bound": "fp32_fma", "achieved": e["tflops"], "peak": fma_peak,
                 "unit": "TFLOP/s", "frac": e["tflops"] / fma_peak, "peak_source": fma_src,
                 "hbm_gbs": e["gbs"], "hbm_frac": e["gbs"] / peaks["hbm_gbs"]}
        else:
            r = {"kernel": e["kernel"], "bound": "hbm", "achieved": e["gbs"], "peak": peaks["hbm_gbs"],
                 "unit": "GB/s", "frac": e["gbs"] / peaks["hbm_gbs"], "peak_source": peaks["source"]}
        r["traffic"] = t["dram_bytes"] if t else None
        r["traffic_shape"] = t["shape"] if t else None
        r["algorithmic"] = e["algorithmic_flop"] if e["tflops"] > 0 else e["algorithmic_bytes"]
        r["avg_launch_us"] = e["avg_us"]
        r["launches_per_step"] = e["launches_per_step"]
        r["share_of_step"] = e["share_of_step"]
        return r

    per_kernel = [entry_of(k) for k in kernel_times]
    roofline = None
    if per_kernel:
        headline = [e for e in per_kernel if e["kernel"] in HEADLINE_KERNELS[args.workload]]
        roofline = roofline_of(headline[0] if headline else per_kernel[0])
        # every headline kernel, flat (scalar keys survive any parser) and as objects
        for e in headline:
            r = roofline_of(e)
            short = e["kernel"].replace("sepconv_", "")
            roofline["%s_frac" % short] = r["frac"]
            roofline["%s_avg_us" % short] = r["avg_launch_us"]
            roofline["%s_traffic" % short] = r["traffic"]
        roofline["headline_kernels"] = {e["kernel"]: roofline_of(e) for e in headline}
        roofline["library_top"] = roofline_of(per_kernel[0])
        roofline["kernels"] = per_kernel

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and not (training and key.startswith("SloMo")):
        torch.cuda.empty_cache()
        clips, _ = cpu_clips_for_budget(args.workload, 12.0, 2, B)   # one warm-up + about 10 s of CPU work
        fn, cframes, sample = cpu_reference_step_factory(args.workload, clips=clips)
        fn()                                   # warm-up (thread pools, oneDNN primitive caches)
        t0 = time.time()
        n = 0
        while n < 1 or (time.time() - t0 < 10.0 and n < 64):   # a bounded sample: about 10 s of CPU work
            fn()
            n += 1
        dt = time.time() - t0
        cpu_baseline = {"value": cframes * n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": sample + "; 1 warm-up + %d timed steps, %.1f s" % (n, dt)}

    line = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, B, world, args.tf32, info["graph"]),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
