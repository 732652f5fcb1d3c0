#!/usr/bin/env python
"""bench.py -- denoising-step frames/s, Zeroscope 24 f 576x320 (72x40 latent, bs=1) on N B200s  (BASELINE.json metric).

A "step" is one iteration of the reference's denoising loop body (controllable_pipeline_text_to_video_synth.py:
836-958) on one prompt: [guidance iterations when enabled] + classifier-free-guidance UNet forward (B=2: uncond, cond) +
CFG combine + DPM-Solver++ update.  frames/s = 24 * (prompts in flight) / step time.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores (oracle port)

Contract notes
  * value   : device-timed (CUDA events), inputs resident in HBM, max over ranks, whole-job aggregate.
  * e2e     : same step through the public API with HOST buffers: pinned latents + prompt embeddings copied H2D and
              the updated latents copied D2H inside the timed region, every step.
  * L2      : the step streams 2.8 GB of weights and >20 GB of activations, far larger than the 126 MB L2, so no
              explicit flush is needed between timed iterations (config.l2 says so).
  * roofline: dominant kernel family = gemm_tc_kernel (Linear + implicit-GEMM Conv2d/Conv3d); algorithmic FLOPs of
              every launch / its CUDA-event duration, measured in an instrumented step after the timed region.
  * cpu_baseline / --impl reference: oracle/ (the CPU restatement of the reference; the reference itself is Python
              that needs diffusers, not installable offline) on a bounded sample: the same CFG step on 2 of the 24
              frames with the full 1.41 B-parameter UNet, fp32, all host threads.
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

METRIC = "denoising-step frames/s, Zeroscope 24f 576x320"
UNIT = "frames/s"
FRAMES, LAT_H, LAT_W = 24, 40, 72
GUIDANCE_SCALE = 9.0
CFG_FWD_TFLOP = 42.79  # algorithmic work of one CFG forward at this shape
GUID_ITER_TFLOP = 32.59  # minimal guidance iteration: forward prefix 15.92 + dgrad 16.67
# lvd_zeroscope runs guidance on the first 10 of 50 steps (max_index_step=10, generation/lvd.py:92-104): the timed
# steps follow the same 1:4 pattern -- step i is guided (one guidance iteration, README.md:68 max_iter=1) iff i % 5 == 0
GUIDED_EVERY = 5
GUIDANCE_KEYS = [("down", 1, 0, 0), ("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0),
                 ("up", 2, 2, 0)]  # generation/lvd.py:66-73
GUIDANCE_HP = dict(fg_top_p=0.25, bg_top_p=0.25, fg_weight=1.0, bg_weight=2.0)  # README.md:68
GUIDANCE_LOSS_SCALE = 2.5


def bear_layout(frames):
    """The in-repo demo layout "A bear walks from the left to the right" (cache/cache_demo_v0.1_gpt-4-1106-preview.json):
    one object sweeping x from 0 to 0.83 at y 0.5..0.695, interpolated to `frames` boxes; prompt token 2."""
    boxes = []
    for f in range(frames):
        x0 = 0.83 * f / max(frames - 1, 1) * 0.8
        boxes.append((x0, 0.5, x0 + 0.17, 0.695))
    return [boxes], [[2]]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--unguided", action="store_true", help="time the unguided CFG step only")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lvd_b200", choices=["lvd_b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=2)
    ap.add_argument("--profile-json", default=None, help="write the per-kernel-family breakdown here")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def pick_cpu_threads():
    """All the host threads the reference can USE: PyTorch's CPU kernels stop scaling (and regress) well before 128
    threads on the small per-layer problems of the bounded sample, so probe a conv + matmul at a few thread counts
    (<2 s) and keep the fastest; the count used is reported as `cores`."""
    import torch
    n = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, n) if c <= n})
    x = torch.randn(2, 320, 40, 72)
    w = torch.randn(320, 320, 3, 3)
    a = torch.randn(5760, 320)
    b = torch.randn(320, 1280)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.conv2d(x, w, padding=1)
            a @ b
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def cpu_reference_step(sample_frames, steps, warmup, guided_every=GUIDED_EVERY):
    """The reference algorithm (oracle/ restatement, fp32, all host threads) on a bounded sample of the workload:
    the same step pattern (guided iff i % 5 == 0) on `sample_frames` of the 24 frames, full-size UNet.
    Returns (frames/s, mean seconds/step, cores, seconds of a guided step, seconds of an unguided step)."""
    import torch
    from oracle import guidance_ref, scheduler_ref
    from oracle import unet3d_ref as R
    from oracle import weights

    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    cfg = weights.make_cfg("zeroscope")
    sd = weights.random_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    lat = torch.randn((1, 4, sample_frames, LAT_H, LAT_W), generator=g)
    ehs = torch.randn((2, 77, 1024), generator=g)
    sched = scheduler_ref.DPMSolverRef()
    sched.set_timesteps(50)
    bboxes, positions = bear_layout(sample_frames)
    t_g, t_u = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        tt = int(sched.timesteps[min(it, 49)])
        guided = guided_every and (it % guided_every == 0)
        if guided:
            lat, _, _, _ = guidance_ref.guidance_step(sd, cfg, lat, tt, ehs[1:], bboxes, positions,
                                                      sched.alphas_cumprod[tt], guidance_attn_keys=GUIDANCE_KEYS,
                                                      loss_scale=GUIDANCE_LOSS_SCALE, base_attn_dim=(LAT_H, LAT_W),
                                                      **GUIDANCE_HP)
        with torch.no_grad():
            x = torch.cat([lat, lat], 0)
            eps = R.unet_forward(sd, cfg, x, tt, ehs)
            e = scheduler_ref.cfg_combine(eps[:1], eps[1:], GUIDANCE_SCALE)
            lat = sched.step(e, lat)
        dt = time.perf_counter() - t0
        if it >= warmup:
            (t_g if guided else t_u).append(dt)
    sec_g = sum(t_g) / len(t_g) if t_g else None
    sec_u = sum(t_u) / len(t_u) if t_u else None
    sec = (sum(t_g) + sum(t_u)) / (len(t_g) + len(t_u))
    return sample_frames / sec, sec, cores, sec_g, sec_u


def _workload(unguided):
    if unguided:
        return "lvd_zeroscope 1 prompt/GPU, 24 frames 576x320 (72x40 latent), unguided CFG denoising step"
    return ("lvd_zeroscope 1 prompt/GPU, 24 frames 576x320 (72x40 latent), 50-step schedule pattern: step i runs one "
            "backward-guidance iteration (6 guided cross-attention keys, demo 'bear' layout) iff i % 5 == 0 "
            "(= 10 guided of 50), then the CFG UNet forward (B=2) + CFG combine + DPM-Solver++ update")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ge = 0 if args.unguided else GUIDED_EVERY
    fps, sec, cores, sec_g, sec_u = cpu_reference_step(args.cpu_sample_frames, max(1, args.steps),
                                                       max(0, min(args.warmup, 1)), ge)
    sample = (f"same step pattern on {args.cpu_sample_frames} of 24 frames (72x40 latent, full 1.41B-param UNet), fp32, "
              f"oracle port of the reference incl. autograd guidance; guided step {sec_g and round(sec_g, 1)} s, "
              f"unguided step {sec_u and round(sec_u, 1)} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload(args.unguided), "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    ge.build()
    from lvd_b200 import _lib, ops, synthetic
    from lvd_b200.engine import UNetEngine
    from lvd_b200.guidance_engine import GRAD_SCALE, GuidanceGraph
    from lvd_b200.scheduler import DPMSolverMultistepScheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl lvd_b200) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = synthetic.make_cfg("zeroscope")
    sd = synthetic.random_state_dict(cfg, seed=0, dtype=torch.float16)
    eng = UNetEngine(cfg, sd, device=dev)
    del sd

    # one independent work item (prompt) per rank: NCCL only carries the seed scatter and the timing reduction
    seeds = torch.arange(world, dtype=torch.int64, device=dev) + 1234
    if world > 1:
        dist.broadcast(seeds, src=0)
    g = torch.Generator().manual_seed(int(seeds[rank].item()))
    lat_host = torch.randn((1, 4, FRAMES, LAT_H, LAT_W), generator=g).half().pin_memory()
    ehs_host = torch.randn((2, 77, 1024), generator=g).half().pin_memory()

    sched = DPMSolverMultistepScheduler()
    sched.set_timesteps(50)
    n_img = 2 * FRAMES
    static_x = torch.empty((2, 4, FRAMES, LAT_H, LAT_W), dtype=torch.float16, device=dev)
    static_t = torch.empty((n_img,), dtype=torch.float32, device=dev)
    static_ehs = ehs_host.to(dev)
    latents = lat_host.to(dev)
    state = {"out": None, "graph": None}

    def forward_eager():
        state["out"] = eng.forward(static_x, static_t, static_ehs)

    graph = GuidanceGraph(eng)
    bboxes, positions = bear_layout(FRAMES)
    guided_every = 0 if args.unguided else GUIDED_EVERY
    cond = static_ehs[1:2]
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def guidance_iteration(i):
        """One iteration of latent_backward_guidance's loop (models/pipelines.py:65-137) on the native pass."""
        nonlocal latents
        tt = int(sched.timesteps[i % 50])
        loss, grad = graph.run(latents, tt, cond, bboxes, positions, GUIDANCE_KEYS, loss_scale=GUIDANCE_LOSS_SCALE,
                               base_attn_dim=(LAT_H, LAT_W), **GUIDANCE_HP)
        scale = (1.0 - float(sched.alphas_cumprod[tt])) ** 0.5
        latents, _ = ops.latent_update(latents, grad, scale, 1.0 / GRAD_SCALE)
        loss_host.copy_(loss, non_blocking=True)  # the 4-byte D2H the reference's while-condition needs

    def denoise_step(i):
        """One denoising-loop iteration on device-resident inputs (latents updated in place)."""
        nonlocal latents
        if guided_every and i % guided_every == 0:
            guidance_iteration(i)
        static_x[0].copy_(latents[0])
        static_x[1].copy_(latents[0])
        static_t.fill_(float(sched.timesteps[i % 50]))
        if state["graph"] is not None and ops.PROF is None:
            state["graph"].replay()
        else:
            forward_eager()
        eps = state["out"]
        if sched._step_index >= 49:
            sched._reset()
        latents = sched.step_cfg(eps[0:1], eps[1:2], latents, GUIDANCE_SCALE)

    # warm-up (also sets kernel attributes before any graph capture)
    for i in range(max(3, args.warmup)):
        denoise_step(i)
    torch.cuda.synchronize()
    use_graph = not args.no_graph
    if use_graph:
        try:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                forward_eager()
            state["graph"] = gr
            denoise_step(0)
            torch.cuda.synchronize()
        except Exception as e:  # graph capture is an optimisation of launch overhead, not a different code path
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e}); running eagerly\n")
            state["graph"] = None
            use_graph = False
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- timed region 1: device-resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    launches0 = _lib.lib().lvd_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        denoise_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib().lvd_launch_count() - launches0
    # a CUDA graph replays the captured launches without re-entering the library: count them from one eager step
    # ---------------- timed region 2: end to end with host buffers
    out_host = torch.empty((1, 4, FRAMES, LAT_H, LAT_W), dtype=torch.float16).pin_memory()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        latents.copy_(lat_host, non_blocking=True)
        static_ehs.copy_(ehs_host, non_blocking=True)
        denoise_step(i)
        out_host.copy_(latents, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    sampler.stop()
    clocks = sampler.summary()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---------------- separate timings of the two step kinds (device events, resident inputs)
    def time_steps(idx):
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in idx:
            denoise_step(i)
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / len(idx)

    ms_unguided = time_steps([1, 2, 3])
    ms_guided = time_steps([0, 5]) if guided_every else None

    # ---------------- instrumented step: per-kernel-family durations (CUDA events around every launch)
    ops.PROF = []
    l0 = _lib.lib().lvd_launch_count()
    static_t.fill_(float(sched.timesteps[0]))
    forward_eager()
    torch.cuda.synchronize()
    launches_per_forward = _lib.lib().lvd_launch_count() - l0
    n_fwd_records = len(ops.PROF)
    launches_per_guidance = 0
    if guided_every:
        l1 = _lib.lib().lvd_launch_count()
        guidance_iteration(0)
        torch.cuda.synchronize()
        launches_per_guidance = _lib.lib().lvd_launch_count() - l1
    fam = {}
    fam_guid = {}
    shapes = ({}, {})
    for k_, (name, fl, nb, s, e, tag) in enumerate(ops.PROF):
        d = (fam if k_ < n_fwd_records else fam_guid).setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0,
                                                                        "launches": 0})
        if tag is not None:
            sd = shapes[0 if k_ < n_fwd_records else 1].setdefault(tag, {"ms": 0.0, "flops": 0.0, "bytes": 0.0,
                                                                                  "launches": 0})
            sd["ms"] += s.elapsed_time(e)
            sd["flops"] += fl
            sd["bytes"] += nb
            sd["launches"] += 1
        d["ms"] += s.elapsed_time(e)
        d["flops"] += fl
        d["bytes"] += nb
        d["launches"] += 1
    ops.PROF = None
    if use_graph:
        n_guided = sum(1 for i in range(args.steps) if guided_every and i % guided_every == 0)
        launches = (launches_per_forward + 1) * args.steps + (launches_per_guidance + 1) * n_guided
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained"
    gm = dict(fam.get("gemm_tc_kernel", {"ms": 1.0, "flops": 0.0, "launches": 1}))
    gg = fam_guid.get("gemm_tc_kernel")
    if gg:  # one guided step = CFG forward + one guidance iteration: the dominant kernel's launches of both
        share = 1.0 / GUIDED_EVERY
        gm = {"ms": gm["ms"] + share * gg["ms"], "flops": gm["flops"] + share * gg["flops"],
              "launches": gm["launches"] + share * gg["launches"], "bytes": gm.get("bytes", 0.0) + share * gg["bytes"]}
    achieved = gm["flops"] / (gm["ms"] * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:  # DRAM bytes per gemm_tc_kernel launch from the committed ncu capture (profiles/, measured, not live)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")))
        traffic, traffic_src = tj["traffic_bytes_per_launch"], "profiles/r1_gemm_traffic.json (ncu dram__bytes_*.sum)"
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": gm.get("bytes", 0.0) / max(1, gm["launches"]),
                "peak_source": peak_src,
                "launches_per_step": gm["launches"], "ms_per_step": gm["ms"],
                "share_of_instrumented_step": gm["ms"] / max(1e-9, sum(v["ms"] for v in fam.values()) + (
                    sum(v["ms"] for v in fam_guid.values()) / GUIDED_EVERY if fam_guid else 0.0)),
                "note": "FLOPs and CUDA-event time summed over all gemm_tc_kernel launches of one CFG forward plus "
                        "1/5 of one guidance iteration (the timed step mix), instrumented pass after the timed region"}
    if args.profile_json and rank == 0:
        json.dump({"cfg_forward": fam, "guidance_iteration": fam_guid, "gemm_shapes_cfg_forward": shapes[0],
                   "gemm_shapes_guidance_iteration": shapes[1]}, open(args.profile_json, "w"), indent=1)

    if rank == 0:
        ms_step = ms / args.steps
        value = world * FRAMES / (ms_step * 1e-3)
        e2e_val = world * FRAMES / (ms_e2e / args.steps * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": _workload(args.unguided),
                       "weights": "seeded random init of the Zeroscope UNet3D config (1411.2 M params), fp16",
                       "parallelism": f"replicas x{world} (independent prompts, no data-path collective)",
                       "l2": "inputs larger than L2 (2.8 GB weights + >20 GB activations per step vs 126 MB)",
                       "cuda_graph": "CFG forward replayed from a CUDA graph; guidance pass eager" if use_graph else False,
                       "algorithmic_tflop_per_step": CFG_FWD_TFLOP + (GUID_ITER_TFLOP / GUIDED_EVERY if guided_every
                                                                      else 0.0),
                       "guidance": None if not guided_every else {
                           "keys": GUIDANCE_KEYS, "layout": "demo bear, 1 object, token 2", "iterations_per_guided_step": 1,
                           "loss_scale": GUIDANCE_LOSS_SCALE, **GUIDANCE_HP}},
            "unguided_step": {"ms": ms_unguided, "frames_per_s": world * FRAMES / (ms_unguided * 1e-3)},
            "guided_step": None if ms_guided is None else {"ms": ms_guided,
                                                           "frames_per_s": world * FRAMES / (ms_guided * 1e-3)},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": lat_host.numel() * 2 + ehs_host.numel() * 2,
                    "d2h_bytes_per_step": out_host.numel() * 2},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "step_tflops": (CFG_FWD_TFLOP + (GUID_ITER_TFLOP / GUIDED_EVERY if guided_every else 0.0)) / (ms_step * 1e-3),
            "kernel_families_ms": {k: round(v["ms"], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
            "guidance_families_ms": {k: round(v["ms"], 3) for k, v in sorted(fam_guid.items(),
                                                                            key=lambda kv: -kv[1]["ms"])},
        }
        if world == 1 and not args.no_cpu_baseline:
            # one guided + one unguided step on the CPU, combined with the schedule weights 1:4
            _, _, cores, sec_g, sec_u = cpu_reference_step(args.cpu_sample_frames, 2, 0, 0 if args.unguided else 2)
            if args.unguided:
                sec = sec_u
            else:
                sec = (sec_g + (GUIDED_EVERY - 1) * sec_u) / GUIDED_EVERY
            line["cpu_baseline"] = {
                "value": args.cpu_sample_frames / sec, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"same step mix on {args.cpu_sample_frames} of 24 frames (72x40 latent, full UNet), fp32, "
                          f"oracle port of the reference (autograd guidance): guided step "
                          f"{sec_g and round(sec_g, 1)} s, unguided {sec_u and round(sec_u, 1)} s, weighted 1:4"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
