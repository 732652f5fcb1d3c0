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
  * warm-up : exactly --warmup steps run before the timed region.  CUDA-graph capture of the step variants happens in
              an untimed SETUP phase ahead of the warm-up (config.setup), it is not counted as warm-up.
  * roofline: dominant kernel family = gemm_tc_kernel (Linear + implicit-GEMM Conv2d/Conv3d); algorithmic FLOPs of
              every launch / its duration measured by CUDA events that are nodes of the replayed graphs themselves
              (`external` events captured around every kernel node), over a region of --steps instrumented replays
              of the same step sequence right after the headline region (roofline.region).
  * cpu_baseline / --impl reference: oracle/ (the CPU restatement of the reference, pinned against the reference
              executed verbatim; the reference itself needs diffusers, not installable offline), fp32, the host threads
              a probe finds fastest, ONE code path for both: the same step mix on a bounded sample = the first
              `sample_frames` of the 24 frames with the full 1.41 B-parameter UNet; sample_frames is the largest of
              {24, 12, 8, 4, 2} whose --steps + --warmup steps fit the CPU time budget (--cpu-budget-s), so a short run
              (--steps 2 --warmup 0) times the full 24-frame configuration.
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


def prompt_batch_layouts(n, frames):
    """Deterministic stand-ins for the cached GPT-4 layouts of `--prompt-type lvd` (55 % one object, 35 %
    two, 5 % three, 5 % four; 6 % of the boxes absent): objects sweep across the frame, phrases of 1-2 tokens."""
    pattern = [1, 2, 1, 1, 2, 1, 3, 2, 1, 2, 1, 4, 1, 2, 1, 1, 2, 1, 2, 1]
    out = []
    for p in range(n):
        n_obj = pattern[p % len(pattern)]
        objs, pos = [], []
        tok = 2
        for o in range(n_obj):
            x0 = 0.05 + 0.2 * o
            boxes = []
            for f in range(frames):
                if (f + 3 * o + p) % 16 == 15:
                    boxes.append((0.0, 0.0, 0.0, 0.0))  # absent in this frame
                else:
                    dx = 0.4 * f / max(frames - 1, 1) * (1 if o % 2 == 0 else -1)
                    boxes.append((min(max(x0 + dx, 0.0), 0.7), 0.15 + 0.15 * o, min(max(x0 + dx, 0.0), 0.7) + 0.28,
                                  0.5 + 0.12 * o))
            objs.append(boxes)
            n_tok = 2 if (o + p) % 3 == 0 else 1
            pos.append(list(range(tok, tok + n_tok)))
            tok += n_tok + 1
        out.append(dict(bboxes=objs, object_positions=pos))
    return out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--unguided", action="store_true", help="time the unguided CFG step only")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lvd_b200", choices=["lvd_b200", "reference"])
    ap.add_argument("--workload", default="lvd_zeroscope", choices=list(WORKLOADS),
                    help="BASELINE.json config: lvd_zeroscope is the headline (config 2)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=0,
                    help="frames of the CPU arm's bounded sample (0 = largest of 24/12/8/4/2 fitting --cpu-budget-s)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="wall-clock budget of the CPU arm")
    ap.add_argument("--no-prompt-batch", action="store_true",
                    help="skip the prompt-batch measurement (parallel.generate_sharded: scatter + videos + gather)")
    ap.add_argument("--batch-prompts-per-gpu", type=int, default=3)
    ap.add_argument("--prompts-per-call", type=int, default=1,
                    help="multi-prompt batching of the prompt batch: prompts per pipeline call (one CFG UNet call of "
                         "2 x this many videos per step, guidance in lock-step)")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel-family breakdown here")
    ap.add_argument("--no-video", action="store_true", help="alias of --no-prompt-batch")
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


def cpu_reference_step(sample_frames, pattern, warmup, budget_s=240.0):
    """The reference algorithm (oracle/ restatement, fp32, host threads) on a bounded sample of the workload: the step
    kinds listed in `pattern` (True = guided step, False = unguided), after `warmup` untimed steps, on the first
    `sample_frames` of the 24 frames with the full-size UNet.  sample_frames == 0: one untimed 2-frame probe step of each
    kind in `pattern` sizes the sample to the largest of {24, 12, 8, 4, 2} frames whose steps fit `budget_s` (cost is
    linear in the frame count: every layer is per-frame work or attention over the frame axis of fixed width).
    Returns dict(frames_per_s, sec_per_step, cores, sec_guided, sec_unguided, sample_frames, steps, warmup)."""
    import torch
    from oracle import guidance_ref, scheduler_ref
    from oracle import unet3d_ref as R
    from oracle import weights

    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    cfg = weights.make_cfg("zeroscope")
    sd = weights.random_state_dict(cfg, seed=0)
    sched = scheduler_ref.DPMSolverRef()

    def run(frames, kinds):
        sched.set_timesteps(50)
        g = torch.Generator().manual_seed(1)
        lat = torch.randn((1, 4, frames, LAT_H, LAT_W), generator=g)
        ehs = torch.randn((2, 77, 1024), generator=g)
        bboxes, positions = bear_layout(frames)
        times = []
        for it, guided in enumerate(kinds):
            t0 = time.perf_counter()
            if sched.step_index is not None and sched.step_index >= 49:
                sched.set_timesteps(50)
            tt = int(sched.timesteps[it % 50])
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
            times.append(time.perf_counter() - t0)
        return times

    all_kinds = [pattern[i % len(pattern)] for i in range(warmup)] + list(pattern)
    if sample_frames <= 0:
        probe_kinds = sorted(set(all_kinds), reverse=True)
        probe = dict(zip(probe_kinds, run(2, probe_kinds)))  # also the first-touch cost of the weights
        probe = dict(zip(probe_kinds, run(2, probe_kinds)))
        per_frame = sum(probe[k] for k in all_kinds) / 2.0
        sample_frames = 2
        try:
            import psutil
            free_gb = psutil.virtual_memory().available / 2 ** 30
        except Exception:
            free_gb = 64.0
        for f in (24, 12, 8, 4):
            # a guided step keeps ~2.6 GB of fp32 autograd state per frame
            if per_frame * f <= budget_s and (not any(all_kinds) or 2.6 * f + 12 <= free_gb):
                sample_frames = f
                break
    times = run(sample_frames, all_kinds)[warmup:]
    t_g = [t for t, k in zip(times, pattern) if k]
    t_u = [t for t, k in zip(times, pattern) if not k]
    sec = sum(times) / len(times)
    return dict(frames_per_s=sample_frames / sec, sec_per_step=sec, cores=cores,
                sec_guided=sum(t_g) / len(t_g) if t_g else None, sec_unguided=sum(t_u) / len(t_u) if t_u else None,
                sample_frames=sample_frames, steps=len(pattern), warmup=warmup)


def cpu_sample_text(r):
    full = "the full 24-frame configuration" if r["sample_frames"] == 24 else \
        f"the first {r['sample_frames']} of the 24 frames"
    return (f"same step mix on {full} (72x40 latent, full 1.41B-param UNet), fp32, oracle port of the reference incl. "
            f"autograd guidance, {r['cores']} threads; guided step {r['sec_guided'] and round(r['sec_guided'], 1)} s, "
            f"unguided step {r['sec_unguided'] and round(r['sec_unguided'], 1)} s; {r['steps']} timed + {r['warmup']} "
            f"warm-up steps")


# BASELINE.json configs 2-4: algorithmic TFLOP per CFG forward / guidance iteration from section 8(d)
WORKLOADS = {
    "lvd_zeroscope": dict(model="Zeroscope", frames=24, h=40, w=72, gated=False, guided=True, cfg_tflop=42.79,
                          cfg_tflop_fusers=None, guid_tflop=32.59),
    "lvd_gligen_zeroscope": dict(model="Zeroscope + GLIGEN", frames=24, h=40, w=72, gated=True, guided=False,
                                 cfg_tflop=42.79, cfg_tflop_fusers=52.885, guid_tflop=0.0),
    "lvd_gligen_zeroscope_guided": dict(model="Zeroscope + GLIGEN", frames=24, h=40, w=72, gated=True, guided=True,
                                        cfg_tflop=42.79, cfg_tflop_fusers=52.885, guid_tflop=32.59),
    "lvd_plus_modelscope256": dict(model="ModelScope-256 + GLIGEN", frames=16, h=32, w=32, gated=True, guided=True,
                                   cfg_tflop=9.77, cfg_tflop_fusers=11.803, guid_tflop=7.48),
}


def _workload(name="lvd_zeroscope", unguided=False):
    wl = WORKLOADS[name]
    shape = f"{wl['frames']} frames {wl['w'] * 8}x{wl['h'] * 8} ({wl['w']}x{wl['h']} latent)"
    parts = [f"{name} 1 prompt/GPU, {shape}, 50-step schedule pattern:"]
    if wl["guided"] and not unguided:
        parts.append("step i runs one backward-guidance iteration (6 guided cross-attention keys, demo 'bear' layout) "
                     "iff i % 5 == 0 (= 10 guided of 50);")
    if wl["gated"]:
        parts.append("GLIGEN fusers enabled iff i % 5 < 2 (= 20 of 50, beta 0.4);")
    parts.append("then the CFG UNet forward (B=2) + CFG combine + DPM-Solver++ update")
    if unguided and not wl["gated"]:
        return f"{name} 1 prompt/GPU, {shape}, unguided CFG denoising step"
    return " ".join(parts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    pattern = [(not args.unguided) and i % GUIDED_EVERY == 0 for i in range(steps)]
    r = cpu_reference_step(args.cpu_sample_frames, pattern, max(0, args.warmup), args.cpu_budget_s)
    sample = cpu_sample_text(r)
    fps = r["frames_per_s"]
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(0, args.warmup), "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload("lvd_zeroscope", args.unguided), "sample": sample,
                   "sample_frames": r["sample_frames"]},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
                         "sample_frames": r["sample_frames"], "sec_guided": r["sec_guided"],
                         "sec_unguided": r["sec_unguided"]},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def gligen_inputs(frames, bboxes, dev, gen):
    """30-slot padded grounding inputs as the pipeline builds them (controllable_pipeline...py:738-814): CFG batch of 2,
    the unconditional half masked out.  Phrase embeddings are synthetic (the text encoder is outside the hot path)."""
    import torch
    n_obj = len(bboxes)
    boxes = torch.zeros((2, frames, 30, 4))
    emb = torch.zeros((2, frames, 30, 1024))
    masks = torch.zeros((2, frames, 30))
    phrase = torch.randn((n_obj, 1024), generator=gen)
    for o in range(n_obj):
        for f in range(frames):
            boxes[:, f, o] = torch.tensor(bboxes[o][f])
        emb[:, :, o] = phrase[o]
        masks[1, :, o] = 1
    return {"boxes": boxes.flatten(0, 1).half().to(dev), "positive_embeddings": emb.flatten(0, 1).half().to(dev),
            "masks": masks.flatten(0, 1).half().to(dev)}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    ge.build()
    from lvd_b200 import _lib, ops, synthetic
    from lvd_b200.models.pipelines import latent_backward_guidance
    from lvd_b200.models.unet_3d_condition import UNet3DConditionModel
    from lvd_b200.scheduler import DPMSolverMultistepScheduler

    # stdout carries exactly ONE line (the JSON): anything else written to fd 1 by libraries (NCCL's version banner
    # ignores NCCL_DEBUG_FILE on some builds) is sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    wl = WORKLOADS[args.workload]
    frames, lat_h, lat_w = wl["frames"], wl["h"], wl["w"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl lvd_b200) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    # the reference's own model class (mirror) with seeded random weights of the Zeroscope / ModelScope architecture
    unet = UNet3DConditionModel(sample_size=32, attention_type="gated" if wl["gated"] else "default")
    unet.to(dev)
    unet.load_state_dict(synthetic.random_state_dict(unet._cfg, seed=0, dtype=torch.float16))
    eng = unet._engine

    # one independent work item (prompt) per rank: NCCL only carries the seed scatter and the timing reduction
    seeds = torch.arange(world, dtype=torch.int64, device=dev) + 1234
    if world > 1:
        dist.broadcast(seeds, src=0)
    g = torch.Generator().manual_seed(int(seeds[rank].item()))
    lat_host = torch.randn((1, 4, frames, lat_h, lat_w), generator=g).half().pin_memory()
    ehs_host = torch.randn((2, 77, 1024), generator=g).half().pin_memory()
    out_host = torch.empty((1, 4, frames, lat_h, lat_w), dtype=torch.float16).pin_memory()

    sched = DPMSolverMultistepScheduler()
    sched.set_timesteps(50)
    prompt_embeds = ehs_host.to(dev)
    latents = lat_host.to(dev)
    bboxes, positions = bear_layout(frames)
    guided_every = GUIDED_EVERY if (wl["guided"] and not args.unguided) else 0
    gl_in = gligen_inputs(frames, bboxes, dev, g) if wl["gated"] else None
    guidance_kw = dict(bboxes=bboxes, object_positions=positions, guidance_attn_keys=GUIDANCE_KEYS,
                       loss_scale=GUIDANCE_LOSS_SCALE, loss_threshold=0.0, max_iter=1, max_index_step=10,
                       base_attn_dim=(lat_h, lat_w), **GUIDANCE_HP)
    import warnings
    warnings.filterwarnings("ignore", message="Using guidance scaled")

    def fusers_on(i):  # GLIGEN scheduled sampling beta = 0.4 (README.md:83): fusers enabled on 2 of every 5 steps
        return wl["gated"] and (i % 5) < 2

    def set_fusers(on):
        for m in unet.modules():
            if type(m).__name__ == "GatedSelfAttentionDense":
                m.enabled = on

    def denoise_step(i):
        """One iteration of the reference's loop body (controllable_pipeline...py:836-958) through the public mirrors:
        latent_backward_guidance -> UNet3DConditionModel.forward (CFG batch) -> fused CFG combine + DPM-Solver++ step."""
        nonlocal latents
        t = sched.timesteps[i % 50]
        if wl["gated"]:
            set_fusers(fusers_on(i))
        if guided_every and i % guided_every == 0:
            latents, _ = latent_backward_guidance(sched, unet, prompt_embeds[1:2], index=0, t=t, latents=latents,
                                                  loss=torch.tensor(10000.0), **guidance_kw)
        kw = {"gligen": gl_in} if wl["gated"] else {}
        eps = unet(torch.cat([latents] * 2), t, encoder_hidden_states=prompt_embeds, cross_attention_kwargs=kw,
                   return_dict=False)[0]
        if sched._step_index >= 49:
            sched._reset()
        latents = sched.step_cfg(eps[0:1], eps[1:2], latents, GUIDANCE_SCALE)

    unet.use_cuda_graph = not args.no_graph
    # SETUP (untimed, not warm-up): two passes over the 5-step pattern -- every step variant (guided / unguided, fusers
    # on / off) runs once eagerly and once under CUDA-graph capture, so that warm-up and timed steps are pure replays
    SETUP_STEPS = 2 * GUIDED_EVERY
    lib = _lib.lib()
    c0 = lib.lvd_launch_count()
    counts = []
    for i in range(SETUP_STEPS):
        denoise_step(i)
        if i < GUIDED_EVERY:  # the eager pass: launches per step variant, counted by the library itself
            torch.cuda.synchronize()
            c1 = lib.lvd_launch_count()
            counts.append(c1 - c0)
            c0 = c1
    torch.cuda.synchronize()
    for i in range(args.warmup):  # exactly --warmup untimed warm-up steps
        denoise_step(i)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- timed region 1: device-resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        denoise_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---------------- timed region 2: end to end with host buffers (same public calls; H2D in, D2H out, every step)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_prof = None
    if os.environ.get("LVD_BENCH_HOST_PROFILE"):  # tuning aid: where the host time of an end-to-end step goes
        import cProfile
        host_prof = cProfile.Profile()
        host_prof.enable()
    f0.record()
    for i in range(args.steps):
        latents = lat_host.to(dev, non_blocking=True)
        prompt_embeds = ehs_host.to(dev, non_blocking=True)
        denoise_step(i)
        out_host.copy_(latents, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    f1.record()
    if host_prof is not None:
        import pstats
        host_prof.disable()
        pstats.Stats(host_prof, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    # ---------------- region 2b (informational, NOT the e2e value): the same copies every step, but the host waits for
    # the device-to-host copy of step i only after it has enqueued step i+1 (two pinned output buffers, one event each)
    # -- what the per-step stream synchronisation of region 2 costs
    barrier()
    out_host2 = [out_host, torch.empty_like(out_host).pin_memory()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(args.steps):
        latents = lat_host.to(dev, non_blocking=True)
        prompt_embeds = ehs_host.to(dev, non_blocking=True)
        denoise_step(i)
        out_host2[i & 1].copy_(latents, non_blocking=True)
        done[i & 1].record()
        if i > 0:
            done[(i - 1) & 1].synchronize()
    done[(args.steps - 1) & 1].synchronize()
    p1.record()
    barrier()
    ms_e2e_pipelined = p0.elapsed_time(p1)
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    sampler.stop()
    clocks = sampler.summary()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---------------- separate timings of the step kinds (device events, resident inputs)
    def time_steps(idx):
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in idx:
            denoise_step(i)
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / len(idx)

    ms_unguided = time_steps([2, 3, 4])          # fusers off, no guidance
    ms_fused = time_steps([1, 6]) if wl["gated"] else None   # fusers on, no guidance
    ms_guided = time_steps([0, 5]) if guided_every else None  # guidance (+ fusers when gated)

    # ---------------- the prompt batch (north_star / SURVEY.md section 8e, BASELINE configs 4-5): a fixed batch of
    # `--batch-prompts-per-gpu` x N prompts with layouts of 1-4 objects driven through parallel.generate_sharded -- rank 0
    # scatters the work table (seeds) and the packed box / token tables (NCCL broadcast), every rank runs its window
    # through TextToVideoSDPipeline.__call__ (50 steps, 10 guided when the workload is guided, GLIGEN beta 0.4 when
    # gated, VAE decode), the uint8 frames are gathered on rank 0.  Scatter and gather are INSIDE the timed region;
    # graph capture is excluded by warm-up items (one per padded slot count of the batch).
    video = None
    prompt_batch = None
    if not (args.no_prompt_batch or args.no_video):
        from lvd_b200 import parallel
        from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline
        from lvd_b200.vae import AutoencoderKL
        vae = AutoencoderKL().to(dev)
        vae.load_state_dict(synthetic.random_vae_state_dict(synthetic.VAE_SD, seed=0, dtype=torch.float16))
        vae.enable_slicing(8)
        pipe = TextToVideoSDPipeline(unet=unet, vae=vae, scheduler=DPMSolverMultistepScheduler()).to(dev)
        call = dict(prompt_embeds=ehs_host[1:2], negative_prompt_embeds=ehs_host[0:1], height=lat_h * 8,
                    width=lat_w * 8, num_frames=frames, num_inference_steps=50, guidance_scale=GUIDANCE_SCALE,
                    output_type="np", gligen_scheduled_sampling_beta=0.4)

        def generate(seed, bboxes_, positions_, steps=50):
            """One pipeline call: a single prompt (scalars) or a group of prompts (lists, multi-prompt batching)."""
            many = isinstance(seed, (list, tuple))
            seeds = list(seed) if many else [seed]
            lays = list(zip(bboxes_, positions_)) if many else [(bboxes_, positions_)]
            nb = len(seeds)
            lat0 = torch.cat([torch.randn((1, 4, frames, lat_h, lat_w),
                                          generator=torch.Generator().manual_seed(int(sd_))).half() for sd_ in seeds])
            kw_ = dict(call, num_inference_steps=steps, latents=lat0,
                       prompt_embeds=ehs_host[1:2].expand(nb, -1, -1).contiguous(),
                       negative_prompt_embeds=ehs_host[0:1].expand(nb, -1, -1).contiguous())
            if guided_every:
                kw_["backward_guidance_kwargs"] = dict(
                    guidance_kw, bboxes=[l[0] for l in lays] if many else lays[0][0],
                    object_positions=[l[1] for l in lays] if many else lays[0][1])
            if wl["gated"]:
                gen = torch.Generator().manual_seed(int(seeds[0]))
                per = [gligen_inputs(frames, l[0], dev, gen) for l in lays]  # each: rows [uncond F | cond F]
                kw_["gligen_inputs"] = {k: torch.cat([p_[k][:frames] for p_ in per] + [p_[k][frames:] for p_ in per])
                                        for k in per[0]}
            fr = pipe(**kw_).frames  # np.float32 [nb, F, H, W, 3] in [0, 1]
            u8 = torch.from_numpy((fr * 255.0 + 0.5).clip(0, 255).astype("uint8"))
            return u8 if many else u8[0]

        n_items = args.batch_prompts_per_gpu * world
        layouts = prompt_batch_layouts(n_items, frames)
        from lvd_b200.utils.guidance import padded_slot_count
        slots_of = lambda l: padded_slot_count(sum(len(t) for t in l["object_positions"]))  # noqa: E731
        ppc = max(1, args.prompts_per_call)
        if ppc == 1:
            for n_pad in sorted({slots_of(l) for l in layouts}):  # warm-up items: one short run per slot-table size
                l = next(l for l in layouts if slots_of(l) == n_pad)
                generate(0, l["bboxes"], l["object_positions"], steps=6)
        else:  # warm-up: this rank's own groups, shortened (same group sizes / padded slot counts as the timed run)
            lo_, hi_ = parallel.shard_items(n_items, rank, world)
            for k in range(lo_, hi_, ppc):
                grp = layouts[k:min(k + ppc, hi_)]
                generate([0] * len(grp), [l["bboxes"] for l in grp], [l["object_positions"] for l in grp], steps=6)
        if world > 1:
            # one-time NCCL connection setup of the collectives the timed region uses (broadcast + point-to-point gather)
            # is excluded like graph capture: a tiny gather / broadcast first
            parallel.gather_results(torch.zeros((1, 16), dtype=torch.uint8, device=dev), world)
            parallel.broadcast_work_table(torch.zeros((1, 3), dtype=torch.int64), dev)
        barrier()
        stats = {}
        t0 = time.perf_counter()
        frames_u8, table = parallel.generate_sharded(generate, layouts if rank == 0 else None, dev, stats=stats,
                                                     batch=ppc)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        tv = torch.tensor([sec, stats["generate_s"], stats["gather_s"], stats["scatter_s"]], dtype=torch.float64,
                          device=dev)
        tmin = tv.clone()
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        sec, gen_max, gather_max, scatter_max = (float(v) for v in tv)
        gen_min = float(tmin[1])
        if rank == 0:
            ok = frames_u8 is not None and tuple(frames_u8.shape) == (n_items, frames, lat_h * 8, lat_w * 8, 3)
            prompt_batch = {
                "items": n_items, "items_per_gpu": args.batch_prompts_per_gpu, "prompts_per_call": ppc, "seconds": sec,
                "videos_per_s": n_items / sec, "frames_per_s": n_items * frames / sec,
                "generate_s_max_rank": gen_max, "generate_s_min_rank": gen_min,
                "imbalance": (gen_max - gen_min) / gen_max if gen_max > 0 else 0.0,
                "scatter_s": scatter_max, "gather_s": gather_max, "gather_share": gather_max / sec,
                "gathered_bytes": int(frames_u8.numel()) if frames_u8 is not None else 0, "frames_ok": bool(ok),
                "objects_per_layout": [len(l["bboxes"]) for l in layouts],
                "path": "parallel.generate_sharded: work table + box/token tables broadcast from rank 0, "
                        "TextToVideoSDPipeline.__call__ per item (50 steps"
                        + (", 10 guided x 1 iteration" if guided_every else "")
                        + (", GLIGEN fusers on the first 20" if wl["gated"] else "")
                        + ", native VAE decode), uint8 frames gathered on rank 0; wall clock, scatter + gather inside"}
            video = {"seconds_per_video": gen_max / args.batch_prompts_per_gpu,
                     "videos_per_s": n_items / sec, "frames_per_s": n_items * frames / sec,
                     "source": "prompt_batch (generate phase of the slowest rank / its items)"}
        del pipe, vae, frames_u8

    # ---------------- instrumented replay region: the same step sequence replayed from graphs that were captured with an
    # `external` CUDA-event pair around every kernel node (ops._Timed); after each step the pairs of the graphs that
    # replayed are read.  Kernel durations therefore come from replay conditions (no launch gaps, the clocks / power
    # state of the steady step loop), K steps long; the event nodes themselves cost ~1.5 us each, which is why this
    # region is separate from the headline one (its own ms/step is reported next to it).
    set_fusers(False)
    gg = getattr(eng, "_guidance_graph", None)
    fam, fam_guid, shapes = {}, {}, ({}, {})
    region = None
    ms_instr = None

    def acc(dst, name, fl, nb, ms_, tag, shp):
        d = dst.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += ms_
        d["flops"] += fl
        d["bytes"] += nb
        d["launches"] += 1
        if tag is not None:
            sd = shp.setdefault(tag, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            sd["ms"] += ms_
            sd["flops"] += fl
            sd["bytes"] += nb
            sd["launches"] += 1

    n_unet_steps = n_guid_steps = 0
    if unet.use_cuda_graph:
        try:
            ops.PROF, ops.PROF_SCOPE = [], "bench"
            unet._graphs.clear()
            if gg is not None:
                gg._graphs.clear()
            for i in range(SETUP_STEPS):
                denoise_step(i)
            torch.cuda.synchronize()
            recs = [r for r in ops.PROF if r[6] is not None]
            ops.PROF, ops.PROF_SCOPE = None, None
            by_scope = {}
            for r in recs:
                by_scope.setdefault(r[6], []).append(r)
            for i in range(min(args.warmup, 3)):
                denoise_step(i)
            torch.cuda.synchronize()
            t_instr = 0.0
            for i in range(args.steps):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                denoise_step(i)
                b_.record()
                torch.cuda.synchronize()
                t_instr += a.elapsed_time(b_)
                guided = bool(guided_every and i % guided_every == 0)
                for r in by_scope.get("unet:fusers" if fusers_on(i) else "unet", []):
                    acc(fam, r[0], r[1], r[2], r[3].elapsed_time(r[4]), r[5], shapes[0])
                n_unet_steps += 1
                if guided:
                    for r in by_scope.get("guidance", []):
                        acc(fam_guid, r[0], r[1], r[2], r[3].elapsed_time(r[4]), r[5], shapes[1])
                    n_guid_steps += 1
            ms_instr = t_instr / args.steps
            region = (f"{args.steps} instrumented graph replays of the timed step sequence (external CUDA-event nodes "
                      f"around every kernel node of the replayed graphs), {ms_instr:.2f} ms/step with the event nodes")
        except Exception as e:  # external-event timing unavailable: fall back to the eager instrumented pass below
            sys.stderr.write(f"[bench] graph-internal event timing failed ({type(e).__name__}: {e}); eager pass instead\n")
            fam, fam_guid, shapes = {}, {}, ({}, {})
            n_unet_steps = n_guid_steps = 0
            region = None
        finally:
            ops.PROF, ops.PROF_SCOPE = None, None
    if region is None:
        x2 = torch.cat([latents] * 2)
        ops.PROF = []
        eng.forward(x2, float(sched.timesteps[0]), prompt_embeds)
        torch.cuda.synchronize()
        n_fwd_records = len(ops.PROF)
        if guided_every:
            latent_backward_guidance(sched, unet, prompt_embeds[1:2], index=0, t=sched.timesteps[0], latents=latents,
                                     loss=torch.tensor(10000.0), **guidance_kw)
            torch.cuda.synchronize()
        prof = ops.PROF
        ops.PROF = None
        for k_, r in enumerate(prof):
            first = k_ < n_fwd_records
            acc(fam if first else fam_guid, r[0], r[1], r[2], r[3].elapsed_time(r[4]), r[5], shapes[0 if first else 1])
        n_unet_steps, n_guid_steps = 1, (1 if guided_every else 0)
        region = "one eager CFG forward + one eager guidance iteration with a CUDA-event pair around every launch"
    for dst, n in ((fam, n_unet_steps), (fam_guid, n_guid_steps), (shapes[0], n_unet_steps), (shapes[1], n_guid_steps)):
        for v in dst.values():  # per CFG forward / per guidance iteration
            for k_ in ("ms", "flops", "bytes", "launches"):
                v[k_] = v[k_] / max(1, n)
    # kernels of this library launched inside timed region 1 (graph replays re-run the captured launches): per-variant
    # launch counts were read from the library's own counter during the eager setup pass
    launches = sum(counts[i % GUIDED_EVERY] for i in range(args.steps))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained"
    gm = dict(fam.get("gemm_tc_kernel", {"ms": 1.0, "flops": 0.0, "launches": 1, "bytes": 0.0}))
    gg_ = fam_guid.get("gemm_tc_kernel")
    if gg_:  # the timed step mix = CFG forward + 1/5 guidance iteration: the dominant kernel's launches of both
        share = 1.0 / GUIDED_EVERY
        gm = {"ms": gm["ms"] + share * gg_["ms"], "flops": gm["flops"] + share * gg_["flops"],
              "launches": gm["launches"] + share * gg_["launches"], "bytes": gm.get("bytes", 0.0) + share * gg_["bytes"]}
    achieved = gm["flops"] / (gm["ms"] * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:  # DRAM bytes per gemm_tc_kernel launch from the committed ncu capture of THIS round's code (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")))
        if args.workload == "lvd_zeroscope":
            traffic = tj["traffic_bytes_per_launch"]
            traffic_src = "profiles/r2_gemm_traffic.json (ncu launch list with dram__bytes_read.sum + " \
                          "dram__bytes_write.sum of every gemm_tc_kernel launch of one CFG forward + one guidance " \
                          "iteration, step-mix weighted; tools/gpu_ncu.sh)"
    except Exception:
        pass
    all_ms = sum(v["ms"] for v in fam.values()) + (sum(v["ms"] for v in fam_guid.values()) / GUIDED_EVERY
                                                     if fam_guid else 0.0)
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": gm.get("bytes", 0.0) / max(1, gm["launches"]),
                "peak_source": peak_src, "launches_per_step": gm["launches"], "ms_per_step": gm["ms"],
                "share_of_step_kernel_time": gm["ms"] / max(1e-9, all_ms), "region": region,
                "instrumented_ms_per_step": ms_instr,
                "note": "FLOPs and durations summed over all gemm_tc_kernel launches of one CFG forward (fusers per the "
                        "step pattern) plus 1/5 of one guidance iteration when the workload is guided"}
    if args.profile_json and rank == 0:
        json.dump({"cfg_forward": fam, "guidance_iteration": fam_guid, "gemm_shapes_cfg_forward": shapes[0],
                   "gemm_shapes_guidance_iteration": shapes[1]}, open(args.profile_json, "w"), indent=1)

    if rank == 0:
        ms_step = ms / args.steps
        value = world * frames / (ms_step * 1e-3)
        e2e_val = world * frames / (ms_e2e / args.steps * 1e-3)
        n_fused = sum(1 for i in range(5) if fusers_on(i)) / 5.0
        tflop_step = (wl["cfg_tflop"] * (1 - n_fused) + (wl["cfg_tflop_fusers"] or 0.0) * n_fused +
                      (wl["guid_tflop"] / GUIDED_EVERY if guided_every else 0.0))
        fps = lambda m: None if m is None else {"ms": m, "frames_per_s": world * frames / (m * 1e-3)}  # noqa: E731
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": _workload(args.workload, args.unguided),
                       "weights": f"seeded random init of the {wl['model']} UNet3D config "
                                  f"({'1623.8' if wl['gated'] else '1411.2'} M params), fp16",
                       "parallelism": f"replicas x{world} (independent prompts, no data-path collective)",
                       "l2": "inputs larger than L2 (2.8 GB weights + >20 GB activations per step vs 126 MB)",
                       "api": "lvd_b200.models: latent_backward_guidance + UNet3DConditionModel.forward + "
                              "DPMSolverMultistepScheduler.step_cfg (the reference pipeline's loop body)",
                       "cuda_graph": "UNet3DConditionModel.forward and the guidance pass (GuidanceGraph.run) each replay a "
                                     "CUDA graph per structure" if unet.use_cuda_graph else False,
                       "setup": f"{SETUP_STEPS} untimed setup steps before the warm-up: every step variant runs once "
                                "eagerly and once under graph capture (not counted as warm-up)",
                       "algorithmic_tflop_per_step": tflop_step,
                       "guidance": None if not guided_every else {
                           "keys": GUIDANCE_KEYS, "layout": "demo bear, 1 object, token 2", "iterations_per_guided_step": 1,
                           "loss_scale": GUIDANCE_LOSS_SCALE, **GUIDANCE_HP},
                       "gligen": None if not wl["gated"] else "30-slot grounding tokens, fusers enabled on 2 of 5 steps "
                                                              "(scheduled sampling beta 0.4)"},
            "unguided_step": fps(ms_unguided), "fuser_step": fps(ms_fused), "guided_step": fps(ms_guided),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": lat_host.numel() * 2 + ehs_host.numel() * 2,
                    "d2h_bytes_per_step": out_host.numel() * 2 + (4 if guided_every else 0),
                    "sync": "the host synchronises the stream after the device-to-host copy of EVERY step",
                    "pipelined_value_informational": world * frames / (ms_e2e_pipelined / args.steps * 1e-3),
                    "pipelined_note": "same copies every step, but the host waits for step i's copy only after it has "
                                      "enqueued step i+1 (region 2b); measured equal to `value` of this block: the "
                                      "1-3 % between the resident region and the end-to-end regions is neither the "
                                      "copies nor the synchronisation (DESIGN.md section 6)"},
            "gpu_launches": int(launches),
            "video": video,
            "prompt_batch": prompt_batch,
            "clocks": clocks,
            "roofline": roofline,
            "step_tflops": tflop_step / (ms_step * 1e-3),
            "kernel_families_ms": {k: round(v["ms"], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
            "guidance_families_ms": {k: round(v["ms"], 3) for k, v in sorted(fam_guid.items(),
                                                                            key=lambda kv: -kv[1]["ms"])},
        }
        if world == 1 and not args.no_cpu_baseline and args.workload == "lvd_zeroscope":
            # the SAME routine as --impl reference (one code path), run as a child process so that a host-memory problem
            # of the fp32 autograd step can never take the GPU line down with it: one guided + one unguided step,
            # combined with the schedule weights 1:4 of the timed step mix
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2" if guided_every else "1",
                   "--warmup", "0", "--cpu-budget-s", str(args.cpu_budget_s / 2),
                   "--cpu-sample-frames", str(args.cpu_sample_frames)] + (["--unguided"] if not guided_every else [])
            try:
                env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=args.cpu_budget_s * 2 + 300, env=env)
                cb = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
                sec = (cb["sec_guided"] + (GUIDED_EVERY - 1) * cb["sec_unguided"]) / GUIDED_EVERY if guided_every \
                    else cb["sec_unguided"]
                line["cpu_baseline"] = {"value": cb["sample_frames"] / sec, "unit": UNIT, "cores": cb["cores"],
                                        "kind": "port",
                                        "sample": cb["sample"] + (", weighted 1:4" if guided_every else "")}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                        "sample": f"CPU leg failed: {type(e).__name__}: {e}"}
        print(json.dumps(line), file=json_out, flush=True)
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
