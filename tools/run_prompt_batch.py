"""A batch of prompts sharded over the GPUs of one box (BASELINE.json config 4 pattern):
  torchrun --nproc-per-node N tools/run_prompt_batch.py [--prompts 8] [--steps 50] [--workload lvd_plus_modelscope256]
Rank 0 scatters seeds / box layouts (NCCL broadcast), every rank runs its window through TextToVideoSDPipeline.__call__
(guidance + GLIGEN + VAE decode), the uint8 frames are gathered on rank 0.  Synthetic weights, given prompt embeddings."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--prompts", type=int, default=8)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--workload", default="lvd_plus_modelscope256", choices=list(bench.WORKLOADS))
args = ap.parse_args()
ge.build()
from lvd_b200 import parallel, synthetic  # noqa: E402
from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline  # noqa: E402
from lvd_b200.models.unet_3d_condition import UNet3DConditionModel  # noqa: E402
from lvd_b200.scheduler import DPMSolverMultistepScheduler  # noqa: E402
from lvd_b200.vae import AutoencoderKL  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS[args.workload]
F, h, w = wl["frames"], wl["h"], wl["w"]
unet = UNet3DConditionModel(sample_size=32, attention_type="gated" if wl["gated"] else "default").to(dev)
unet.load_state_dict(synthetic.random_state_dict(unet._cfg, seed=0, dtype=torch.float16))
vae = AutoencoderKL().to(dev)
vae.load_state_dict(synthetic.random_vae_state_dict(synthetic.VAE_SD, seed=0, dtype=torch.float16))
vae.enable_slicing(8)
pipe = TextToVideoSDPipeline(unet=unet, vae=vae, scheduler=DPMSolverMultistepScheduler()).to(dev)
emb = torch.randn((2, 77, 1024), generator=torch.Generator().manual_seed(7)).half()

layouts = None
if rank == 0:  # 1-2 objects per prompt sweeping across the frame (stand-ins for the cached GPT-4 layouts)
    layouts = []
    for p in range(args.prompts):
        objs = [[(0.6 * f / (F - 1) + 0.02 * p, 0.3, 0.6 * f / (F - 1) + 0.3, 0.7) for f in range(F)]]
        pos = [[2]]
        if p % 2:
            objs.append([(0.1, 0.05 * (p % 5), 0.4, 0.5) for _ in range(F)])
            pos.append([5, 6])
        layouts.append(dict(bboxes=objs, object_positions=pos))


def generate(seed, bboxes, positions):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn((1, 4, F, h, w), generator=g)
    kw = dict(prompt_embeds=emb[1:], negative_prompt_embeds=emb[:1], height=h * 8, width=w * 8, num_frames=F,
              num_inference_steps=args.steps, guidance_scale=bench.GUIDANCE_SCALE, latents=lat, output_type="np",
              gligen_scheduled_sampling_beta=0.4)
    if wl["guided"]:
        kw["backward_guidance_kwargs"] = dict(bboxes=bboxes, object_positions=positions,
                                              guidance_attn_keys=bench.GUIDANCE_KEYS,
                                              loss_scale=bench.GUIDANCE_LOSS_SCALE, loss_threshold=0.0, max_iter=1,
                                              max_index_step=10, base_attn_dim=(h, w), **bench.GUIDANCE_HP)
    if wl["gated"]:
        kw["gligen_inputs"] = bench.gligen_inputs(F, bboxes, dev, g)
    frames = pipe(**kw).frames[0]
    return torch.from_numpy((frames * 255.0 + 0.5).astype("uint8"))


if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
frames, table = parallel.generate_sharded(generate, layouts, dev)
torch.cuda.synchronize()
sec = parallel.max_over_ranks(time.perf_counter() - t0, dev)
if rank == 0:
    print(json.dumps({"workload": args.workload, "n_gpus": world, "videos": int(frames.shape[0]),
                      "frames_shape": list(frames.shape), "seconds": sec, "videos_per_s": frames.shape[0] / sec,
                      "steps": args.steps, "seeds": table[:, 2].tolist(),
                      "frame_means": [round(float(v), 3) for v in frames.float().mean(dim=(1, 2, 3, 4)).tolist()]}),
          flush=True)
if world > 1:
    dist.destroy_process_group()
