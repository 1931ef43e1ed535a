"""Measured whole-clip time through the public pipeline call (reduced to 2 chained clips):
SVIVideoPipeline.__call__ on an 81-frame 480x832 clip, 50 CFG steps, with every stage the reference runs per clip —
umT5 prompt encoding (positive + negative), CLIP image encoding, VAE encode of the conditioning video, 100 DiT forwards,
VAE decode, uint8 frame conversion — and the clip loop of test_svi.py:424-485 (last frames recycled as the next clip's
conditioning).  Random-init weights: a 1.3B-width image-to-video DiT (d 1536, 30 layers, in_dim 36, CLIP branch), the
full-size umT5-XXL and CLIP ViT-H encoders, the Wan VAE.

    python tools/clip_loop_bench.py [--clips 2] [--steps 50] [--motion-frames 5]

Prints one JSON object with per-clip wall seconds (time.perf_counter around pipe(...), synchronised) and clips/hour."""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-video-infinity_b200")):
    sys.path.insert(0, p)
from tools import synth, synth_enc, synth_vae  # noqa: E402
from tools.enc_bench import device_init  # noqa: E402


class TokenPrompter:
    """WanPrompter front end without a tokenizer directory (none in this image): seeded token ids of a fixed prompt
    length go through the real text encoder + zero fill, exactly like WanPrompter.encode_prompt after tokenisation."""

    def __init__(self, prompter, vocab):
        self.prompter, self.vocab = prompter, vocab

    def __call__(self, prompt, positive=True):
        import zlib
        g = torch.Generator().manual_seed(zlib.crc32(prompt.encode()) % 100000 + (0 if positive else 1))   # same ids on every rank
        n = 60 if positive else 90
        ids = torch.zeros(1, 512, dtype=torch.int64)
        ids[0, :n] = torch.randint(2, self.vocab, (n,), generator=g)
        mask = (ids > 0).long()
        return self.prompter.encode_ids(ids, mask, device="cuda").to(torch.float32)


def main_from_bench(bargs):
    """bench.py --workload cfg4: the chained-clip loop (BASELINE configs[3]) with bench's --clips; 14B-I2V when launched on
    several GPUs (torchrun), the 1.3B-width I2V model on one."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return main(["--clips", str(bargs.clips)] + (["--model", "14b"] if world > 1 else []))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--motion-frames", type=int, default=5)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--model", default="1.3b-i2v", choices=["1.3b-i2v", "14b"])
    a = ap.parse_args(argv)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from diffsynth import ModelManager, SVIVideoPipeline
    from diffsynth.models.wan_video_dit import WanModel, precompute_freqs_cis_3d
    from diffsynth.models.wan_video_image_encoder import WanImageEncoder
    from diffsynth.models.wan_video_text_encoder import WanTextEncoder
    from diffsynth.models.wan_video_vae import WanVideoVAE
    from diffsynth.prompters import WanPrompter
    dev = f"cuda:{local}"
    cfg = dict(synth.CFG_T2V_1_3B, has_image_input=True, in_dim=36) if a.model == "1.3b-i2v" else dict(synth.CFG_I2V_14B)
    with torch.device("meta"):
        dit = WanModel(**cfg)
    dit.load_state_dict(synth.make_dit_state_dict_fast(cfg, seed=0, device=dev, dtype=torch.bfloat16), assign=True)
    dit.freqs = precompute_freqs_cis_3d(128)
    dit.eval()
    vae = WanVideoVAE().eval()
    vae.load_state_dict(synth_vae.make_vae_state_dict(seed=0))
    vae.to(dev)
    with torch.device("meta"):
        te = WanTextEncoder(**synth_enc.TEXT_UMT5_XXL).to(torch.bfloat16)
        ie = WanImageEncoder(**synth_enc.CLIP_VIT_H).to(torch.bfloat16)
    te, ie = device_init(te, 0), device_init(ie, 1)
    mm = ModelManager(torch_dtype=torch.bfloat16, device=dev)
    mm.add_model("wan_video_dit", dit)
    mm.add_model("wan_video_vae", vae)
    mm.add_model("wan_video_image_encoder", ie)
    pipe = SVIVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device=dev, is_test=True, use_usp=world > 1)
    wp = WanPrompter()
    wp.fetch_models(te)
    pipe.text_encoder = te
    pipe.prompter = TokenPrompter(wp, synth_enc.TEXT_UMT5_XXL["vocab"])
    img = Image.fromarray(np.random.default_rng(5).integers(0, 255, size=(a.height, a.width, 3), dtype=np.uint8))
    ref = torch.from_numpy(np.array(img))
    args = types.SimpleNamespace(ref_pad_cfg=False, ref_pad_num=-1, sequential_cfg="none")
    cond, video_list, secs = img, [], []
    for k in range(a.clips):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        video = pipe(prompt=f"prompt {k}", negative_prompt="negative", input_image=cond, num_inference_steps=a.steps,
                     cfg_scale={"text": 5.0}, seed=42 * k, tiled=False, random_ref_frame=ref, height=a.height, width=a.width,
                     num_frames=81, args=args, progress_bar_cmd=lambda x: x)
        torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
        cond = video[-a.motion_frames:]
        video_list += video[:-a.motion_frames] if k < a.clips - 1 else video
    steady = secs[1:] if len(secs) > 1 else secs
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([sum(steady) / len(steady)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steady_s = t.item()
        plan = pipe.sp_group().describe() + " + VAE row bands over all ranks"
    else:
        steady_s, plan = sum(steady) / len(steady), "single"
    if rank == 0:
        print(json.dumps({"metric": "SVI clips per hour (81f, 50 CFG steps, encoders + VAE + clip hand-off included)",
                          "value": 3600.0 / steady_s, "unit": "clips/h", "n_gpus": world, "higher_is_better": True,
                          "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": "cfg4", "description": f"{a.clips} chained SVI clips, 81f x {a.height}x{a.width}, "
                                     f"{a.steps} CFG steps, {a.motion_frames} recycled motion frames (BASELINE configs[3])",
                                     "parallelism": plan,
                                     "models": ("Wan2.1-I2V-14B" if a.model == "14b" else "1.3B-width I2V DiT") +
                                               " + umT5-XXL + CLIP ViT-H + Wan VAE, random init"},
                          "clip_seconds": secs, "steady_clip_s": steady_s, "clips_per_hour": 3600.0 / steady_s,
                          "frames_kept": len(video_list), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                          "api": "SVIVideoPipeline.__call__ (use_usp=True on N > 1)",
                          "note": "first clip includes one-time engine builds, graph capture and kernel attribute setup"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
