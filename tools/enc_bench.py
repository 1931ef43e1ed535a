"""Conditioning-encoder timing on one GPU: full-size umT5-XXL encoder (24 layers, d 4096, 512 tokens)
and CLIP ViT-H/14 visual tower (31 of 32 blocks, 257 tokens), random-init weights created on the device.

    python tools/enc_bench.py [--cpu-layers 1]

Prints one JSON object: per-call milliseconds (CUDA events, warm), algorithmic TFLOP/s, kernel launches, and the CPU
oracle's time for `--cpu-layers` layers of each encoder on the host cores (bounded sample, extrapolated)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-video-infinity_b200")):
    sys.path.insert(0, p)
from tools import synth_enc as SE  # noqa: E402


def device_init(model, seed):
    model.to_empty(device="cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel() if "embedding" not in name else 1
                p.copy_((torch.randn(p.shape, generator=g, device="cuda") * (1.0 / max(fan_in, 1) ** 0.5)).to(p.dtype))
            elif name.endswith("bias"):
                p.zero_()
            else:
                p.fill_(1.0)
    return model.eval()


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-layers", type=int, default=1)
    a = ap.parse_args()
    from diffsynth.models.wan_video_image_encoder import WanImageEncoder
    from diffsynth.models.wan_video_text_encoder import WanTextEncoder
    res = {}
    # ---- umT5-XXL
    cfg = SE.TEXT_UMT5_XXL
    with torch.device("meta"):
        te = WanTextEncoder(**cfg).to(torch.bfloat16)
    te = device_init(te, 0)
    ids, mask = SE.make_text_inputs(cfg, seq_len=512, valid=120, seed=0)
    ids, mask = ids.cuda(), mask.cuda()
    eng = te.engine(ids.device)
    l0 = eng.launches
    ms = timed(lambda: te(ids, mask))
    L, d, da, f, n = 512, cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_layers"]
    flops = n * (2 * L * d * 3 * da + 2 * L * da * d + 3 * 2 * L * d * f + 4 * L * L * da)
    res["umt5_xxl"] = {"ms": ms, "tflops": flops / ms / 1e9, "algorithmic_tflop": flops / 1e12, "launches": (eng.launches - l0) // 4,
                       "tokens": L, "params_b": sum(p.numel() for p in te.parameters()) / 1e9}
    del te, eng
    torch.cuda.empty_cache()
    # ---- CLIP ViT-H
    cfg = SE.CLIP_VIT_H
    with torch.device("meta"):
        ie = WanImageEncoder(**cfg).to(torch.bfloat16)
    ie = device_init(ie, 1)
    img = SE.make_clip_image(480, 832, seed=0).cuda()
    eng = ie.engine(img.device)
    l0 = eng.launches
    ms = timed(lambda: ie.encode_image([img]))
    L, d, m, n = 257, cfg["dim"], cfg["dim"] * cfg["mlp_ratio"], cfg["num_layers"] - 1
    flops = n * (2 * L * d * 3 * d + 2 * L * d * d + 2 * 2 * L * d * m + 4 * L * L * d) + 2 * 256 * 588 * d
    res["clip_vit_h"] = {"ms": ms, "tflops": flops / ms / 1e9, "algorithmic_tflop": flops / 1e12, "launches": (eng.launches - l0) // 4,
                         "tokens": L}
    # ---- CPU oracle (bounded sample)
    if a.cpu_layers > 0:
        from oracle import wan_encoders_oracle as E
        torch.set_num_threads(os.cpu_count())
        cfg = dict(SE.TEXT_UMT5_XXL, vocab=1000, num_layers=a.cpu_layers)
        sd = SE.make_text_state_dict(cfg, seed=0)
        ids_c, mask_c = SE.make_text_inputs(cfg, seq_len=512, valid=120, seed=0)
        t0 = time.perf_counter()
        E.text_encode(sd, cfg, ids_c, mask_c)
        t_text = (time.perf_counter() - t0) / a.cpu_layers
        cfg = dict(SE.CLIP_VIT_H, num_layers=a.cpu_layers + 1)
        sd = SE.make_clip_state_dict(cfg, seed=0)
        t0 = time.perf_counter()
        E.image_encode(sd, cfg, SE.make_clip_image(480, 832, seed=0), SE.CLIP_MEAN, SE.CLIP_STD)
        t_clip = (time.perf_counter() - t0) / a.cpu_layers
        res["cpu_oracle"] = {"cores": os.cpu_count(), "umt5_s_per_layer": t_text, "umt5_s_extrapolated_24_layers": 24 * t_text,
                             "clip_s_per_layer": t_clip, "clip_s_extrapolated_31_layers": 31 * t_clip,
                             "sample": f"{a.cpu_layers} layer(s) of each encoder at full width, fp32"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
