"""One-GPU timing of the 14B image-to-video DiT (BASELINE configs[2] model, SURVEY.md §8 cfg-3 family): Wan2.1-I2V-14B
(d 5120, 40 heads, 40 layers, ffn 13824, 257 CLIP tokens + 512 text tokens) on an 81-frame clip, random-init weights
created on the device.

    python tools/dit14b_bench.py [--height 480 --width 832] [--steps 2]

Prints one JSON object: milliseconds per CFG step (2 forwards + fused CFG/Euler), algorithmic TFLOP/s, launches,
self-attention time per launch.  The 8-GPU sequence-parallel run of cfg-3 needs 8 GPUs; this is the single-GPU point."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-video-infinity_b200")):
    sys.path.insert(0, p)
from tools import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    from diffsynth.models.wan_video_dit import WanModel, precompute_freqs_cis_3d
    from oracle import wan_dit_oracle as O      # FLOP formula only
    cfg = synth.CFG_I2V_14B
    with torch.device("meta"):
        m = WanModel(**cfg).to(torch.bfloat16)
    m.to_empty(device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() >= 2 and "modulation" not in name:
                p.copy_((torch.randn(p.shape, generator=g, device="cuda") * (1.0 / p[0].numel() ** 0.5)).to(p.dtype))
            elif name.endswith("bias"):
                p.zero_()
            elif "modulation" in name:
                p.copy_((torch.randn(p.shape, generator=g, device="cuda") * 0.02).to(p.dtype))
            else:
                p.fill_(1.0)
    m.freqs = precompute_freqs_cis_3d(cfg["dim"] // cfg["num_heads"])     # built under the meta context above: redo on the CPU
    m.eval()
    eng = m.engine("cuda")
    f, h, w = (a.frames - 1) // 4 + 1, a.height // 8, a.width // 8
    inp = synth.make_dit_inputs(cfg, f, h, w, seed=0, ctx_len=512)
    x = inp["x"].cuda().float()
    y, clip = inp["y"].cuda(), inp["clip_feature"].cuda()
    cp = eng.context_state(inp["context"].cuda(), clip)
    cn = eng.context_state(torch.randn(inp["context"].shape, generator=torch.Generator().manual_seed(2)).cuda(), clip)
    v_c, v_u = torch.empty_like(x), torch.empty_like(x)

    def step(t):
        eng.forward(x, t, cp, y=y, out=v_c)
        eng.forward(x, t, cn, y=y, out=v_u)
        eng.k.cfg_euler_step(x, v_c, v_u, 5.0, 0.9, 0.88)

    step(900.0)
    step(880.0)                      # second call of the geometry: captured into a CUDA graph
    torch.cuda.synchronize()
    n0 = eng.k.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(860.0 - 20.0 * i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    L = f * (h // 2) * (w // 2)
    flops = 2 * O.dit_forward_flops(cfg, L, 512)
    # self-attention time per launch (eager, bracketed)
    eng.attn_events = []
    eng.forward(x, 700.0, cp, y=y, out=v_c)
    torch.cuda.synchronize()
    att = [s.elapsed_time(e) for s, e in eng.attn_events]
    eng.attn_events = None
    print(json.dumps({"model": "Wan2.1-I2V-14B (random init)", "clip": f"{a.frames}f x {a.height}x{a.width}", "tokens": L,
                      "ms_per_step": ms, "latent_fps_50_steps": f / (50 * ms / 1e3), "dit_tflops": flops / ms / 1e9,
                      "algorithmic_tflop_per_step": flops / 1e12, "launches_per_step": (eng.k.launches - n0) // a.steps,
                      "self_attn_ms": sum(att) / len(att), "self_attn_tflops": 4.0 * L * L * cfg["dim"] / (sum(att) / len(att)) / 1e9,
                      "params_b": sum(p.numel() for p in m.parameters()) / 1e9,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
