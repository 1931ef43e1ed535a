"""VAE throughput on one GPU (BASELINE configs[4] style sweep, reduced): decode / encode wall time (CUDA events),
algorithmic TFLOP/s from the per-position formulas of BASELINE.md §3, and per-kernel launch counts.

    python tools/vae_bench.py [--frames 81] [--height 480] [--width 832]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-video-infinity_b200")):
    sys.path.insert(0, p)
from tools import synth_vae  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    from diffsynth.models.wan_video_vae import WanVideoVAE
    vae = WanVideoVAE().eval()
    vae.load_state_dict(synth_vae.make_vae_state_dict(seed=0))
    vae.to("cuda")
    eng = vae.engine("cuda")
    T, H, W = a.frames, a.height, a.width
    tl, h, w = (T - 1) // 4 + 1, H // 8, W // 8
    g = torch.Generator().manual_seed(0)
    z = torch.randn(16, tl, h, w, generator=g).cuda()
    video = (torch.rand(3, T, H, W, generator=g) * 2 - 1).cuda()
    res = {}
    for name, fn, flops in (("decode", lambda: eng.decode(z), (0.688 + 2.162 * (tl - 1)) * 1e9 * h * w),
                            ("encode", lambda: eng.encode(video), (6.66 + 5.00 * (tl - 1)) * 1e6 * H * W)):
        fn()
        torch.cuda.synchronize()
        l0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        res[name] = {"ms": ms, "tflops": flops / ms / 1e9, "algorithmic_tflop": flops / 1e12,
                     "launches": (eng.launches - l0) // a.iters, "frames_per_s": T / ms * 1e3}
    print(json.dumps({"workload": f"vae {T}f x {H}x{W}", "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, **res}))


if __name__ == "__main__":
    main()
