"""Multi-GPU check (torchrun, 1 process per GPU): the spatially sharded Wan VAE (bands of image rows per rank, one border
row exchanged per conv input, tokens gathered for the attention block) against the single-GPU engine on every rank.
Every output pixel is the same sequence of tensor-core accumulations in both runs, so the results must agree to the last
bit (the only reassociation would be a different K order, and there is none).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/vae_shard_check.py [--bench]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-video-infinity_b200")):
    sys.path.insert(0, p)
from tools import synth_vae  # noqa: E402


def timed(fn):
    fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, t.item()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from diffsynth.models.wan_video_vae import WanVideoVAE
    vae = WanVideoVAE().eval()
    vae.load_state_dict(synth_vae.make_vae_state_dict(seed=0))
    vae.to(dev)
    ok = True
    # (frames, H, W): latent rows 12 (even split), 9 (uneven split over 2 / 4 ranks), 60 (480p)
    cases = [(9, 96, 64), (5, 72, 80)] + ([(17, 480, 832)] if "--bench" in sys.argv else [])
    for T, H, W in cases:
        g = torch.Generator().manual_seed(T + H)
        video = (torch.rand(3, T, H, W, generator=g) * 2 - 1).to(dev)
        z = torch.randn(1, 16, (T - 1) // 4 + 1, H // 8, W // 8, generator=g).to(dev)
        if H // 8 < world:
            continue
        vae.shard_group = None
        enc1, t_e1 = timed(lambda: vae.encode([video], device=dev))
        dec1, t_d1 = timed(lambda: vae.decode(z, device=dev))
        vae.enable_spatial_sharding()
        encP, t_eP = timed(lambda: vae.encode([video], device=dev))
        decP, t_dP = timed(lambda: vae.decode(z, device=dev))
        ee, de = (enc1 - encP).abs().max().item(), (dec1 - decP).abs().max().item()
        good = ee < 1e-5 and de < 1e-5 and encP.shape == enc1.shape and decP.shape == dec1.shape
        ok &= good
        if rank == 0:
            print(json.dumps({"case": f"{T}f x {H}x{W}", "ranks": world, "encode_max_abs_diff": ee, "decode_max_abs_diff": de,
                              "encode_ms_1gpu": t_e1, "encode_ms_sharded": t_eP, "decode_ms_1gpu": t_d1, "decode_ms_sharded": t_dP,
                              "halo_exchanges_decode": vae.engine(dev).halo_exchanges, "ok": good}), flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("VAE_SHARD_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
