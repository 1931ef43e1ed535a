"""Multi-GPU check (run under torchrun, 1 process per GPU): the native sequence-parallel / CFG-parallel DiT step
against the single-GPU native result on every rank.  xfuser USP parity is unpinned in the reference
(SURVEY.md §8c), so the single-GPU kernel path is the oracle here.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-video-infinity_b200")):
    sys.path.insert(0, p)
from tools import synth  # noqa: E402


def _stats(out, ref):
    err = (out - ref).abs()
    return (err <= 1e-3 + 1e-2 * ref.abs()).float().mean().item(), err.max().item(), err.mean().item() / ref.std().item()


def oracle_cases(rank, world, dev):
    """The plan against the CPU ORACLE (not only against the single-GPU kernels): 1.3B-width models whose rows per rank are
    NOT a multiple of the 128-row attention tile, so K/V tiles straddle two owners' chunks of the peer buffer —
    L = 1260 (sp 2: 630 rows, sp 4: 315) with two layers, and the BENCHMARK shape L = 32760 (sp 2: 16380, sp 4: 8190 rows)
    with one layer.  Rank 0 computes the oracle; the verdict is shared."""
    from diffsynth.distributed.sequence_parallel import SequenceParallelGroup
    from diffsynth.models.wan_video_dit import WanModel
    from oracle import wan_dit_oracle as O
    ok = True
    for layers, (f, h, w), bar in ((2, (5, 28, 36), 0.95), (1, (21, 60, 104), 0.95)):
        cfg = dict(synth.CFG_T2V_1_3B, num_layers=layers)
        sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=3).items()}
        m = WanModel(**cfg).eval()
        m.load_state_dict(sd)
        m.to(dev)
        eng = m.engine(dev)
        inp = synth.make_dit_inputs(cfg, f, h, w, seed=3, ctx_len=512)
        ts = torch.tensor([800.0])
        ref = None
        if rank == 0:
            torch.set_num_threads(os.cpu_count() or 1)
            with torch.no_grad():
                ref = O.dit_forward(sd, cfg, inp["x"], ts, inp["context"])
        x, ctx = inp["x"].to(dev), inp["context"].to(dev)
        L = f * (h // 2) * (w // 2)
        for cfg_parallel in (True, False):
            sp = SequenceParallelGroup(world, rank, cfg_parallel=cfg_parallel)
            if sp.sp_size == 1 or L % sp.sp_size:
                continue
            out = eng.forward(x, ts, ctx, sp=sp).float().cpu()
            torch.cuda.synchronize()
            if rank == 0:
                inside, mx, rel = _stats(out, ref)
                good = inside > bar and mx < 0.02
                ok &= good
                print(f"[oracle] {sp.describe()} layers={layers} L={L} rows/rank={L // sp.sp_size} (mod 128 = {(L // sp.sp_size) % 128}) "
                      f"peer={sp._peer is not None}: inside={inside:.4f} max={mx:.4e} mean/std={rel:.4e} {'OK' if good else 'BAD'}", flush=True)
            if sp._peer is not None:
                sp._peer.close()
                sp._peer = None
        del m, eng
        torch.cuda.empty_cache()
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from diffsynth.distributed.sequence_parallel import SequenceParallelGroup
    from diffsynth.models.wan_video_dit import WanModel
    ok = True
    # token counts: 128 (slabs < one attention tile -> NCCL all-gather path), 1024 and the ragged 540 (peer exchange)
    for cfg, (f, h, w) in ((synth.CFG_TINY_T2V, (4, 8, 16)), (synth.CFG_TINY_T2V, (4, 32, 32)), (synth.CFG_TINY_I2V, (3, 20, 36))):
        sd = synth.make_dit_state_dict(cfg, seed=0)
        m = WanModel(**cfg).eval()
        m.load_state_dict(sd)
        m.to(dev)
        eng = m.engine(dev)
        inp = synth.make_dit_inputs(cfg, f, h, w, seed=0, ctx_len=40)
        x = inp["x"].to(dev)
        kw = {k: inp[k].to(dev) for k in ("clip_feature", "y") if k in inp}
        ref = eng.forward(x, 500.0, inp["context"].to(dev), **kw).clone()
        for cfg_parallel in (False, True):
            sp = SequenceParallelGroup(world, rank, cfg_parallel=cfg_parallel)
            if sp.sp_size > 1:
                out = eng.forward(x, 500.0, inp["context"].to(dev), sp=sp, **kw)
                err = (out - ref).abs().max().item()
                good = err < 2e-2
                ok &= good
                print(f"[rank {rank}] {sp.describe()} L={f * (h // 2) * (w // 2)} peer={sp._peer is not None} has_image={cfg['has_image_input']} forward max|sp - single| = {err:.3e} {'OK' if good else 'BAD'}", flush=True)
            # full step under the plan vs two local forwards + fused update
            cp = eng.context_state(inp["context"].to(dev), kw.get("clip_feature"))
            ctx2 = torch.randn(inp["context"].shape, generator=torch.Generator().manual_seed(77)).to(dev)   # same on every rank
            cn = eng.context_state(ctx2, kw.get("clip_feature"))
            lat_a, lat_b = x.clone().float(), x.clone().float()
            vc, vu = torch.empty_like(lat_a), torch.empty_like(lat_a)
            sp.cfg_parallel_step(eng, lat_a, 500.0, cp, cn, vc, vu, 5.0, 0.9, 0.8, y=kw.get("y"))
            eng.forward(lat_b, 500.0, cp, y=kw.get("y"), out=vc)
            eng.forward(lat_b, 500.0, cn, y=kw.get("y"), out=vu)
            eng.k.cfg_euler_step(lat_b, vc, vu, 5.0, 0.9, 0.8)
            err = (lat_a - lat_b).abs().max().item()
            good = err < 2e-2
            ok &= good
            print(f"[rank {rank}] {sp.describe()} step max|plan - single| = {err:.3e} {'OK' if good else 'BAD'}", flush=True)
    ok &= oracle_cases(rank, world, dev)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("SP_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
