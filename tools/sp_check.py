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
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("SP_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
