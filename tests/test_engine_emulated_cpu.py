"""CPU: the HOST logic of WanDiTEngine — operation order, fused-weight layouts, buffer reuse, RoPE row offsets, the
sequence-parallel / CFG-parallel plan under a 2-process gloo group — with the kernels replaced by the torch stand-ins of
tests/nv_emulation.py (the product itself refuses to run without CUDA; see test_host_logic.py).  The GPU tests check the
kernels; these check everything around them on the CPU-only build box, including the N > 1 data path."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tools import synth


def _stats(out, ref):
    err = (out - ref).abs()
    return (err <= 1e-3 + 1e-2 * ref.abs()).float().mean().item(), err.max().item(), err.mean().item() / ref.std().item()


def _model(cfg, sd):
    from diffsynth.models.wan_video_dit import WanModel
    m = WanModel(**cfg).eval()
    m.load_state_dict(sd)
    return m.to(torch.bfloat16)


@pytest.mark.parametrize("cfg,f,h,w,seed", [(synth.CFG_TINY_T2V, 3, 8, 12, 0), (synth.CFG_TINY_I2V, 2, 6, 10, 1),
                                            (synth.CFG_TINY_T2V, 3, 16, 24, 2), (synth.CFG_TINY_I2V, 2, 20, 28, 3)])   # > 128 rows: LayerNorm fold
def test_engine_orchestration_matches_oracle(monkeypatch, cfg, f, h, w, seed):
    from oracle import wan_dit_oracle as O
    import nv_emulation
    nv_emulation.install(monkeypatch)
    monkeypatch.setenv("SVI_CUDA_GRAPHS", "0")
    monkeypatch.setenv("SVI_LN_FOLD", "1" if f * (h // 2) * (w // 2) > 128 else "0")     # the optional fold has its own host logic
    sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=seed).items()}
    m = _model(cfg, sd)
    inp = synth.make_dit_inputs(cfg, f, h, w, seed=seed, ctx_len=24)
    kw = {k: inp[k] for k in ("clip_feature", "y") if k in inp}
    ts = torch.tensor([700.0])
    eng = m.engine("cpu")
    out = eng.forward(inp["x"], ts, inp["context"], **kw)
    ref = O.dit_forward(sd, cfg, inp["x"], ts, inp["context"], kw.get("clip_feature"), kw.get("y"))
    inside, mx, rel = _stats(out, ref)
    assert inside > 0.80 and rel < 4e-3, (inside, mx, rel)
    # second call with another prompt / timestep re-uses every work buffer: must not leak state from the first
    inp2 = synth.make_dit_inputs(cfg, f, h, w, seed=seed + 10, ctx_len=24)
    kw2 = {k: inp2[k] for k in ("clip_feature", "y") if k in inp2}
    out2 = eng.forward(inp2["x"], torch.tensor([300.0]), inp2["context"], **kw2)
    ref2 = O.dit_forward(sd, cfg, inp2["x"], torch.tensor([300.0]), inp2["context"], kw2.get("clip_feature"), kw2.get("y"))
    inside2, _, rel2 = _stats(out2, ref2)
    assert inside2 > 0.80 and rel2 < 4e-3
    # TeaCache and add_condition run through the same engine entry
    cond = torch.randn(1, f * (h // 2) * (w // 2), cfg["dim"], generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).float()
    out3 = eng.forward(inp["x"], ts, inp["context"], add_condition=cond, **kw)
    ref3 = O.dit_forward(sd, cfg, inp["x"], ts, inp["context"], kw.get("clip_feature"), kw.get("y"), add_condition=cond)
    assert _stats(out3, ref3)[2] < 4e-3


def test_talk_branch_orchestration_matches_oracle(monkeypatch):
    """SVI-Talk host logic on the emulated kernels: audio projector, per-layer audio K|V, the per-frame attention slicing."""
    import nv_emulation
    from diffsynth.pipelines.svi_video_talk import preprocess_audio
    from oracle import wan_dit_oracle as O
    nv_emulation.install(monkeypatch)
    monkeypatch.setenv("SVI_CUDA_GRAPHS", "0")
    cfg = synth.CFG_TINY_TALK
    sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=5).items()}
    m = _model(cfg, sd)
    f, h, w = 3, 8, 8
    inp = synth.make_dit_inputs(cfg, f, h, w, seed=5, ctx_len=24)
    tup = preprocess_audio(synth.make_audio_embed(4 * (f - 1) + 1, seed=5).to(torch.bfloat16).float())
    ts = torch.tensor([600.0])
    out = m.engine("cpu").forward(inp["x"], ts, inp["context"], inp["clip_feature"], inp["y"], audio=tup)
    ref = O.dit_forward(sd, cfg, inp["x"], ts, inp["context"], inp["clip_feature"], inp["y"], audio_embed_tuple=tup)
    inside, mx, rel = _stats(out, ref)
    assert inside > 0.80 and rel < 4e-3, (inside, mx, rel)


def _sp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      SVI_CUDA_GRAPHS="0")
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "stable-video-infinity_b200"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    import nv_emulation
    nv_emulation.install()
    from diffsynth.distributed.sequence_parallel import SequenceParallelGroup
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        for cfg, (f, h, w) in ((synth.CFG_TINY_T2V, (2, 8, 8)), (synth.CFG_TINY_I2V, (2, 4, 12))):
            sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=0).items()}
            m = _model(cfg, sd)
            eng = m.engine("cpu")
            inp = synth.make_dit_inputs(cfg, f, h, w, seed=0, ctx_len=24)
            kw = {k: inp[k] for k in ("clip_feature", "y") if k in inp}
            ref = eng.forward(inp["x"], 500.0, inp["context"], **kw).clone()
            tag = "i2v" if cfg["has_image_input"] else "t2v"
            for cfg_parallel in (False, True):
                sp = SequenceParallelGroup(world, rank, cfg_parallel=cfg_parallel)
                if sp.sp_size > 1:      # token-axis split: every rank must reproduce the full single-process output
                    out = eng.forward(inp["x"], 500.0, inp["context"], sp=sp, **kw)
                    res[f"{tag} {sp.describe()} forward"] = (out - ref).abs().max().item()
                cp = eng.context_state(inp["context"], kw.get("clip_feature"))
                ctx2 = torch.randn(inp["context"].shape, generator=torch.Generator().manual_seed(77))
                cn = eng.context_state(ctx2, kw.get("clip_feature"))
                lat_a, lat_b = inp["x"].clone().float(), inp["x"].clone().float()
                vc, vu = torch.empty_like(lat_a), torch.empty_like(lat_a)
                sp.cfg_parallel_step(eng, lat_a, 500.0, cp, cn, vc, vu, 5.0, 0.9, 0.8, y=kw.get("y"))
                eng.forward(lat_b, 500.0, cp, y=kw.get("y"), out=vc)
                eng.forward(lat_b, 500.0, cn, y=kw.get("y"), out=vu)
                eng.k.cfg_euler_step(lat_b, vc, vu, 5.0, 0.9, 0.8)
                res[f"{tag} {sp.describe()} step"] = (lat_a - lat_b).abs().max().item()
        # SVI-Talk under the token split: 3 latent frames x 16 tokens over 2 ranks -> rank rows start / end inside frame 1
        from diffsynth.pipelines.svi_video_talk import preprocess_audio
        cfg = synth.CFG_TINY_TALK
        sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=5).items()}
        eng = _model(cfg, sd).engine("cpu")
        inp = synth.make_dit_inputs(cfg, 3, 8, 8, seed=5, ctx_len=24)
        tup = preprocess_audio(synth.make_audio_embed(9, seed=5).to(torch.bfloat16).float())
        ref = eng.forward(inp["x"], 500.0, inp["context"], inp["clip_feature"], inp["y"], audio=tup).clone()
        sp = SequenceParallelGroup(world, rank, cfg_parallel=False)
        out = eng.forward(inp["x"], 500.0, inp["context"], inp["clip_feature"], inp["y"], sp=sp, audio=tup)
        res["talk cfg1xsp2 forward"] = (out - ref).abs().max().item()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_sequence_and_cfg_parallel_plan_two_ranks():
    """world_size 2 over gloo: the sequence-parallel forward (rows split over the ranks, K|V all-gathered per layer, RoPE
    with the rank's row offset, head rows gathered) and the CFG-parallel step (one branch per rank, velocity all-gather)
    against the single-process result, on both ranks, for a T2V and an I2V model."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_sp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in got:
        assert len(res) == 7, res                     # 2 models x (sp2 forward, sp2 step, cfg2 step) + talk sp2 forward
        for name, err in res.items():
            assert err < 2e-2, (rank, name, err)


def test_sequence_and_cfg_parallel_plan_four_ranks():
    """world_size 4 over gloo: the plan the public pipeline installs on 4 GPUs — two guidance branches x 2-way token split
    (cfg2 x sp2: two K|V groups, velocity exchange between the branch groups) — and the pure 4-way token split, against the
    single-process result on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 150)
    procs = [ctx.Process(target=_sp_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in got:
        # 2 models x (sp4 forward, sp4 step, cfg2xsp2 forward, cfg2xsp2 step) + talk sp4 forward
        assert len(res) == 9, res
        for name, err in res.items():
            assert err < 2e-2, (rank, name, err)
