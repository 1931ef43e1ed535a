"""CPU: the public pipeline call (SVIVideoPipeline.__call__, the call shape of test_svi.py:457-470) on kernel stand-ins
(tests/nv_emulation.py): image conditioning -> VAE encode -> CFG denoise loop -> VAE decode -> uint8 frames against the CPU
oracles, and the multi-GPU plan `use_usp=True` installs (CFG-parallel x sequence-parallel denoise + row-band VAE) under a 2-process
gloo group against the single-process frames.  The GPU tests check the kernels; this checks everything the pipeline does around
them, including the N > 1 path, on the CPU-only build box."""
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tools import synth, synth_vae

H, W, FRAMES, STEPS = 32, 48, 5, 2


def _build(device="cpu", use_usp=False):
    import test_pipeline_gpu as P
    from diffsynth import ModelManager, SVIVideoPipeline
    from diffsynth.models.wan_video_dit import WanModel
    from diffsynth.models.wan_video_vae import WanVideoVAE
    cfg = synth.CFG_TINY_I2V
    dit_sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=2).items()}
    vae_sd = {k: v.to(torch.bfloat16).float() for k, v in synth_vae.make_vae_state_dict(seed=0).items()}
    dit = WanModel(**cfg).eval()
    dit.load_state_dict(dit_sd)
    vae = WanVideoVAE().eval()
    vae.load_state_dict(vae_sd)
    mm = ModelManager(torch_dtype=torch.bfloat16, device=device)
    mm.add_model("wan_video_dit", dit)
    mm.add_model("wan_video_vae", vae)
    pipe = SVIVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device=device, is_test=True, use_usp=use_usp)
    pipe.prompter = P._prompter
    pipe.image_encoder = P._ClipStub()
    return pipe, cfg, dit_sd, vae_sd


def _build_dance(use_usp=False):
    import test_dance_gpu as D
    import test_pipeline_gpu as P
    from diffsynth import ModelManager, SVIDanceVideoPipeline
    from diffsynth.models.wan_video_dit import WanModel
    from diffsynth.models.wan_video_vae import WanVideoVAE
    cfg = synth.CFG_TINY_I2V
    dit = WanModel(**cfg).eval()
    dit.load_state_dict({k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=2).items()})
    vae = WanVideoVAE().eval()
    vae.load_state_dict({k: v.to(torch.bfloat16).float() for k, v in synth_vae.make_vae_state_dict(seed=0).items()})
    mm = ModelManager(torch_dtype=torch.bfloat16, device="cpu")
    mm.add_model("wan_video_dit", dit)
    mm.add_model("wan_video_vae", vae)
    mm.state_dict_new_module = {"pipe.dwpose_embedding." + k: v for k, v in D._stem_sd(cfg["dim"], seed=7).items()}
    pipe = SVIDanceVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device="cpu", is_test=True, use_usp=use_usp)
    pipe.prompter = P._prompter
    pipe.image_encoder = P._ClipStub()
    return pipe


def _call_dance(pipe, img, seed):
    import test_dance_gpu as D
    args = types.SimpleNamespace(ref_pad_cfg=False, ref_pad_num=-1, sequential_cfg="none")
    return pipe(prompt="p", negative_prompt="n", input_image=img, num_inference_steps=STEPS, cfg_scale={"text": 5.0}, seed=seed,
                tiled=False, random_ref_frame=torch.from_numpy(np.array(img)), height=H, width=W, num_frames=FRAMES,
                humanpose_data=D._pose(FRAMES, H, W, seed=3), cond_wo_pose=False, args=args, progress_bar_cmd=lambda x: x)


def _build_t2v(use_usp=False):
    import test_pipeline_gpu as P
    from diffsynth import ModelManager, WanVideoPipeline
    from diffsynth.models.wan_video_dit import WanModel
    from diffsynth.models.wan_video_vae import WanVideoVAE
    cfg = synth.CFG_TINY_T2V
    dit = WanModel(**cfg).eval()
    dit.load_state_dict({k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=4).items()})
    vae = WanVideoVAE().eval()
    vae.load_state_dict({k: v.to(torch.bfloat16).float() for k, v in synth_vae.make_vae_state_dict(seed=0).items()})
    mm = ModelManager(torch_dtype=torch.bfloat16, device="cpu")
    mm.add_model("wan_video_dit", dit)
    mm.add_model("wan_video_vae", vae)
    pipe = WanVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device="cpu", use_usp=use_usp)
    td = cfg["text_dim"]
    pipe.prompter = lambda prompt, positive=True: torch.randn(1, 24, td, generator=torch.Generator().manual_seed(31 if positive else 32))
    return pipe


def _call_t2v(pipe, seed):
    return pipe(prompt="p", negative_prompt="n", num_inference_steps=STEPS, cfg_scale=5.0, seed=seed, tiled=False, height=H, width=W,
                num_frames=FRAMES, progress_bar_cmd=lambda x: x)


def _call(pipe, img, seed, steps=STEPS, **extra):
    args = types.SimpleNamespace(ref_pad_cfg=False, ref_pad_num=-1, sequential_cfg="none")
    return pipe(prompt="p", negative_prompt="n", input_image=img, num_inference_steps=steps, cfg_scale={"text": 5.0}, seed=seed,
                tiled=False, random_ref_frame=torch.from_numpy(np.array(img)), height=H, width=W, num_frames=FRAMES,
                args=args, progress_bar_cmd=lambda x: x, **extra)


def _image():
    from PIL import Image
    g = np.random.default_rng(5)
    return Image.fromarray(g.integers(0, 255, size=(H, W, 3), dtype=np.uint8))


def test_pipeline_call_matches_oracle_clip(monkeypatch):
    import nv_emulation
    import test_pipeline_gpu as P
    nv_emulation.install(monkeypatch)
    monkeypatch.setenv("SVI_CUDA_GRAPHS", "0")
    monkeypatch.setattr(P, "H", H), monkeypatch.setattr(P, "W", W), monkeypatch.setattr(P, "FRAMES", FRAMES), monkeypatch.setattr(P, "STEPS", STEPS)
    pipe, cfg, dit_sd, vae_sd = _build()
    img = _image()
    frames = _call(pipe, img, seed=42)
    assert len(frames) == FRAMES and frames[0].size == (W, H)
    got = np.stack([np.array(f) for f in frames]).astype(np.float32)
    _, ref = P._oracle_clip(dit_sd, vae_sd, cfg, img, seed=42)
    diff = np.abs(got - ref.astype(np.float32))
    assert diff.mean() < 2.0 and np.percentile(diff, 99) < 12, (diff.mean(), np.percentile(diff, 99))


def _usp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), SVI_CUDA_GRAPHS="0")
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "stable-video-infinity_b200"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    import nv_emulation
    nv_emulation.install()
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = _image()
        ref = np.stack([np.array(f) for f in _call(_build()[0], img, seed=7)]).astype(np.int32)          # this process alone
        pipe = _build(use_usp=True)[0]
        plan = pipe.sp_group().describe() if hasattr(pipe, "sp_group") and pipe.sp_group() is not None else "none"
        got = np.stack([np.array(f) for f in _call(pipe, img, seed=7)]).astype(np.int32)
        eng = pipe.vae.engine("cpu")
        res = {"plan": plan, "max_diff": int(np.abs(got - ref).max()), "halo_exchanges": eng.halo_exchanges}
        # SVI-Dance: the pose condition (add_condition rows) must follow the token split of the plan
        dref = np.stack([np.array(f) for f in _call_dance(_build_dance(), img, seed=9)]).astype(np.int32)
        dgot = np.stack([np.array(f) for f in _call_dance(_build_dance(use_usp=True), img, seed=9)]).astype(np.int32)
        res["dance_max_diff"] = int(np.abs(dgot - dref).max())
        # the plain Wan text-to-video pipeline (reference pipelines/wan_video.py) takes the same plan
        wp = _build_t2v(use_usp=True)
        wref = np.stack([np.array(f) for f in _call_t2v(_build_t2v(), seed=5)]).astype(np.int32)
        wgot = np.stack([np.array(f) for f in _call_t2v(wp, seed=5)]).astype(np.int32)
        res["t2v_max_diff"] = int(np.abs(wgot - wref).max())
        res["t2v_vae_sharded"] = wp.vae.engine("cpu").halo_exchanges > 0
        res["dance_differs_from_svi"] = bool(np.abs(dref - ref).max() > 0)
        # TeaCache under the plan: every rank must take the same skip decisions (they depend on the timestep embedding only) and
        # keep its own rows' residual
        tea = dict(steps=8, tea_cache_l1_thresh=22000.0, tea_cache_model_id="Wan2.1-I2V-14B-720P")   # random weights: a threshold that skips
        single = _build()[0]
        tref = np.stack([np.array(f) for f in _call(single, img, seed=11, **tea)]).astype(np.int32)
        tgot = np.stack([np.array(f) for f in _call(pipe, img, seed=11, **tea)]).astype(np.int32)
        plain = np.stack([np.array(f) for f in _call(single, img, seed=11, steps=8)]).astype(np.int32)
        res["tea_max_diff"] = int(np.abs(tgot - tref).max())
        res["tea_skips_steps"] = bool(np.abs(tref - plain).max() > 0)
        q.put((rank, res))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, {"error": traceback.format_exc()[-2500:] + repr(ex)}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_pipeline_call_with_use_usp(world):
    """`from_model_manager(use_usp=True)` (reference pipelines/svi_video.py:259-275) on 2 and 4 gloo ranks: the guidance branches on
    two halves of the ranks (x 2-way token split at 4), the VAE on row bands of all ranks; every rank must return the frames of the
    single-process pipeline."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + (os.getpid() % 150) + world
    procs = [ctx.Process(target=_usp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in got:
        assert "error" not in res, res["error"]
        assert res["plan"].startswith("cfg2"), res
        assert res["halo_exchanges"] > 0, res                 # the VAE really ran on row bands
        assert res["max_diff"] <= 1, res                      # bf16 re-rounding of a token-split forward may flip a grey level
        assert res["dance_max_diff"] <= 1 and res["dance_differs_from_svi"], res
        assert res["tea_max_diff"] <= 1 and res["tea_skips_steps"], res
        assert res["t2v_max_diff"] <= 1 and res["t2v_vae_sharded"], res
