"""GPU: one whole SVI clip through the public pipeline API (SVIVideoPipeline.__call__, same call shape as
test_svi.py:457-470) on tiny random-init models, against the same clip computed by the CPU oracles
(noise -> image conditioning (VAE encode) -> CFG denoise -> VAE decode -> uint8 frames)."""
import types

import numpy as np
import pytest
import torch
from PIL import Image

from tools import synth, synth_vae

pytestmark = pytest.mark.gpu

H, W, FRAMES, STEPS, CTX = 64, 96, 9, 3, 24


class _ClipStub:
    """Stands in for CLIP ViT-H (out of scope): a fixed random projection of the pooled image."""

    def __init__(self):
        g = torch.Generator().manual_seed(99)
        self.proj = torch.randn(3, 257 * 1280, generator=g) * 0.5

    def encode_image(self, images):
        pooled = images[0].float().cpu().mean(dim=(0, 2, 3))            # [3]
        return (pooled @ self.proj).reshape(1, 257, 1280)


def _prompter(prompt, positive=True):
    g = torch.Generator().manual_seed(11 if positive else 12)
    return torch.randn(1, CTX, synth.CFG_TINY_I2V["text_dim"], generator=g)


def _image():
    g = np.random.default_rng(5)
    return Image.fromarray(g.integers(0, 255, size=(H, W, 3), dtype=np.uint8))


def _oracle_clip(dit_sd, vae_sd, cfg, img, seed, first_frames=None, ref_pad_num=-1, ref_pad_cfg=False):
    """One clip on the CPU oracles.  first_frames: conditioning frames (default: the reference image alone); the
    conditioning video is [first_frames ++ padding] as in reference svi_video.py:328-350."""
    from oracle import wan_dit_oracle as O
    from oracle import wan_vae_oracle as V
    bf = lambda t: t.to(torch.bfloat16).float()
    noise = bf(torch.randn((1, 16, (FRAMES - 1) // 4 + 1, H // 8, W // 8), generator=torch.Generator().manual_seed(seed)))
    prep = lambda im: torch.from_numpy(np.array(im, dtype=np.float32) * (2 / 255) - 1).permute(2, 0, 1).unsqueeze(0)   # [1,3,H,W]
    x = prep(img)
    firsts = [x] if first_frames is None else [prep(f) for f in first_frames]
    clip = bf(_ClipStub().encode_image([firsts[0]]))
    msk = torch.zeros(1, FRAMES, H // 8, W // 8)
    msk[:, :(len(firsts) if ref_pad_cfg else 1)] = 1
    msk = torch.cat([torch.repeat_interleave(msk[:, 0:1], 4, dim=1), msk[:, 1:]], dim=1)
    msk = msk.view(1, msk.shape[1] // 4, 4, H // 8, W // 8).transpose(1, 2)[0]
    cond = torch.cat(firsts, dim=0).permute(1, 0, 2, 3)                                                   # [3,k,H,W]
    rest = FRAMES - len(firsts)
    pad = x.transpose(0, 1).repeat(1, rest, 1, 1) if ref_pad_num == -1 else torch.zeros(3, rest, H, W)
    vae_in = torch.cat([cond, pad], dim=1)
    with torch.no_grad():
        y = bf(torch.cat([msk, V.vae_encode(vae_sd, vae_in.unsqueeze(0))[0]]).unsqueeze(0))
        lat = O.denoise(dit_sd, cfg, noise, bf(_prompter("p", True)), bf(_prompter("n", False)), steps=STEPS, cfg_scale=5.0,
                        clip_feature=clip, y=y)
        vid = V.vae_decode(vae_sd, lat)
    return lat, ((vid[0].permute(1, 2, 3, 0) + 1) * 127.5).clip(0, 255).numpy().astype(np.uint8)


def test_svi_clip_matches_oracle_clip():
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
    mm = ModelManager(torch_dtype=torch.bfloat16, device="cuda")
    mm.add_model("wan_video_dit", dit.to("cuda"))
    mm.add_model("wan_video_vae", vae.to("cuda"))
    pipe = SVIVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device="cuda", is_test=True)
    pipe.enable_vram_management(num_persistent_param_in_dit=6 * 10 ** 9)
    pipe.prompter = _prompter
    pipe.image_encoder = _ClipStub()
    img = _image()
    args = types.SimpleNamespace(ref_pad_cfg=False, ref_pad_num=-1, sequential_cfg="none")
    frames = pipe(prompt="p", negative_prompt="n", input_image=img, num_inference_steps=STEPS, cfg_scale={"text": 5.0}, seed=42,
                  tiled=False, random_ref_frame=torch.from_numpy(np.array(img)), height=H, width=W, num_frames=FRAMES,
                  args=args, progress_bar_cmd=lambda x: x)
    assert len(frames) == FRAMES and frames[0].size == (W, H)
    got = np.stack([np.array(f) for f in frames]).astype(np.float32)
    _, ref = _oracle_clip(dit_sd, vae_sd, cfg, img, seed=42)
    diff = np.abs(got - ref.astype(np.float32))
    print(f"clip parity: mean {diff.mean():.3f} levels, p99 {np.percentile(diff, 99):.1f}, max {diff.max():.0f}")
    assert diff.mean() < 2.0 and np.percentile(diff, 99) < 12


def test_svi_clip_chain_recycles_motion_frames():
    """Clip loop of test_svi.py:424-485: clip k+1 is conditioned on the last `num_motion_frames` uint8 frames of clip k
    (multi-frame conditioning video, zero padding, mask over the motion frames) with a new seed.  Clip 2 is checked
    against the oracle fed with the SAME recycled frames, so only clip 2's own arithmetic is compared."""
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
    mm = ModelManager(torch_dtype=torch.bfloat16, device="cuda")
    mm.add_model("wan_video_dit", dit.to("cuda"))
    mm.add_model("wan_video_vae", vae.to("cuda"))
    pipe = SVIVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device="cuda", is_test=True)
    pipe.prompter = _prompter
    pipe.image_encoder = _ClipStub()
    img = _image()
    ref = torch.from_numpy(np.array(img))
    num_motion_frames = 5
    args = types.SimpleNamespace(ref_pad_cfg=True, ref_pad_num=0, sequential_cfg="none")
    kw = dict(prompt="p", negative_prompt="n", num_inference_steps=STEPS, cfg_scale={"text": 5.0}, tiled=False,
              random_ref_frame=ref, height=H, width=W, num_frames=FRAMES, args=args, progress_bar_cmd=lambda x: x)
    video_list = []
    cond = img
    for k, seed in enumerate((0, 42)):                         # seeds[chunk_idx] * seed_times
        video = pipe(input_image=cond, seed=seed, **kw)
        assert len(video) == FRAMES
        if k == 1:
            _, want = _oracle_clip(dit_sd, vae_sd, cfg, img, seed, first_frames=cond, ref_pad_num=0, ref_pad_cfg=True)
            got = np.stack([np.array(f) for f in video]).astype(np.float32)
            diff = np.abs(got - want.astype(np.float32))
            print(f"clip 2 (5 motion frames) parity: mean {diff.mean():.3f} levels, p99 {np.percentile(diff, 99):.1f}")
            assert diff.mean() < 2.0 and np.percentile(diff, 99) < 12
        cond = video[-num_motion_frames:]
        video_list += video[:-num_motion_frames] if k == 0 else video
    assert len(video_list) == 2 * FRAMES - num_motion_frames   # n*81 - (n-1)*m frames kept


def test_wan_t2v_pipeline_matches_oracle():
    """WanVideoPipeline (scalar cfg_scale, no image conditioning) — reference pipelines/wan_video.py:25-286."""
    from diffsynth import ModelManager, WanVideoPipeline
    from diffsynth.models.wan_video_dit import WanModel
    from diffsynth.models.wan_video_vae import WanVideoVAE
    from oracle import wan_dit_oracle as O
    from oracle import wan_vae_oracle as V
    cfg = synth.CFG_TINY_T2V
    dit_sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=5).items()}
    vae_sd = {k: v.to(torch.bfloat16).float() for k, v in synth_vae.make_vae_state_dict(seed=0).items()}
    dit = WanModel(**cfg).eval()
    dit.load_state_dict(dit_sd)
    vae = WanVideoVAE().eval()
    vae.load_state_dict(vae_sd)
    mm = ModelManager(torch_dtype=torch.bfloat16, device="cuda")
    mm.add_model("wan_video_dit", dit.to("cuda"))
    mm.add_model("wan_video_vae", vae.to("cuda"))
    pipe = WanVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device="cuda")
    pipe.prompter = lambda prompt, positive=True: torch.randn(1, CTX, cfg["text_dim"], generator=torch.Generator().manual_seed(21 if positive else 22))
    frames = pipe(prompt="p", negative_prompt="n", num_inference_steps=STEPS, cfg_scale=5.0, seed=7, tiled=False,
                  height=H, width=W, num_frames=FRAMES, progress_bar_cmd=lambda x: x)
    got = np.stack([np.array(f) for f in frames]).astype(np.float32)
    bf = lambda t: t.to(torch.bfloat16).float()
    noise = bf(torch.randn((1, 16, (FRAMES - 1) // 4 + 1, H // 8, W // 8), generator=torch.Generator().manual_seed(7)))
    with torch.no_grad():
        lat = O.denoise(dit_sd, cfg, noise, bf(pipe.prompter("p", True)), bf(pipe.prompter("n", False)), steps=STEPS, cfg_scale=5.0)
        vid = V.vae_decode(vae_sd, lat)
    ref = ((vid[0].permute(1, 2, 3, 0) + 1) * 127.5).clip(0, 255).numpy().astype(np.uint8).astype(np.float32)
    diff = np.abs(got - ref)
    print(f"t2v clip parity: mean {diff.mean():.3f} levels, p99 {np.percentile(diff, 99):.1f}, max {diff.max():.0f}")
    assert diff.mean() < 2.0 and np.percentile(diff, 99) < 12


def test_lora_merge_on_native_gemm_and_engine_invalidation():
    """load_lora_v2: W += alpha * B @ A (reference lora.py:246-267) computed through the GEMM epilogue."""
    from diffsynth import ModelManager
    from diffsynth.models.wan_video_dit import WanModel
    cfg = synth.CFG_TINY_T2V
    sd = synth.make_dit_state_dict(cfg, seed=9)
    dit = WanModel(**cfg).eval()
    dit.load_state_dict(sd)
    dit.to(device="cuda", dtype=torch.bfloat16)
    mm = ModelManager(torch_dtype=torch.bfloat16, device="cuda")
    mm.add_model("wan_video_dit", dit)
    inp = synth.make_dit_inputs(cfg, 2, 8, 8, seed=9, ctx_len=16)
    before = dit(inp["x"].cuda(), torch.tensor([500.0]), inp["context"].cuda()).float().clone()
    g = torch.Generator().manual_seed(1)
    r, alpha = 12, 0.5
    lora = {}
    targets = ["blocks.0.self_attn.q", "blocks.1.ffn.0", "blocks.1.ffn.2"]
    for t in targets:
        w = dict(dit.named_parameters())[t + ".weight"]
        lora[f"pipe.dit.{t}.lora_A.default.weight"] = torch.randn(r, w.shape[1], generator=g) * 0.1
        lora[f"pipe.dit.{t}.lora_B.default.weight"] = torch.randn(w.shape[0], r, generator=g) * 0.1
    w0 = {t: dict(dit.named_parameters())[t + ".weight"].detach().float().clone() for t in targets}
    mm.load_lora_v2("synthetic.safetensors", state_dict=lora, lora_alpha=alpha)
    for t in targets:
        want = w0[t] + alpha * (lora[f"pipe.dit.{t}.lora_B.default.weight"].cuda() @ lora[f"pipe.dit.{t}.lora_A.default.weight"].cuda())
        got = dict(dit.named_parameters())[t + ".weight"].detach().float()
        assert (got - want).abs().max().item() < 2e-2 * want.abs().max().item()       # one bf16 rounding of the merged weight
    after = dit(inp["x"].cuda(), torch.tensor([500.0]), inp["context"].cuda()).float()
    assert (after - before).abs().max().item() > 1e-3      # the engine picked up the merged weights
