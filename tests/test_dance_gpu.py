"""GPU: SVI-Dance — the pose stem on the native kernels against the oracle, and one pose-conditioned clip through
SVIDanceVideoPipeline against the same clip computed by the CPU oracles (reference pipelines/svi_video_dance.py)."""
import types

import numpy as np
import pytest
import torch
from PIL import Image

from tools import synth, synth_vae

pytestmark = pytest.mark.gpu

H, W, FRAMES, STEPS, CTX = 64, 96, 9, 3, 24


def _stem_sd(dim, seed):
    from diffsynth.models.dwpose_embedding import make_dwpose_embedding
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in make_dwpose_embedding(dim=dim).state_dict().items():
        if k.endswith("weight"):
            fan_in = v[0].numel()
            sd[k] = (torch.randn(v.shape, generator=g) * (1.5 / fan_in ** 0.5)).to(torch.bfloat16).float()
        else:
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
    return sd


def _pose(T, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (3, T, h, w), generator=g).float()


def test_pose_stem_matches_oracle():
    from diffsynth.models.dwpose_embedding import DWPoseEmbeddingEngine, make_dwpose_embedding, pose_condition
    from oracle import dwpose_oracle as DO
    for (dim, T, h, w) in ((256, 9, 32, 48), (5120, 5, 64, 96)):
        sd = _stem_sd(dim, seed=dim)
        seq = make_dwpose_embedding(dim=dim)
        seq.load_state_dict(sd)
        pose = _pose(T, h, w, seed=T)
        got = pose_condition(DWPoseEmbeddingEngine(seq, "cuda"), pose).cpu()
        want = DO.pose_condition(sd, pose)
        assert got.shape == want.shape == (1, ((T + 3) // 4) * (h // 16) * (w // 16), dim)
        err = (got - want).abs()
        rel = err.mean().item() / want.std().item()
        print(f"pose stem dim={dim}: max={err.max().item():.4e} mean/std={rel:.4e}")
        assert rel < 1e-2 and err.max().item() < 0.1 * want.abs().max().item()


class _ClipStub:
    def __init__(self):
        self.proj = torch.randn(3, 257 * 1280, generator=torch.Generator().manual_seed(99)) * 0.5

    def encode_image(self, images):
        return (images[0].float().cpu().mean(dim=(0, 2, 3)) @ self.proj).reshape(1, 257, 1280)


def _prompter(prompt, positive=True):
    return torch.randn(1, CTX, synth.CFG_TINY_I2V["text_dim"], generator=torch.Generator().manual_seed(11 if positive else 12))


@pytest.mark.parametrize("cond_wo_pose", [False, True])
def test_dance_clip_matches_oracle(cond_wo_pose):
    from diffsynth import ModelManager, SVIDanceVideoPipeline
    from diffsynth.models.wan_video_dit import WanModel
    from diffsynth.models.wan_video_vae import WanVideoVAE
    from oracle import dwpose_oracle as DO
    from oracle import wan_dit_oracle as O
    from oracle import wan_vae_oracle as V
    cfg = synth.CFG_TINY_I2V
    bf = lambda t: t.to(torch.bfloat16).float()
    dit_sd = {k: bf(v) for k, v in synth.make_dit_state_dict(cfg, seed=2).items()}
    vae_sd = {k: bf(v) for k, v in synth_vae.make_vae_state_dict(seed=0).items()}
    stem_sd = _stem_sd(cfg["dim"], seed=7)
    dit = WanModel(**cfg).eval()
    dit.load_state_dict(dit_sd)
    vae = WanVideoVAE().eval()
    vae.load_state_dict(vae_sd)
    mm = ModelManager(torch_dtype=torch.bfloat16, device="cuda")
    mm.add_model("wan_video_dit", dit.to("cuda"))
    mm.add_model("wan_video_vae", vae.to("cuda"))
    mm.state_dict_new_module = {"pipe.dwpose_embedding." + k: v for k, v in stem_sd.items()}      # as load_lora_v2 leaves them
    pipe = SVIDanceVideoPipeline.from_model_manager(mm, torch_dtype=torch.bfloat16, device="cuda", is_test=True)
    assert pipe.dwpose_embedding is not None
    pipe.prompter = _prompter
    pipe.image_encoder = _ClipStub()
    img = Image.fromarray(np.random.default_rng(5).integers(0, 255, size=(H, W, 3), dtype=np.uint8))
    pose = _pose(FRAMES, H, W, seed=3)
    args = types.SimpleNamespace(ref_pad_cfg=False, ref_pad_num=-1, sequential_cfg="none")
    frames = pipe(prompt="p", negative_prompt="n", input_image=img, num_inference_steps=STEPS, cfg_scale={"text": 5.0}, seed=42,
                  tiled=False, random_ref_frame=torch.from_numpy(np.array(img)), height=H, width=W, num_frames=FRAMES,
                  humanpose_data=pose, cond_wo_pose=cond_wo_pose, args=args, progress_bar_cmd=lambda x: x)
    got = np.stack([np.array(f) for f in frames]).astype(np.float32)
    # ---- the same clip on the CPU oracles
    noise = bf(torch.randn((1, 16, (FRAMES - 1) // 4 + 1, H // 8, W // 8), generator=torch.Generator().manual_seed(42)))
    x = torch.from_numpy(np.array(img, dtype=np.float32) * (2 / 255) - 1).permute(2, 0, 1).unsqueeze(0)
    clip = bf(_ClipStub().encode_image([x]))
    msk = torch.zeros(1, FRAMES, H // 8, W // 8)
    msk[:, 0] = 1
    msk = torch.cat([torch.repeat_interleave(msk[:, 0:1], 4, dim=1), msk[:, 1:]], dim=1)
    msk = msk.view(1, msk.shape[1] // 4, 4, H // 8, W // 8).transpose(1, 2)[0]
    vae_in = torch.cat([x.transpose(0, 1), x.transpose(0, 1).repeat(1, FRAMES - 1, 1, 1)], dim=1)
    with torch.no_grad():
        y = bf(torch.cat([msk, V.vae_encode(vae_sd, vae_in.unsqueeze(0))[0]]).unsqueeze(0))
        cond = bf(DO.pose_condition(stem_sd, pose))
        sig = O.flow_match_sigmas(STEPS, 5.0)
        lat = noise
        for i in range(STEPS):
            ts = (sig[i] * 1000).reshape(1)
            vc = O.dit_forward(dit_sd, cfg, lat, ts, bf(_prompter("p", True)), clip, y, add_condition=cond)
            vu = O.dit_forward(dit_sd, cfg, lat, ts, bf(_prompter("n", False)), clip, y, add_condition=cond if cond_wo_pose else None)
            lat = O.flow_match_step(sig, i, O.cfg_combine(vc, vu, 5.0), lat)
        vid = V.vae_decode(vae_sd, lat)
    ref = ((vid[0].permute(1, 2, 3, 0) + 1) * 127.5).clip(0, 255).numpy().astype(np.uint8).astype(np.float32)
    diff = np.abs(got - ref)
    print(f"dance clip (cond_wo_pose={cond_wo_pose}) parity: mean {diff.mean():.3f} levels, p99 {np.percentile(diff, 99):.1f}")
    assert diff.mean() < 2.0 and np.percentile(diff, 99) < 12
