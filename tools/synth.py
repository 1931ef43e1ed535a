"""Deterministic synthetic weights / inputs shared by the golden generator, the tests, smoke() and bench.py.

No checkpoints exist in the build or GPU containers, so every parity and benchmark run uses random-init
weights of the real architecture.  Generation is pure CPU torch.Generator arithmetic (identical on the build
box and the GPU box); keys and shapes follow the reference's WanModel parameter names
(reference diffsynth/models/wan_video_dit.py:408-452).
"""
import math

import torch

# tiny configs for goldens / fast tests (head_dim is always 128 like the real models)
CFG_TINY_T2V = dict(has_image_input=False, patch_size=(1, 2, 2), in_dim=16, dim=256, ffn_dim=512,
                    freq_dim=64, text_dim=96, out_dim=16, num_heads=2, num_layers=2, eps=1e-6)
CFG_TINY_I2V = dict(has_image_input=True, patch_size=(1, 2, 2), in_dim=36, dim=256, ffn_dim=512,
                    freq_dim=64, text_dim=96, out_dim=16, num_heads=2, num_layers=2, eps=1e-6)
CFG_TINY_TALK = dict(CFG_TINY_I2V, enable_multitalk=True)
CFG_T2V_1_3B = dict(has_image_input=False, patch_size=(1, 2, 2), in_dim=16, dim=1536, ffn_dim=8960,
                    freq_dim=256, text_dim=4096, out_dim=16, num_heads=12, num_layers=30, eps=1e-6)
CFG_I2V_14B = dict(has_image_input=True, patch_size=(1, 2, 2), in_dim=36, dim=5120, ffn_dim=13824,
                   freq_dim=256, text_dim=4096, out_dim=16, num_heads=40, num_layers=40, eps=1e-6)


def dit_param_shapes(cfg):
    d, ffn, L = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    sh = {}

    def lin(name, o, i):
        sh[name + ".weight"] = (o, i)
        sh[name + ".bias"] = (o,)

    sh["patch_embedding.weight"] = (d, cfg["in_dim"], *cfg["patch_size"])
    sh["patch_embedding.bias"] = (d,)
    lin("text_embedding.0", d, cfg["text_dim"])
    lin("text_embedding.2", d, d)
    lin("time_embedding.0", d, cfg["freq_dim"])
    lin("time_embedding.2", d, d)
    lin("time_projection.1", 6 * d, d)
    for i in range(L):
        p = f"blocks.{i}"
        for att in ("self_attn", "cross_attn"):
            for n in "qkvo":
                lin(f"{p}.{att}.{n}", d, d)
            sh[f"{p}.{att}.norm_q.weight"] = (d,)
            sh[f"{p}.{att}.norm_k.weight"] = (d,)
        if cfg["has_image_input"]:
            lin(f"{p}.cross_attn.k_img", d, d)
            lin(f"{p}.cross_attn.v_img", d, d)
            sh[f"{p}.cross_attn.norm_k_img.weight"] = (d,)
        sh[f"{p}.norm3.weight"] = (d,)
        sh[f"{p}.norm3.bias"] = (d,)
        lin(f"{p}.ffn.0", ffn, d)
        lin(f"{p}.ffn.2", d, ffn)
        sh[f"{p}.modulation"] = (1, 6, d)
        if cfg.get("enable_multitalk"):          # SVI-Talk audio cross-attention (wan_video_dit.py:339-352)
            lin(f"{p}.audio_cross_attn.q_linear", d, d)
            lin(f"{p}.audio_cross_attn.proj", d, d)
            lin(f"{p}.audio_cross_attn.kv_linear", 2 * d, 768)
            sh[f"{p}.norm_x.weight"] = (d,)
            sh[f"{p}.norm_x.bias"] = (d,)
    lin("head.head", cfg["out_dim"] * math.prod(cfg["patch_size"]), d)
    sh["head.modulation"] = (1, 2, d)
    if cfg["has_image_input"]:
        sh["img_emb.proj.0.weight"] = (1280,)
        sh["img_emb.proj.0.bias"] = (1280,)
        lin("img_emb.proj.1", 1280, 1280)
        lin("img_emb.proj.3", d, 1280)
        sh["img_emb.proj.4.weight"] = (d,)
        sh["img_emb.proj.4.bias"] = (d,)
    if cfg.get("enable_multitalk"):              # AudioProjModel (wan_video_dit.py:455-470: windows 5 / 8, 12 x 768 wav2vec blocks)
        lin("audio_proj.proj1", 512, 5 * 12 * 768)
        lin("audio_proj.proj1_vf", 512, 8 * 12 * 768)
        lin("audio_proj.proj2", 512, 512)
        lin("audio_proj.proj3", 32 * 768, 512)
        sh["audio_proj.norm.weight"] = (768,)
        sh["audio_proj.norm.bias"] = (768,)
    return sh


def make_dit_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32, init="normal"):
    """Seeded random-init state dict.  init="normal": matrices ~ N(0, 1/fan_in), norm weights ~ 1 + 0.1 N, biases ~ 0.02 N,
    modulation ~ N(0,1)/sqrt(dim) (as wan_video_dit.py:336,399).  init="torch_default": the distribution the reference's
    constructors produce (nn.Linear / nn.Conv3d: weights and biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); norm weights 1,
    norm biases 0) — a sqrt(3) smaller weight std, i.e. the random-init model SURVEY.md section 7 measured its parity bars on.
    Per-tensor generators keyed by name order so that a subset (e.g. fewer layers) is reproducible."""
    sd = {}
    shapes = dit_param_shapes(cfg)
    for idx, (name, shape) in enumerate(shapes.items()):
        g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + idx)
        if init == "torch_default" and not name.endswith("modulation"):
            if "norm" in name or name.startswith("img_emb.proj.0") or name.startswith("img_emb.proj.4"):
                t = torch.ones(shape) if name.endswith(".weight") else torch.zeros(shape)
            else:
                wshape = shape if name.endswith(".weight") else shapes[name[:-5] + ".weight"]
                bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith("modulation"):
            t = torch.randn(shape, generator=g) / cfg["dim"] ** 0.5
        elif "norm" in name and name.endswith(".weight") or name in ("img_emb.proj.0.weight", "img_emb.proj.4.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:])
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[name] = t.to(device=device, dtype=dtype)
    return sd


def make_audio_embed(num_frames, seed=0, scale=1.0):
    """wav2vec-style window features of a clip: [1, num_frames, 5, 12, 768] (svi_video_talk.py:412-430), seeded."""
    g = torch.Generator(device="cpu").manual_seed(4321 + seed)
    return torch.randn(1, num_frames, 5, 12, 768, generator=g) * scale


def make_dit_state_dict_fast(cfg, seed=0, device="cuda", dtype=torch.bfloat16):
    """Same distribution as make_dit_state_dict but generated on `device` (for 1.3B / 14B benchmark models,
    where CPU generation + upload would take minutes).  NOT bit-identical to the CPU version."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, shape in dit_param_shapes(cfg).items():
        if name.endswith("modulation"):
            t = torch.randn(shape, generator=g, device=device) / cfg["dim"] ** 0.5
        elif "norm" in name and name.endswith(".weight") or name in ("img_emb.proj.0.weight", "img_emb.proj.4.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        else:
            t = torch.randn(shape, generator=g, device=device) / math.sqrt(math.prod(shape[1:]))
        sd[name] = t.to(dtype)
    return sd


def make_dit_inputs(cfg, f, h, w, seed=0, ctx_len=512, zero_ctx_after=None):
    """Seeded latents [1,16,f,h,w], context [1,ctx_len,text_dim], and (I2V) y [1,20,f,h,w] with the SVI mask
    pattern (svi_video.py:319-326) + clip_feature [1,257,1280].  h, w are LATENT sizes (even)."""
    g = torch.Generator(device="cpu").manual_seed(1234 + seed)
    out = {"x": torch.randn(1, 16, f, h, w, generator=g),
           "context": torch.randn(1, ctx_len, cfg["text_dim"], generator=g)}
    if zero_ctx_after is not None:  # real umT5 embeddings are zero-filled past the prompt (wan_prompter.py:107-108)
        out["context"][:, zero_ctx_after:] = 0
    if cfg["has_image_input"]:
        msk = torch.zeros(1, 4, f, h, w)
        msk[:, :, 0] = 1
        out["y"] = torch.cat([msk, torch.randn(1, 16, f, h, w, generator=g)], dim=1)
        out["clip_feature"] = torch.randn(1, 257, 1280, generator=g)
    return out
