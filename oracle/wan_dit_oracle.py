"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference's Wan-DiT forward.

Functional, state-dict driven restatement of ``model_fn_wan_video`` / ``WanModel`` in plain torch CPU ops
(fp32 by default, fp64 on request).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg may import this module; the product path
(``stable-video-infinity_b200/diffsynth``) never does.

Parity pinning: ``tests/golden/make_golden.py`` imports the real reference modules from /root/reference
(in the build container), runs them on seeded weights/inputs and stores the outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against those files.

Every function cites the reference lines it restates (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# Wan2.1-T2V-1.3B, reference diffsynth/models/wan_video_dit.py:657-669
CFG_T2V_1_3B = dict(has_image_input=False, patch_size=(1, 2, 2), in_dim=16, dim=1536, ffn_dim=8960,
                    freq_dim=256, text_dim=4096, out_dim=16, num_heads=12, num_layers=30, eps=1e-6)
# Wan2.1-I2V-14B, reference diffsynth/models/wan_video_dit.py:700-712
CFG_I2V_14B = dict(has_image_input=True, patch_size=(1, 2, 2), in_dim=36, dim=5120, ffn_dim=13824,
                   freq_dim=256, text_dim=4096, out_dim=16, num_heads=40, num_layers=40, eps=1e-6)


def sinusoidal_embedding(dim, position):
    """wan_video_dit.py:154-158 — fp64 outer product, [cos | sin], cast back to position dtype."""
    pos = position.to(torch.float64)
    inv = torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                    -torch.arange(dim // 2, dtype=torch.float64) / (dim // 2))
    ang = pos[:, None] * inv[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1).to(position.dtype)


def rope_angles_1d(dim, end=1024, theta=10000.0):
    """wan_video_dit.py:169-175 — angles whose polar form is freqs_cis (fp64)."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].double() / dim))
    return torch.outer(torch.arange(end, dtype=torch.float64), inv)  # [end, dim/2]


def rope_angles_3d(head_dim, f, h, w):
    """wan_video_dit.py:161-166 + svi_video.py:106-110 — per-token angle table [f*h*w, head_dim/2] (fp64).

    The head_dim/2 complex pairs are split (frame | height | width) = (hd/2 - 2*(hd//3)/2 ... ) exactly as
    precompute_freqs_cis_3d does: dims (hd - 2*(hd//3), hd//3, hd//3) -> 22 | 21 | 21 pairs for hd=128.
    """
    d_h = head_dim // 3
    d_f = head_dim - 2 * d_h
    af, ah, aw = rope_angles_1d(d_f), rope_angles_1d(d_h), rope_angles_1d(d_h)
    ang = torch.cat([af[:f].view(f, 1, 1, -1).expand(f, h, w, -1),
                     ah[:h].view(1, h, 1, -1).expand(f, h, w, -1),
                     aw[:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1)
    return ang.reshape(f * h * w, -1)


def rope_apply(x, angles, num_heads):
    """wan_video_dit.py:178-183 — interleaved pairs (x[2i], x[2i+1]) rotated in fp64. x: [B, L, H*hd]."""
    B, L, D = x.shape
    xc = x.to(torch.float64).reshape(B, L, num_heads, -1, 2)
    c, s = angles.cos()[None, :, None, :], angles.sin()[None, :, None, :]
    re = xc[..., 0] * c - xc[..., 1] * s
    im = xc[..., 0] * s + xc[..., 1] * c
    return torch.stack([re, im], dim=-1).reshape(B, L, D).to(x.dtype)


def rms_norm(x, weight, eps):
    """wan_video_dit.py:186-197 — over the FULL last dim (all heads jointly), fp32 internal."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).to(x.dtype) * weight


def attention(q, k, v, num_heads):
    """wan_video_dit.py:141-146 (SDPA branch) — non-causal softmax(q k^T / sqrt(hd)) v."""
    B, Lq, D = q.shape
    hd = D // num_heads
    qh = q.view(B, Lq, num_heads, hd).transpose(1, 2)
    kh = k.view(B, -1, num_heads, hd).transpose(1, 2)
    vh = v.view(B, -1, num_heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh)
    return o.transpose(1, 2).reshape(B, Lq, D)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def self_attention(sd, pre, x, angles, H, eps):
    """wan_video_dit.py:226-242."""
    q = rms_norm(_lin(sd, pre + ".q", x), sd[pre + ".norm_q.weight"], eps)
    k = rms_norm(_lin(sd, pre + ".k", x), sd[pre + ".norm_k.weight"], eps)
    v = _lin(sd, pre + ".v", x)
    q = rope_apply(q, angles, H)
    k = rope_apply(k, angles, H)
    return _lin(sd, pre + ".o", attention(q, k, v, H))


def cross_attention(sd, pre, x, ctx, H, eps, has_image_input):
    """wan_video_dit.py:266-303 — first 257 context rows are CLIP image tokens when has_image_input."""
    if has_image_input:
        img, ctx = ctx[:, :257], ctx[:, 257:]
    q = rms_norm(_lin(sd, pre + ".q", x), sd[pre + ".norm_q.weight"], eps)
    k = rms_norm(_lin(sd, pre + ".k", ctx), sd[pre + ".norm_k.weight"], eps)
    v = _lin(sd, pre + ".v", ctx)
    o = attention(q, k, v, H)
    if has_image_input:
        k_img = rms_norm(_lin(sd, pre + ".k_img", img), sd[pre + ".norm_k_img.weight"], eps)
        v_img = _lin(sd, pre + ".v_img", img)
        o = o + attention(q, k_img, v_img, H)
    return _lin(sd, pre + ".o", o)


def preprocess_audio(audio_embed, audio_window=5, vae_scale=4):
    """svi_video_talk.py:432-446: [1, 4n+1, 5, 12, 768] window features -> (first frame [1,1,5,12,768],
    latter frames [1, n, 8, 12, 768]): per latent frame the first video frame keeps its left half window + centre, the
    last its centre + right half, the middle ones their centre."""
    first = audio_embed[:, :1]
    rest = audio_embed[:, 1:]
    b, n4, w, s_, c = rest.shape
    rest = rest.reshape(b, n4 // vae_scale, vae_scale, w, s_, c)
    mid = audio_window // 2
    a = rest[:, :, :1, :mid + 1].reshape(b, n4 // vae_scale, -1, s_, c)
    z = rest[:, :, -1:, mid:].reshape(b, n4 // vae_scale, -1, s_, c)
    m = rest[:, :, 1:-1, mid:mid + 1].reshape(b, n4 // vae_scale, -1, s_, c)
    return first, torch.cat([a, m, z], dim=2)


def audio_proj(sd, first, latter):
    """AudioProjModel.forward wan_video_dit.py:82-112 -> audio tokens [n_latent_frames, 32, 768]."""
    x0 = F.relu(_lin(sd, "audio_proj.proj1", first.reshape(first.shape[1], -1)))
    x1 = F.relu(_lin(sd, "audio_proj.proj1_vf", latter.reshape(latter.shape[1], -1)))
    x = F.relu(_lin(sd, "audio_proj.proj2", torch.cat([x0, x1], dim=0)))
    tok = _lin(sd, "audio_proj.proj3", x).reshape(x.shape[0], 32, 768)
    return F.layer_norm(tok, (768,), sd["audio_proj.norm.weight"], sd["audio_proj.norm.bias"])


def audio_cross_attention(sd, pre, x, audio_tokens, H):
    """SingleStreamAttention.forward models/attention.py:318-371 (human_num = 1, qk_norm off): every latent frame's
    tokens attend to that frame's 32 audio tokens; scale 1/sqrt(head_dim)."""
    n_t = audio_tokens.shape[0]
    B, L, d = x.shape
    xf = x.reshape(n_t, L // n_t, d)
    q = _lin(sd, pre + ".q_linear", xf).view(n_t, -1, H, d // H).transpose(1, 2)
    kv = _lin(sd, pre + ".kv_linear", audio_tokens).view(n_t, -1, 2, H, d // H)
    k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n_t, -1, d)
    return _lin(sd, pre + ".proj", o).reshape(B, L, d)


def dit_block(sd, i, x, ctx, t_mod, angles, cfg, audio_tokens=None):
    """wan_video_dit.py:354-374; audio_tokens [n_t, 32, 768] enables the multitalk branch (:361-366)."""
    pre = f"blocks.{i}"
    d, H, eps = cfg["dim"], cfg["num_heads"], cfg["eps"]
    mod = sd[pre + ".modulation"].to(t_mod.dtype) + t_mod  # [1,6,d]
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
    h = F.layer_norm(x, (d,), eps=eps) * (1 + sc_a) + sh_a
    x = x + g_a * self_attention(sd, pre + ".self_attn", h, angles, H, eps)
    h = F.layer_norm(x, (d,), sd[pre + ".norm3.weight"], sd[pre + ".norm3.bias"], eps)
    x = x + cross_attention(sd, pre + ".cross_attn", h, ctx, H, eps, cfg["has_image_input"])
    if audio_tokens is not None:
        h = F.layer_norm(x, (d,), sd[pre + ".norm_x.weight"], sd[pre + ".norm_x.bias"], eps)
        x = x + audio_cross_attention(sd, pre + ".audio_cross_attn", h, audio_tokens, H)
    h = F.layer_norm(x, (d,), eps=eps) * (1 + sc_m) + sh_m
    h = _lin(sd, pre + ".ffn.2", F.gelu(_lin(sd, pre + ".ffn.0", h), approximate="tanh"))
    return x + g_m * h


def embed_conditions(sd, cfg, timestep, context, clip_feature, dtype):
    """svi_video.py:90-97 — time MLP, time projection, text MLP, CLIP MLP (wan_video_dit.py:377-389)."""
    d = cfg["dim"]
    te = sinusoidal_embedding(cfg["freq_dim"], timestep).to(dtype)
    t = _lin(sd, "time_embedding.2", F.silu(_lin(sd, "time_embedding.0", te)))
    t_mod = _lin(sd, "time_projection.1", F.silu(t)).unflatten(1, (6, d))
    ctx = _lin(sd, "text_embedding.2", F.gelu(_lin(sd, "text_embedding.0", context), approximate="tanh"))
    if cfg["has_image_input"]:
        c = F.layer_norm(clip_feature, (clip_feature.shape[-1],), sd["img_emb.proj.0.weight"], sd["img_emb.proj.0.bias"])
        c = _lin(sd, "img_emb.proj.3", F.gelu(_lin(sd, "img_emb.proj.1", c)))
        c = F.layer_norm(c, (d,), sd["img_emb.proj.4.weight"], sd["img_emb.proj.4.bias"])
        ctx = torch.cat([c, ctx], dim=1)
    return t, t_mod, ctx


def patchify(sd, cfg, x):
    """wan_video_dit.py:473-477 — Conv3d k=s=patch_size then 'b c f h w -> b (f h w) c'."""
    x = F.conv3d(x, sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=cfg["patch_size"])
    f, h, w = x.shape[2:]
    return x.flatten(2).transpose(1, 2).contiguous(), (f, h, w)


def head_unpatchify(sd, cfg, x, t, grid):
    """wan_video_dit.py:401-404 (Head) and :479-484 (unpatchify)."""
    d, eps = cfg["dim"], cfg["eps"]
    mod = sd["head.modulation"].to(t.dtype) + t.unsqueeze(1)  # [1,2,d] + [B,1,d]
    shift, scale = mod.chunk(2, dim=1)
    x = _lin(sd, "head.head", F.layer_norm(x, (d,), eps=eps) * (1 + scale) + shift)
    f, h, w = grid
    pt, ph, pw = cfg["patch_size"]
    B = x.shape[0]
    x = x.view(B, f, h, w, pt, ph, pw, cfg["out_dim"])
    return x.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, cfg["out_dim"], f * pt, h * ph, w * pw)


class TeaCacheOracle:
    """svi_video.py:23-72 restated: accumulate poly(rel-L1 change of t_mod) and skip the block stack while the sum stays
    below the threshold; first and last step of a clip always compute.  Pinned to the reference class by
    tests/golden/teacache.npz (tests/test_oracle_golden.py)."""

    COEFFICIENTS = {   # svi_video.py:34-39 (fitted constants of the TeaCache release), highest power first
        "Wan2.1-T2V-1.3B": [-5.21862437e+04, 9.23041404e+03, -5.28275948e+02, 1.36987616e+01, -4.99875664e-02],
        "Wan2.1-T2V-14B": [-3.03318725e+05, 4.90537029e+04, -2.65530556e+03, 5.87365115e+01, -3.15583525e-01],
        "Wan2.1-I2V-14B-480P": [2.57151496e+05, -3.54229917e+04, 1.40286849e+03, -1.35890334e+01, 1.32517977e-01],
        "Wan2.1-I2V-14B-720P": [8.10705460e+03, 2.13393892e+03, -3.72934672e+02, 1.66203073e+01, -4.17769401e-02],
    }

    def __init__(self, num_inference_steps, rel_l1_thresh, model_id):
        self.n, self.thresh = num_inference_steps, rel_l1_thresh
        self.poly = np.poly1d(self.COEFFICIENTS[model_id])
        self.step, self.acc = 0, 0.0
        self.prev_mod = self.prev_in = self.residual = None
        self.skipped = []

    def check(self, x, t_mod):            # :41-65
        if self.step == 0 or self.step == self.n - 1:
            compute, self.acc = True, 0.0
        else:
            rel = ((t_mod - self.prev_mod).abs().mean() / self.prev_mod.abs().mean()).item()
            self.acc += float(self.poly(rel))
            compute = not (self.acc < self.thresh)
            if compute:
                self.acc = 0.0
        self.prev_mod = t_mod.clone()
        if not compute:
            self.skipped.append(self.step)
        self.step = (self.step + 1) % self.n
        if compute:
            self.prev_in = x.clone()
        return not compute

    def store(self, x):                   # :67-69
        self.residual = x - self.prev_in
        self.prev_in = None

    def update(self, x):                  # :71-72
        return x + self.residual


def dit_forward(sd, cfg, x, timestep, context, clip_feature=None, y=None, dtype=torch.float32, return_tokens=False,
                tea_cache=None, add_condition=None, audio_embed_tuple=None):
    """svi_video.py:74-137 (model_fn_wan_video; no USP) on CPU in `dtype`; tea_cache: a TeaCacheOracle or None;
    add_condition: [1, L, dim] added to the patch embedding (:102-103); audio_embed_tuple: the (first, latter) window
    features of preprocess_audio for an enable_multitalk model (svi_video_talk.py:123-124)."""
    sd = {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}
    x = x.to(dtype)
    context = context.to(dtype)
    t, t_mod, ctx = embed_conditions(sd, cfg, timestep.to(dtype), context,
                                     None if clip_feature is None else clip_feature.to(dtype), dtype)
    if cfg["has_image_input"]:
        x = torch.cat([x, y.to(dtype)], dim=1)
    tok, (f, h, w) = patchify(sd, cfg, x)
    if add_condition is not None:
        tok = add_condition.to(dtype) + tok
    angles = rope_angles_3d(cfg["dim"] // cfg["num_heads"], f, h, w)
    audio_tokens = None
    if audio_embed_tuple is not None:
        audio_tokens = audio_proj(sd, audio_embed_tuple[0].to(dtype), audio_embed_tuple[1].to(dtype))
    if tea_cache is not None and tea_cache.check(tok, t_mod):          # :114-126
        tok = tea_cache.update(tok)
    else:
        for i in range(cfg["num_layers"]):
            tok = dit_block(sd, i, tok, ctx, t_mod, angles, cfg, audio_tokens)
        if tea_cache is not None:
            tea_cache.store(tok)
    if return_tokens:
        return tok
    return head_unpatchify(sd, cfg, tok, t, (f, h, w))


# ---------------------------------------------------------------------------------------------------
# flow-matching scheduler (reference diffsynth/schedulers/flow_match.py)
# ---------------------------------------------------------------------------------------------------
def flow_match_sigmas(num_inference_steps, shift=5.0, sigma_max=1.0, sigma_min=0.0, denoising_strength=1.0,
                      extra_one_step=True):
    """flow_match.py:31-44 — SVI uses shift=5, sigma_min=0, extra_one_step=True (svi_video.py:144)."""
    start = sigma_min + (sigma_max - sigma_min) * denoising_strength
    if extra_one_step:
        s = torch.linspace(start, sigma_min, num_inference_steps + 1)[:-1]
    else:
        s = torch.linspace(start, sigma_min, num_inference_steps)
    return shift * s / (1 + (shift - 1) * s)


def flow_match_step(sigmas, i, model_output, sample):
    """flow_match.py:53-64 — Euler update; the last step integrates to sigma = 0."""
    sigma = sigmas[i]
    sigma_next = sigmas[i + 1] if i + 1 < len(sigmas) else 0.0
    return sample + model_output * (sigma_next - sigma)


def cfg_combine(v_cond, v_uncond, scale):
    """svi_video.py:410."""
    return v_uncond + scale * (v_cond - v_uncond)


def denoise(sd, cfg, latents, ctx_pos, ctx_neg, steps, cfg_scale=5.0, shift=5.0, num_train_timesteps=1000,
            clip_feature=None, y=None, dtype=torch.float32, tea_cache_l1_thresh=None, tea_cache_model_id=""):
    """svi_video.py:392-421 — the CFG + Euler loop (test-size only: every step is 2 full DiT forwards).
    tea_cache_l1_thresh: one TeaCacheOracle per CFG branch as in svi_video.py:500-501."""
    sig = flow_match_sigmas(steps, shift)
    x = latents.to(dtype)
    mk = lambda: TeaCacheOracle(steps, tea_cache_l1_thresh, tea_cache_model_id) if tea_cache_l1_thresh is not None else None
    tc_pos, tc_neg = mk(), mk()
    for i in range(steps):
        ts = (sig[i] * num_train_timesteps).reshape(1)
        vc = dit_forward(sd, cfg, x, ts, ctx_pos, clip_feature, y, dtype, tea_cache=tc_pos)
        if cfg_scale != 1.0:
            vu = dit_forward(sd, cfg, x, ts, ctx_neg, clip_feature, y, dtype, tea_cache=tc_neg)
            vc = cfg_combine(vc, vu, cfg_scale)
        x = flow_match_step(sig, i, vc, x)
    return x


def dit_forward_flops(cfg, L, Lc=512):
    """Algorithmic FLOPs of one forward (multiply-add = 2), SURVEY.md §8(d) / BASELINE.md §3."""
    d, ffn, n = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    blk = 8 * L * d * d + 4 * L * L * d + (4 * L * d * d + 4 * Lc * d * d + 4 * L * Lc * d) + 4 * L * d * ffn
    if cfg["has_image_input"]:
        blk += 4 * 257 * d * d + 4 * L * 257 * d
    pre = 2 * L * (4 * cfg["in_dim"]) * d + 2 * Lc * cfg["text_dim"] * d + 2 * Lc * d * d + 2 * L * d * 64
    return n * blk + pre
