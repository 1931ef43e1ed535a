"""Deterministic synthetic weights / inputs of the two conditioning encoders (umT5 text encoder, CLIP ViT image encoder).

Shared by tests/golden/make_golden.py, the tests and tools/enc_bench.py.  Keys and shapes follow the reference's
parameter names: WanTextEncoder (diffsynth/models/wan_video_text_encoder.py:209-255) and the visual tower of
WanImageEncoder (diffsynth/models/wan_video_image_encoder.py:386-478,852-880; keys carry the `model.visual.` prefix the
reference's converter produces, :894-901).
"""
import torch

# tiny configs for goldens / fast tests: the head widths are the real ones (umT5: 64, CLIP ViT-H: 80)
TEXT_TINY = dict(vocab=97, dim=128, dim_attn=128, dim_ffn=320, num_heads=2, num_layers=2, num_buckets=32)
TEXT_UMT5_XXL = dict(vocab=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32)
CLIP_TINY = dict(image_size=42, patch_size=14, dim=160, mlp_ratio=4, out_dim=64, num_heads=2, num_layers=3)
CLIP_VIT_H = dict(image_size=224, patch_size=14, dim=1280, mlp_ratio=4, out_dim=1024, num_heads=16, num_layers=32)

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def text_param_shapes(cfg):
    d, da, f = cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"]
    sh = {"token_embedding.weight": (cfg["vocab"], d), "norm.weight": (d,)}
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}"
        sh[f"{p}.norm1.weight"] = (d,)
        for n in "qkv":
            sh[f"{p}.attn.{n}.weight"] = (da, d)
        sh[f"{p}.attn.o.weight"] = (d, da)
        sh[f"{p}.norm2.weight"] = (d,)
        sh[f"{p}.ffn.gate.0.weight"] = (f, d)
        sh[f"{p}.ffn.fc1.weight"] = (f, d)
        sh[f"{p}.ffn.fc2.weight"] = (d, f)
        sh[f"{p}.pos_embedding.embedding.weight"] = (cfg["num_buckets"], cfg["num_heads"])
    return sh


def clip_param_shapes(cfg, prefix="model.visual."):
    d, m = cfg["dim"], int(cfg["dim"] * cfg["mlp_ratio"])
    n_tok = (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
    sh = {"cls_embedding": (1, 1, d), "pos_embedding": (1, n_tok, d), "head": (d, cfg["out_dim"]),
          "patch_embedding.weight": (d, 3, cfg["patch_size"], cfg["patch_size"]),
          "pre_norm.weight": (d,), "pre_norm.bias": (d,), "post_norm.weight": (d,), "post_norm.bias": (d,)}
    for i in range(cfg["num_layers"]):
        p = f"transformer.{i}"
        for nm, shape in (("norm1.weight", (d,)), ("norm1.bias", (d,)), ("attn.to_qkv.weight", (3 * d, d)),
                          ("attn.to_qkv.bias", (3 * d,)), ("attn.proj.weight", (d, d)), ("attn.proj.bias", (d,)),
                          ("norm2.weight", (d,)), ("norm2.bias", (d,)), ("mlp.0.weight", (m, d)), ("mlp.0.bias", (m,)),
                          ("mlp.2.weight", (d, m)), ("mlp.2.bias", (d,))):
            sh[f"{p}.{nm}"] = shape
    return {prefix + k: v for k, v in sh.items()}


def _fill(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        if name.endswith("norm.weight") or ".norm1.weight" in name or ".norm2.weight" in name or "pre_norm.weight" in name:
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif "pos_embedding.embedding" in name:
            sd[name] = 0.5 * torch.randn(shape, generator=g)
        elif name.endswith("token_embedding.weight"):
            sd[name] = torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if name.endswith("cls_embedding") or name.endswith("pos_embedding"):
                fan_in = shape[-1]
            # T5 does not scale its attention scores: a q projection as large as the others would give score std ~17 and a
            # one-hot softmax that no trained model has; 0.25 keeps the scores at std ~3
            gain = 0.25 if name.endswith(".attn.q.weight") else 1.5
            sd[name] = torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)
        else:
            sd[name] = torch.randn(shape, generator=g)
    return sd


def make_text_state_dict(cfg, seed=0):
    return _fill(text_param_shapes(cfg), seed)


def make_clip_state_dict(cfg, seed=0, prefix="model.visual."):
    return _fill(clip_param_shapes(cfg, prefix), seed)


def make_text_inputs(cfg, seq_len=24, valid=17, seed=0):
    """ids int64 [1, seq_len] (pad id 0 past `valid`), mask int64 [1, seq_len]."""
    g = torch.Generator().manual_seed(1000 + seed)
    ids = torch.randint(2, cfg["vocab"], (1, seq_len), generator=g)
    mask = torch.zeros(1, seq_len, dtype=torch.int64)
    mask[:, :valid] = 1
    ids = ids * mask
    return ids, mask


def make_clip_image(height=50, width=70, seed=0):
    """image f32 [1, 3, H, W] in [-1, 1] (what the pipelines hand to encode_image)."""
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.rand(1, 3, height, width, generator=g) * 2 - 1
