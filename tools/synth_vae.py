"""Deterministic synthetic weights for the Wan 3-D VAE (key names / shapes of the reference's WanVideoVAE,
``wan_video_vae.py:599-808``; the unprefixed key set hashes to the reference's checkpoint fingerprint
``ccc42284ea13e1ad04693284c7a09be6``, model_config.py:125)."""
import math

import torch

ENC_RES = {0: (96, 96), 1: (96, 96), 3: (96, 192), 4: (192, 192), 6: (192, 384), 7: (384, 384), 9: (384, 384), 10: (384, 384)}
ENC_DOWN = {2: (96, False), 5: (192, True), 8: (384, True)}
DEC_RES = {0: (384, 384), 1: (384, 384), 2: (384, 384), 4: (192, 384), 5: (384, 384), 6: (384, 384),
           8: (192, 192), 9: (192, 192), 10: (192, 192), 12: (96, 96), 13: (96, 96), 14: (96, 96)}
DEC_UP = {3: (384, True), 7: (384, True), 11: (192, False)}


def vae_param_shapes(prefix="model."):
    sh = {}

    def conv(name, o, i, *k):
        sh[name + ".weight"] = (o, i, *k)
        sh[name + ".bias"] = (o,)

    def res(name, i, o):
        sh[name + ".residual.0.gamma"] = (i, 1, 1, 1)
        conv(name + ".residual.2", o, i, 3, 3, 3)
        sh[name + ".residual.3.gamma"] = (o, 1, 1, 1)
        conv(name + ".residual.6", o, o, 3, 3, 3)
        if i != o:
            conv(name + ".shortcut", o, i, 1, 1, 1)

    def attn(name, c):
        sh[name + ".norm.gamma"] = (c, 1, 1)
        conv(name + ".to_qkv", 3 * c, c, 1, 1)
        conv(name + ".proj", c, c, 1, 1)

    def middle(name, c):
        res(name + ".0", c, c)
        attn(name + ".1", c)
        res(name + ".2", c, c)

    e = "encoder"
    conv(e + ".conv1", 96, 3, 3, 3, 3)
    for idx in range(11):
        n = f"{e}.downsamples.{idx}"
        if idx in ENC_RES:
            res(n, *ENC_RES[idx])
        else:
            c, temporal = ENC_DOWN[idx]
            conv(n + ".resample.1", c, c, 3, 3)
            if temporal:
                conv(n + ".time_conv", c, c, 3, 1, 1)
    middle(e + ".middle", 384)
    sh[e + ".head.0.gamma"] = (384, 1, 1, 1)
    conv(e + ".head.2", 32, 384, 3, 3, 3)
    conv("conv1", 32, 32, 1, 1, 1)
    conv("conv2", 16, 16, 1, 1, 1)
    d = "decoder"
    conv(d + ".conv1", 384, 16, 3, 3, 3)
    middle(d + ".middle", 384)
    for idx in range(15):
        n = f"{d}.upsamples.{idx}"
        if idx in DEC_RES:
            res(n, *DEC_RES[idx])
        else:
            c, temporal = DEC_UP[idx]
            conv(n + ".resample.1", c // 2, c, 3, 3)
            if temporal:
                conv(n + ".time_conv", 2 * c, c, 3, 1, 1)
    sh[d + ".head.0.gamma"] = (96, 1, 1, 1)
    conv(d + ".head.2", 3, 96, 3, 3, 3)
    return {prefix + k: v for k, v in sh.items()}


def make_vae_state_dict(seed=0, device="cpu", dtype=torch.float32, gain=1.0):
    """conv weights ~ N(0, gain/fan_in), gamma ~ 1 + 0.1 N, bias ~ 0.02 N.  The attention `proj` weights are
    NON-zero (the reference zero-initialises them, wan_video_vae.py:250, which would leave attention untested)."""
    sd = {}
    for idx, (name, shape) in enumerate(vae_param_shapes().items()):
        g = torch.Generator(device="cpu").manual_seed(777 + seed * 100003 + idx)
        if name.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g) * math.sqrt(gain / math.prod(shape[1:]))
        sd[name] = t.to(device=device, dtype=dtype)
    return sd
