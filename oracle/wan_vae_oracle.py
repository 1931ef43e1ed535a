"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference's Wan 3-D causal VAE.

Streaming formulation: the reference walks a list of cache slots with a shared mutable counter
(``wan_video_vae.py:198-232, 328-376, 432-481, 525-596``).  Unrolling that protocol shows every cached
``CausalConv3d`` is a causal convolution over the frame STREAM with two frames of zero-initialised history
(chunk 0: zero pad 2; chunk 1: [0, x0, x1]; later: last two input frames), with these exceptions:

* decoder ``upsample3d`` (``:122-156``): the first chunk bypasses ``time_conv`` (sentinel 'Rep'), so its
  history starts EMPTY at chunk 1 (x0 is not part of it) and the first latent frame yields 1 pixel frame;
* encoder ``downsample3d`` (``:162-173``): the first chunk bypasses ``time_conv``; later chunks run a
  stride-2 (3,1,1) conv over [last frame of the previous chunk, x] (history = 1 frame, including x0);
* ``shortcut`` 1x1x1 convs and the per-frame attention block carry no history.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
Pinned against the real reference by tests/golden/vae_tiny.npz (tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F

LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]   # wan_video_vae.py:604-607
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]       # :608-611


class Stream:
    """Per-clip streaming state: name -> history tensor; absent = zero history / first chunk."""

    def __init__(self):
        self.hist = {}
        self.chunk = 0


def causal_conv(sd, name, x, st, pad_hw=1):
    """CausalConv3d with k_t = 3 (wan_video_vae.py:33-52) as a streaming conv: y = conv(cat[hist(2), x])."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    h = st.hist.get(name)
    if h is None:
        h = x.new_zeros(x.shape[0], x.shape[1], 2, *x.shape[3:])
    xin = torch.cat([h, x], dim=2)
    st.hist[name] = xin[:, :, -2:].clone()
    return F.conv3d(F.pad(xin, (pad_hw, pad_hw, pad_hw, pad_hw, 0, 0)), w, b)


def conv_1x1x1(sd, name, x):
    return F.conv3d(x, sd[name + ".weight"], sd[name + ".bias"])


def rms_norm(x, gamma):
    """RMS_norm (wan_video_vae.py:55-70): F.normalize over channels * sqrt(C) * gamma."""
    c = x.shape[1]
    return F.normalize(x, dim=1) * (c ** 0.5) * gamma.reshape(1, c, *([1] * (x.dim() - 2)))


def residual_block(sd, pre, x, st):
    """ResidualBlock (:198-232): shortcut(x) + conv(silu(norm(conv(silu(norm(x))))))."""
    h = conv_1x1x1(sd, pre + ".shortcut", x) if (pre + ".shortcut.weight") in sd else x
    y = causal_conv(sd, pre + ".residual.2", F.silu(rms_norm(x, sd[pre + ".residual.0.gamma"])), st)
    y = causal_conv(sd, pre + ".residual.6", F.silu(rms_norm(y, sd[pre + ".residual.3.gamma"])), st)
    return y + h


def attention_block(sd, pre, x):
    """AttentionBlock (:235-273): per-frame single-head attention over h*w tokens, scale 1/sqrt(C)."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm(y, sd[pre + ".norm.gamma"])
    qkv = F.conv2d(y, sd[pre + ".to_qkv.weight"], sd[pre + ".to_qkv.bias"]).reshape(b * t, 3 * c, h * w)
    q, k, v = (u.transpose(1, 2) for u in qkv.chunk(3, dim=1))           # [bt, hw, c]
    a = torch.softmax(q @ k.transpose(1, 2) / c ** 0.5, dim=-1) @ v
    y = a.transpose(1, 2).reshape(b * t, c, h, w)
    y = F.conv2d(y, sd[pre + ".proj.weight"], sd[pre + ".proj.bias"])
    return y.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x


def _per_frame(x, fn):
    b, c, t, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def upsample(sd, pre, x, st, temporal):
    """Resample upsample2d/3d (:82-156): [time_conv doubling T by channel split] + nearest-exact x2 + Conv2d 3x3."""
    if temporal:
        if st.chunk > 0:                       # chunk 0: 'Rep' sentinel, time_conv skipped
            b, c, t, h, w = x.shape
            y = causal_conv(sd, pre + ".time_conv", x, st, pad_hw=0)          # [b, 2c, t, h, w]
            y = y.reshape(b, 2, c, t, h, w)
            x = torch.stack((y[:, 0], y[:, 1]), dim=3).reshape(b, c, 2 * t, h, w)
    return _per_frame(x, lambda f: F.conv2d(F.interpolate(f, scale_factor=2.0, mode="nearest-exact"),
                                            sd[pre + ".resample.1.weight"], sd[pre + ".resample.1.bias"], padding=1))


def downsample(sd, pre, x, st, temporal):
    """Resample downsample2d/3d (:108-173): ZeroPad2d((0,1,0,1)) + Conv2d 3x3 s2 [+ stride-2 (3,1,1) time conv]."""
    x = _per_frame(x, lambda f: F.conv2d(F.pad(f, (0, 1, 0, 1)), sd[pre + ".resample.1.weight"],
                                         sd[pre + ".resample.1.bias"], stride=2))
    if temporal:
        key = pre + ".time_conv"
        last = st.hist.get(key)
        st.hist[key] = x[:, :, -1:].clone()
        if last is not None:                   # first chunk: time_conv skipped
            x = F.conv3d(torch.cat([last, x], dim=2), sd[key + ".weight"], sd[key + ".bias"], stride=(2, 1, 1))
    return x


# channel plans (wan_video_vae.py:276-376 / :379-481 with dim=96, dim_mult=[1,2,4,4], num_res_blocks=2,
# temperal_downsample=[False, True, True]); entries: ("res", index) | ("down"/"up", index, temporal)
ENC_PLAN = [("res", 0), ("res", 1), ("down", 2, False), ("res", 3), ("res", 4), ("down", 5, True),
            ("res", 6), ("res", 7), ("down", 8, True), ("res", 9), ("res", 10)]
DEC_PLAN = [("res", 0), ("res", 1), ("res", 2), ("up", 3, True), ("res", 4), ("res", 5), ("res", 6), ("up", 7, True),
            ("res", 8), ("res", 9), ("res", 10), ("up", 11, False), ("res", 12), ("res", 13), ("res", 14)]


def encoder_chunk(sd, x, st):
    """Encoder3d.forward (:328-376) for one chunk [1,3,T,H,W] (T = 1 for the first chunk, then 4)."""
    p = "model.encoder"
    x = causal_conv(sd, p + ".conv1", x, st)
    for item in ENC_PLAN:
        n = f"{p}.downsamples.{item[1]}"
        x = residual_block(sd, n, x, st) if item[0] == "res" else downsample(sd, n, x, st, item[2])
    x = residual_block(sd, p + ".middle.0", x, st)
    x = attention_block(sd, p + ".middle.1", x)
    x = residual_block(sd, p + ".middle.2", x, st)
    x = causal_conv(sd, p + ".head.2", F.silu(rms_norm(x, sd[p + ".head.0.gamma"])), st)
    st.chunk += 1
    return x


def decoder_chunk(sd, x, st):
    """Decoder3d.forward (:432-481) for one latent frame [1,16,1,h,w] (already through conv2)."""
    p = "model.decoder"
    x = causal_conv(sd, p + ".conv1", x, st)
    x = residual_block(sd, p + ".middle.0", x, st)
    x = attention_block(sd, p + ".middle.1", x)
    x = residual_block(sd, p + ".middle.2", x, st)
    for item in DEC_PLAN:
        n = f"{p}.upsamples.{item[1]}"
        x = residual_block(sd, n, x, st) if item[0] == "res" else upsample(sd, n, x, st, item[2])
    x = causal_conv(sd, p + ".head.2", F.silu(rms_norm(x, sd[p + ".head.0.gamma"])), st)
    st.chunk += 1
    return x


def vae_encode(sd, video):
    """VideoVAE_.encode (:525-550) + WanVideoVAE scaling: video [1,3,T,H,W] in [-1,1] (T = 4k+1) -> mu [1,16,k+1,H/8,W/8]."""
    st = Stream()
    t = video.shape[2]
    outs = [encoder_chunk(sd, video[:, :, :1], st)]
    for i in range(1, 1 + (t - 1) // 4):
        outs.append(encoder_chunk(sd, video[:, :, 1 + 4 * (i - 1):1 + 4 * i], st))
    mu = conv_1x1x1(sd, "model.conv1", torch.cat(outs, dim=2))[:, :16]
    mean = torch.tensor(LATENT_MEAN, dtype=mu.dtype).view(1, 16, 1, 1, 1)
    std = torch.tensor(LATENT_STD, dtype=mu.dtype).view(1, 16, 1, 1, 1)
    return (mu - mean) * (1.0 / std)


def vae_decode(sd, z):
    """VideoVAE_.decode (:552-575) + clamp (:753-756): z [1,16,T,h,w] -> video [1,3,4T-3,8h,8w] in [-1,1]."""
    st = Stream()
    mean = torch.tensor(LATENT_MEAN, dtype=z.dtype).view(1, 16, 1, 1, 1)
    std = torch.tensor(LATENT_STD, dtype=z.dtype).view(1, 16, 1, 1, 1)
    x = conv_1x1x1(sd, "model.conv2", z / (1.0 / std) + mean)
    outs = [decoder_chunk(sd, x[:, :, i:i + 1], st) for i in range(x.shape[2])]
    return torch.cat(outs, dim=2).clamp(-1, 1)
