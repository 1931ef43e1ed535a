"""CPU oracle of the SVI-Dance pose stem — TEST INFRASTRUCTURE ONLY (never imported by the product path).

The reference defines the stem inline as an nn.Sequential of seven nn.Conv3d with SiLU in between
(diffsynth/pipelines/svi_video_dance.py:256-268) and feeds it cat([first frame x3, pose video]) / 255 (:526-528); the
result is flattened 'b c f h w -> b (f h w) c'.  This is that computation with torch.nn.functional in fp32.  Parity
pinned by construction: the layers ARE torch's own conv3d / silu with the reference's hyper-parameters, there is no
restated arithmetic to pin.
"""
import torch
import torch.nn.functional as F

# (stride, padding) of the seven convolutions, svi_video_dance.py:256-268
LAYERS = [((1, 1, 1), (1, 1, 1)), ((1, 1, 1), (1, 1, 1)), ((1, 1, 1), (1, 1, 1)), ((1, 2, 2), (1, 1, 1)),
          ((2, 2, 2), (1, 1, 1)), ((2, 2, 2), (1, 1, 1)), ((1, 2, 2), (0, 0, 0))]


def stem(sd, x):
    """sd: keys '0.weight', '0.bias', '2.weight', ... '12.bias'; x f32 [1, 3, T, H, W] -> [1, dim, f, h, w]."""
    for i, (stride, pad) in enumerate(LAYERS):
        x = F.conv3d(x, sd[f"{2 * i}.weight"].float(), sd[f"{2 * i}.bias"].float(), stride=stride, padding=pad)
        if i < len(LAYERS) - 1:
            x = F.silu(x)
    return x


def pose_condition(sd, humanpose_data):
    """humanpose_data [3, T, H, W] in 0..255 -> [1, f*h*w, dim] (:526-528)."""
    x = humanpose_data.float().unsqueeze(0)
    x = torch.cat([x[:, :, :1].repeat(1, 1, 3, 1, 1), x], dim=2) / 255.0
    y = stem(sd, x)
    return y.flatten(2).transpose(1, 2).contiguous()
