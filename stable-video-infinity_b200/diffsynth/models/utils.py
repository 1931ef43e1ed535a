"""Checkpoint I/O helpers (reference diffsynth/models/utils.py): safetensors / torch loaders, meta-device
initialisation and the hash-of-keys fingerprint the model detector uses."""
import hashlib
from contextlib import contextmanager

import torch


@contextmanager
def init_weights_on_device(device=torch.device("meta"), include_buffers: bool = False):
    """Construct modules without allocating real storage (reference utils.py:6-53)."""
    with torch.device(device):
        yield


def load_state_dict_from_folder(file_path, torch_dtype=None):
    import os
    sd = {}
    for name in sorted(os.listdir(file_path)):
        if name.rsplit(".", 1)[-1] in ("safetensors", "bin", "ckpt", "pth", "pt"):
            sd.update(load_state_dict(os.path.join(file_path, name), torch_dtype=torch_dtype))
    return sd


def load_state_dict(file_path, torch_dtype=None, device="cpu"):
    if file_path.endswith(".safetensors"):
        return load_state_dict_from_safetensors(file_path, torch_dtype=torch_dtype, device=device)
    return load_state_dict_from_bin(file_path, torch_dtype=torch_dtype, device=device)


def load_state_dict_from_safetensors(file_path, torch_dtype=None, device="cpu"):
    from safetensors import safe_open
    sd = {}
    with safe_open(file_path, framework="pt", device=str(device)) as f:
        for k in f.keys():
            t = f.get_tensor(k)
            sd[k] = t.to(torch_dtype) if torch_dtype is not None else t
    return sd


def load_state_dict_from_bin(file_path, torch_dtype=None, device="cpu"):
    sd = torch.load(file_path, map_location=device, weights_only=True)
    if torch_dtype is not None:
        sd = {k: (v.to(torch_dtype) if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}
    return sd


def search_for_embeddings(state_dict):
    out = []
    for v in state_dict.values():
        if isinstance(v, torch.Tensor):
            out.append(v)
        elif isinstance(v, dict):
            out += search_for_embeddings(v)
    return out


def convert_state_dict_keys_to_single_str(state_dict, with_shape=True):
    """reference utils.py:158-176: sorted 'key[:shape]' strings joined by ',' (nested dicts recurse with '|')."""
    keys = []
    for k, v in state_dict.items():
        if isinstance(k, str):
            if isinstance(v, torch.Tensor):
                if with_shape:
                    keys.append(k + ":" + "_".join(str(s) for s in v.shape))
                keys.append(k)  # the reference appends the bare key as well (utils.py:153-156)
            elif isinstance(v, dict):
                keys.append(k + "|" + convert_state_dict_keys_to_single_str(v, with_shape=with_shape))
    keys.sort()
    return ",".join(keys)


def split_state_dict_with_prefix(state_dict):
    groups = {}
    for k in sorted(k for k in state_dict if isinstance(k, str)):
        groups.setdefault(k if "." not in k else k.split(".")[0], []).append(k)
    return [{k: state_dict[k] for k in ks} for ks in groups.values()]


def hash_state_dict_keys(state_dict, with_shape=True):
    s = convert_state_dict_keys_to_single_str(state_dict, with_shape=with_shape)
    return hashlib.md5(s.encode("UTF-8")).hexdigest()
