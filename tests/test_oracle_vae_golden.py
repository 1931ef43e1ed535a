"""CPU: the streaming VAE oracle (oracle/wan_vae_oracle.py) against outputs of the REAL reference VAE
(tests/golden/vae_tiny.npz, produced by tests/golden/make_golden.py with the reference's chunked cache protocol)."""
import os

import numpy as np
import torch

from oracle import wan_vae_oracle as V
from tools import synth_vae

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae_tiny.npz")


def test_vae_key_set_matches_reference_checkpoint_hash():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "stable-video-infinity_b200"))
    from diffsynth.models.utils import hash_state_dict_keys
    sh = synth_vae.vae_param_shapes(prefix="")
    assert hash_state_dict_keys({k: torch.empty(v, device="meta") for k, v in sh.items()}) == "ccc42284ea13e1ad04693284c7a09be6"


def test_vae_encode_decode_match_reference():
    g = np.load(GOLD)
    sd = synth_vae.make_vae_state_dict(seed=0)
    with torch.no_grad():
        lat = V.vae_encode(sd, torch.from_numpy(g["video"]).unsqueeze(0))
        dec = V.vae_decode(sd, torch.from_numpy(g["z"]))
    assert lat.shape == g["lat"].shape and dec.shape == g["dec"].shape
    torch.testing.assert_close(lat, torch.from_numpy(g["lat"]), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dec, torch.from_numpy(g["dec"]), rtol=1e-4, atol=1e-4)
