"""CPU: the HOST logic of WanVAEEngine — frame rings and their slot tables, the chunked causal schedule, which conv epilogue feeds
which ring, up / down-sampling bookkeeping, the attention block, and the SPATIALLY SHARDED path (row bands, halo exchange per conv
input, token gather) under a 2-process gloo group — with the kernels replaced by the torch stand-ins of tests/nv_emulation.py
(the conv stand-in reads the very descriptor the C ABI receives, through its raw pointers).  The GPU tests check the kernels."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tools import synth_vae


def _vae():
    from diffsynth.models.wan_video_vae import WanVideoVAE
    sd = {k: v.to(torch.bfloat16).float() if v.dim() > 1 and "gamma" not in k else v
          for k, v in synth_vae.make_vae_state_dict(seed=0).items()}
    m = WanVideoVAE().eval()
    m.load_state_dict(sd)
    return m, sd


def _levels(a, b):
    d = (a - b).abs() * 127.5
    return d.max().item(), d.mean().item()


def test_vae_engine_orchestration_matches_oracle(monkeypatch):
    from oracle import wan_vae_oracle as V
    import nv_emulation
    nv_emulation.install(monkeypatch)
    m, sd = _vae()
    g = torch.Generator().manual_seed(5)
    T, H, W = 5, 32, 48
    z = torch.randn(1, 16, (T - 1) // 4 + 1, H // 8, W // 8, generator=g)
    video = torch.rand(3, T, H, W, generator=g) * 2 - 1
    with torch.no_grad():
        ref_dec = V.vae_decode(sd, z)
        ref_enc = V.vae_encode(sd, video.unsqueeze(0))
    dec = m.decode(z, device="cpu")
    assert dec.shape == ref_dec.shape
    mx, mean = _levels(dec, ref_dec)
    assert mx < 6.0 and mean < 0.6, (mx, mean)
    enc = m.encode([video], device="cpu")
    err = (enc - ref_enc).abs()
    assert err.max().item() < 0.05 * ref_enc.std().item() + 0.02 and err.mean().item() < 6e-3 * ref_enc.std().item()
    # a second video of the same geometry re-uses the frame rings (rewound, not cleared): must not see the first one
    z2 = torch.randn(1, 16, (T - 1) // 4 + 1, H // 8, W // 8, generator=g)
    with torch.no_grad():
        ref2 = V.vae_decode(sd, z2)
    mx2, mean2 = _levels(m.decode(z2, device="cpu"), ref2)
    assert mx2 < 6.0 and mean2 < 0.6, (mx2, mean2)


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "stable-video-infinity_b200"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    import nv_emulation
    nv_emulation.install()
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        m, _ = _vae()
        # (frames, H, W): 6 latent rows = 3 + 3 (or 1 + 2 + 1 + 2 over 4 ranks), 5 latent rows = 2 + 3 (or 1 + 1 + 1 + 2): uneven bands,
        # odd band heights at every level
        for T, H, W in ((5, 48, 32), (5, 40, 48)):
            g = torch.Generator().manual_seed(T + H)
            video = torch.rand(3, T, H, W, generator=g) * 2 - 1
            z = torch.randn(1, 16, (T - 1) // 4 + 1, H // 8, W // 8, generator=g)
            m.shard_group = None
            enc1, dec1 = m.encode([video], device="cpu").clone(), m.decode(z, device="cpu").clone()
            m.enable_spatial_sharding()
            encP, decP = m.encode([video], device="cpu"), m.decode(z, device="cpu")
            assert encP.shape == enc1.shape and decP.shape == dec1.shape
            res[f"{T}x{H}x{W} encode"] = (enc1 - encP).abs().max().item()
            res[f"{T}x{H}x{W} decode"] = (dec1 - decP).abs().max().item()
            eng = m.engine("cpu")
            res[f"{T}x{H}x{W} halo exchanges"] = eng.halo_exchanges
        q.put((rank, res))
    except Exception as ex:  # noqa: BLE001 — report instead of letting the parent wait for its queue timeout
        import traceback
        q.put((rank, {"error": traceback.format_exc()[-1500:] + repr(ex)}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_spatially_sharded_vae(world):
    """world_size 2 and 4 over gloo (4: bands of ONE latent row, whose 3x3 convolutions read nothing but halo rows above and below): every rank encodes / decodes its band of image rows — one border row exchanged with the neighbour
    per convolution input, the attention block's tokens gathered, the result bands gathered on every rank — and must reproduce
    the single-process result exactly, as the GPU kernels do (tools/vae_shard_check.py): the stand-in kernels accumulate in fp64
    and round once, so a pixel's value does not depend on how the CPU library blocks a band vs a whole frame."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90) + world
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in got:
        assert "error" not in res, res["error"]
        assert len(res) == 6, res
        for name, val in res.items():
            if name.endswith("halo exchanges"):
                assert val > 0, (rank, name)
            else:
                assert val < 1e-6, (rank, name, val)
