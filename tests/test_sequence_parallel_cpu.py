"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: rank partition, group formation, in-place gathers."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_partition_plans():
    from diffsynth.distributed.sequence_parallel import partition
    assert partition(1, 0) == (1, 1, 0, 0)
    assert partition(2, 1) == (2, 1, 1, 0)            # pure CFG-parallel
    assert partition(4, 3) == (2, 2, 1, 1)
    assert partition(8, 5) == (2, 4, 1, 1)            # cfg2 x sp4: 32760 tokens -> 8190 per rank
    assert partition(8, 5, cfg_parallel=False) == (1, 8, 0, 5)
    assert partition(3, 2) == (1, 3, 0, 2)            # odd world: no CFG split


def test_peer_exchange_host_logic():
    from diffsynth.distributed.sequence_parallel import epoch_of, push_order
    assert push_order(1, 4) == [0, 3, 2] and push_order(0, 2) == [1] and push_order(0, 1) == []
    # consumer c walks chunks c+1, c+2, ...: chunk c+k is owned by rank c+k, whose k-th push goes to rank c
    for P in (2, 4, 8):
        for c in range(P):
            for k in range(1, P):
                assert push_order((c + k) % P, P)[k - 1] == c
    eps = [epoch_of(i, 8) for i in range(1, 30)]
    assert 0 not in eps and eps[:8] == [1, 2, 3, 4, 5, 6, 7, 1]
    assert all(a != b for a, b in zip(eps, eps[2:]))     # same-parity neighbours always differ


def _worker(rank, world, port, cfg_parallel, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "stable-video-infinity_b200"))
    from diffsynth.distributed.sequence_parallel import SequenceParallelGroup, all_gather_inplace
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = SequenceParallelGroup(world, rank, cfg_parallel=cfg_parallel)
        L = 12
        ok = True
        if sp.sp_size > 1:
            sp.set_tokens(L)
            full = torch.zeros(L, 3)
            n = sp.local_rows(L)
            full[sp.row_offset:sp.row_offset + n] = torch.arange(sp.row_offset, sp.row_offset + n, dtype=torch.float32)[:, None]
            sp.all_gather_rows(full)
            ok &= torch.equal(full[:, 0], torch.arange(L, dtype=torch.float32))
        if sp.cfg_groups > 1:
            vp = torch.zeros(2, 4)
            vp[sp.cfg_idx] = float(sp.cfg_idx + 1)
            all_gather_inplace(vp, vp[sp.cfg_idx], sp.cfg_group)
            ok &= torch.equal(vp, torch.tensor([[1.0] * 4, [2.0] * 4]))
        ok &= sp.peer_exchange(1024, 64, torch.device("cpu")) is None      # CPU tensors: gather path, no CUDA IPC
        with_err = sp.sp_size == 1
        if sp.sp_size > 1:
            try:
                sp.local_rows(7)
            except RuntimeError:
                with_err = True
        q.put((rank, bool(ok), with_err, sp.describe()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg_parallel,want", [(True, "cfg2xsp1"), (False, "cfg1xsp2")])
def test_two_rank_groups_and_gathers(cfg_parallel, want):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + (7 if cfg_parallel else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg_parallel, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err_ok, desc in res:
        assert ok and err_ok and desc == want
