"""Native multi-GPU plan for the DiT hot loop: CFG-parallel x token-axis sequence parallelism over NCCL.

Replaces the reference's xfuser "USP" wiring (``diffsynth/distributed/xdit_context_parallel.py``:
``usp_dit_forward`` :42-105 chunks tokens on dim 1 and all-gathers the head output; ``usp_attn_forward``
:108-129 runs Ulysses/ring attention).  xfuser/yunchang are unvendored and unpinned, so their numerics are
"parity unpinned" (SURVEY.md §8c): this implementation is checked against the native single-GPU result.

Plan for N ranks (one process per GPU):
  * the two classifier-free-guidance branches are independent -> ranks [0, N/2) run the conditional forward,
    ranks [N/2, N) the unconditional one (zero communication inside a forward); one all-gather of the two
    velocity fields (8 MB) per step joins them;
  * inside a branch the token axis is split over sp_size = N/2 ranks: every row-local op (LayerNorm, GEMMs,
    RMSNorm, RoPE with a row offset, cross-attention, FFN) runs on L/sp rows; self-attention needs all K/V:
    one in-place NCCL all-gather of the [L, 2d] K|V buffer per layer (NVLink 5 / NVSwitch);
  * the head output rows are all-gathered once per forward (reference: get_sp_group().all_gather).
Requires L % sp_size == 0 like the reference (torch.chunk + equal-size gather).
"""
import torch
import torch.distributed as dist

_GROUP = None


def partition(world, rank, cfg_parallel=True):
    """Pure host logic: (cfg_groups, sp_size, cfg_idx, sp_rank) for a rank."""
    cfg_groups = 2 if (cfg_parallel and world % 2 == 0) else 1
    sp_size = world // cfg_groups
    return cfg_groups, sp_size, rank // sp_size, rank % sp_size


def all_gather_inplace(full, local, group=None):
    """full[r*n:(r+1)*n] = local of rank r; `local` may alias its own slice of `full`."""
    try:
        dist.all_gather_into_tensor(full, local, group=group)
    except (RuntimeError, NotImplementedError):   # gloo (CPU tests): list form
        n = dist.get_world_size(group)
        dist.all_gather([c.view_as(local) for c in full.chunk(n, dim=0)], local.clone(), group=group)
    return full


class SequenceParallelGroup:
    def __init__(self, world, rank, cfg_parallel=True):
        self.world, self.rank = world, rank
        self.cfg_groups, self.sp_size, self.cfg_idx, self.sp_rank = partition(world, rank, cfg_parallel)
        self.sp_group = None
        self.cfg_group = None
        # every rank creates every group, in the same order
        for c in range(self.cfg_groups):
            ranks = list(range(c * self.sp_size, (c + 1) * self.sp_size))
            g = dist.new_group(ranks) if self.sp_size > 1 else None
            if c == self.cfg_idx:
                self.sp_group = g
        for s in range(self.sp_size):
            ranks = [c * self.sp_size + s for c in range(self.cfg_groups)]
            g = dist.new_group(ranks) if self.cfg_groups > 1 else None
            if s == self.sp_rank:
                self.cfg_group = g
        self._rows = None
        self._bufs = {}

    def describe(self):
        return f"cfg{self.cfg_groups}xsp{self.sp_size}"

    # ---- token partition
    def local_rows(self, L):
        if L % self.sp_size != 0:
            raise RuntimeError(f"sequence parallelism needs L % sp_size == 0 (L={L}, sp_size={self.sp_size}); "
                               f"the reference has the same constraint (torch.chunk + equal all_gather)")
        return L // self.sp_size

    def set_tokens(self, L):
        self._rows = self.local_rows(L)

    @property
    def row_offset(self):
        return self.sp_rank * self._rows

    def local_slice(self, full):
        n = full.shape[0] // self.sp_size
        return full[self.sp_rank * n:(self.sp_rank + 1) * n]

    # ---- collectives
    def all_gather_rows(self, full):
        """in-place all-gather of a row-sharded [L, *] buffer inside the sequence-parallel group"""
        if self.sp_size > 1:
            all_gather_inplace(full, self.local_slice(full), self.sp_group)
        return full

    def cfg_parallel_step(self, eng, lat, t, cp, cn, v_c, v_u, cfg_scale, sigma, nxt, y=None):
        """One denoise step under the plan above; lat is replicated on every rank and updated identically."""
        inner = self if self.sp_size > 1 else None
        if self.cfg_groups == 2:
            key = ("vpair", tuple(lat.shape))
            vp = self._bufs.get(key)
            if vp is None:
                vp = torch.empty((2,) + tuple(lat.shape), device=lat.device, dtype=torch.float32)
                self._bufs[key] = vp
            eng.forward(lat, t, cp if self.cfg_idx == 0 else cn, y=y, sp=inner, out=vp[self.cfg_idx])
            all_gather_inplace(vp, vp[self.cfg_idx], self.cfg_group)
            eng.k.cfg_euler_step(lat, vp[0], vp[1], cfg_scale, sigma, nxt)
        else:
            eng.forward(lat, t, cp, y=y, sp=inner, out=v_c)
            eng.forward(lat, t, cn, y=y, sp=inner, out=v_u)
            eng.k.cfg_euler_step(lat, v_c, v_u, cfg_scale, sigma, nxt)
        return lat


def init_sp_groups(world=None, rank=None, cfg_parallel=True):
    global _GROUP
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    _GROUP = SequenceParallelGroup(world, rank, cfg_parallel)
    return _GROUP


def get_sp_group():
    if _GROUP is None:
        if not dist.is_initialized():
            raise RuntimeError("sequence parallelism requested but torch.distributed is not initialised")
        return init_sp_groups(cfg_parallel=False)
    return _GROUP
