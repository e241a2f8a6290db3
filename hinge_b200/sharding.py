"""Sharding of the hot path across the GPUs of one box (one process per GPU).

Overlap records of one A-read are contiguous in a LAsort-ed .las
(thirdparty/DALIGNER/LAsort.c:28-67), so reads — and with them their pile-ups —
shard by A-read id.  Two per-read arrays cross shards inside `hinge filter`:

  * the per-read mean coverage, whose global median sets MIN_COV
    (filter.cpp:642-678).  The median is found by counting, so the ranks only sum
    their 4096-bin histograms of it                         -> all-reduce, 16 KB
  * the mask of every B read a pile-up touches
    (filter.cpp:884-889)                                    -> all-gather, 4 B/read
    (both bounds are multiples of gcd(40, tspace) and travel as 16-bit units; 8 B/read
    when a read is too long for that)

Both live in padded torch tensors bound into the context (hg_bind_buffer), so
`torch.distributed.all_gather_into_tensor` moves them over NCCL/NVLink in place.
This module holds only the host-side plumbing; it runs unchanged on the gloo
backend with CPU tensors (tests/test_sharding_gloo.py).
"""
import torch
import torch.distributed as dist


def shard_ranges(n_read, world):
    """Equal read-count slices [lo, hi) per rank plus the padded slice length.

    Equal counts (not equal record counts) keep every rank's slice of the
    exchanged arrays the same size, which all_gather_into_tensor needs; read ids
    are in sequencing order, so record counts balance statistically."""
    chunk = (n_read + world - 1) // world
    return [(min(n_read, r * chunk), min(n_read, (r + 1) * chunk)) for r in range(world)], chunk


class ShardedArrays:
    """The two exchanged per-read arrays of one rank."""

    def __init__(self, n_read, rank, world, device):
        self.rank, self.world = rank, world
        self.ranges, self.chunk = shard_ranges(n_read, world)
        self.lo, self.hi = self.ranges[rank]
        self.mean_cov = torch.full((world * self.chunk,), -1, dtype=torch.int32, device=device)
        self.mask = torch.zeros((world * self.chunk, 2), dtype=torch.int32, device=device)
        self.hist = torch.zeros((4098,), dtype=torch.int32, device=device)  # HG_BUF_MEDIAN_HIST
        self.mask_pk = torch.zeros((world * self.chunk,), dtype=torch.int32, device=device)
        self.packed = False

    def bind(self, ctx):
        """Makes the context compute straight into the exchanged arrays."""
        from . import api

        if self.world > 1:
            ctx.bind_buffer(api.HG_BUF_MEDIAN_HIST, self.hist)
            try:
                ctx.bind_buffer(api.HG_BUF_MASK_PACKED, self.mask_pk)
                self.packed = True
            except Exception:  # a read too long for 16-bit bounds: exchange the full masks
                self.packed = False
        if not self.packed:
            ctx.bind_buffer(api.HG_BUF_MASK, self.mask)

    def exchange(self, t):
        """All-gather the rank's own slice of `t` into every rank's copy of `t`."""
        if self.world == 1:
            return
        mine = t[self.rank * self.chunk:(self.rank + 1) * self.chunk].clone()
        dist.all_gather_into_tensor(t, mine)


def run_filter_sharded(ctx, params, arrays):
    """hg_filter split at its two global dependencies (include/hinge_b200.h)."""
    ctx.filter_phase1(params)
    if arrays.world > 1:
        dist.all_reduce(arrays.hist)  # the context was bound to arrays.hist (HG_BUF_MEDIAN_HIST)
    ctx.filter_phase2()
    arrays.exchange(arrays.mask_pk if arrays.packed else arrays.mask)
    return ctx.filter_phase3()
