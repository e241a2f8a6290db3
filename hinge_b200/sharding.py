"""Sharding of the hot path across the GPUs of one box (one process per GPU).

Overlap records of one A-read are contiguous in a LAsort-ed .las
(thirdparty/DALIGNER/LAsort.c:28-67), so reads — and with them their pile-ups —
shard by A-read id: every rank owns a contiguous range of A-reads holding about the
same volume of records.  What crosses shards:

  inside `hinge filter`
  * the per-read mean coverage, whose global median sets MIN_COV (filter.cpp:642-678).
    The median is found by counting, so only a 4096-bin histogram travels (16 KB)
  * the mask of every B read a pile-up touches (filter.cpp:884-889): 4 B per read
    (both bounds are multiples of gcd(40, tspace) and travel as 16-bit units)
  Two ways to move them (`exchange`):
    "peer"  the kernels do it themselves over NVLink peer memory (hg_peer_export /
            hg_peer_connect): the histogram kernel stores its part into every rank's
            block, K2 stores each packed mask word into every rank's array, arrival flags
            replace the barrier — hg_filter stays ONE stream of kernels with no host or
            NCCL call in between
    "nccl"  phase-level calls with an all-reduce and an all-gather between them

  before `hinge layout` (north_star's "single NCCL all-gather of the hinge / maximal-read
  bitmaps")
  * hinge and repeat lists of every read, the maximal-read bitmap

This module holds only the host-side plumbing; the NCCL paths run unchanged on the gloo
backend with CPU tensors (tests/test_sharding_gloo.py).
"""
import math

import numpy as np
import torch
import torch.distributed as dist


def shard_ranges(n_read, world, weights=None):
    """Contiguous read ranges [lo, hi) per rank.

    weights=None: equal read counts (what `all_gather_into_tensor` needs).
    weights=w[n_read]: equal cumulative weight — pass the per-read record counts
    (np.diff(read_off)) or, before the records exist, the read lengths (pile-up depth
    goes with read length); cuts fall on read boundaries."""
    if weights is None:
        chunk = (n_read + world - 1) // world
        return [(min(n_read, r * chunk), min(n_read, (r + 1) * chunk)) for r in range(world)], chunk
    cum = np.cumsum(np.asarray(weights, dtype=np.float64))
    total = float(cum[-1]) if n_read else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(n_read)
    for r in range(1, world + 1):  # monotone, non-empty where possible
        cuts[r] = max(cuts[r], cuts[r - 1])
    ranges = [(cuts[r], cuts[r + 1]) for r in range(world)]
    return ranges, max(hi - lo for lo, hi in ranges)


class ShardedArrays:
    """The exchanged per-read arrays of one rank (NCCL exchange) or the peer-memory
    connection of its context (peer exchange)."""

    def __init__(self, n_read, rank, world, device, weights=None, equal_slices=False):
        self.rank, self.world, self.n_read, self.device = rank, world, n_read, device
        self.ranges, self.chunk = shard_ranges(n_read, world, None if equal_slices else weights)
        self.lo, self.hi = self.ranges[rank]
        self.exchange = "none"
        self.packed = False
        self.mask = self.hist = self.mask_pk = None
        self.g = 1

    # ---- binding
    def bind(self, ctx, exchange="nccl", group=None):
        """Makes the context compute straight into the exchanged arrays (nccl) or connects
        the contexts of all ranks through peer memory (peer)."""
        from . import api

        self.ctx = ctx
        if self.world == 1:
            return
        if exchange == "peer":
            handle = ctx.peer_export(self.rank, self.world)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle, group=group)
            ctx.peer_connect(b"".join(handles))
            dist.barrier(group=group)  # every rank has mapped every block before anyone writes
            self.exchange = "peer"
            return
        if any(hi - lo != self.chunk for lo, hi in self.ranges[:-1]):
            raise ValueError("the NCCL exchange needs equal read-count slices (equal_slices=True)")
        self.exchange = "nccl"
        n = self.world * self.chunk
        self.hist = torch.zeros((4098,), dtype=torch.int32, device=self.device)  # HG_BUF_MEDIAN_HIST
        ctx.bind_buffer(api.HG_BUF_MEDIAN_HIST, self.hist)
        self.mask_pk = torch.zeros((n,), dtype=torch.int32, device=self.device)
        try:
            ctx.bind_buffer(api.HG_BUF_MASK_PACKED, self.mask_pk)
            self.packed = True
        except api.HingeError:  # a read too long for 16-bit bounds: exchange the full masks
            self.packed = False
            self.mask = torch.zeros((n, 2), dtype=torch.int32, device=self.device)
            ctx.bind_buffer(api.HG_BUF_MASK, self.mask)

    def set_global_range(self, ctx, first_aread, last_aread, group=None):
        """The reference's r_begin / r_end are the first and last A-read of the WHOLE .las
        (filter.cpp:516-517); a shard only sees its own records."""
        if self.world > 1:
            t = torch.tensor([-first_aread, last_aread], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            first_aread, last_aread = -int(t[0].item()), int(t[1].item())
        ctx.set_global_range(first_aread, last_aread)

    def exchange_slices(self, t):
        """All-gather the rank's own slice of `t` into every rank's copy of `t`."""
        if self.world == 1:
            return
        mine = t[self.rank * self.chunk:(self.rank + 1) * self.chunk].clone()
        dist.all_gather_into_tensor(t, mine)

    def gathered_mask(self, n_read):
        """The masks of ALL reads as this rank holds them after a run (numpy, n_read x 2)."""
        if self.exchange == "peer":
            return self.ctx.peer_masks()
        if self.packed:
            pk = self.mask_pk[:n_read].cpu().numpy().view(np.uint32)
            g = math.gcd(40, self.ctx.tspace)
            return np.stack([(pk & 0xffff).astype(np.int32) * g, (pk >> 16).astype(np.int32) * g], axis=1)
        return self.mask[:n_read].cpu().numpy()


def run_filter_sharded(ctx, params, arrays, max_attempts=8):
    """hg_filter on a shard.  Returns (rc, summary)."""
    from . import api

    if arrays.world == 1 or arrays.exchange == "peer":
        # one stream of kernels; with peer exchange the annotation-pool retry is agreed on between
        # the ranks on the device, so every rank reruns together
        return 0, ctx.filter(params)
    for _ in range(max_attempts):
        ctx.filter_phase1(params)
        dist.all_reduce(arrays.hist)  # the context was bound to arrays.hist (HG_BUF_MEDIAN_HIST)
        ctx.filter_phase2()
        arrays.exchange_slices(arrays.mask_pk if arrays.packed else arrays.mask)
        rc, s = ctx.filter_phase3()
        # a pool overflow on one rank reruns the stage on all of them (phase 3 has grown the pool)
        t = torch.tensor([rc], dtype=torch.int32, device=arrays.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if int(t.item()) != api.HG_RETRY_POOL:
            return rc, s
    return api.HG_RETRY_POOL, s


def gather_filter_lists(result, arrays, group=None):
    """All-gather-v of the shards' repeat-annotation / hinge lists ("after K4"):
    `result` is this rank's hg_filter_fetch; returns the GLOBAL CSRs (rep, hin), each (off[n+1], pos, type),
    identical on every rank -- what `.repeat.txt` / `.hinges.txt` hold in a single-process run."""
    n, lo, hi = arrays.n_read, arrays.lo, arrays.hi
    a0, a1 = int(result["anno_off"][lo]), int(result["anno_off"][hi])
    mine = (lo, hi, np.diff(result["anno_off"][lo:hi + 1]).astype(np.int64), result["anno_pos"][a0:a1],
            result["anno_type"][a0:a1], result["hinge_keep"][a0:a1].astype(bool))
    parts = [mine]
    if arrays.world > 1:
        parts = [None] * arrays.world
        dist.all_gather_object(parts, mine, group=group)
    counts = np.zeros(n, np.int64)
    for plo, phi, cnt, _, _, _ in parts:
        counts[plo:phi] = cnt
    pos = np.concatenate([p[3] for p in parts]).astype(np.int32)
    typ = np.concatenate([p[4] for p in parts]).astype(np.int32)
    keep = np.concatenate([p[5] for p in parts])
    rep_off = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=rep_off[1:])
    per_read = np.repeat(np.arange(n), counts)
    hin_off = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(per_read[keep], minlength=n), out=hin_off[1:])
    return (rep_off, pos, typ), (hin_off, pos[keep], typ[keep])


def run_maximal_sharded(ctx, params, arrays, mask, group=None):
    """hg_maximal on shards: local classification + containment lists, then ONE exchange round over
    NCCL (MAX all-reduce of the per-read states, all-gather of the unknown reads' lists) and the
    resolve on every rank.  Returns the maximal-read bitmap of ALL reads (numpy uint8[n_read])."""
    dev, world, n = arrays.device, arrays.world, arrays.n_read
    state = torch.zeros((n,), dtype=torch.uint8, device=dev)
    unk_cap, pool_cap = max(1024, (arrays.hi - arrays.lo) // 4), max(4096, arrays.hi - arrays.lo)
    while True:
        unk = torch.zeros((unk_cap, 4), dtype=torch.int32, device=dev)
        pool = torch.zeros((pool_cap,), dtype=torch.int32, device=dev)
        n_unk, n_pool = ctx.maximal_phase1(params, mask, state, unk, pool)
        if n_unk <= unk_cap and n_pool <= pool_cap:
            break
        unk_cap, pool_cap = max(unk_cap, n_unk), max(pool_cap, n_pool)
    if world == 1:
        return ctx.maximal_phase2(state, unk, [n_unk, n_pool], 1, unk_cap, pool, pool_cap)
    counts = torch.tensor([n_unk, n_pool], dtype=torch.int64, device=dev)
    counts_all = torch.zeros((world, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_all, counts, group=group)
    counts_all = counts_all.cpu().numpy()
    unk_stride, pool_stride = max(1, int(counts_all[:, 0].max())), max(1, int(counts_all[:, 1].max()))
    mine_unk = torch.zeros((unk_stride, 4), dtype=torch.int32, device=dev)
    mine_unk[:n_unk] = unk[:n_unk]
    mine_pool = torch.zeros((pool_stride,), dtype=torch.int32, device=dev)
    mine_pool[:n_pool] = pool[:n_pool]
    unk_all = torch.empty((world * unk_stride, 4), dtype=torch.int32, device=dev)
    pool_all = torch.empty((world * pool_stride,), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(unk_all, mine_unk, group=group)
    dist.all_gather_into_tensor(pool_all, mine_pool, group=group)
    dist.all_reduce(state, op=dist.ReduceOp.MAX, group=group)  # slices are disjoint; 0 elsewhere
    return ctx.maximal_phase2(state, unk_all, counts_all.reshape(-1), world, unk_stride, pool_all, pool_stride)


def run_layout_sharded(ctx, params, arrays, mask, maximal, rep, hin, group=None):
    """hg_layout on shards.  mask / maximal / rep / hin are the GLOBAL arrays (the filter's results of all
    shards gathered, the bitmap of run_maximal_sharded): north_star's "single NCCL all-gather of the hinge /
    maximal-read bitmaps before layout".  Between the phases two small flag arrays are reduced and the
    hinge-graph records gathered.  Returns this rank's edges (list of EdgeC, read order) and the device ms."""
    from . import api

    dev, world = arrays.device, arrays.world
    n_hinges = int(hin[0][-1])
    contained = api.layout_phase1(ctx, params, mask, maximal, rep, hin)
    if world > 1:
        t = torch.from_numpy(contained).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        contained = t.cpu().numpy()
    alive, graph = api.layout_phase2(ctx, contained, n_hinges)
    if world > 1:
        t = torch.from_numpy(np.ascontiguousarray(alive)).to(dev) if n_hinges else None
        if t is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            alive = t.cpu().numpy()
        parts = [None] * world
        dist.all_gather_object(parts, graph, group=group)
        graph = np.vstack([p.reshape(-1, 10) for p in parts])
    return api.layout_phase3(ctx, alive, graph)
