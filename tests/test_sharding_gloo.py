"""Host-side sharding logic on the gloo backend, world size 2, CPU tensors: the
slices tile the read range, and the in-place all-gather assembles on every rank
exactly the array a single process would hold."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_ranges_tile_the_reads():
    from hinge_b200.sharding import shard_ranges

    for n in (1, 7, 64, 1000, 683870):
        for world in (1, 2, 3, 4, 8):
            ranges, chunk = shard_ranges(n, world)
            assert len(ranges) == world and chunk * world >= n
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (l0, h0), (l1, h1) in zip(ranges, ranges[1:]):
                assert h0 == l1 and l0 <= h0
            assert all(h - l <= chunk for l, h in ranges)


def test_weighted_shard_ranges_balance_the_record_volume():
    import numpy as np

    from hinge_b200.sharding import shard_ranges

    rng = np.random.default_rng(5)
    for n, world in ((1, 2), (5, 8), (1000, 3), (50000, 8)):
        # a few very deep pile-ups among many shallow ones
        w = rng.integers(1, 200, n).astype(np.int64)
        w[rng.integers(0, n, max(1, n // 100))] += 20000
        ranges, chunk = shard_ranges(n, world, weights=w)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == n
        for (l0, h0), (l1, h1) in zip(ranges, ranges[1:]):
            assert h0 == l1 and l0 <= h0
        assert chunk == max(h - l for l, h in ranges)
        if n >= 1000:  # no shard exceeds its share by more than the heaviest read
            share = w.sum() / world
            assert all(w[l:h].sum() <= share + w.max() for l, h in ranges)


def _worker(rank, world, port, n_read, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hinge_b200.sharding import ShardedArrays

    arr = ShardedArrays(n_read, rank, world, torch.device("cpu"), equal_slices=True)
    n = world * arr.chunk
    arr.mean_cov = torch.full((n,), -1, dtype=torch.int32)
    arr.mask = torch.zeros((n, 2), dtype=torch.int32)
    arr.hist = torch.zeros((4098,), dtype=torch.int32)
    arr.mask_pk = torch.zeros((n,), dtype=torch.int32)
    # every rank fills only the reads it owns, like hg_filter_phase1 / phase2 do
    ids = torch.arange(arr.lo, arr.hi, dtype=torch.int32)
    arr.mean_cov[arr.lo:arr.hi] = ids * 3 + 1
    arr.mask[arr.lo:arr.hi, 0] = ids
    arr.mask[arr.lo:arr.hi, 1] = ids + 1000
    # the two exchanges run_filter_sharded makes: histogram of the owned reads' mean coverage summed
    # over the ranks, masks gathered as one packed word per read
    arr.hist[:4096] += torch.bincount((ids % 4096).to(torch.int64), minlength=4096).to(torch.int32)
    arr.hist[4096] += len(ids)
    arr.mask_pk[arr.lo:arr.hi] = (ids // 20) | ((ids // 20 + 7) << 16)
    arr.exchange_slices(arr.mean_cov)
    arr.exchange_slices(arr.mask)
    arr.exchange_slices(arr.mask_pk)
    dist.all_reduce(arr.hist)
    full = torch.arange(n_read, dtype=torch.int32)
    want_hist = torch.bincount((full % 4096).to(torch.int64), minlength=4096).to(torch.int32)
    ok = bool(torch.equal(arr.mean_cov[:n_read], full * 3 + 1) and torch.equal(arr.mask[:n_read, 0], full)
              and torch.equal(arr.mask[:n_read, 1], full + 1000)
              and torch.equal(arr.mask_pk[:n_read], (full // 20) | ((full // 20 + 7) << 16))
              and torch.equal(arr.hist[:4096], want_hist) and int(arr.hist[4096]) == n_read)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_read", [10, 1001])
def test_all_gather_assembles_the_global_arrays(n_read):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_read % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_read, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


class _FakeContext:
    """Stands in for hinge_b200.api.Context on the CPU: computes into the bound tensors what the
    phases of the real context leave there, and records the order of the calls."""

    def __init__(self, arrays, refuse_packed):
        self.arrays, self.refuse_packed, self.calls, self.bound = arrays, refuse_packed, [], {}

    def bind_buffer(self, which, tensor):
        from hinge_b200 import api

        if which == api.HG_BUF_MASK_PACKED and self.refuse_packed:
            raise api.HingeError("reads too long for 16-bit mask bounds")
        self.bound[which] = tensor

    def filter_phase1(self, params):
        from hinge_b200 import api

        a = self.arrays
        self.calls.append("phase1")
        self.bound[api.HG_BUF_MEDIAN_HIST][7] += a.hi - a.lo  # every owned read has mean coverage 7

    def filter_phase2(self):
        from hinge_b200 import api

        a = self.arrays
        self.calls.append("phase2")
        self.hist_seen = int(self.bound[api.HG_BUF_MEDIAN_HIST][7])
        ids = torch.arange(a.lo, a.hi, dtype=torch.int32)
        if api.HG_BUF_MASK_PACKED in self.bound:
            self.bound[api.HG_BUF_MASK_PACKED][a.lo:a.hi] = ids | ((ids + 5) << 16)
        else:
            self.bound[api.HG_BUF_MASK][a.lo:a.hi, 0] = ids
            self.bound[api.HG_BUF_MASK][a.lo:a.hi, 1] = ids + 5

    def filter_phase3(self):
        from hinge_b200 import api

        self.calls.append("phase3")
        if api.HG_BUF_MASK_PACKED in self.bound:
            m = self.bound[api.HG_BUF_MASK_PACKED]
            self.masks_seen = torch.stack([m & 0xffff, m >> 16], dim=1)
        else:
            self.masks_seen = self.bound[api.HG_BUF_MASK].clone()
        return 0, None


def _flow_worker(rank, world, port, n_read, refuse_packed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hinge_b200.sharding import ShardedArrays, run_filter_sharded

    arr = ShardedArrays(n_read, rank, world, torch.device("cpu"), equal_slices=True)
    ctx = _FakeContext(arr, refuse_packed)
    arr.bind(ctx, exchange="nccl")
    rc, _ = run_filter_sharded(ctx, None, arr)
    full = torch.arange(n_read, dtype=torch.int32)
    ok = (rc == 0 and ctx.calls == ["phase1", "phase2", "phase3"]
          and ctx.hist_seen == n_read  # phase 2 saw the histogram of ALL ranks
          and arr.packed == (not refuse_packed)
          and bool(torch.equal(ctx.masks_seen[:n_read, 0], full))  # phase 3 saw the masks of ALL reads
          and bool(torch.equal(ctx.masks_seen[:n_read, 1], full + 5)))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("refuse_packed", [False, True])
def test_sharded_filter_flow_exchanges_between_the_phases(refuse_packed):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 11 + int(refuse_packed)
    procs = [ctx.Process(target=_flow_worker, args=(r, 2, port, 1001, refuse_packed, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def _lists_worker(rank, world, port, n_read, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np

    from hinge_b200.sharding import ShardedArrays, gather_filter_lists

    # the whole result as one process would see it: read i carries i % 3 annotations, every second one a hinge
    counts = np.arange(n_read) % 3
    off = np.zeros(n_read + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    pos = (np.arange(off[-1]) * 40).astype(np.int32)
    typ = np.where(np.arange(off[-1]) % 2 == 0, 1, -1).astype(np.int32)
    keep = (np.arange(off[-1]) % 2 == 1).astype(np.uint8)
    arr = ShardedArrays(n_read, rank, world, torch.device("cpu"), weights=np.arange(1, n_read + 1))
    # this rank's fetch: offsets are global-sized, but only the annotations of its own reads are there
    a0, a1 = off[arr.lo], off[arr.hi]
    mine_off = np.clip(off, a0, a1) - a0
    mine = {"anno_off": mine_off, "anno_pos": pos[a0:a1], "anno_type": typ[a0:a1], "hinge_keep": keep[a0:a1]}
    (rep_off, rep_pos, rep_typ), (hin_off, hin_pos, hin_typ) = gather_filter_lists(mine, arr)
    k = keep.astype(bool)
    per_read = np.repeat(np.arange(n_read), counts)
    want_hin_off = np.zeros(n_read + 1, np.int64)
    np.cumsum(np.bincount(per_read[k], minlength=n_read), out=want_hin_off[1:])
    ok = (np.array_equal(rep_off, off) and np.array_equal(rep_pos, pos) and np.array_equal(rep_typ, typ)
          and np.array_equal(hin_off, want_hin_off) and np.array_equal(hin_pos, pos[k]) and np.array_equal(hin_typ, typ[k]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_annotation_lists_are_gathered_into_global_csrs():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 23
    procs = [ctx.Process(target=_lists_worker, args=(r, 2, port, 203, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
