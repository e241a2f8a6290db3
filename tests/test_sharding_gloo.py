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


def _worker(rank, world, port, n_read, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hinge_b200.sharding import ShardedArrays

    arr = ShardedArrays(n_read, rank, world, torch.device("cpu"))
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
    arr.exchange(arr.mean_cov)
    arr.exchange(arr.mask)
    arr.exchange(arr.mask_pk)
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
