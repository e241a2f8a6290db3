"""Run under torchrun with N >= 2 ranks: the sharded filter (reads split by A-read id, balanced on
record volume) must reproduce the single-context result exactly, with both forms of the phase
exchange: inside the kernels through NVLink peer memory, and NCCL calls between phase-level calls.

  torchrun --nproc-per-node 2 scripts/sharded_check.py [--shape c5]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import hgsynth  # noqa: E402
from hinge_b200 import api  # noqa: E402
from hinge_b200.sharding import (ShardedArrays, gather_filter_lists, run_filter_sharded, run_layout_sharded,  # noqa: E402
                                 run_maximal_sharded)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if "c5" in sys.argv:  # long reads, several local alignments per pair (BASELINE configs[4] shape)
    syn = hgsynth.Synth(genome_len=6_000_000, coverage=40.0, read_mean=24000, read_sd=8000, read_min=2000,
                        seed=99, frag_prob=1.2, n_families=12)
else:
    syn = hgsynth.Synth(genome_len=3_000_000, coverage=40.0, seed=321, n_families=12)
FIELDS = ("cmask", "flags", "anno_off", "anno_pos", "anno_type", "hinge_keep")
want = None
ok = True
for exchange in ("peer", "nccl"):
    arrays = ShardedArrays(syn.n_read, rank, world, dev, weights=syn.rlen, equal_slices=(exchange == "nccl"))
    novl = syn.generate(arrays.lo, arrays.hi, want_trace=False, threads=4)
    cols = {k: v.copy() for k, v in syn.cols().items()}
    ctx = api.Context(local, torch.cuda.current_stream().cuda_stream)
    ctx.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
    arrays.bind(ctx, exchange=exchange)
    ctx.set_overlaps(novl, cols, a_lo=arrays.lo, a_hi=arrays.hi)
    arrays.set_global_range(ctx, int(cols["aread"][0]), int(cols["aread"][-1]))
    for rep in range(3):  # several epochs of the arrival flags
        rc, summ = run_filter_sharded(ctx, api.FilterParams(), arrays)
        assert rc == 0
    mine = ctx.filter_fetch(int(summ.n_annotations))
    lists = gather_filter_lists(mine, arrays)  # the last exchange's lists feed the layout check below
    masks_all = arrays.gathered_mask(syn.n_read)
    part = {k: mine[k] for k in FIELDS}
    part.update(lo=arrays.lo, hi=arrays.hi, cov_est=summ.cov_est, min_cov=summ.min_cov,
                mask_all=arrays.gathered_mask(syn.n_read))
    parts = [None] * world
    dist.all_gather_object(parts, part)
    if rank == 0:
        if want is None:
            novl_all = syn.generate(0, syn.n_read, want_trace=False, threads=8)
            ref = api.Context(local, torch.cuda.current_stream().cuda_stream)
            ref.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
            ref.set_overlaps(novl_all, {k: v.copy() for k, v in syn.cols().items()})
            s1 = ref.filter(api.FilterParams())
            want = ref.filter_fetch(int(s1.n_annotations))
            ref.close()
        good = True
        for p in parts:
            lo, hi = p["lo"], p["hi"]
            good &= (p["cov_est"], p["min_cov"]) == (s1.cov_est, s1.min_cov)
            good &= bool(np.array_equal(p["mask_all"], want["mask"]))
            good &= bool(np.array_equal(p["cmask"][lo:hi], want["cmask"][lo:hi]))
            good &= bool(np.array_equal(p["flags"][lo:hi], want["flags"][lo:hi]))
            a0, a1 = want["anno_off"][lo], want["anno_off"][hi]
            good &= bool(np.array_equal(np.diff(p["anno_off"][lo:hi + 1]), np.diff(want["anno_off"][lo:hi + 1])))
            g0, g1 = p["anno_off"][lo], p["anno_off"][hi]
            for k in ("anno_pos", "anno_type", "hinge_keep"):
                good &= bool(np.array_equal(p[k][g0:g1], want[k][a0:a1]))
        print("SHARDED_CHECK", exchange, "OK" if good else "MISMATCH", "world", world, "ranges",
              [(p["lo"], p["hi"]) for p in parts], "reads", syn.n_read, "overlaps", novl_all, "annotations",
              int(s1.n_annotations), "hinges", int(want["hinge_keep"].sum()), flush=True)
        ok &= good
    ctx.close()
    dist.barrier()

# ---- maximal reads: classification and containment lists per shard, one NCCL exchange round
# (states + lists), resolve on every rank == hg_maximal on one context
arrays = ShardedArrays(syn.n_read, rank, world, dev, weights=syn.rlen)
novl = syn.generate(arrays.lo, arrays.hi, want_trace=True, threads=4)
cols = {k: v.copy() for k, v in syn.cols().items()}
toff, tr = syn.trace()
toff, tr = toff.copy(), tr.copy()
ctx = api.Context(local, torch.cuda.current_stream().cuda_stream)
ctx.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
ctx.set_overlaps(novl, cols, trace_off=toff, trace=tr, a_lo=arrays.lo, a_hi=arrays.hi)
masks = [masks_all]  # as every rank holds them after the sharded filter (checked against `want` above)
lp = api.LayoutParams()
got_max = run_maximal_sharded(ctx, lp, arrays, masks[0])
all_max = [None] * world
dist.all_gather_object(all_max, got_max)

# ---- layout on shards: the annotation / hinge lists all-gathered from the shards' filter results, the
# maximal bitmap from above
rep_csr, hin_csr = lists
my_edges, _ = run_layout_sharded(ctx, lp, arrays, masks[0], got_max, rep_csr, hin_csr)
ctx.close()
edge_key = lambda e: (e.a, e.b, e.length, e.comp, e.type, e.weight, tuple(e.eff_a), tuple(e.eff_b), tuple(e.raw_a),
                      tuple(e.raw_b), e.hinge_pos)
all_edges = [None] * world
dist.all_gather_object(all_edges, [edge_key(e) for e in my_edges])
if rank == 0:
    novl_all = syn.generate(0, syn.n_read, want_trace=True, threads=8)
    cols_all = {k: v.copy() for k, v in syn.cols().items()}
    toff_all, tr_all = syn.trace()
    ref = api.Context(local, torch.cuda.current_stream().cuda_stream)
    ref.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
    ref.set_overlaps(novl_all, cols_all, trace_off=toff_all, trace=tr_all)
    want_max, _, _ = ref.maximal(lp, masks[0])
    good = all(bool(np.array_equal(m, want_max)) for m in all_max)
    print("SHARDED_CHECK maximal", "OK" if good else "MISMATCH", "world", world, "maximal reads", int(want_max.sum()),
          "of", syn.n_read, flush=True)
    ok &= good
    keep = want["hinge_keep"].astype(bool)
    per_read = np.repeat(np.arange(syn.n_read), np.diff(want["anno_off"]))
    hin_off = np.zeros(syn.n_read + 1, np.int64)
    np.cumsum(np.bincount(per_read[keep], minlength=syn.n_read), out=hin_off[1:])
    want_edges, _ = ref.layout(lp, want["mask"], want_max, (want["anno_off"], want["anno_pos"], want["anno_type"]),
                               (hin_off, want["anno_pos"][keep], want["anno_type"][keep]))
    ref.close()
    got_edges = [e for part in all_edges for e in part]  # rank order = read order
    good = got_edges == [edge_key(e) for e in want_edges]
    print("SHARDED_CHECK layout", "OK" if good else "MISMATCH", "world", world, "edges", len(want_edges),
          "hinged", sum(1 for e in want_edges if e.hinge_pos >= 0), flush=True)
    ok &= good
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
