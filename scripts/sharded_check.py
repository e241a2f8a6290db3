"""Run under torchrun with N >= 2 ranks: the sharded filter (reads split by A-read id,
NCCL all-gathers between the phases) must reproduce the single-context result exactly."""
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
from hinge_b200.sharding import ShardedArrays, run_filter_sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
names = ["aread", "bread", "abpos", "aepos", "bbpos", "bepos", "flags"]
syn = hgsynth.Synth(genome_len=3_000_000, coverage=40.0, seed=321, n_families=12)
arrays = ShardedArrays(syn.n_read, rank, world, dev)
novl = syn.generate(arrays.lo, arrays.hi, want_trace=False, threads=4)
cols = {k: v.copy() for k, v in syn.cols().items()}
ctx = api.Context(local, torch.cuda.current_stream().cuda_stream)
ctx.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
arrays.bind(ctx)
ctx.set_overlaps(novl, cols, a_lo=arrays.lo, a_hi=arrays.hi)
rc, summ = run_filter_sharded(ctx, api.FilterParams(), arrays)
assert rc == 0
mine = ctx.filter_fetch(int(summ.n_annotations))
parts = [None] * world
dist.all_gather_object(parts, {k: mine[k] for k in ("cmask", "flags", "anno_off", "anno_pos", "anno_type", "hinge_keep")}
                       | {"lo": arrays.lo, "hi": arrays.hi, "cov_est": summ.cov_est, "min_cov": summ.min_cov})
ok = True
if rank == 0:
    novl_all = syn.generate(0, syn.n_read, want_trace=False, threads=8)
    ref = api.Context(local, torch.cuda.current_stream().cuda_stream)
    ref.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
    ref.set_overlaps(novl_all, {k: v.copy() for k, v in syn.cols().items()})
    s1 = ref.filter(api.FilterParams())
    want = ref.filter_fetch(int(s1.n_annotations))
    if arrays.packed:  # both bounds in units of gcd(40, tspace), 16 bits each
        import math

        pk = arrays.mask_pk[:syn.n_read].cpu().numpy().view(np.uint32)
        g = math.gcd(40, 100)
        got_mask = np.stack([(pk & 0xffff).astype(np.int32) * g, (pk >> 16).astype(np.int32) * g], axis=1)
    else:
        got_mask = arrays.mask[:syn.n_read].cpu().numpy()
    ok &= bool(np.array_equal(got_mask, want["mask"]))
    for p in parts:
        lo, hi = p["lo"], p["hi"]
        ok &= (p["cov_est"], p["min_cov"]) == (s1.cov_est, s1.min_cov)
        ok &= bool(np.array_equal(p["cmask"][lo:hi], want["cmask"][lo:hi]))
        ok &= bool(np.array_equal(p["flags"][lo:hi], want["flags"][lo:hi]))
        a0, a1 = want["anno_off"][lo], want["anno_off"][hi]
        ok &= bool(np.array_equal(np.diff(p["anno_off"][lo:hi + 1]), np.diff(want["anno_off"][lo:hi + 1])))
        for k in ("anno_pos", "anno_type", "hinge_keep"):
            ok &= bool(np.array_equal(p[k], want[k][a0:a1]))
    print("SHARDED_CHECK", "OK" if ok else "MISMATCH", "world", world, "reads", syn.n_read, "overlaps", novl_all,
          "annotations", int(s1.n_annotations), "hinges", int(want["hinge_keep"].sum()))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
