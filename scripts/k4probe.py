import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import numpy as np, torch
import hgsynth
from hinge_b200 import api
s = hgsynth.Synth(genome_len=50_000_000, coverage=50.0, read_mean=3500, read_sd=1500, read_min=1000, seed=1234)
novl = s.generate(want_trace=False, threads=16)
cols = s.cols()
ctx = api.Context(0, torch.cuda.current_stream().cuda_stream)
ctx.set_reads(s.rlen, s.qv_off, s.qv, 100)
ctx.set_option(api.HG_OPT_PROFILE, 1)
dev = {k: torch.from_numpy(cols[k]).cuda() for k in ["aread","bread","abpos","aepos","bbpos","bepos","flags"]}
ctx.set_overlaps(novl, dev, where=api.HG_MEM_DEVICE)
for i in range(3):
    summ = ctx.filter(api.FilterParams())
print("novl", novl, "reads", s.n_read, "anno", summ.n_annotations, "exact", summ.n_exact_order, ctx.filter_kernel_times())
res = ctx.filter_fetch(int(summ.n_annotations))
cnt = np.diff(res["anno_off"])
print("reads with anno", (cnt>0).sum(), "hinges", res["hinge_keep"].sum())
pile = np.bincount(cols["aread"], minlength=s.n_read)
print("pileup max", pile.max(), "mean", pile.mean(), "p99", np.percentile(pile, 99))

import ctypes as C
from hinge_b200._lib import lib
lib.hg_debug_item_log.argtypes=[C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
n=C.c_int64()
lib.hg_debug_item_log(ctx._h, None, 0, C.byref(n))
log=np.zeros((n.value,4),np.int32)
lib.hg_debug_item_log(ctx._h, C.c_void_p(log.ctypes.data), n.value, C.byref(n))
o=np.argsort(-log[:,1])
print("items", n.value, "total cycles", log[:,1].astype(np.int64).sum(), "median", np.median(log[:,1]))
print("top items (read, cycles, support, exact_n):")
print(log[o[:15]])
print("pileup of top:", pile[log[o[:15],0]])
