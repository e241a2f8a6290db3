"""Python mirror of the C ABI (include/hinge_b200.h).  Arrays are numpy arrays
(host) or anything with a `data_ptr()` (torch tensors; device or pinned host);
nothing here computes — every call goes straight into libhinge_b200.so."""
import ctypes as C
import sys

import numpy as np

from ._lib import (EdgeC, FilterParamsC, FilterSummaryC, GraphRecC, HingeError, LayoutParamsC, lib)

HG_MEM_HOST, HG_MEM_DEVICE = 0, 1
HG_OPT_KEEP_COVERAGE, HG_OPT_PROFILE, HG_OPT_SCATTER_SPREAD, HG_OPT_PROFILE_KERNEL, HG_OPT_KEEP_MASKS, HG_OPT_ANNO_POOL = 1, 2, 3, 4, 5, 6
HG_BUF_MEAN_COV, HG_BUF_MASK, HG_BUF_READ_FLAGS, HG_BUF_MEDIAN_HIST, HG_BUF_MASK_PACKED = 1, 2, 3, 4, 5
HG_RETRY_POOL = 1
HG_PEER_HANDLE_BYTES = 64


def FilterParams(**kw):
    """Defaults of the reference when a key is absent from the INI (filter.cpp:377-406)
    with utils/nominal.ini's values for the keys it sets."""
    p = FilterParamsC(min_cov=5, cut_off=300, theta=300, est_cov=0, reso=40, use_qv_mask=1,
                      use_coverage_mask=1, coverage_fraction=3, min_repeat_annotation_threshold=10,
                      max_repeat_annotation_threshold=20, repeat_annotation_gap_threshold=300,
                      no_hinge_region=500, hinge_min_support=7, hinge_bin_pileup_threshold=7,
                      hinge_read_unbridged_threshold=6, hinge_bin_length=200, hinge_tolerance_length=100,
                      delete_telomere=0)
    for k, v in kw.items():
        setattr(p, k, int(v))
    return p


def LayoutParams(**kw):
    p = LayoutParamsC(length_threshold=1000, aln_threshold=1000, theta=300, theta2=0, use_two_matches=1,
                      hinge_slack=1000, hinge_tolerance=150, kill_hinge_overlap=300, kill_hinge_internal=40,
                      matching_hinge_slack=200, num_events_telomere=7, min_connected_component_size=8,
                      keep_only_maximal=1, delete_telomeres=0)
    for k, v in kw.items():
        setattr(p, k, int(v))
    return p


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))


class Context:
    """One hg_ctx: one device, one stream."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        rc = lib.hg_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise HingeError("hg_ctx_create failed with status %d (no CUDA device? there is no CPU fallback)" % rc)
        self._h = h
        self.n_read = 0

    def close(self):
        if self._h:
            lib.hg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise HingeError("%s: status %d: %s" % (what, rc, lib.hg_last_error(self._h).decode()))
        return rc

    def set_option(self, option, value):
        self._check(lib.hg_set_option(self._h, option, int(value)), "hg_set_option")

    def set_reads(self, rlen, qv_off=None, qv=None, tspace=100):
        rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        self.n_read = len(rlen)
        self.tspace = int(tspace)
        if qv_off is not None:
            qv_off = np.ascontiguousarray(qv_off, dtype=np.int64)
            qv = np.ascontiguousarray(qv, dtype=np.uint8)
        self._check(lib.hg_set_reads(self._h, self.n_read, _ptr(rlen), _ptr(qv_off), _ptr(qv), int(tspace)),
                    "hg_set_reads")

    def set_overlaps(self, novl, cols, trace_off=None, trace=None, tbytes=1, where=HG_MEM_HOST, a_lo=0,
                     a_hi=None):
        """cols: dict with aread, bread, abpos, aepos, bbpos, bepos, flags (and optionally diffs)."""
        a_hi = self.n_read if a_hi is None else a_hi
        rc = lib.hg_set_overlaps(self._h, int(novl), _ptr(cols["aread"]), _ptr(cols["bread"]),
                                 _ptr(cols["abpos"]), _ptr(cols["aepos"]), _ptr(cols["bbpos"]),
                                 _ptr(cols["bepos"]), _ptr(cols.get("diffs")), _ptr(cols["flags"]),
                                 _ptr(trace_off), _ptr(trace), int(tbytes), int(where), int(a_lo), int(a_hi))
        self._check(rc, "hg_set_overlaps")
        self._keep = (cols, trace_off, trace)  # adopted device memory must outlive the context's use

    def set_global_range(self, first_aread, last_aread):
        self._check(lib.hg_set_global_range(self._h, int(first_aread), int(last_aread)), "hg_set_global_range")

    def peer_export(self, rank, world):
        """Allocates this rank's exchange block; returns its CUDA IPC handle (bytes)."""
        buf = C.create_string_buffer(HG_PEER_HANDLE_BYTES)
        self._check(lib.hg_peer_export(self._h, int(rank), int(world), buf), "hg_peer_export")
        return buf.raw

    def peer_connect(self, handles):
        """handles: the concatenated handles of all ranks, in rank order."""
        buf = C.create_string_buffer(bytes(handles), len(handles))
        self._check(lib.hg_peer_connect(self._h, buf), "hg_peer_connect")

    def peer_masks(self):
        out = np.zeros((self.n_read, 2), np.int32)
        self._check(lib.hg_peer_masks(self._h, _ptr(out)), "hg_peer_masks")
        return out

    def filter(self, params):
        s = FilterSummaryC()
        self._check(lib.hg_filter(self._h, C.byref(params), C.byref(s)), "hg_filter")
        return s

    def filter_phase1(self, params):
        self._check(lib.hg_filter_phase1(self._h, C.byref(params)), "hg_filter_phase1")

    def filter_phase2(self):
        self._check(lib.hg_filter_phase2(self._h), "hg_filter_phase2")

    def filter_phase3(self):
        s = FilterSummaryC()
        rc = self._check(lib.hg_filter_phase3(self._h, C.byref(s)), "hg_filter_phase3")
        return rc, s

    def filter_enqueue(self, params):
        """Launches the stage and returns (no wait); filter_finish() waits for the run enqueued last."""
        self._check(lib.hg_filter_enqueue(self._h, C.byref(params)), "hg_filter_enqueue")

    def filter_finish(self):
        s = FilterSummaryC()
        rc = self._check(lib.hg_filter_finish(self._h, C.byref(s)), "hg_filter_finish")
        return rc, s

    def filter_fetch(self, n_annotations):
        n = self.n_read
        out = {
            "mask": np.zeros((n, 2), np.int32), "cmask": np.zeros((n, 2), np.int32),
            "flags": np.zeros(n, np.uint8), "anno_off": np.zeros(n + 1, np.int64),
            "anno_pos": np.zeros(n_annotations + 1, np.int32), "anno_type": np.zeros(n_annotations + 1, np.int32),
            "hinge_keep": np.zeros(n_annotations + 1, np.uint8),
        }
        self._check(lib.hg_filter_fetch(self._h, _ptr(out["mask"]), _ptr(out["cmask"]), _ptr(out["flags"]),
                                        _ptr(out["anno_off"]), _ptr(out["anno_pos"]), _ptr(out["anno_type"]),
                                        _ptr(out["hinge_keep"])), "hg_filter_fetch")
        for k in ("anno_pos", "anno_type", "hinge_keep"):
            out[k] = out[k][:n_annotations]
        return out

    def filter_kernel_times(self):
        ms = (C.c_float * 4)()
        self._check(lib.hg_filter_kernel_times(self._h, ms, 4), "hg_filter_kernel_times")
        return {"profile": ms[0], "median": ms[1], "mask_anno": ms[2], "hinge_call": ms[3]}

    def device_buffer(self, which):
        p, b = C.c_void_p(), C.c_int64()
        self._check(lib.hg_device_buffer(self._h, which, C.byref(p), C.byref(b)), "hg_device_buffer")
        return p.value, b.value

    def bind_buffer(self, which, tensor):
        self._check(lib.hg_bind_buffer(self._h, which, _ptr(tensor), tensor.numel() * tensor.element_size()),
                    "hg_bind_buffer")
        self._bound = getattr(self, "_bound", []) + [tensor]

    def maximal(self, params, mask=None, want_contained_by=False):
        n = self.n_read
        mx = np.zeros(n, np.uint8)
        by = np.zeros(n, np.int32) if want_contained_by else None
        ms = C.c_float()
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
        self._check(lib.hg_maximal(self._h, C.byref(params), _ptr(mask), _ptr(mx), _ptr(by), C.byref(ms)),
                    "hg_maximal")
        return mx, by, ms.value

    def maximal_phase1(self, params, mask, state, unk, pool):
        """state: uint8[n_read] device tensor; unk: int32[cap, 4]; pool: int32[cap].  Returns (unknown reads,
        pool entries) -- if either exceeds the tensors' capacity, enlarge them and call again."""
        counts = (C.c_int32 * 2)()
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
        self._check(lib.hg_maximal_phase1(self._h, C.byref(params), _ptr(mask), _ptr(state), _ptr(unk),
                                          unk.shape[0], _ptr(pool), pool.shape[0], counts), "hg_maximal_phase1")
        return int(counts[0]), int(counts[1])

    def maximal_phase2(self, state_all, unk_all, counts_all, world, unk_stride, pool_all, pool_stride):
        mx = np.zeros(self.n_read, np.uint8)
        cc = (C.c_int32 * (2 * world))(*[int(x) for x in counts_all])
        self._check(lib.hg_maximal_phase2(self._h, _ptr(state_all), _ptr(unk_all), cc, int(world), int(unk_stride),
                                          _ptr(pool_all), int(pool_stride), _ptr(mx)), "hg_maximal_phase2")
        return mx

    def layout(self, params, mask, maximal, rep, hin):
        """rep / hin: (off[int64 n+1], pos[int32], type[int32]) CSR triples."""
        ms = C.c_float()
        args = [np.ascontiguousarray(mask, np.int32), np.ascontiguousarray(maximal, np.uint8)]
        for off, pos, typ in (rep, hin):
            args += [np.ascontiguousarray(off, np.int64), np.ascontiguousarray(pos, np.int32),
                     np.ascontiguousarray(typ, np.int32)]
        self._check(lib.hg_layout(self._h, C.byref(params), *[_ptr(a) for a in args], C.byref(ms)), "hg_layout")
        n = C.c_int64()
        self._check(lib.hg_layout_edges(self._h, None, 0, C.byref(n)), "hg_layout_edges")
        edges = (EdgeC * max(n.value, 1))()
        self._check(lib.hg_layout_edges(self._h, edges, n.value, C.byref(n)), "hg_layout_edges")
        return list(edges)[:n.value], ms.value


def _layout_args(mask, maximal, rep, hin):
    args = [np.ascontiguousarray(mask, np.int32), np.ascontiguousarray(maximal, np.uint8)]
    for off, pos, typ in (rep, hin):
        # pointers must be valid even for empty lists
        args += [np.ascontiguousarray(off, np.int64), np.ascontiguousarray(np.append(pos, 0), np.int32),
                 np.ascontiguousarray(np.append(typ, 0), np.int32)]
    return args


def layout_phase1(ctx, params, mask, maximal, rep, hin):
    """Sharded hg_layout, phase 1.  Returns the "[contained]" flags of the context's own reads (uint8[n_read])."""
    args = _layout_args(mask, maximal, rep, hin)
    contained = np.zeros(ctx.n_read, np.uint8)
    ctx._check(lib.hg_layout_phase1(ctx._h, C.byref(params), *[_ptr(a) for a in args], _ptr(contained)), "hg_layout_phase1")
    ctx._layout_keep = args
    return contained


def layout_phase2(ctx, contained_all, n_hinges):
    """Returns (alive flags of all hinges after this rank's kill pass, this rank's hinge-graph records as a
    structured numpy array)."""
    alive = np.ones(max(n_hinges, 1), np.uint8)
    n = C.c_int64()
    ctx._check(lib.hg_layout_phase2(ctx._h, _ptr(np.ascontiguousarray(contained_all, np.uint8)), _ptr(alive), C.byref(n)),
               "hg_layout_phase2")
    graph = np.zeros((max(n.value, 1), 10), np.int32)
    ctx._check(lib.hg_layout_graph(ctx._h, _ptr(graph), n.value), "hg_layout_graph")
    return alive[:n_hinges], graph[:n.value]


def layout_phase3(ctx, alive_all, graph_all):
    ms = C.c_float()
    alive_all = np.ascontiguousarray(np.append(alive_all, 1), np.uint8)
    graph_all = np.ascontiguousarray(graph_all, np.int32).reshape(-1, 10)
    g = np.ascontiguousarray(np.vstack([graph_all, np.zeros((1, 10), np.int32)]))
    ctx._check(lib.hg_layout_phase3(ctx._h, _ptr(alive_all), _ptr(g), len(graph_all), C.byref(ms)), "hg_layout_phase3")
    n = C.c_int64()
    ctx._check(lib.hg_layout_edges(ctx._h, None, 0, C.byref(n)), "hg_layout_edges")
    edges = (EdgeC * max(n.value, 1))()
    ctx._check(lib.hg_layout_edges(ctx._h, edges, n.value, C.byref(n)), "hg_layout_edges")
    return list(edges)[:n.value], ms.value


def launch_count():
    return int(lib.hg_launch_count())


def _main(fn, argv):
    argv = [a.encode() for a in argv]
    arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
    return fn(len(argv), arr)


def main_filter(argv):
    """`hinge filter` / Reads_filter (filter.cpp:168): argv as for the executable, argv[0] included."""
    return _main(lib.hg_main_filter, argv)


def main_maximal(argv):
    return _main(lib.hg_main_maximal, argv)


def main_layout(argv):
    return _main(lib.hg_main_layout, argv)


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else ""
    fn = {"filter": main_filter, "maximal": main_maximal, "layout": main_layout}.get(stage)
    if fn is None:
        sys.exit("usage: python -m hinge_b200.api filter|maximal|layout <flags of the reference executable>")
    sys.exit(fn([stage] + sys.argv[2:]))
