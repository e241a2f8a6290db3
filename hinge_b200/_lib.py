"""ctypes loader for libhinge_b200.so with the prototypes of include/hinge_b200.h."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libhinge_b200.so")


class HingeError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "hinge_b200: %s is missing. Build it with `python -m hinge_b200.build` (nvcc, sm_100a). "
        "There is no Python/CPU fallback for the CUDA path." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

i32, i64, u8p = C.c_int32, C.c_int64, C.POINTER(C.c_uint8)
i32p, i64p, vp, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_float)


class FilterParamsC(C.Structure):
    _fields_ = [(n, i32) for n in (
        "min_cov", "cut_off", "theta", "est_cov", "reso", "use_qv_mask", "use_coverage_mask",
        "coverage_fraction", "min_repeat_annotation_threshold", "max_repeat_annotation_threshold",
        "repeat_annotation_gap_threshold", "no_hinge_region", "hinge_min_support",
        "hinge_bin_pileup_threshold", "hinge_read_unbridged_threshold", "hinge_bin_length",
        "hinge_tolerance_length", "delete_telomere")]


class LayoutParamsC(C.Structure):
    _fields_ = [(n, i32) for n in (
        "length_threshold", "aln_threshold", "theta", "theta2", "use_two_matches", "hinge_slack",
        "hinge_tolerance", "kill_hinge_overlap", "kill_hinge_internal", "matching_hinge_slack",
        "num_events_telomere", "min_connected_component_size", "keep_only_maximal", "delete_telomeres")]


class FilterSummaryC(C.Structure):
    _fields_ = [("r_begin", i32), ("r_end", i32), ("cov_est", i32), ("min_cov", i32),
                ("n_annotations", i64), ("n_hinges", i64), ("ms_device", C.c_float),
                ("n_exact_order", i32)]


class GraphRecC(C.Structure):
    _fields_ = [("owner", i32), ("seq", i32), ("f", i32 * 4), ("flag", i32), ("rev", i32), ("u", i32), ("v", i32)]


class EdgeC(C.Structure):
    _fields_ = [(n, i32) for n in ("a", "b", "length", "comp", "type", "weight")] + \
               [(n, i32 * 2) for n in ("eff_a", "eff_b", "read_a", "read_b", "raw_a", "raw_b")] + \
               [("hinge_pos", i32)]


# every symbol include/hinge_b200.h declares: (restype, argtypes)
PROTOTYPES = {
    "hg_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "hg_ctx_destroy": (None, [vp]),
    "hg_last_error": (C.c_char_p, [vp]),
    "hg_version": (C.c_char_p, []),
    "hg_set_option": (C.c_int, [vp, C.c_int, i64]),
    "hg_device_buffer": (C.c_int, [vp, C.c_int, C.POINTER(vp), i64p]),
    "hg_bind_buffer": (C.c_int, [vp, C.c_int, vp, i64]),
    "hg_filter_kernel_times": (C.c_int, [vp, f32p, C.c_int]),
    "hg_launch_count": (i64, []),
    "hg_set_reads": (C.c_int, [vp, i32, vp, vp, vp, i32]),
    "hg_set_overlaps": (C.c_int, [vp, i64] + [vp] * 8 + [vp, vp, i32, i32, i32, i32]),
    "hg_set_global_range": (C.c_int, [vp, i32, i32]),
    "hg_peer_export": (C.c_int, [vp, i32, i32, vp]),
    "hg_peer_connect": (C.c_int, [vp, vp]),
    "hg_peer_connect_local": (C.c_int, [C.POINTER(vp), i32]),
    "hg_peer_masks": (C.c_int, [vp, vp]),
    "hg_filter": (C.c_int, [vp, C.POINTER(FilterParamsC), C.POINTER(FilterSummaryC)]),
    "hg_filter_phase1": (C.c_int, [vp, C.POINTER(FilterParamsC)]),
    "hg_filter_phase2": (C.c_int, [vp]),
    "hg_filter_phase3": (C.c_int, [vp, C.POINTER(FilterSummaryC)]),
    "hg_filter_enqueue": (C.c_int, [vp, C.POINTER(FilterParamsC)]),
    "hg_filter_finish": (C.c_int, [vp, C.POINTER(FilterSummaryC)]),
    "hg_filter_fetch": (C.c_int, [vp] + [vp] * 7),
    "hg_filter_coverage": (C.c_int, [vp, vp, vp, i64p]),
    "hg_maximal": (C.c_int, [vp, C.POINTER(LayoutParamsC), vp, vp, vp, f32p]),
    "hg_maximal_phase1": (C.c_int, [vp, C.POINTER(LayoutParamsC), vp, vp, vp, i64, vp, i64, i32p]),
    "hg_maximal_phase2": (C.c_int, [vp, vp, vp, i32p, i32, i64, vp, i64, vp]),
    "hg_layout": (C.c_int, [vp, C.POINTER(LayoutParamsC)] + [vp] * 8 + [f32p]),
    "hg_layout_edges": (C.c_int, [vp, C.POINTER(EdgeC), i64, i64p]),
    "hg_layout_phase1": (C.c_int, [vp, C.POINTER(LayoutParamsC)] + [vp] * 8 + [vp]),
    "hg_layout_phase2": (C.c_int, [vp, vp, vp, i64p]),
    "hg_layout_graph": (C.c_int, [vp, vp, i64]),
    "hg_layout_phase3": (C.c_int, [vp, vp, vp, i64, f32p]),
    "hg_main_filter": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "hg_main_maximal": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "hg_main_layout": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "hg_main_exit_after": (None, [C.c_int]),
}

MISSING = []
for _name, (_res, _args) in PROTOTYPES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)
        continue
    _fn.restype = _res
    _fn.argtypes = _args
