// Parameter blocks shared by the host front-end, the C-ABI and the kernels.
//
// The fields mirror the INI keys the reference stages read
//   filter : /root/reference/src/filter/filter.cpp:377-406
//   maximal: /root/reference/src/maximal/maximal.cpp:443-474
//   layout : /root/reference/src/layout/hinging.cpp:775-803
// and keep the reference's defaults (value when the key is absent).
#ifndef HG_PARAMS_H
#define HG_PARAMS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Overlap classes, numbered like the reference's enum MatchType
// (/root/reference/src/include/LAInterface.h:30-32) so debug dumps that print
// the raw enum value (edges.g_out.txt, hinging.cpp:1080-1090) stay identical.
enum hg_match_type {
    HG_FORWARD = 0,
    HG_BACKWARD = 1,
    HG_ACOVERB = 2,
    HG_BCOVERA = 3,
    HG_UNDEFINED = 4,
    HG_INTERNAL = 5,
    HG_NOT_ACTIVE = 6,
    HG_FORWARD_INTERNAL = 12,
    HG_BACKWARD_INTERNAL = 13,
    HG_NOT_CLASSIFIED = 255  // record was not among its pair's top two
};

typedef struct hg_filter_params {
    int32_t min_cov;        // [filter] min_cov (default -1); raised to cov_est/3
    int32_t cut_off;        // [filter] cut_off (default -1)
    int32_t theta;          // [filter] theta (default -1)
    int32_t est_cov;        // [filter] ec (default 0 = estimate)
    int32_t reso;           // fixed 40 (filter.cpp:386)
    int32_t use_qv_mask;    // [filter] use_qv && qual track present
    int32_t use_coverage_mask;  // [filter] coverage
    int32_t coverage_fraction;  // coverage_frac_repeat_annotation (3)
    int32_t min_repeat_annotation_threshold;  // 10
    int32_t max_repeat_annotation_threshold;  // 20
    int32_t repeat_annotation_gap_threshold;  // 300
    int32_t no_hinge_region;                  // 500
    int32_t hinge_min_support;                // 7
    int32_t hinge_bin_pileup_threshold;       // 7
    int32_t hinge_read_unbridged_threshold;   // 6
    int32_t hinge_bin_length;                 // 2 * hinge_tolerance_length
    int32_t hinge_tolerance_length;           // 100
    int32_t delete_telomere;                  // [layout] del_telomere (sic)
} hg_filter_params;

typedef struct hg_layout_params {
    int32_t length_threshold;  // [filter] length_threshold (-1)
    int32_t aln_threshold;     // [filter] aln_threshold (-1)
    int32_t theta;             // [filter] theta (-1)
    int32_t theta2;            // [filter] theta2 (0)
    int32_t use_two_matches;   // [layout] use_two_matches (1)
    int32_t hinge_slack;       // [layout] hinge_slack (1000)
    int32_t hinge_tolerance;   // [layout] hinge_tolerance (150)
    int32_t kill_hinge_overlap;   // [layout] kill_hinge_overlap (300)
    int32_t kill_hinge_internal;  // [layout] kill_hinge_internal (40)
    int32_t matching_hinge_slack;  // [layout] matching_hinge_slack (200)
    int32_t num_events_telomere;   // [layout] num_events_telomere (7)
    int32_t min_connected_component_size;  // [layout] (8)
    int32_t keep_only_maximal;     // keep_only_matches_between_maximal_reads (1)
    int32_t delete_telomeres;      // [layout] del_telomeres
} hg_layout_params;

#ifdef __cplusplus
}
#endif
#endif
