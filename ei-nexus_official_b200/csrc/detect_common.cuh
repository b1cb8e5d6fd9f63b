// Shared between the two detection kernels (detect.cu: shared-memory resident bands; detect_large.cu: maps
// too large for a cluster's shared memory).
#pragma once
#include "common.cuh"

// fp32 emulation of q = (n-k)/n, rank = q*(n-1) (detector_util.py:113-124; torch divides an int64 tensor by
// a Python int in fp32 and quantile scales q in the input dtype)
void einx_topk_ranks(int n, int k, int* lo, int* hi);

// Same contract as einx_detect (include/einx.h); used when no cluster of row bands can hold the map.
// `only_if` (optional, device, [B] int32): images whose entry is 0 are skipped -- the conditional redo behind the tiled kernel.
int einx_detect_large(einx_ctx* ctx, float* score, const uint8_t* mask, int B, int Hp, int Wp, int nms_radius,
                      int border, float prob_thresh, int top_k, float* nms_map, float* kpts, int kcap,
                      int32_t* counts, einx_stream stream_, const int32_t* only_if = nullptr);
