// BoxManager.record / record_corr replayed on the device (box_manager.py:40-129): ONE implementation, included by every
// kernel that edits fusion lists (sparse and dense NMS paths in bf_assoc.cu, the engine's correspondence tail in
// bf_engine.cu).
#pragma once
#include "bf_internal.cuh"

// box_manager.py:188-215 with the test of :55 / :71 (record) or :100-102 (record_corr: no centre term)
__device__ __forceinline__ bool bf_views_differ(const float* __restrict__ p1, const float* __restrict__ p2,
                                                float translation_gap, float rotation_gap, bool use_center,
                                                float center_dis, float center_gap) {
    const float dx = p2[3] - p1[3], dy = p2[7] - p1[7], dz = p2[11] - p1[11];
    const float baseline = sqrtf(dx * dx + dy * dy + dz * dz);
    float tr = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) tr += p2[4 * r] * p1[4 * r] + p2[4 * r + 1] * p1[4 * r + 1] + p2[4 * r + 2] * p1[4 * r + 2];
    const float c = fminf(fmaxf((tr - 1.f) * 0.5f, -1.f), 1.f);
    const float angle = acosf(c) * 180.f / 3.14159265358979323846f;
    return (baseline > translation_gap || angle > rotation_gap) || (use_center && center_dis > center_gap);
}

// sorted insert of `val` into list[0..len) (ascending; duplicates kept, like list.sort())
__device__ __forceinline__ void bf_sorted_insert(int32_t* list, int& len, int32_t val) {
    int k = len;
    while (k > 0 && list[k - 1] > val) { list[k] = list[k - 1]; --k; }
    list[k] = val;
    ++len;
}

// THE record function: BoxManager.record for one suppressed box `idx` of head `cur` (box_manager.py:48-86) and, with
// ctx.centers == nullptr, BoxManager.record_corr (box_manager.py:98-127: same list logic, no centre-distance term).
// lc/len_c: the head's list.  Returns true when the reference would swap `cur` out of keep in favour of `idx`
// (keep.remove(cur); keep.append(idx) / keep[keep == cur] = idx); the caller applies that to its keep representation.
// Used by every kernel that replays the fusion-list bookkeeping (sparse and dense NMS paths, the engine's
// correspondence tail).  A merge that would exceed BF_FUSION_CAP is skipped and reported in *status (sticky).
__device__ __forceinline__ bool bf_record_one(const bf_record_ctx& c, int cur, int idx, int32_t* lc, int& len_c) {
    float cdis = 0.f;
    const bool use_center = c.centers != nullptr;
    if (use_center) {
        const float* cc = c.centers + 3 * cur;
        const float* ci = c.centers + 3 * idx;
        const float ex = cc[0] - ci[0], ey = cc[1] - ci[1], ez = cc[2] - ci[2];
        cdis = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
    }
    const int len_i = c.flen[idx];
    const int32_t* li = c.fl + (size_t)idx * BF_FUSION_CAP;
    bool swap = false;
    if (len_i == 1) {                                     // box_manager.py:50-62
        const float* pi = c.poses + 16 * (size_t)c.init_id[idx];
        int cnt = 0;
        for (int k = 0; k < len_c; ++k)
            cnt += bf_views_differ(c.poses + 16 * (size_t)lc[k], pi, c.translation_gap, c.rotation_gap, use_center, cdis, c.center_gap);
        if (cnt == len_c && len_c < 5) {
            if (len_c + 1 > BF_FUSION_CAP) atomicExch(c.status, BF_ERR_CAPACITY);
            else bf_sorted_insert(lc, len_c, c.init_id[idx]);
        }
    } else {                                              // box_manager.py:65-86
        const float* pc = c.poses + 16 * (size_t)c.init_id[cur];
        int cnt = 0;
        for (int k = 0; k < len_i; ++k)
            cnt += bf_views_differ(c.poses + 16 * (size_t)li[k], pc, c.translation_gap, c.rotation_gap, use_center, cdis, c.center_gap);
        if (cnt == len_i && len_i < 5) {
            if (len_c + len_i > BF_FUSION_CAP) atomicExch(c.status, BF_ERR_CAPACITY);
            else for (int k = 0; k < len_i; ++k) bf_sorted_insert(lc, len_c, li[k]);
        } else swap = true;
        if (c.fflag[idx] == 1) c.fflag[cur] = 1;
    }
    return swap;
}

