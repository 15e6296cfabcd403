// Pieces shared by the row-wise softmax-KL kernels (kl_rows.cu: register-resident single pass;
// kl_rows_stream.cu: streaming two-phase kernel for rows split over several CTAs).
#pragma once

#include "common.cuh"
#include "params.h"

namespace sd {

constexpr unsigned kSpinLimit = 1u << 22;
constexpr int kRedFloats = 8;                     // per-warp record of kl_rows.cu: Ms, Mt, Zs, Zt, A, SQ, DD
constexpr float kPadValue = -1.0e30f;             // stands in for elements a partial chunk does not have
constexpr float kMaxFloor = -1.0e29f;             // floor of a thread's local maximum: a thread that holds only
                                                  // padding then exponentiates to exact zeros (not to exp2 of
                                                  // the rounding error of kPadValue * c2)

struct Unit {
    int b, grp, ck, nch;
    int e0;   // first logical row element of this chunk
    int len;  // elements in this chunk
};

// unit r of sample b (r < units_per_sample)
__device__ __forceinline__ Unit decode_unit(const RowsParams& p, int b, int r) {
    Unit x;
    x.b = b;
    const int full_units = p.G_full * p.nch_full;
    int g_real;
    if (r < full_units) {
        if (p.nch_full == 1) {
            x.grp = r;
            x.ck = 0;
        } else {
            x.grp = r / p.nch_full;
            x.ck = r - x.grp * p.nch_full;
        }
        x.nch = p.nch_full;
        g_real = p.l[0].g;
    } else {
        x.grp = p.G_full;
        x.ck = r - full_units;
        x.nch = p.nch_last;
        g_real = p.g_last;
    }
    const int L = g_real * p.HW;
    x.e0 = x.ck * p.chunk_elems;
    x.len = min(p.chunk_elems, L - x.e0);
    return x;
}

// walks the units of one CTA (u = blockIdx.x, += gridDim.x) without a 64-bit division per unit
struct UnitCursor {
    long long u;
    int b, r;
    __device__ __forceinline__ void init(const RowsParams& p, long long u0) {
        u = u0;
        b = (int)(u0 / p.units_per_sample);
        r = (int)(u0 - (long long)b * p.units_per_sample);
    }
    __device__ __forceinline__ void advance(const RowsParams& p, int step) {
        u += step;
        r += step;
        while (r >= p.units_per_sample) {
            r -= p.units_per_sample;
            ++b;
        }
    }
};

// first unit (within the sample) of l[0] row j; j == number of rows gives the end
__device__ __forceinline__ int unit_start(const RowsParams& p, int j) {
    return j <= p.G_full ? j * p.nch_full : p.units_per_sample;
}

// global element offset of logical row element e of a gathered row
__device__ __forceinline__ size_t perm_elem_offset(const RowsParams& p, const Unit& x, int e) {
    const int j = e / p.HW;
    const int pos = e - j * p.HW;
    const int ch = p.perm[x.grp * p.l[0].g + j];
    return ((size_t)x.b * p.C + ch) * p.HW + pos;
}

// softmax statistics of a piece of a row: raw-value maxima (ms, mt), sums relative to them, and the two sums the
// KL is built from (common.cuh, "KL without cancellation"): a = sum et (at - as), dd = sum (et - es)
struct RowStat {
    float ms, zs, mt, zt, a, dd;
};
__device__ __forceinline__ RowStat rowstat_empty() { return RowStat{-INFINITY, 0.f, -INFINITY, 0.f, 0.f, 0.f}; }
__device__ __forceinline__ RowStat rowstat_merge(const RowStat& x, const RowStat& y, float c2) {
    RowStat r;
    r.ms = fmaxf(x.ms, y.ms);
    r.mt = fmaxf(x.mt, y.mt);
    const bool hx = x.zs > 0.f || x.zt > 0.f, hy = y.zs > 0.f || y.zt > 0.f;     // (an empty part has infinite references)
    const float fxs = x.zs > 0.f ? ref_factor(x.ms, r.ms, c2) : 0.f;
    const float fys = y.zs > 0.f ? ref_factor(y.ms, r.ms, c2) : 0.f;
    const float fxt = x.zt > 0.f ? ref_factor(x.mt, r.mt, c2) : 0.f;
    const float fyt = y.zt > 0.f ? ref_factor(y.mt, r.mt, c2) : 0.f;
    const float gx = hx ? merge_shift(x.ms, x.mt, r.ms, r.mt, c2) : 0.f;
    const float gy = hy ? merge_shift(y.ms, y.mt, r.ms, r.mt, c2) : 0.f;
    r.zs = __fadd_rn(__fmul_rn(x.zs, fxs), __fmul_rn(y.zs, fys));
    r.zt = __fadd_rn(__fmul_rn(x.zt, fxt), __fmul_rn(y.zt, fyt));
    r.a = __fadd_rn(fmaf(__fmul_rn(x.zt, fxt), gx, __fmul_rn(x.a, fxt)), fmaf(__fmul_rn(y.zt, fyt), gy, __fmul_rn(y.a, fyt)));
    r.dd = __fadd_rn(fmaf(x.zs, factor_diff(fxs, fxt, gx), __fmul_rn(x.dd, fxt)),
                     fmaf(y.zs, factor_diff(fys, fyt, gy), __fmul_rn(y.dd, fyt)));
    return r;
}

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float sum16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float max16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}


}  // namespace sd
