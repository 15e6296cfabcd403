// Softmax-KL losses on maps that the reference first RESIZES to the label size - channel mode (CD / CGD: this
// comment) and pixel mode (PD: kl_pixels_up_kernel further down):
// KLDLoss.resize (mmseg/models/distillation/losses.py:25-33, F.interpolate(..., mode='bilinear',
// align_corners=False), ops/wrappers.py:8-29) followed by the transform + KL chain (:50-58, :108-112) and autograd's
// backward through both.  Every shipped preset resizes logits at 1/4 or 1/8 resolution to 512x512 (SURVEY.md 8 a8 /
// f1): the reference then streams 16-64x the data through ~20 kernels and keeps several full-resolution temporaries.
//
// Here the up-sampled maps never exist.  For an integer scale s (2, 4, 8) the s x s block of up-sampled values
// that belongs to one low-resolution cell depends on the 3 x 3 cells around it with CONSTANT weights
// ((k + 1/2)/s -+ 1/2), so a thread regenerates the block in registers from shared memory:
//
//   kernel 1 (statistics)  unit = (sample, channel, strip of low-res rows): load the strip (+ one halo row each
//                          side) of S and T, reference = the strip maximum (a unit whose sums underflow against
//                          it is redone against the exact maximum of its up-sampled values), sums of exp2 over
//                          the up-sampled block of every cell -> one partial record per unit.
//   kernel 2 (gradient)    merges the records of its row (g channels x strips) by log-sum-exp into the row statistics,
//                          regenerates the block, g = coef (q - p), and applies the TRANSPOSED stencil: a cell's
//                          block contributes a 3 x 3 matrix W_y^T g W_x to the cells around it; every thread
//                          stores its nine contributions to nine shared-memory planes (no atomics), then every
//                          output cell sums its nine planes in a fixed order.  A strip also computes the cell rows
//                          just outside it, so no gradient crosses CTAs.  HBM traffic: the low-resolution maps
//                          twice (L2-resident the second time) + dS once.
//
// Deterministic (no floating-point atomics); fp32 arithmetic; dS has the dtype and the (low) resolution of S.
#include <type_traits>

#include "up_common.cuh"
#include "launch.h"

namespace sd {

constexpr int kUpThreads = 256;

// unit -> plane and strip
struct UpUnit {
    int b, lc, grp, j;  // sample, logical channel, row of the loss within the sample, channel within the row
    int ch;             // physical channel (through the permutation)
    int i0, i1;         // low-res rows [i0, i1) of the strip
};
__device__ __forceinline__ UpUnit up_unit(const UpParams& p, long long u) {
    UpUnit x;
    const int plane = (int)(u / p.NS);
    const int strip = (int)(u - (long long)plane * p.NS);
    x.b = plane / p.C;
    x.lc = plane - x.b * p.C;
    x.grp = x.lc / p.g;
    x.j = x.lc - x.grp * p.g;
    x.ch = p.perm ? p.perm[x.lc] : x.lc;
    x.i0 = strip * p.SR;
    x.i1 = min(p.Hl, x.i0 + p.SR);
    return x;
}

// strip rows [i0 - halo, i1 + halo) (clamped to the plane) of S and T -> shared memory; returns the maxima of what
// was loaded.  SCALED: stored as (x - ref) * c2, the exponent the kernels need (interpolation is affine-invariant:
// the weights of a pixel sum to one).
template <typename T, bool SCALED>
__device__ __forceinline__ void up_load_strip(const UpParams& p, const UpUnit& x, int halo, float* sS, float* sT, int& r0,
                                              int& nr, float& ms, float& mt, float ref_s = 0.f, float ref_t = 0.f) {
    r0 = max(0, x.i0 - halo);
    const int r1 = min(p.Hl, x.i1 + halo);
    nr = r1 - r0;
    const size_t base = ((size_t)x.b * p.C + x.ch) * (size_t)p.Hl * p.Wl + (size_t)r0 * p.Wl;
    const T* gs = static_cast<const T*>(p.S) + base;
    const T* gt = static_cast<const T*>(p.T) + base;
    ms = kUpFloor;
    mt = kUpFloor;
    for (int e = threadIdx.x; e < nr * p.Wl; e += kUpThreads) {
        const float a = up_load<T>(gs + e), b = up_load<T>(gt + e);
        sS[e] = SCALED ? (a - ref_s) * p.c2 : a;
        sT[e] = SCALED ? (b - ref_t) * p.c2 : b;
        ms = fmaxf(ms, a);
        mt = fmaxf(mt, b);
    }
}
// cell index -> (row, column) without an integer division (c < 2^20)
__device__ __forceinline__ void up_cell(int c, int Wl, float inv_Wl, int& r, int& j) {
    r = (int)(((float)c + 0.5f) * inv_Wl);
    j = c - r * Wl;
}

// the 3 x 3 neighbourhood of cell (i, j) with clamped indices, from the strip in shared memory
__device__ __forceinline__ void up_nbhd(const float* sm, int Wl, int Hl, int r0, int i, int j, float (&a)[3][3]) {
    const int im = max(i - 1, 0) - r0, ic = i - r0, ip = min(i + 1, Hl - 1) - r0;
    const int jm = max(j - 1, 0), jp = min(j + 1, Wl - 1);
    a[0][0] = sm[im * Wl + jm]; a[0][1] = sm[im * Wl + j]; a[0][2] = sm[im * Wl + jp];
    a[1][0] = sm[ic * Wl + jm]; a[1][1] = sm[ic * Wl + j]; a[1][2] = sm[ic * Wl + jp];
    a[2][0] = sm[ip * Wl + jm]; a[2][1] = sm[ip * Wl + j]; a[2][2] = sm[ip * Wl + jp];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < kUpThreads / 32; ++w) r = fmaxf(r, red[w]);
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < kUpThreads / 32; ++w) r += red[w];
    return r;
}

// ------------------------------------------------------------------------------------------ kernel 1
template <typename T, int S>
__global__ void __launch_bounds__(kUpThreads) kl_rows_up_stats_kernel(const UpParams p) {
    extern __shared__ __align__(16) float up_smem[];
    float* sS = up_smem;
    float* sT = sS + (p.SR + 2) * p.Wl;
    __shared__ float red[kUpThreads / 32];
    for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
        const UpUnit x = up_unit(p, u);
        int r0, nr;
        float ms, mt;
        __syncthreads();   // the previous unit's readers are done with the strip
        up_load_strip<T, false>(p, x, 1, sS, sT, r0, nr, ms, mt);
        // reference of the exponentials: the maximum of the strip - an up-sampled value is a convex combination of
        // cells, so never larger.  (If it is so much smaller everywhere that the sums underflow - an isolated spike
        // between cells - the unit is redone below against the exact maximum.)
        ms = block_max(ms, red);
        mt = block_max(mt, red);
        float zs, zt, acc, dd;     // dd = sum (et - es) term by term (common.cuh: KL without cancellation)
        for (int attempt = 0;; ++attempt) {
            // sweep over the (small) strip: values -> exponents relative to the references (the redo reloads the raw
            // values: exponents thousands below the old reference have lost their low bits)
            if (attempt > 0) {
                float d0, d1;
                up_load_strip<T, false>(p, x, 1, sS, sT, r0, nr, d0, d1);
                __syncthreads();
            }
            for (int e = threadIdx.x; e < nr * p.Wl; e += kUpThreads) {
                sS[e] = (sS[e] - ms) * p.c2;
                sT[e] = (sT[e] - mt) * p.c2;
            }
            __syncthreads();
            zs = 0.f;
            zt = 0.f;
            acc = 0.f;
            dd = 0.f;
            const int ncell = (x.i1 - x.i0) * p.Wl;
            for (int c = threadIdx.x; c < ncell; c += kUpThreads) {
                int i, j;
                up_cell(c, p.Wl, p.inv_Wl, i, j);
                i += x.i0;
                float a[3][3], hs[3][S], ht[3][S], ds_[2][S], dt_[2][S];
                up_nbhd(sS, p.Wl, p.Hl, r0, i, j, a);
                up_hrows<S>(a, hs);
                up_nbhd(sT, p.Wl, p.Hl, r0, i, j, a);
                up_hrows<S>(a, ht);
                up_vdiff<S>(hs, ds_);
                up_vdiff<S>(ht, dt_);
#pragma unroll
                for (int ky = 0; ky < S; ++ky) {
#pragma unroll
                    for (int kx = 0; kx < S; ++kx) {
                        const float vs = up_value<S>(hs, ds_, ky, kx), vt = up_value<S>(ht, dt_, ky, kx);
                        const float es = fast_exp2(vs), et = fast_exp2(vt);
                        zs += es;
                        zt += et;
                        acc = fmaf(et, vt - vs, acc);
                        dd += et - es;
                    }
                }
            }
            zs = block_sum(zs, red);
            zt = block_sum(zt, red);
            acc = block_sum(acc, red);
            dd = block_sum(dd, red);
            if (attempt > 0 || (zs >= 1e-20f && zt >= 1e-20f)) break;      // (uniform over the CTA)
            // ---- rare: EXACT maximum of the up-sampled values of my cells.  Between four cell centres the interpolant
            // is bilinear, and a bilinear patch takes its extremes at the corners of any axis-aligned rectangle of
            // samples: only the samples next to the cell centres can be the maximum - plus, on the first and last row
            // of the strip, the outermost sample rows (the rectangle is cut there).  (Values in shared memory are
            // exponents relative to the old references: maxima and shifts are taken in that domain.)
            float es_max = -3.0e38f, et_max = -3.0e38f;
            constexpr int KC0 = S / 2 - 1, KC1 = S / 2;
            for (int c = threadIdx.x; c < ncell; c += kUpThreads) {
                int i, j;
                up_cell(c, p.Wl, p.inv_Wl, i, j);
                i += x.i0;
                float as[3][3], at[3][3];
                up_nbhd(sS, p.Wl, p.Hl, r0, i, j, as);
                up_nbhd(sT, p.Wl, p.Hl, r0, i, j, at);
                auto sample = [](const float (&a)[3][3], int ky, int kx) {
                    const int fy = UpW<S>::first(ky) + 1, fx = UpW<S>::first(kx) + 1;
                    const float wy = UpW<S>::w1(ky), wx = UpW<S>::w1(kx);
                    const float top = fmaf(wx, a[fy][fx + 1] - a[fy][fx], a[fy][fx]);
                    const float bot = fmaf(wx, a[fy + 1][fx + 1] - a[fy + 1][fx], a[fy + 1][fx]);
                    return fmaf(wy, bot - top, top);
                };
#pragma unroll
                for (int kx = KC0; kx <= KC1; ++kx) {
#pragma unroll
                    for (int ky = KC0; ky <= KC1; ++ky) {
                        es_max = fmaxf(es_max, sample(as, ky, kx));
                        et_max = fmaxf(et_max, sample(at, ky, kx));
                    }
                    if (i == x.i0) {
                        es_max = fmaxf(es_max, sample(as, 0, kx));
                        et_max = fmaxf(et_max, sample(at, 0, kx));
                    }
                    if (i == x.i1 - 1) {
                        es_max = fmaxf(es_max, sample(as, S - 1, kx));
                        et_max = fmaxf(et_max, sample(at, S - 1, kx));
                    }
                }
            }
            es_max = block_max(es_max, red);
            et_max = block_max(et_max, red);
            ms += es_max * p.inv_c2;          // the new references, back in the value domain (any value close to the
            mt += et_max * p.inv_c2;          // true maximum serves: what matters is that it is used consistently)
            __syncthreads();
        }
        // (acc = sum et (at - as) stays in the exponent domain: common.cuh, KL without cancellation)
        if (threadIdx.x == 0) {
            float4* rec = reinterpret_cast<float4*>(p.part + u * 8);
            rec[0] = make_float4(ms, mt, zs, zt);
            rec[1] = make_float4(acc, dd, 0.f, 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------ kernel 2
template <typename T, int S>
__global__ void __launch_bounds__(kUpThreads) kl_rows_up_grad_kernel(const UpParams p) {
    extern __shared__ __align__(16) float up_smem[];
    float* sS = up_smem;
    float* sT = sS + (p.SR + 4) * p.Wl;
    float* planes = sT + (p.SR + 4) * p.Wl;          // [9][SR][Wl + 2]: column tj lives at tj + 1, so tj = -1 and Wl land in the pad
    __shared__ float rowstat[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pstride = p.Wl + 2;
    const int plane_sz = p.SR * pstride;
    for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
        const UpUnit x = up_unit(p, u);
        __syncthreads();   // the previous unit is done with the shared memory
        // ---- the statistics of my row: merge the records of its channels x strips (warp 0, fixed order)
        if (warp == 0) {
            const int g_real = min(p.g, p.C - x.grp * p.g);
            const long long first = ((long long)x.b * p.C + (long long)x.grp * p.g) * p.NS;
            const int nrec = g_real * p.NS;
            // row reference = the largest log-sum-exp of its units: the merged sums then lie in [1, number of units]
            // whatever reference a unit used, and no up-sampled value exceeds it.  Taken relative to the largest unit
            // reference so that nothing is rounded at the magnitude of the values themselves.
            float bs = kUpFloor, bt = kUpFloor;
            for (int r = lane; r < nrec; r += 32) {
                const float4 q = *reinterpret_cast<const float4*>(p.part + (first + r) * 8);
                bs = fmaxf(bs, q.x);
                bt = fmaxf(bt, q.y);
            }
            bs = warp_max(bs);
            bt = warp_max(bt);
            float ls = -3.0e38f, lt = -3.0e38f;
            for (int r = lane; r < nrec; r += 32) {
                const float4 q = *reinterpret_cast<const float4*>(p.part + (first + r) * 8);
                ls = fmaxf(ls, fmaf(q.x - bs, p.c2, log2f(q.z)));
                lt = fmaxf(lt, fmaf(q.y - bt, p.c2, log2f(q.w)));
            }
            ls = warp_max(ls);
            lt = warp_max(lt);
            const float ms = bs + ls * p.inv_c2, mt = bt + lt * p.inv_c2;     // used consistently from here on
            float zs = 0.f, zt = 0.f, acc = 0.f, dd = 0.f;
            for (int r = lane; r < nrec; r += 32) {
                const float4 q0 = *reinterpret_cast<const float4*>(p.part + (first + r) * 8);
                const float2 q1 = *reinterpret_cast<const float2*>(p.part + (first + r) * 8 + 4);
                const float fs = fast_exp2((q0.x - ms) * p.c2), ft = fast_exp2((q0.y - mt) * p.c2);
                // shift of the exponent gap from the unit's references to the row's (compensated: exact)
                const float gx = merge_shift2(q0.x, q0.y, ms, mt) * p.c2;
                const float ztf = q0.w * ft;
                zs = fmaf(q0.z, fs, zs);
                zt += ztf;
                acc += fmaf(ztf, gx, q1.x * ft);                               // a2 = sum ft (a2_u + zt_u x_u)
                dd += fmaf(q0.z, factor_diff(fs, ft, gx), q1.y * ft);          // zt ft - zs fs = dd ft + zs (ft - fs)
            }
            zs = warp_sum(zs);
            zt = warp_sum(zt);
            acc = warp_sum(acc);
            dd = warp_sum(dd);
            if (lane == 0) {
                rowstat[0] = ms;
                rowstat[1] = mt;
                rowstat[2] = p.coef / zs;
                rowstat[3] = p.coef / zt;
                if (x.j == 0 && x.i0 == 0) {
                    // KL(p||q) = sum p (t - s)/tau - (lse_t - lse_s), once per row
                    const float kl = kl_from_stats(zs, zt, acc, dd);
                    p.row_kl[x.b * p.G + x.grp] = kl;
                }
            }
        }
        __syncthreads();
        const float gsc = rowstat[2], gtc = rowstat[3];
        int r0, nr;
        float dum0, dum1;
        up_load_strip<T, true>(p, x, 2, sS, sT, r0, nr, dum0, dum1, rowstat[0], rowstat[1]);
        __syncthreads();
        // ---- cells of the strip and of the cell rows just outside it: block gradient -> nine contributions
        const int ca = max(0, x.i0 - 1), cb = min(p.Hl, x.i1 + 1);
        const int ncell = (cb - ca) * p.Wl;
        for (int c = threadIdx.x; c < ncell; c += kUpThreads) {
            int i, j;
            up_cell(c, p.Wl, p.inv_Wl, i, j);
            i += ca;
            float a[3][3], hs[3][S], ht[3][S], ds_[2][S], dt_[2][S];
            up_nbhd(sS, p.Wl, p.Hl, r0, i, j, a);
            up_hrows<S>(a, hs);
            up_nbhd(sT, p.Wl, p.Hl, r0, i, j, a);
            up_hrows<S>(a, ht);
            up_vdiff<S>(hs, ds_);
            up_vdiff<S>(ht, dt_);
            float m[3][3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int e = 0; e < 3; ++e) m[d][e] = 0.f;
#pragma unroll
            for (int ky = 0; ky < S; ++ky) {
                float tr[3] = {0.f, 0.f, 0.f};      // this block row folded onto the three low-res columns
#pragma unroll
                for (int kx = 0; kx < S; ++kx) {
                    const float es = fast_exp2(up_value<S>(hs, ds_, ky, kx)), et = fast_exp2(up_value<S>(ht, dt_, ky, kx));
                    const float gv = es * gsc - et * gtc;
                    const int f = UpW<S>::first(kx) + 1;
                    const float w1 = UpW<S>::w1(kx);
                    tr[f] = fmaf(1.f - w1, gv, tr[f]);
                    tr[f + 1] = fmaf(w1, gv, tr[f + 1]);
                }
                const int f = UpW<S>::first(ky) + 1;
                const float w1 = UpW<S>::w1(ky);
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    m[f][e] = fmaf(1.f - w1, tr[e], m[f][e]);
                    m[f + 1][e] = fmaf(w1, tr[e], m[f + 1][e]);
                }
            }
            // taps clamped at the border of the plane fall onto the cell itself
            if (i == 0) {
#pragma unroll
                for (int e = 0; e < 3; ++e) { m[1][e] += m[0][e]; m[0][e] = 0.f; }
            }
            if (i == p.Hl - 1) {
#pragma unroll
                for (int e = 0; e < 3; ++e) { m[1][e] += m[2][e]; m[2][e] = 0.f; }
            }
            if (j == 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) { m[d][1] += m[d][0]; m[d][0] = 0.f; }
            }
            if (j == p.Wl - 1) {
#pragma unroll
                for (int d = 0; d < 3; ++d) { m[d][1] += m[d][2]; m[d][2] = 0.f; }
            }
            // (contributions that left the plane were folded above and are zero; the pad columns swallow them)
            float* pl = planes + (i - 1 - x.i0) * pstride + j;           // target (i - 1, j - 1) at column j - 1 + 1
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int ti = i + d - 1;
                if (ti >= x.i0 && ti < x.i1) {                             // (uniform over a warp when Wl % 32 == 0)
#pragma unroll
                    for (int e = 0; e < 3; ++e) pl[(d * 3 + e) * plane_sz + d * pstride + e] = m[d][e];
                }
            }
        }
        __syncthreads();
        // ---- every output cell: its nine contributions in a fixed order
        T* out = static_cast<T*>(p.dS) + ((size_t)x.b * p.C + x.ch) * (size_t)p.Hl * p.Wl + (size_t)x.i0 * p.Wl;
        const int nout = (x.i1 - x.i0) * p.Wl;
        for (int c = threadIdx.x; c < nout; c += kUpThreads) {
            int ti, tj;
            up_cell(c, p.Wl, p.inv_Wl, ti, tj);
            ti += x.i0;
            // plane (d, e) holds what cell (ti - d + 1, tj - e + 1) sent here; where that cell does not exist
            // (outside the map) nothing was stored
            const float* pl = planes + (ti - x.i0) * pstride + tj + 1;
            const bool okd[3] = {ti + 1 < p.Hl, true, ti >= 1};
            const bool oke[3] = {tj + 1 < p.Wl, true, tj >= 1};
            float v = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const float q = pl[(d * 3 + e) * plane_sz];
                    v += (okd[d] && oke[e]) ? q : 0.f;
                }
            }
            up_store<T>(out + c, v);
        }
    }
    // ---- loss: the last CTA sums the row terms in a fixed order
    __shared__ unsigned ticket_s;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        ticket_s = atomicAdd(&p.ctrl[0], 1u);
    }
    __syncthreads();
    if (ticket_s == gridDim.x - 1) {
        __threadfence();
        double acc = 0.0;
        for (int r = threadIdx.x; r < p.R; r += kUpThreads) acc += (double)__ldcg(&p.row_kl[r]);
        // block reduction in double, fixed order
        __shared__ double dred[kUpThreads / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) dred[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kUpThreads / 32; ++w) t += dred[w];
            *p.loss = (float)((double)p.loss_scale * t);
            atomicExch(&p.ctrl[0], 0u);
        }
    }
}

// ====================================================================================================
// Pixel mode (PDLoss behind the resize, losses.py:25-33 + :47-49 + :108-112): softmax over the C channels of every
// UP-SAMPLED pixel.  One thread per low-resolution cell keeps the statistics of the s x s pixels of its block in
// registers while it walks the channels twice - once for the sums (running per-cell reference, sums rescaled when
// it moves), once for the gradient - and the transposed stencil runs per channel through nine shared-memory planes
// (double buffered: one CTA barrier per channel).  A CTA computes a 16 x 16 tile of cells and owns the 14 x 14
// inside (gradients never cross CTAs).  Channels stream through shared memory in chunks of 4.
struct PxSmem {
    float st[2][2][kPxCh][kPxLoad * kPxLoad];  // [stage][S|T][channel][cell]
    float planes[2][9][kPxPlane];
    float red[kPxThreads / 32];
};
// The fast path takes one reference per CELL (the maximum of its 3 x 3 neighbourhood over the channels).  When
// neighbouring cells differ by more than ~87 tau somewhere, a pixel far below that reference underflows: its sums come
// out 0.  The CTA then redoes the window EXACTLY: a sweep for the maximum of every up-sampled pixel over the channels,
// and the two sweeps again with those per-pixel references - kept in (dynamic) shared memory, [S|T][pixel][thread], so
// that the fast path's registers are untouched.  Slower (three sweeps, two shared-memory reads per value), rare.
template <int NP>
struct PxRefs {
    float r2[2][NP][kPxThreads];               // fl(max * c2): the references of the exponents, as the sweeps use them
};

template <typename T, int S>
__global__ void __launch_bounds__(kPxThreads, 2) kl_pixels_up_kernel(const UpParams p) {
    constexpr int SB = S > 4 ? 4 : S;              // window side
    constexpr int NWIN = (S / SB) * (S / SB);
    __shared__ PxSmem sm;
    extern __shared__ __align__(16) unsigned char px_dyn[];
    PxRefs<SB * SB>& refs = *reinterpret_cast<PxRefs<SB * SB>*>(px_dyn);
    const int tid = threadIdx.x;
    const int ty = tid / kPxTile, tx = tid % kPxTile;
    const int tiles_x = (p.Wl + kPxOwn - 1) / kPxOwn, tiles_y = (p.Hl + kPxOwn - 1) / kPxOwn;
    const long long n_tiles = (long long)p.B * tiles_y * tiles_x;
    const size_t plane_elems = (size_t)p.Hl * p.Wl;
    float kl_acc = 0.f;
    // with several windows per block (s = 8) a unit is (tile, window): four times the parallelism, and the windows'
    // gradients go to separate fp32 planes of the workspace that up_sum_windows_kernel adds in a fixed order
    for (long long unit = blockIdx.x; unit < n_tiles * NWIN; unit += gridDim.x) {
        const long long tile = unit / NWIN;
        const int my_win = (int)(unit - tile * NWIN);
        const int b = (int)(tile / (tiles_y * tiles_x));
        const int trem = (int)(tile - (long long)b * tiles_y * tiles_x);
        const int i0 = (trem / tiles_x) * kPxOwn - 1, j0 = (trem % tiles_x) * kPxOwn - 1;   // first computed cell
        const int i = i0 + ty, j = j0 + tx;                                                 // my cell
        const bool in_map = i >= 0 && i < p.Hl && j >= 0 && j < p.Wl;
        const bool owned = in_map && ty >= 1 && ty <= kPxOwn && tx >= 1 && tx <= kPxOwn;
        const T* gS = static_cast<const T*>(p.S) + (size_t)b * p.C * plane_elems;
        const T* gT = static_cast<const T*>(p.T) + (size_t)b * p.C * plane_elems;
        // chunk of channels [c0, c0 + kPxCh) -> stage: loaded cell (ly, lx) = map cell clamp(i0 - 1 + ly, j0 - 1 + lx)
        auto load_chunk = [&](int c0, int stage) {
            const int nch = min(kPxCh, p.C - c0);
            for (int e = tid; e < nch * kPxLoad * kPxLoad; e += kPxThreads) {
                const int ch = e / (kPxLoad * kPxLoad), cell = e - ch * (kPxLoad * kPxLoad);
                const int ly = cell / kPxLoad, lx = cell - ly * kPxLoad;
                const int yi = min(max(i0 - 1 + ly, 0), p.Hl - 1), xj = min(max(j0 - 1 + lx, 0), p.Wl - 1);
                const size_t off = (size_t)(c0 + ch) * plane_elems + (size_t)yi * p.Wl + xj;
                sm.st[stage][0][ch][cell] = up_load<T>(gS + off);
                sm.st[stage][1][ch][cell] = up_load<T>(gT + off);
            }
        };
        // neighbourhood of my cell in the stage (clamping happened at load time; the border cells of the MAP must
        // see their own row/column instead of the clamped neighbour only when the neighbour is outside the map -
        // which the clamped load already delivers)
        auto nbhd = [&](const float* cellp, float (&a)[3][3]) {
            const float* q = cellp + ty * kPxLoad + tx;          // loaded cell (ty, tx) = my cell (-1, -1)
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int e = 0; e < 3; ++e) a[d][e] = q[d * kPxLoad + e];
        };
        const int n_chunks = (p.C + kPxCh - 1) / kPxCh;

        T* gD = static_cast<T*>(p.dS) + (size_t)b * p.C * plane_elems;
        // the s x s block of a cell is walked in windows of SB x SB pixels (one window for s <= 4, four for s = 8):
        // per window two sweeps over the channels
#pragma unroll
        for (int win = 0; win < NWIN; ++win) {
            if (NWIN > 1 && win != my_win) continue;          // (uniform over the CTA)
            const int ky0 = (win / (S / SB)) * SB, kx0 = (win % (S / SB)) * SB;
            // ---------------- pass 1: per-pixel sums over the channels (reference = running maximum of the cell).
            // dd = sum (et - es) term by term (common.cuh: KL without cancellation); the sum p (t - s) term of the KL is
            // collected in pass 2, where the probabilities exist (a scalar per thread instead of a register per pixel)
            float zs[SB * SB], zt[SB * SB], dd[SB * SB];
#pragma unroll
            for (int q = 0; q < SB * SB; ++q) zs[q] = zt[q] = dd[q] = 0.f;
            float ref_s = kUpFloor, ref_t = kUpFloor;
            __syncthreads();
            load_chunk(0, 0);
            for (int ck = 0; ck < n_chunks; ++ck) {
                __syncthreads();
                if (ck + 1 < n_chunks) load_chunk((ck + 1) * kPxCh, (ck + 1) & 1);
                const int nch = min(kPxCh, p.C - ck * kPxCh);
                if (in_map) {
                    for (int ch = 0; ch < nch; ++ch) {
                        float a[3][3], hs[3][SB], ht[3][SB], ds_[2][SB], dt_[2][SB];
                        nbhd(sm.st[ck & 1][0][ch], a);
                        float ms = a[0][0];
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int e = 0; e < 3; ++e) ms = fmaxf(ms, a[d][e]);
                        up_hrows_win<S, SB>(a, kx0, hs);
                        nbhd(sm.st[ck & 1][1][ch], a);
                        float mt = a[0][0];
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int e = 0; e < 3; ++e) mt = fmaxf(mt, a[d][e]);
                        up_hrows_win<S, SB>(a, kx0, ht);
                        if (ms > ref_s || mt > ref_t) {
                            const float nrs = fmaxf(ref_s, ms), nrt = fmaxf(ref_t, mt);
                            const float fs = ref_factor(ref_s, nrs, p.c2), ft = ref_factor(ref_t, nrt, p.c2);
                            const float df = factor_diff(fs, ft, merge_shift(ref_s, ref_t, nrs, nrt, p.c2));
#pragma unroll
                            for (int q = 0; q < SB * SB; ++q) {
                                dd[q] = fmaf(zs[q], df, dd[q] * ft);
                                zs[q] *= fs;
                                zt[q] *= ft;
                            }
                            ref_s = nrs;
                            ref_t = nrt;
                        }
                        up_vdiff<SB>(hs, ds_);
                        up_vdiff<SB>(ht, dt_);
                        const float rs2 = __fmul_rn(ref_s, p.c2), rt2 = __fmul_rn(ref_t, p.c2);
#pragma unroll
                        for (int ky = 0; ky < SB; ++ky) {
#pragma unroll
                            for (int kx = 0; kx < SB; ++kx) {
                                const float vs = up_value_win<S, SB>(hs, ds_, ky0 + ky, kx);
                                const float vt = up_value_win<S, SB>(ht, dt_, ky0 + ky, kx);
                                const float es = fast_exp2(fmaf(vs, p.c2, -rs2)), et = fast_exp2(fmaf(vt, p.c2, -rt2));
                                zs[ky * SB + kx] += es;
                                zt[ky * SB + kx] += et;
                                dd[ky * SB + kx] += et - es;
                            }
                        }
                    }
                }
            }
            // ---------------- a pixel whose sums vanished against the cell's reference (or went non-finite): redo the
            // window with one reference per pixel (uniform over the CTA; see PxRefs)
            bool bad = false;
            if (in_map) {
#pragma unroll
                for (int q = 0; q < SB * SB; ++q) bad = bad || !(zs[q] >= 1e-30f && zs[q] < 3e38f && zt[q] >= 1e-30f && zt[q] < 3e38f);
            }
            const bool exact = __syncthreads_or(bad) != 0;
            if (exact) {
                // sweep 0: the maximum of every pixel over the channels (zs / zt hold the maxima for now)
#pragma unroll
                for (int q = 0; q < SB * SB; ++q) zs[q] = zt[q] = -3.0e38f;
                load_chunk(0, 0);
                for (int ck = 0; ck < n_chunks; ++ck) {
                    __syncthreads();
                    if (ck + 1 < n_chunks) load_chunk((ck + 1) * kPxCh, (ck + 1) & 1);
                    const int nch = min(kPxCh, p.C - ck * kPxCh);
                    if (in_map) {
                        for (int ch = 0; ch < nch; ++ch) {
                            float a[3][3], hs[3][SB], ht[3][SB], ds_[2][SB], dt_[2][SB];
                            nbhd(sm.st[ck & 1][0][ch], a);
                            up_hrows_win<S, SB>(a, kx0, hs);
                            nbhd(sm.st[ck & 1][1][ch], a);
                            up_hrows_win<S, SB>(a, kx0, ht);
                            up_vdiff<SB>(hs, ds_);
                            up_vdiff<SB>(ht, dt_);
#pragma unroll
                            for (int ky = 0; ky < SB; ++ky) {
#pragma unroll
                                for (int kx = 0; kx < SB; ++kx) {
                                    zs[ky * SB + kx] = fmaxf(zs[ky * SB + kx], up_value_win<S, SB>(hs, ds_, ky0 + ky, kx));
                                    zt[ky * SB + kx] = fmaxf(zt[ky * SB + kx], up_value_win<S, SB>(ht, dt_, ky0 + ky, kx));
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < SB * SB; ++q) {
                    refs.r2[0][q][tid] = __fmul_rn(zs[q], p.c2);
                    refs.r2[1][q][tid] = __fmul_rn(zt[q], p.c2);
                    zs[q] = zt[q] = dd[q] = 0.f;
                }
                // sweep 1 again, against the per-pixel references (only this thread reads its column of refs)
                __syncthreads();
                load_chunk(0, 0);
                for (int ck = 0; ck < n_chunks; ++ck) {
                    __syncthreads();
                    if (ck + 1 < n_chunks) load_chunk((ck + 1) * kPxCh, (ck + 1) & 1);
                    const int nch = min(kPxCh, p.C - ck * kPxCh);
                    if (in_map) {
                        for (int ch = 0; ch < nch; ++ch) {
                            float a[3][3], hs[3][SB], ht[3][SB], ds_[2][SB], dt_[2][SB];
                            nbhd(sm.st[ck & 1][0][ch], a);
                            up_hrows_win<S, SB>(a, kx0, hs);
                            nbhd(sm.st[ck & 1][1][ch], a);
                            up_hrows_win<S, SB>(a, kx0, ht);
                            up_vdiff<SB>(hs, ds_);
                            up_vdiff<SB>(ht, dt_);
#pragma unroll
                            for (int ky = 0; ky < SB; ++ky) {
#pragma unroll
                                for (int kx = 0; kx < SB; ++kx) {
                                    const int q = ky * SB + kx;
                                    const float vs = up_value_win<S, SB>(hs, ds_, ky0 + ky, kx);
                                    const float vt = up_value_win<S, SB>(ht, dt_, ky0 + ky, kx);
                                    const float es = fast_exp2(fmaf(vs, p.c2, -refs.r2[0][q][tid]));
                                    const float et = fast_exp2(fmaf(vt, p.c2, -refs.r2[1][q][tid]));
                                    zs[q] += es;
                                    zt[q] += et;
                                    dd[q] += et - es;
                                }
                            }
                        }
                    }
                }
            }
            // the ln(zt / zs) part of the KL of my pixels (owned cells only; kl_from_stats with a2 = 0), then the sums
            // become the gradient factors coef / Z
            if (owned) {
#pragma unroll
                for (int q = 0; q < SB * SB; ++q) kl_acc += kl_from_stats(zs[q], zt[q], 0.f, dd[q]);
            }
            float kl_a = 0.f;              // coef * sum over my pixels and the channels of p (at - as), exponent domain
            if (in_map) {
#pragma unroll
                for (int q = 0; q < SB * SB; ++q) {
                    zs[q] = __fdividef(p.coef, zs[q]);
                    zt[q] = __fdividef(p.coef, zt[q]);
                }
            }
            const float rs2 = __fmul_rn(ref_s, p.c2), rt2 = __fmul_rn(ref_t, p.c2);

            // ---------------- pass 2: per channel, window gradient -> nine contributions -> owned cells
            // (two compiled copies: the fast one with the cell's references in registers, the exact one reading a
            // reference per pixel from shared memory)
            auto pass2 = [&](auto exact_tag) {
                constexpr bool EXACT = decltype(exact_tag)::value;
                __syncthreads();
                load_chunk(0, 0);
                int cglob = 0;
                for (int ck = 0; ck < n_chunks; ++ck) {
                    __syncthreads();
                    if (ck + 1 < n_chunks) load_chunk((ck + 1) * kPxCh, (ck + 1) & 1);
                    const int nch = min(kPxCh, p.C - ck * kPxCh);
                    for (int ch = 0; ch < nch; ++ch, ++cglob) {
                        float* pl = sm.planes[cglob & 1][0];
                        if (in_map) {
                            float a[3][3], hs[3][SB], ht[3][SB], ds_[2][SB], dt_[2][SB];
                            nbhd(sm.st[ck & 1][0][ch], a);
                            up_hrows_win<S, SB>(a, kx0, hs);
                            nbhd(sm.st[ck & 1][1][ch], a);
                            up_hrows_win<S, SB>(a, kx0, ht);
                            up_vdiff<SB>(hs, ds_);
                            up_vdiff<SB>(ht, dt_);
                            float m[3][3];
    #pragma unroll
                            for (int d = 0; d < 3; ++d)
    #pragma unroll
                                for (int e = 0; e < 3; ++e) m[d][e] = 0.f;
    #pragma unroll
                            for (int ky = 0; ky < SB; ++ky) {
                                float tr[3] = {0.f, 0.f, 0.f};
    #pragma unroll
                                for (int kx = 0; kx < SB; ++kx) {
                                    const float vs = up_value_win<S, SB>(hs, ds_, ky0 + ky, kx);
                                    const float vt = up_value_win<S, SB>(ht, dt_, ky0 + ky, kx);
                                    const float as = fmaf(vs, p.c2, EXACT ? -refs.r2[0][ky * SB + kx][tid] : -rs2);
                                    const float at = fmaf(vt, p.c2, EXACT ? -refs.r2[1][ky * SB + kx][tid] : -rt2);
                                    const float es = fast_exp2(as);
                                    const float pt = fast_exp2(at) * zt[ky * SB + kx];                       // coef * p
                                    const float gv = fmaf(es, zs[ky * SB + kx], -pt);
                                    kl_a = fmaf(pt, at - as, kl_a);
                                    const int f = UpW<S>::first(kx0 + kx) + 1;
                                    const float w1 = UpW<S>::w1(kx0 + kx);
                                    tr[f] = fmaf(1.f - w1, gv, tr[f]);
                                    tr[f + 1] = fmaf(w1, gv, tr[f + 1]);
                                }
                                const int f = UpW<S>::first(ky0 + ky) + 1;
                                const float w1 = UpW<S>::w1(ky0 + ky);
    #pragma unroll
                                for (int e = 0; e < 3; ++e) {
                                    m[f][e] = fmaf(1.f - w1, tr[e], m[f][e]);
                                    m[f + 1][e] = fmaf(w1, tr[e], m[f + 1][e]);
                                }
                            }
                            // taps clamped at the border of the map fall onto the cell itself
                            if (i == 0) {
    #pragma unroll
                                for (int e = 0; e < 3; ++e) { m[1][e] += m[0][e]; m[0][e] = 0.f; }
                            }
                            if (i == p.Hl - 1) {
    #pragma unroll
                                for (int e = 0; e < 3; ++e) { m[1][e] += m[2][e]; m[2][e] = 0.f; }
                            }
                            if (j == 0) {
    #pragma unroll
                                for (int d = 0; d < 3; ++d) { m[d][1] += m[d][0]; m[d][0] = 0.f; }
                            }
                            if (j == p.Wl - 1) {
    #pragma unroll
                                for (int d = 0; d < 3; ++d) { m[d][1] += m[d][2]; m[d][2] = 0.f; }
                            }
                            // contribution (d, e) goes to tile cell (ty + d - 1, tx + e - 1): plane (d, e), row ty + d - 1,
                            // padded column tx + e
    #pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                const int ry = ty + d - 1;
                                if (ry >= 0 && ry < kPxTile) {
    #pragma unroll
                                    for (int e = 0; e < 3; ++e) pl[(d * 3 + e) * kPxPlane + ry * (kPxTile + 2) + tx + e] = m[d][e];
                                }
                            }
                        }
                        __syncthreads();
                        if (owned) {
                            // plane (d, e) at my position holds what cell (i - d + 1, j - e + 1) sent here
                            const float* q = pl + ty * (kPxTile + 2) + tx + 1;
                            const bool okd[3] = {i + 1 < p.Hl, true, i >= 1};
                            const bool oke[3] = {j + 1 < p.Wl, true, j >= 1};
                            float v = 0.f;
    #pragma unroll
                            for (int d = 0; d < 3; ++d)
    #pragma unroll
                                for (int e = 0; e < 3; ++e) {
                                    const float x = q[(d * 3 + e) * kPxPlane];
                                    v += (okd[d] && oke[e]) ? x : 0.f;
                                }
                            const size_t off = (size_t)(ck * kPxCh + ch) * plane_elems + (size_t)i * p.Wl + j;
                            if (NWIN > 1) p.wpart[((size_t)win * p.B + b) * p.C * plane_elems + off] = v;
                            else up_store<T>(gD + off, v);
                        }
                    }
                }
            };
            if (exact) pass2(std::true_type{});
            else pass2(std::false_type{});
            if (owned) kl_acc = fmaf(kl_a, kLn2 * p.inv_coef, kl_acc);
        }
    }
    // ---- loss: CTA partial, the last CTA sums the partials in a fixed order
    kl_acc = block_sum_n<kPxThreads>(kl_acc, sm.red);
    __shared__ unsigned ticket_s;
    if (tid == 0) {
        __stcg(&p.part[blockIdx.x], kl_acc);
        __threadfence();
        ticket_s = atomicAdd(&p.ctrl[0], 1u);
    }
    __syncthreads();
    if (ticket_s == gridDim.x - 1) {
        __threadfence();
        double acc = 0.0;
        for (int r = tid; r < (int)gridDim.x; r += kPxThreads) acc += (double)__ldcg(&p.part[r]);
        __shared__ double dred[kPxThreads / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) dred[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < kPxThreads / 32; ++w) t += dred[w];
            *p.loss = (float)((double)p.loss_scale * t);
            atomicExch(&p.ctrl[0], 0u);
        }
    }
}

// dS = sum of the window planes (s = 8), fixed order
template <typename T>
__global__ void __launch_bounds__(256) up_sum_windows_kernel(const float* __restrict__ wpart, T* __restrict__ dS,
                                                             long long n, int nwin) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        float v = 0.f;
        for (int w = 0; w < nwin; ++w) v += wpart[(size_t)w * n + i];
        up_store<T>(dS + i, v);
    }
}

template <typename T, int S>
static cudaError_t launch_px_up_t(const UpParams& p, int sms, cudaStream_t stream, int* grid_out) {
    auto k = kl_pixels_up_kernel<T, S>;
    constexpr int SB = S > 4 ? 4 : S;
    const size_t dyn = sizeof(PxRefs<SB * SB>);           // per-pixel references of the exact redo
    static std::atomic<bool> configured[kMaxDevices];
    const int dev = device_slot();
    if (!configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kPxThreads, dyn);
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    constexpr int nwin = S > 4 ? (S / 4) * (S / 4) : 1;
    const long long units = (long long)p.B * ((p.Hl + kPxOwn - 1) / kPxOwn) * ((p.Wl + kPxOwn - 1) / kPxOwn) * nwin;
    long long grid = (long long)sms * occ;
    if (grid > units) grid = units;
    if (grid > kMaxGrid) grid = kMaxGrid;
    if (grid_out) *grid_out = (int)grid;
    k<<<(unsigned)grid, kPxThreads, dyn, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || nwin == 1) return e;
    const long long n = (long long)p.B * p.C * p.Hl * p.Wl;
    long long g2 = (n + 255) / 256;
    if (g2 > (long long)sms * 8) g2 = (long long)sms * 8;
    up_sum_windows_kernel<T><<<(unsigned)g2, 256, 0, stream>>>(p.wpart, static_cast<T*>(p.dS), n, nwin);
    return cudaGetLastError();
}

cudaError_t launch_kl_pixels_up(const UpParams& p, bool bf16, int sms, cudaStream_t stream) {
    switch (p.scale) {
        case 2: return bf16 ? launch_px_up_t<__nv_bfloat16, 2>(p, sms, stream, nullptr) : launch_px_up_t<float, 2>(p, sms, stream, nullptr);
        case 4: return bf16 ? launch_px_up_t<__nv_bfloat16, 4>(p, sms, stream, nullptr) : launch_px_up_t<float, 4>(p, sms, stream, nullptr);
        case 8: return bf16 ? launch_px_up_t<__nv_bfloat16, 8>(p, sms, stream, nullptr) : launch_px_up_t<float, 8>(p, sms, stream, nullptr);
        default: return cudaErrorInvalidValue;
    }
}

// ====================================================================================================
size_t up_smem_bytes(int SR, int Wl) { return sizeof(float) * ((size_t)2 * (SR + 4) * Wl + (size_t)9 * SR * (Wl + 2)); }

template <typename T, int S>
static cudaError_t launch_up_t(const UpParams& p, int sms, cudaStream_t stream) {
    const size_t smem1 = sizeof(float) * (size_t)2 * (p.SR + 2) * p.Wl;
    const size_t smem2 = up_smem_bytes(p.SR, p.Wl);
    auto k1 = kl_rows_up_stats_kernel<T, S>;
    auto k2 = kl_rows_up_grad_kernel<T, S>;
    // the largest dynamic shared memory opted into so far, per instantiation and device
    static std::atomic<size_t> cfg1_dev[kMaxDevices], cfg2_dev[kMaxDevices];
    const int dev = device_slot();
    int occ1 = 1, occ2 = 1;
    if (smem1 > cfg1_dev[dev].load() || smem2 > cfg2_dev[dev].load()) {
        cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return e;
        if (smem1 > cfg1_dev[dev].load()) cfg1_dev[dev].store(smem1);
        if (smem2 > cfg2_dev[dev].load()) cfg2_dev[dev].store(smem2);
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k1, kUpThreads, smem1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k2, kUpThreads, smem2);
    if (occ1 < 1 || occ2 < 1) return cudaErrorLaunchOutOfResources;
    long long g1 = (long long)sms * occ1, g2 = (long long)sms * occ2;
    if (g1 > p.units) g1 = p.units;
    if (g2 > p.units) g2 = p.units;
    k1<<<(unsigned)g1, kUpThreads, smem1, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k2<<<(unsigned)g2, kUpThreads, smem2, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_kl_rows_up(const UpParams& p, bool bf16, int sms, cudaStream_t stream) {
    switch (p.scale) {
        case 2: return bf16 ? launch_up_t<__nv_bfloat16, 2>(p, sms, stream) : launch_up_t<float, 2>(p, sms, stream);
        case 4: return bf16 ? launch_up_t<__nv_bfloat16, 4>(p, sms, stream) : launch_up_t<float, 4>(p, sms, stream);
        case 8: return bf16 ? launch_up_t<__nv_bfloat16, 8>(p, sms, stream) : launch_up_t<float, 8>(p, sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sd
