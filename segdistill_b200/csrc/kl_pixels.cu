// Per-pixel softmax-KL over the channel axis, forward + backward fused (PDLoss, ATLoss).
//
// Replaces mmseg/models/distillation/losses.py:47-49 (permute(0,2,3,1).reshape - an NHWC copy of
// both maps) + :108-112, and ATLoss :190-196, plus their autograd backward.  The softmax runs over
// C at stride HW directly on the NCHW tensors: no permuted copy is ever materialised.
//
//   kl_pixels_tma_kernel   persistent.  One elected thread fetches a [C x 64-or-128 pixel] tile of
//                          S and of T per step with ONE 3-D tiled TMA each (tensor maps over the
//                          (B, C, HW) view) into a shared-memory ring, as soon as a slot is drained;
//                          512 threads = 64 pixel columns x 8 channel groups hold their
//                          elements in registers, combine per-pixel (max, sum) partials through
//                          shared memory and write dS from registers.  12 B/elem fp32, 6 B/elem bf16.
//   kl_pixels_generic      one thread per pixel, three strided passes; any alignment, any C.
#include <cuda.h>

#include "common.cuh"
#include "params.h"

namespace sd {

constexpr int kPixCons = 512;
constexpr int kPixThreads = kPixCons;
constexpr int kPixCols = 64;   // thread columns of a tile (1 fp32 pixel or 2 bf16 pixels each)
constexpr int kPixCG = 8;      // channel groups: thread (col, cg) owns channels cg, cg+8, cg+16, ...
constexpr int kPixRowBytes = 256;  // bytes of one channel row of a tile (64 fp32 or 128 bf16 pixels)

template <typename T>
struct PixTraits;
template <>
struct PixTraits<float> {
    static constexpr int PXT = 1;
    static __device__ __forceinline__ void unpack(uint32_t w, float* f) { f[0] = __uint_as_float(w); }
    static __device__ __forceinline__ uint32_t pack(const float* f) { return __float_as_uint(f[0]); }
};
template <>
struct PixTraits<__nv_bfloat16> {
    static constexpr int PXT = 2;
    static __device__ __forceinline__ void unpack(uint32_t w, float* f) {
        Elem<__nv_bfloat16>::unpack2(w, f[0], f[1]);
    }
    static __device__ __forceinline__ uint32_t pack(const float* f) { return Elem<__nv_bfloat16>::pack2(f[0], f[1]); }
};

// cols: thread columns of a tile - 64 (512 threads, one CTA per SM) or 32 (256 threads, two CTAs per SM: their phases
// - loads, maxima, exponentials, gradient - overlap; the bf16 kernel is not HBM-bound)
size_t pix_tma_smem_bytes(int C, int pxt, int nstages, int cols) {
    return (size_t)nstages * 2 * (size_t)C * cols * 4              // ring
           + 2 * (size_t)kPixCG * cols * pxt * sizeof(float4)      // red_max, red_sum
           + 128;                                                  // barriers
}

// EXACT: C > 8 * (CPT - 1), i.e. only the last of a thread's CPT channel slots can be missing - the channel guards
// of the other slots fold away at compile time (C = 150 with CPT = 19; the guards were a quarter of the instructions)
template <typename T, int CPT, bool AT, bool EXACT, int COLS>
__global__ void __launch_bounds__(COLS * kPixCG, COLS == 32 ? 2 : 1)
kl_pixels_tma_kernel(const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapT,
                     const PixParams p) {
    constexpr int kPixCols = COLS;     // (shadows the default: this instantiation's thread columns)
    constexpr int PXT = PixTraits<T>::PXT;
    constexpr int P = kPixCols * PXT;  // pixels per tile

    extern __shared__ __align__(128) unsigned char smem[];
    const size_t ring_bytes = (size_t)p.nstages * 2 * p.stage_bytes;
    float4* red_max = reinterpret_cast<float4*>(smem + ring_bytes);
    float4* red_sum = red_max + kPixCG * kPixCols * PXT;
    uint64_t* full = reinterpret_cast<uint64_t*>(red_sum + kPixCG * kPixCols * PXT);

    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // ============================ TMA issue (thread 0 only) ============================
    // a slot is free again once every thread copied its elements to registers, i.e. after the
    // first CTA barrier of the tile that occupied it
    long long prod_tile = blockIdx.x;
    int prod_stage = 0, free_slots = p.nstages;
    uint64_t pol = 0;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapS);
        tma_prefetch_desc(&mapT);
        pol = l2_policy_evict_first();
    }
    auto issue_loads = [&]() {
        while (free_slots > 0 && prod_tile < p.total_tiles) {
            const int b = (int)(prod_tile / p.tiles_per_sample);
            const int px0 = (int)(prod_tile - (long long)b * p.tiles_per_sample) * P;
            mbar_arrive_expect_tx(&full[prod_stage], 2u * p.stage_bytes);
            unsigned char* dst = smem + (size_t)prod_stage * 2 * p.stage_bytes;
            tma_tile3d_g2s(dst, &mapS, px0, 0, b, &full[prod_stage], pol);
            tma_tile3d_g2s(dst + p.stage_bytes, &mapT, px0, 0, b, &full[prod_stage], pol);
            prod_tile += gridDim.x;
            if (++prod_stage == p.nstages) prod_stage = 0;
            --free_slots;
        }
    };
    if (threadIdx.x == 0) issue_loads();

    // ============================ compute ============================
    const int tid = threadIdx.x;
    const int col = tid & (kPixCols - 1);
    const int cg = tid / kPixCols;
    const float c2 = p.c2;
    float s[CPT * PXT], t[CPT * PXT];
    float acc_kl = 0.f, acc_at = 0.f;  // cg == 0 threads only, in tile order
    int stage = 0;
    uint32_t phase = 0;

    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = (int)(tile / p.tiles_per_sample);
        const int px = (int)(tile - (long long)b * p.tiles_per_sample) * P + col * PXT;  // first pixel of this thread

        mbar_wait(&full[stage], phase);
        const uint32_t* ws = reinterpret_cast<const uint32_t*>(smem + (size_t)stage * 2 * p.stage_bytes);
        const uint32_t* wt = ws + p.stage_bytes / 4;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int c = cg + kPixCG * k;
            if ((EXACT && k < CPT - 1) || c < p.C) {
                PixTraits<T>::unpack(ws[c * kPixCols + col], &s[k * PXT]);
                PixTraits<T>::unpack(wt[c * kPixCols + col], &t[k * PXT]);
            }
        }
        if (++stage == p.nstages) {
            stage = 0;
            phase ^= 1u;
        }

        // ---- per-pixel max (and channel sums for the AT term) over this thread's channels
        float mxs[PXT], mxt[PXT], sus[PXT], sut[PXT];
#pragma unroll
        for (int q = 0; q < PXT; ++q) {
            mxs[q] = -INFINITY;
            mxt[q] = -INFINITY;
            sus[q] = 0.f;
            sut[q] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            if ((EXACT && k < CPT - 1) || cg + kPixCG * k < p.C) {
#pragma unroll
                for (int q = 0; q < PXT; ++q) {
                    mxs[q] = fmaxf(mxs[q], s[k * PXT + q]);
                    mxt[q] = fmaxf(mxt[q], t[k * PXT + q]);
                    if (AT) {
                        sus[q] += s[k * PXT + q];
                        sut[q] += t[k * PXT + q];
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PXT; ++q)
            red_max[(q * kPixCG + cg) * kPixCols + col] = make_float4(mxs[q], mxt[q], sus[q], sut[q]);
        __syncthreads();
        if (tid == 0) {
            ++free_slots;
            issue_loads();
        }
        float dm[PXT];
#pragma unroll
        for (int q = 0; q < PXT; ++q) {
            float a = -INFINITY, bb = -INFINITY, ss = 0.f, st = 0.f;
#pragma unroll
            for (int g = 0; g < kPixCG; ++g) {
                const float4 r = red_max[(q * kPixCG + g) * kPixCols + col];
                a = fmaxf(a, r.x);
                bb = fmaxf(bb, r.y);
                ss += r.z;
                st += r.w;
            }
            mxs[q] = a;
            mxt[q] = bb;
            dm[q] = (ss - st) * p.inv_C;  // difference of the channel means
        }

        // ---- exponentials stay in registers; per-pixel partial sums
        // (dd = sum (et - es) term by term against the pixel's references: common.cuh, KL without cancellation)
        float zs[PXT], zt[PXT], ac[PXT], dd[PXT], rs2[PXT], rt2[PXT];
#pragma unroll
        for (int q = 0; q < PXT; ++q) {
            zs[q] = 0.f;
            zt[q] = 0.f;
            ac[q] = 0.f;
            dd[q] = 0.f;
            rs2[q] = __fmul_rn(mxs[q], c2);
            rt2[q] = __fmul_rn(mxt[q], c2);
        }
        if constexpr (PXT == 2) {
            // bf16: the thread's two pixels side by side, two fp32 per instruction (FFMA2 / FADD2, common.cuh) - the
            // bf16 kernel is bound by instruction issue, not by HBM
            const F2 C2 = f2_dup(c2), NRS = f2_make(-rs2[0], -rs2[1]), NRT = f2_make(-rt2[0], -rt2[1]), neg1 = f2_dup(-1.f);
            F2 ZS = f2_dup(0.f), ZT = f2_dup(0.f), AC = f2_dup(0.f), DD = f2_dup(0.f);
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                if ((EXACT && k < CPT - 1) || cg + kPixCG * k < p.C) {
                    const int i = k * PXT;
                    const F2 as2 = f2_fma(f2_make(s[i], s[i + 1]), C2, NRS), at2 = f2_fma(f2_make(t[i], t[i + 1]), C2, NRT);
                    float as0, as1, at0, at1;
                    f2_split(as2, as0, as1);
                    f2_split(at2, at0, at1);
                    s[i] = fast_exp2(as0);
                    s[i + 1] = fast_exp2(as1);
                    t[i] = fast_exp2(at0);
                    t[i + 1] = fast_exp2(at1);
                    const F2 es2 = f2_make(s[i], s[i + 1]), et2 = f2_make(t[i], t[i + 1]);
                    ZS = f2_add(ZS, es2);
                    ZT = f2_add(ZT, et2);
                    AC = f2_fma(et2, f2_fma(as2, neg1, at2), AC);
                    DD = f2_add(DD, f2_fma(es2, neg1, et2));
                }
            }
            f2_split(ZS, zs[0], zs[PXT - 1]);
            f2_split(ZT, zt[0], zt[PXT - 1]);
            f2_split(AC, ac[0], ac[PXT - 1]);
            f2_split(DD, dd[0], dd[PXT - 1]);
        } else {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                if ((EXACT && k < CPT - 1) || cg + kPixCG * k < p.C) {
#pragma unroll
                    for (int q = 0; q < PXT; ++q) {
                        const int i = k * PXT + q;
                        const float as = fmaf(s[i], c2, -rs2[q]), at = fmaf(t[i], c2, -rt2[q]);
                        const float es = fast_exp2(as);
                        const float et = fast_exp2(at);
                        zs[q] += es;
                        zt[q] += et;
                        ac[q] = fmaf(et, at - as, ac[q]);
                        dd[q] += et - es;
                        s[i] = es;
                        t[i] = et;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PXT; ++q)
            red_sum[(q * kPixCG + cg) * kPixCols + col] = make_float4(zs[q], zt[q], ac[q], dd[q]);
        __syncthreads();
        float ks[PXT], kt[PXT], ga[PXT];
#pragma unroll
        for (int q = 0; q < PXT; ++q) {
            float Zs = 0.f, Zt = 0.f, A = 0.f, DD = 0.f;
#pragma unroll
            for (int g = 0; g < kPixCG; ++g) {
                const float4 r = red_sum[(q * kPixCG + g) * kPixCols + col];
                Zs += r.x;
                Zt += r.y;
                A += r.z;
                DD += r.w;
            }
            ks[q] = p.coef / Zs;
            kt[q] = p.coef / Zt;
            ga[q] = AT ? p.at_gcoef * dm[q] : 0.f;
            if (cg == 0 && px + q < p.HW) {
                const float kl = kl_from_stats(Zs, Zt, A, DD);
                p.row_kl[(size_t)b * p.HW + px + q] = kl;
                acc_kl += kl;
                if (AT) acc_at = fmaf(dm[q], dm[q], acc_at);
            }
        }

        // ---- gradient from registers, coalesced along the pixel axis
        if (px < p.HW) {
            uint32_t* out = reinterpret_cast<uint32_t*>(static_cast<T*>(p.dS) + ((size_t)b * p.C) * p.HW + px);
            const size_t cstride = (size_t)p.HW * sizeof(T) / 4;  // words per channel plane
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int c = cg + kPixCG * k;
                if ((EXACT && k < CPT - 1) || c < p.C) {
                    float o[PXT];
                    if constexpr (PXT == 2) {
                        const F2 u = f2_fma(f2_make(t[k * PXT], t[k * PXT + 1]), f2_make(-kt[0], -kt[PXT - 1]), f2_make(ga[0], ga[PXT - 1]));
                        f2_split(f2_fma(f2_make(s[k * PXT], s[k * PXT + 1]), f2_make(ks[0], ks[PXT - 1]), u), o[0], o[PXT - 1]);
                    } else {
#pragma unroll
                        for (int q = 0; q < PXT; ++q) o[q] = fmaf(s[k * PXT + q], ks[q], -t[k * PXT + q] * kt[q]) + ga[q];
                    }
                    out[(size_t)c * cstride] = PixTraits<T>::pack(o);
                }
            }
        }
    }

    // ============================ loss ============================
    // cg == 0 threads are consumer warps 0 and 1
    float* scratch = reinterpret_cast<float*>(red_max);
    __syncthreads();  // red_max is free again
    if (cg == 0) {
        const float wk = warp_sum(acc_kl);
        const float wa = warp_sum(acc_at);
        if (lane == 0) {
            scratch[2 * (tid >> 5)] = wk;
            scratch[2 * (tid >> 5) + 1] = wa;
        }
    }
    __syncthreads();
    if (tid < 32) {
        unsigned ticket = 0;
        if (lane == 0) {
            __stcg(&p.cta_part[blockIdx.x], scratch[0] + (kPixCols == 64 ? scratch[2] : 0.f));
            __stcg(&p.cta_part[p.nparts + blockIdx.x], scratch[1] + (kPixCols == 64 ? scratch[3] : 0.f));
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double kl = 0.0, at = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
                kl += (double)__ldcg(&p.cta_part[i]);
                at += (double)__ldcg(&p.cta_part[p.nparts + i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kl += __shfl_down_sync(0xffffffffu, kl, o);
                at += __shfl_down_sync(0xffffffffu, at, o);
            }
            if (lane == 0) {
                *p.loss = (float)((double)p.loss_scale * kl);
                if (AT && p.at_loss) *p.at_loss = (float)((double)p.at_scale * at);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

// ====================================================================================================
// generic: one thread per pixel
// ====================================================================================================
template <typename T, bool AT>
__global__ void __launch_bounds__(256) kl_pixels_generic(const PixParams p) {
    using E = Elem<T>;
    __shared__ float sh[2][8];
    const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long R = (long long)p.B * p.HW;
    float kl = 0.f, atsq = 0.f;
    if (r < R) {
        const int b = (int)(r / p.HW);
        const int px = (int)(r - (long long)b * p.HW);
        const size_t base = ((size_t)b * p.C) * p.HW + px;
        const T* s = static_cast<const T*>(p.S) + base;
        const T* t = static_cast<const T*>(p.T) + base;
        T* o = static_cast<T*>(p.dS) + base;
        float ms = -INFINITY, mt = -INFINITY, ss = 0.f, st = 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float a = E::load(s + (size_t)c * p.HW), bb = E::load(t + (size_t)c * p.HW);
            ms = fmaxf(ms, a);
            mt = fmaxf(mt, bb);
            ss += a;
            st += bb;
        }
        const float ms2 = __fmul_rn(ms, p.c2), mt2 = __fmul_rn(mt, p.c2);
        float zs = 0.f, zt = 0.f, ac = 0.f, dd = 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float a = E::load(s + (size_t)c * p.HW), bb = E::load(t + (size_t)c * p.HW);
            const float as = fmaf(a, p.c2, -ms2), at = fmaf(bb, p.c2, -mt2);
            const float es = fast_exp2(as);
            const float et = fast_exp2(at);
            zs += es;
            zt += et;
            ac = fmaf(et, at - as, ac);
            dd += et - es;
        }
        const float dm = (ss - st) * p.inv_C;
        const float ks = p.coef / zs, kt = p.coef / zt;
        const float ga = AT ? p.at_gcoef * dm : 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float a = E::load(s + (size_t)c * p.HW), bb = E::load(t + (size_t)c * p.HW);
            const float es = fast_exp2(fmaf(a, p.c2, -ms2));
            const float et = fast_exp2(fmaf(bb, p.c2, -mt2));
            E::store(o + (size_t)c * p.HW, fmaf(es, ks, -et * kt) + ga);
        }
        kl = kl_from_stats(zs, zt, ac, dd);
        p.row_kl[r] = kl;
        if (AT) atsq = dm * dm;
    }
    kl = warp_sum(kl);
    atsq = warp_sum(atsq);
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = kl;
        sh[1][threadIdx.x >> 5] = atsq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, bb = 0.f;
        for (int w = 0; w < 8; ++w) {
            a += sh[0][w];
            bb += sh[1][w];
        }
        p.cta_part[blockIdx.x] = a;
        p.cta_part[p.nparts + blockIdx.x] = bb;
    }
}

__global__ void __launch_bounds__(1024) kl_pixels_generic_finalize(const PixParams p, int nblocks) {
    __shared__ double sh[2][32];
    double kl = 0.0, at = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 1024) {
        kl += (double)p.cta_part[i];
        at += (double)p.cta_part[p.nparts + i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kl += __shfl_down_sync(0xffffffffu, kl, o);
        at += __shfl_down_sync(0xffffffffu, at, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = kl;
        sh[1][threadIdx.x >> 5] = at;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; ++w) {
            a += sh[0][w];
            b += sh[1][w];
        }
        *p.loss = (float)((double)p.loss_scale * a);
        if (p.at_loss) *p.at_loss = (float)((double)p.at_scale * b);
    }
}

// ====================================================================================================
// host launchers
// ====================================================================================================
template <typename T, int CPT, bool AT, int COLS>
static cudaError_t launch_pix_tma_t(const CUtensorMap& mS, const CUtensorMap& mT, const PixParams& p, int grid,
                                    size_t smem, cudaStream_t stream) {
    const bool exact = p.C > kPixCG * (CPT - 1);
    cudaError_t e;
    if (exact) {
        auto kern = kl_pixels_tma_kernel<T, CPT, AT, true, COLS>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, COLS * kPixCG, smem, stream>>>(mS, mT, p);
    } else {
        auto kern = kl_pixels_tma_kernel<T, CPT, AT, false, COLS>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, COLS * kPixCG, smem, stream>>>(mS, mT, p);
    }
    return cudaGetLastError();
}

template <typename T, bool AT, int COLS>
static cudaError_t launch_pix_tma_c(const CUtensorMap& mS, const CUtensorMap& mT, const PixParams& p, int grid,
                                    size_t smem, cudaStream_t stream) {
    if (p.C <= 3 * kPixCG) return launch_pix_tma_t<T, 3, AT, COLS>(mS, mT, p, grid, smem, stream);
    if (p.C <= 8 * kPixCG) return launch_pix_tma_t<T, 8, AT, COLS>(mS, mT, p, grid, smem, stream);
    if (p.C <= 19 * kPixCG) return launch_pix_tma_t<T, 19, AT, COLS>(mS, mT, p, grid, smem, stream);
    if (sizeof(T) == 4 && p.C <= 32 * kPixCG) return launch_pix_tma_t<T, (sizeof(T) == 4 ? 32 : 19), AT, COLS>(mS, mT, p, grid, smem, stream);
    return cudaErrorInvalidValue;
}

// largest channel count the TMA kernel holds in registers
int kl_pixels_tma_max_channels(bool bf16) { return (bf16 ? 19 : 32) * kPixCG; }
int kl_pixels_tile_pixels(bool bf16, int cols) { return cols * (bf16 ? 2 : 1); }

cudaError_t launch_kl_pixels_tma(const void* mapS, const void* mapT, const PixParams& p, bool bf16, int cols, int grid,
                                 size_t smem, cudaStream_t stream) {
    const CUtensorMap& mS = *static_cast<const CUtensorMap*>(mapS);
    const CUtensorMap& mT = *static_cast<const CUtensorMap*>(mapT);
    const bool at = p.at_gcoef != 0.f || p.at_loss != nullptr;
    if (bf16 && cols == 32) {
        return at ? launch_pix_tma_c<__nv_bfloat16, true, 32>(mS, mT, p, grid, smem, stream)
                  : launch_pix_tma_c<__nv_bfloat16, false, 32>(mS, mT, p, grid, smem, stream);
    }
    if (cols != 64) return cudaErrorInvalidValue;
    if (bf16) {
        return at ? launch_pix_tma_c<__nv_bfloat16, true, 64>(mS, mT, p, grid, smem, stream)
                  : launch_pix_tma_c<__nv_bfloat16, false, 64>(mS, mT, p, grid, smem, stream);
    }
    return at ? launch_pix_tma_c<float, true, 64>(mS, mT, p, grid, smem, stream)
              : launch_pix_tma_c<float, false, 64>(mS, mT, p, grid, smem, stream);
}

cudaError_t launch_kl_pixels_generic(const PixParams& p, bool bf16, cudaStream_t stream) {
    const long long R = (long long)p.B * p.HW;
    const int nblocks = (int)((R + 255) / 256);
    const bool at = p.at_gcoef != 0.f || p.at_loss != nullptr;
    if (bf16) {
        if (at) kl_pixels_generic<__nv_bfloat16, true><<<nblocks, 256, 0, stream>>>(p);
        else kl_pixels_generic<__nv_bfloat16, false><<<nblocks, 256, 0, stream>>>(p);
    } else {
        if (at) kl_pixels_generic<float, true><<<nblocks, 256, 0, stream>>>(p);
        else kl_pixels_generic<float, false><<<nblocks, 256, 0, stream>>>(p);
    }
    kl_pixels_generic_finalize<<<1, 1024, 0, stream>>>(p, nblocks);
    return cudaGetLastError();
}

}  // namespace sd
