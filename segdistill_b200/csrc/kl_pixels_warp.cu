// Per-pixel softmax-KL over the channel axis for bf16 maps (PDLoss on BASELINE config 3), every warp on its own.
//
// Same mathematics as kl_pixels.cu (mmseg/models/distillation/losses.py:47-49 + :108-112 and the backward).  The bf16
// launch of kl_pixels_tma_kernel is not HBM-bound: 24 instructions per element (one 32-bit shared-memory load per
// element pair, word-sized global stores), and two CTA barriers per tile that line up the phases of all warps - loads,
// maxima, exponentials (MUFU), gradient - one after the other.  Here
//   * a tile is [C x 64 pixels] of S and of T (128-byte rows, 128-byte TMA swizzle), fetched by a PRODUCER warp into a
//     ring of 5 stages;
//   * compute warp q (of 16) owns the 4-pixel column q of every tile: LANE = channel (c = lane, lane + 32, ...), so the
//     softmax reductions over the channels are WARP reductions (redux.sync for the maxima, a transposed butterfly for
//     the 16 sums of a column) - no CTA barrier anywhere, the warps drift apart and their phases overlap.  The swizzle
//     makes the 8-byte loads of 32 consecutive rows conflict-free per quarter;
//   * the gradient is written IN PLACE over the warp's column of the S tile (nobody else reads or writes it) and the
//     whole tile leaves with one tensor store (cp.async.bulk.tensor shared -> global) issued by the producer warp when
//     the 16 warps have signalled the stage's `done` barrier; the store's read of the stage frees it for the next load.
// 16 instructions per element instead of 24; statistics as everywhere (common.cuh, KL without cancellation).
#include <cuda.h>

#include "common.cuh"
#include "launch.h"
#include "params.h"

namespace sd {

constexpr int kPwWarps = 16;                       // compute warps = 4-pixel columns of a tile
constexpr int kPwThreads = 32 * (kPwWarps + 1);    // + the producer warp
constexpr int kPwStages = 5;                       // ring stages at most (p.nstages: as many as fit, C = 150: 5, C = 256: 3)
constexpr uint32_t kNegInf2 = 0xff80ff80u;

// bytes of one tensor's tile in a stage: C rows of 128 bytes, padded to the swizzle atom (1024 bytes)
__host__ __device__ inline unsigned pw_tile_bytes(int C) { return ((unsigned)C * 128u + 1023u) & ~1023u; }
size_t pix_warp_smem_bytes(int C, int nstages) {
    return 1024 /* alignment slack */ + (size_t)nstages * 2 * pw_tile_bytes(C) + 2 * kPwStages * sizeof(uint64_t) +
           kPwWarps * sizeof(float);
}
int pix_warp_stages(int C) {
    for (int n = kPwStages; n >= 2; --n)
        if (pix_warp_smem_bytes(C, n) <= 227u * 1024u) return n;
    return 0;
}

// the 16 per-thread sums v[4 p + {zs, zt, a, dd}] of a column's 4 pixels over the 32 lanes in 16 shuffles: lanes L and
// L ^ 1 return the total of v[L >> 1].  Fixed order: deterministic.
__device__ __forceinline__ float warp_sum16_transposed(const float (&v)[16], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0;
    float w[8], x[4], y[2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = b4 ? v[j + 8] : v[j], send = b4 ? v[j] : v[j + 8];
        w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = b3 ? w[j + 4] : w[j], send = b3 ? w[j] : w[j + 4];
        x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = b2 ? x[j + 2] : x[j], send = b2 ? x[j] : x[j + 2];
        y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float keep = b1 ? y[1] : y[0], send = b1 ? y[0] : y[1];
    float t = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// what differs between the element types: pixels of a lane's 8-byte column (4 bf16 / 2 fp32), i.e. H = 2 / 1 pixel pairs
template <typename T>
struct PwTraits;
template <>
struct PwTraits<__nv_bfloat16> {
    static constexpr int kPairs = 2;
    static __device__ __forceinline__ void pair(const uint2& w, int h, float& a, float& b) {
        Elem<__nv_bfloat16>::unpack2(h ? w.y : w.x, a, b);
    }
    // running maxima of the column's pixels, packed like the data
    static __device__ __forceinline__ uint2 max_init() { return make_uint2(kNegInf2, kNegInf2); }
    static __device__ __forceinline__ void max_acc(uint2& m, const uint2& w) {
        asm("max.bf16x2 %0, %1, %2;" : "=r"(m.x) : "r"(m.x), "r"(w.x));
        asm("max.bf16x2 %0, %1, %2;" : "=r"(m.y) : "r"(m.y), "r"(w.y));
    }
    static __device__ __forceinline__ void store_pair(uint2& o, int h, float a, float b) {
        (h ? o.y : o.x) = Elem<__nv_bfloat16>::pack2(a, b);
    }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }   // (bf16 results)
};
template <>
struct PwTraits<float> {
    static constexpr int kPairs = 1;
    static __device__ __forceinline__ void pair(const uint2& w, int, float& a, float& b) {
        a = __uint_as_float(w.x);
        b = __uint_as_float(w.y);
    }
    static __device__ __forceinline__ uint2 max_init() { return make_uint2(0xff800000u, 0xff800000u); }
    static __device__ __forceinline__ void max_acc(uint2& m, const uint2& w) {
        m.x = __float_as_uint(fmaxf(__uint_as_float(m.x), __uint_as_float(w.x)));
        m.y = __float_as_uint(fmaxf(__uint_as_float(m.y), __uint_as_float(w.y)));
    }
    static __device__ __forceinline__ void store_pair(uint2& o, int, float a, float b) {
        o.x = __float_as_uint(a);
        o.y = __float_as_uint(b);
    }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
};

// CPT = channels per lane = ceil(C / 32); only the last slot of a lane can be missing (C > 32 (CPT - 1))
template <typename T, int CPT>
__global__ void __launch_bounds__(kPwThreads, 1)
kl_pixels_warp_kernel(const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapT,
                      const __grid_constant__ CUtensorMap mapD, const PixParams p) {
    using W = PwTraits<T>;
    constexpr int H = W::kPairs;               // pixel pairs of a column
    constexpr int PXW = 2 * H;                 // pixels of a column
    constexpr int NVAL = 4 * PXW;              // sums of a column: v[4 px + {zs, zt, a, dd}]
    constexpr int SH = NVAL == 16 ? 1 : 2;     // the transposed reduction leaves value L >> SH on lane L
    constexpr int LPP = 4 << SH;               // lanes per pixel (8 / 16): as many tiles share one evaluation of the logarithms
    constexpr int kTilePx = PXW * kPwWarps;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const unsigned tile_bytes = pw_tile_bytes(p.C);
    const unsigned stage_bytes = 2 * tile_bytes;
    const int nst = p.nstages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)nst * stage_bytes);
    uint64_t* done = full + kPwStages;
    float* warp_kl = reinterpret_cast<float*>(done + kPwStages);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&done[s], kPwWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // tiles of this CTA: blockIdx.x, + gridDim.x, ...  (32-bit arithmetic: the caller checks total_tiles < 2^31)
    const unsigned total = (unsigned)p.total_tiles, tps = (unsigned)p.tiles_per_sample;
    const int my_tiles = total > blockIdx.x ? (int)((total - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto tile_coords = [&](int k, int& b, int& px0) {
        const unsigned tile = blockIdx.x + (unsigned)k * gridDim.x;
        b = (int)(tile / tps);
        px0 = (int)(tile - (unsigned)b * tps) * kTilePx;
    };

    if (warp == kPwWarps) {
        // ============================ producer: loads, and the tensor store of every finished stage ============================
        if (lane == 0) {
            tma_prefetch_desc(&mapS);
            tma_prefetch_desc(&mapT);
            tma_prefetch_desc(&mapD);
            const uint64_t pol = l2_policy_evict_first();
            int st = 0;
            uint32_t par = 1;                  // parity of the stage's previous use
            for (int k = 0; k < my_tiles + nst; ++k, ++st) {
                if (st == nst) {
                    st = 0;
                    par ^= 1u;
                }
                unsigned char* dst = smem + (size_t)st * stage_bytes;
                if (k >= nst) {
                    // tile k - nst lived here: its gradient is complete when the 16 warps have arrived
                    const int kp = k - nst;
                    mbar_wait_sleep<200>(&done[st], par);
                    int b, px0;
                    tile_coords(kp, b, px0);
                    tma_tile3d_s2g(&mapD, px0, 0, b, dst);
                    tma_bulk_commit();
                    if (k < my_tiles) tma_bulk_wait_read0();      // the stage is read: it may be loaded into again
                }
                if (k < my_tiles) {
                    int b, px0;
                    tile_coords(k, b, px0);
                    mbar_arrive_expect_tx(&full[st], (uint32_t)p.C * 256u);
                    tma_tile3d_g2s(dst, &mapS, px0, 0, b, &full[st], pol);
                    tma_tile3d_g2s(dst + tile_bytes, &mapT, px0, 0, b, &full[st], pol);
                }
            }
            tma_bulk_wait0();
        }
    } else {
        // ============================ 16 compute warps: column `warp` of every tile ============================
        const float c2 = p.c2;
        // my 8 bytes of row c = lane + 32 j: chunk (warp >> 1) ^ (c & 7) of the swizzled row, half (warp & 1); c & 7 = lane & 7
        const unsigned col_off = (unsigned)lane * 128u + ((((unsigned)warp >> 1) ^ ((unsigned)lane & 7u)) << 4) + ((unsigned)warp & 1u) * 8u;
        const bool last_ok = lane + 32 * (CPT - 1) < p.C;
        float acc_kl = 0.f;
        // the four statistics of one pixel, collected over LPP tiles: lane LPP px + i keeps pixel px of the column in
        // tile (k & ~(LPP - 1)) + i; the logarithms of kl_from_stats then run once per LPP tiles with all lanes busy
        float q_zs = 1.f, q_zt = 1.f, q_a = 0.f, q_dd = 0.f;
        int st = 0;
        uint32_t par = 0;
        auto flush_kl = [&](int k_last) {      // tiles (k_last & ~(LPP - 1)) .. k_last
            const int kk = (k_last & ~(LPP - 1)) + (lane & (LPP - 1));
            if (kk <= k_last) {
                int b, px0;
                tile_coords(kk, b, px0);
                const int px = px0 + PXW * warp + lane / LPP;
                if (px < p.HW) {
                    const float kl = kl_from_stats(q_zs, q_zt, q_a, q_dd);
                    p.row_kl[(size_t)b * p.HW + px] = kl;
                    acc_kl += kl;
                }
            }
        };
        for (int k = 0; k < my_tiles; ++k, ++st) {
            if (st == nst) {
                st = 0;
                par ^= 1u;
            }
            mbar_wait(&full[st], par);
            unsigned char* ts = smem + (size_t)st * stage_bytes + col_off;
            const unsigned char* tt = ts + tile_bytes;

            // ---- my channels of the column as they lie in the tile; maxima over them
            uint2 rs[CPT], rt[CPT];
            uint2 ms2 = W::max_init(), mt2 = W::max_init();
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                if (j < CPT - 1 || last_ok) {
                    rs[j] = *reinterpret_cast<const uint2*>(ts + j * 4096);
                    rt[j] = *reinterpret_cast<const uint2*>(tt + j * 4096);
                    W::max_acc(ms2, rs[j]);
                    W::max_acc(mt2, rt[j]);
                }
            }
            // ---- pixel maxima over the 32 lanes -> references of the exponents
            F2 NS[H], NT[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float a0, a1;
                W::pair(ms2, h, a0, a1);
                NS[h] = f2_make(-__fmul_rn(warp_max_uniform(a0), c2), -__fmul_rn(warp_max_uniform(a1), c2));
                W::pair(mt2, h, a0, a1);
                NT[h] = f2_make(-__fmul_rn(warp_max_uniform(a0), c2), -__fmul_rn(warp_max_uniform(a1), c2));
            }

            // ---- exponentials (kept in registers) and the thread's sums per pixel pair
            F2 es[CPT][H], et[CPT][H];
            float v[NVAL];
            {
                const F2 C2 = f2_dup(c2), neg1 = f2_dup(-1.f);
                F2 ZS[H], ZT[H], A[H], DD[H];
#pragma unroll
                for (int h = 0; h < H; ++h) ZS[h] = ZT[h] = A[h] = DD[h] = f2_dup(0.f);
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    if (j < CPT - 1 || last_ok) {
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            float s0, s1, t0, t1;
                            W::pair(rs[j], h, s0, s1);
                            W::pair(rt[j], h, t0, t1);
                            const F2 as2 = f2_fma(f2_make(s0, s1), C2, NS[h]), at2 = f2_fma(f2_make(t0, t1), C2, NT[h]);
                            float as0, as1, at0, at1;
                            f2_split(as2, as0, as1);
                            f2_split(at2, at0, at1);
                            es[j][h] = f2_make(fast_exp2(as0), fast_exp2(as1));
                            et[j][h] = f2_make(fast_exp2(at0), fast_exp2(at1));
                            ZS[h] = f2_add(ZS[h], es[j][h]);
                            ZT[h] = f2_add(ZT[h], et[j][h]);
                            DD[h] = f2_add(DD[h], f2_fma(es[j][h], neg1, et[j][h]));
                            A[h] = f2_fma(et[j][h], f2_fma(as2, neg1, at2), A[h]);
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < H; ++h) es[j][h] = et[j][h] = f2_dup(0.f);
                    }
                }
#pragma unroll
                for (int h = 0; h < H; ++h) {          // pixel 2 h + {0, 1}: v[4 px + {zs, zt, a, dd}]
                    f2_split(ZS[h], v[8 * h + 0], v[8 * h + 4]);
                    f2_split(ZT[h], v[8 * h + 1], v[8 * h + 5]);
                    f2_split(A[h], v[8 * h + 2], v[8 * h + 6]);
                    f2_split(DD[h], v[8 * h + 3], v[8 * h + 7]);
                }
            }
            // ---- column sums: lane L holds value L >> SH = 4 px + stat
            float tot;
            if constexpr (NVAL == 16) tot = warp_sum16_transposed(v, lane);
            else tot = warp_sum8_transposed(v, lane);
            // the statistics of pixel px (lanes LPP px + {0, 1, 2, 3} << SH) -> lane LPP px + (k & (LPP - 1))
            {
                const int src = lane & ~(LPP - 1);
                const float zs = __shfl_sync(0xffffffffu, tot, src), zt = __shfl_sync(0xffffffffu, tot, src + (1 << SH)),
                            a = __shfl_sync(0xffffffffu, tot, src + (2 << SH)), dd = __shfl_sync(0xffffffffu, tot, src + (3 << SH));
                if ((lane & (LPP - 1)) == (k & (LPP - 1))) {
                    q_zs = zs;
                    q_zt = zt;
                    q_a = a;
                    q_dd = dd;
                }
                if ((k & (LPP - 1)) == LPP - 1 || k == my_tiles - 1) flush_kl(k);
            }
            // gradient factors: lanes holding zs / zt divide, everybody fetches the pixels' pairs
            const float kq = W::div(p.coef, tot);
            F2 KS[H], NKT[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                KS[h] = f2_make(__shfl_sync(0xffffffffu, kq, (2 * h) * LPP), __shfl_sync(0xffffffffu, kq, (2 * h + 1) * LPP));
                NKT[h] = f2_make(-__shfl_sync(0xffffffffu, kq, (2 * h) * LPP + (1 << SH)),
                                 -__shfl_sync(0xffffffffu, kq, (2 * h + 1) * LPP + (1 << SH)));
            }

            // ---- gradient in place over my column of the S tile
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                if (j < CPT - 1 || last_ok) {
                    uint2 o;
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        float o0, o1;
                        f2_split(f2_fma(es[j][h], KS[h], f2_mul(et[j][h], NKT[h])), o0, o1);
                        W::store_pair(o, h, o0, o1);
                    }
                    *reinterpret_cast<uint2*>(ts + j * 4096) = o;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[st]);
        }
        // ---- loss: warp partials (fixed order)
        const float wk = warp_sum(acc_kl);
        if (lane == 0) warp_kl[warp] = wk;
    }
    __syncthreads();
    if (tid < 32) {
        unsigned ticket = 0;
        if (lane == 0) {
            float s = 0.f;
            for (int w = 0; w < kPwWarps; ++w) s += warp_kl[w];
            __stcg(&p.cta_part[blockIdx.x], s);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double kl = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) kl += (double)__ldcg(&p.cta_part[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kl += __shfl_down_sync(0xffffffffu, kl, o);
            if (lane == 0) {
                *p.loss = (float)((double)p.loss_scale * kl);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

template <typename T, int CPT>
static cudaError_t launch_pw_t(const CUtensorMap& mS, const CUtensorMap& mT, const CUtensorMap& mD, const PixParams& p, int grid,
                               cudaStream_t stream) {
    auto kern = kl_pixels_warp_kernel<T, CPT>;
    const size_t smem = pix_warp_smem_bytes(p.C, p.nstages);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kPwThreads, smem, stream>>>(mS, mT, mD, p);
    return cudaGetLastError();
}

int kl_pixels_warp_tile_pixels(bool bf16) { return (bf16 ? 4 : 2) * kPwWarps; }

template <typename T>
static cudaError_t launch_pw_c(const CUtensorMap& mS, const CUtensorMap& mT, const CUtensorMap& mD, const PixParams& p, int grid,
                               cudaStream_t stream) {
    switch ((p.C + 31) / 32) {
        case 1: return launch_pw_t<T, 1>(mS, mT, mD, p, grid, stream);
        case 2: return launch_pw_t<T, 2>(mS, mT, mD, p, grid, stream);
        case 3: return launch_pw_t<T, 3>(mS, mT, mD, p, grid, stream);
        case 4: return launch_pw_t<T, 4>(mS, mT, mD, p, grid, stream);
        case 5: return launch_pw_t<T, 5>(mS, mT, mD, p, grid, stream);
        case 6: return launch_pw_t<T, 6>(mS, mT, mD, p, grid, stream);
        case 7: return launch_pw_t<T, 7>(mS, mT, mD, p, grid, stream);
        case 8: return launch_pw_t<T, 8>(mS, mT, mD, p, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

// no AT term, C <= 256, maps encoded with 128-byte swizzle and boxes of (128 bytes of pixels, C, 1)
cudaError_t launch_kl_pixels_warp(const void* mapS, const void* mapT, const void* mapD, const PixParams& p, bool bf16, int grid,
                                  cudaStream_t stream) {
    const CUtensorMap& mS = *static_cast<const CUtensorMap*>(mapS);
    const CUtensorMap& mT = *static_cast<const CUtensorMap*>(mapT);
    const CUtensorMap& mD = *static_cast<const CUtensorMap*>(mapD);
    return bf16 ? launch_pw_c<__nv_bfloat16>(mS, mT, mD, p, grid, stream) : launch_pw_c<float>(mS, mT, mD, p, grid, stream);
}

}  // namespace sd
