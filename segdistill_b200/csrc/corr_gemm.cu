// CGD per-group correlation (Gram) loss on the 5th-generation tensor cores - a builder-defined
// EXTENSION: the reference has no contraction on this path (SURVEY.md box 1, 8 a7), so parity is
// unpinned; the oracle is oracle/kld_oracle.py:corr_loss_torch.
//
//   per (sample, group of g channels):  X in R^{g x HW},  G = X X^T / HW
//   loss = alpha * mean_{b,grp,i,j} (G_S - G_T)^2,        dS = 4 alpha/(N HW) (G_S - G_T) X_S
//
// Channels are cut into BLOCKS of whole groups (<= 256 channels, balanced over the sample); the block is
// the GEMM problem: the full CBxCB Gram of the block runs on tcgen05 and only its block-diagonal (the
// groups) is used - the contraction is HBM-bound for every shipped g (AI = g flop/B), the spare tensor
// throughput pays for the off-diagonal waste.
//
//   K1 corr_gram_kernel   (block, HW split): E_part = X_S X_S^T - X_T X_T^T over its HW range.  TMA (2-D
//                         tiled, 128-byte swizzle) -> shared ring -> tcgen05.mma (tf32 / bf16, K-major A and
//                         B from the same tile; the teacher pass sets the descriptor's negate-A bit so one
//                         TMEM accumulator holds the difference) -> tcgen05.ld -> partial E to the workspace.
//   K2 corr_mask_kernel   sums the splits in a fixed order, keeps the block-diagonal, accumulates the loss and
//                         writes coef*(G_S-G_T) as the A operand of K3 in the canonical no-swizzle layout.
//   K3 corr_grad_kernel   (block, HW split): dX_S = D X_S: A = D (K-major, shared), B = the X_S tile again, now
//                         read MN-major (HW contiguous), accumulator in TMEM, epilogue TMEM -> registers -> dS.
//
// Warp roles in K1/K3: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warps 2-5 epilogue (one
// TMEM lane quarter each).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstring>
#include <type_traits>

#include "../../include/segdistill.h"
#include "common.cuh"
#include "launch.h"
#include "params.h"

namespace sd {

constexpr int kGThreads = 192;
constexpr int kGTileBytes = 128 * 128;          // one operand tile: 128 rows x 128 bytes (swizzle-128B atom rows)
constexpr int kGTmemCols = 512;

// ---------------------------------------------------------------- tcgen05 / TMA primitives
__device__ __forceinline__ void tma_tile2d_g2s(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread for the CTA
template <bool BF16>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (BF16) {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// mbarrier arrives once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets
// (all >> 4), version 1 (sm_100), layout type in bits 61-63 (0 = no swizzle, 2 = 128-byte swizzle)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), fp32 accumulation
__host__ __device__ constexpr uint32_t instr_desc(bool bf16, int M, int N, bool a_neg, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | ((bf16 ? 1u : 2u) << 7) | ((bf16 ? 1u : 2u) << 10) | ((a_neg ? 1u : 0u) << 13) |
           ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct CorrParams {
    int B, C, HW;
    int g;            // channels per group
    int nb;           // groups per block
    int CB;           // channels per (complete) block = nb*g
    int CBp;          // ... rounded up to 16: N of the Gram, K of the gradient GEMM
    int MT;           // 128-row tiles of a block: 1 or 2
    int blocks_per_sample;
    int nblocks;
    int ksplit;       // HW splits of K1
    int nsplit;       // HW splits of K3
    int kbox;         // elements of HW per TMA box (128 bytes)
    int rows0, rows1; // rows of the TMA boxes of tile 0 / tile 1
    int nstages;
    int gboxes;       // 128-byte HW boxes per K3 tile
    float inv_hw;
    float dcoef;      // grad_scale * 4 alpha / (N HW)
    float loss_scale; // alpha / N
    float* partial;   // [nblocks][ksplit][MT*128][CBp]
    unsigned char* dmask;   // [nblocks][dmask_bytes]  A operand of K3, canonical no-swizzle K-major layout
    int dmask_bytes;
    float* blk_loss;  // [nblocks]
    unsigned* ctrl;
    float* loss;
    void* dS;
};

// ====================================================================================================
// K1: partial Gram difference of one block over one HW range
// ====================================================================================================
template <bool BF16>
__global__ void __launch_bounds__(kGThreads, 1)
corr_gram_kernel(const __grid_constant__ CUtensorMap mapS0, const __grid_constant__ CUtensorMap mapT0,
                 const __grid_constant__ CUtensorMap mapS1, const __grid_constant__ CUtensorMap mapT1, const CorrParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle-128B tiles: 1024-byte aligned
    // stage layout: [S tile 0][S tile 1][T tile 0][T tile 1] (tile 1 only when MT == 2), 16 KB each
    const int stage_bytes = 2 * p.MT * kGTileBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * stage_bytes);
    uint64_t* empty = full + p.nstages;
    uint64_t* acc_full = empty + p.nstages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x / p.ksplit, split = blockIdx.x - blk * p.ksplit;
    const int b = blk / p.blocks_per_sample, bi = blk - b * p.blocks_per_sample;
    const int row0 = b * p.C + bi * p.CB;                       // first channel row of the block in the (B*C, HW) view
    const int nk_total = (p.HW + p.kbox - 1) / p.kbox;          // K steps of the whole HW
    const int k_lo = (int)((long long)nk_total * split / p.ksplit), k_hi = (int)((long long)nk_total * (split + 1) / p.ksplit);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tc_alloc(tmem_slot, (uint32_t)(p.MT * 256));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&mapS0);
            tma_prefetch_desc(&mapT0);
            const uint32_t bytes = (uint32_t)(2 * (p.rows0 + (p.MT == 2 ? p.rows1 : 0))) * 128u;
            int s = 0;
            uint32_t ph = 0;
            for (int k = k_lo; k < k_hi; ++k) {
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full[s], bytes);
                unsigned char* st = smem + (size_t)s * stage_bytes;
                tma_tile2d_g2s(st, &mapS0, k * p.kbox, row0, &full[s]);
                tma_tile2d_g2s(st + p.MT * kGTileBytes, &mapT0, k * p.kbox, row0, &full[s]);
                if (p.MT == 2) {
                    tma_tile2d_g2s(st + kGTileBytes, &mapS1, k * p.kbox, row0 + 128, &full[s]);
                    tma_tile2d_g2s(st + 3 * kGTileBytes, &mapT1, k * p.kbox, row0 + 128, &full[s]);
                }
                if (++s == p.nstages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_pos = instr_desc(BF16, 128, p.CBp, false, false, false);
            const uint32_t id_neg = instr_desc(BF16, 128, p.CBp, true, false, false);
            int s = 0;
            uint32_t ph = 0;
            uint32_t acc = 0;
            for (int k = k_lo; k < k_hi; ++k) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(smem + (size_t)s * stage_bytes);
                for (int t = 0; t < 2; ++t) {                      // student, then teacher with A negated
                    const uint32_t tb = sb + (uint32_t)(t * p.MT * kGTileBytes);
                    for (int m = 0; m < p.MT; ++m) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {               // 4 x 32 bytes of K per 128-byte box
                            const uint64_t ad = smem_desc(tb + (uint32_t)(m * kGTileBytes + kk * 32), 16, 1024, 2);
                            const uint64_t bd = smem_desc(tb + (uint32_t)(kk * 32), 16, 1024, 2);
                            // the very first MMA into an accumulator tile overwrites it, every later one accumulates
                            tc_mma<BF16>(tmem + (uint32_t)(m * 256), ad, bd, t ? id_neg : id_pos, acc | (uint32_t)(t | kk));
                        }
                    }
                }
                acc = 1;
                tc_commit(&empty[s]);                                  // the stage may be refilled once these MMAs are done
                if (++s == p.nstages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            tc_commit(acc_full);
        }
    } else {
        // epilogue: TMEM -> registers -> partial E in the workspace (row-major [MT*128][CBp])
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int q = warp & 3;                                         // TMEM lane quarter of this warp
        float* dst = p.partial + ((size_t)blk * p.ksplit + split) * (size_t)(p.MT * 128) * p.CBp;
        for (int m = 0; m < p.MT; ++m) {
            const int row = m * 128 + q * 32 + lane;
            for (int c0 = 0; c0 < p.CBp; c0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * 256 + c0), v);
                tc_wait_ld(v);
                float4* o = reinterpret_cast<float4*>(dst + (size_t)row * p.CBp + c0);
                if (k_hi > k_lo) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, (uint32_t)(p.MT * 256));
}

// ====================================================================================================
// K2: sum the splits, keep the block-diagonal, loss, A operand of K3
// ====================================================================================================
constexpr int kMaskPerCta = 2048;   // Gram entries per CTA of K2

template <bool BF16>
__global__ void __launch_bounds__(256) corr_mask_kernel(const CorrParams p) {
    __shared__ double red[8];
    const int M = p.MT * 128;
    const int chunks = (M * p.CBp + kMaskPerCta - 1) / kMaskPerCta;
    const int blk = blockIdx.x / chunks, chunk = blockIdx.x - blk * chunks;
    const int bi = blk % p.blocks_per_sample;
    const int c_real = min(p.CB, p.C - bi * p.CB);                 // real channels of this block
    const float* part = p.partial + (size_t)blk * p.ksplit * (size_t)M * p.CBp;
    unsigned char* dm = p.dmask + (size_t)blk * p.dmask_bytes;
    constexpr int ES = BF16 ? 2 : 4;
    constexpr int CH = 16 / ES;                                    // elements per 16-byte chunk of the A operand
    const size_t tile_bytes = (size_t)(p.CBp / CH) * 2048;         // A operand of one 128-row tile
    double acc = 0.0;
    // one thread = one 16-byte chunk of the operand = CH consecutive columns of one row (CBp % 16 == 0: never straddles)
    const int vend = min(M * p.CBp, (chunk + 1) * kMaskPerCta) / CH;
    for (int v = chunk * (kMaskPerCta / CH) + threadIdx.x; v < vend; v += 256) {
        const int idx = v * CH;
        const int i = idx / p.CBp, j0 = idx - i * p.CBp;
        float e[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) e[q] = 0.f;
        for (int s = 0; s < p.ksplit; ++s) {                        // fixed order
            const float4* src = reinterpret_cast<const float4*>(part + (size_t)s * M * p.CBp + idx);
#pragma unroll
            for (int q = 0; q < CH / 4; ++q) {
                const float4 x = __ldcs(src + q);
                e[4 * q] += x.x;
                e[4 * q + 1] += x.y;
                e[4 * q + 2] += x.z;
                e[4 * q + 3] += x.w;
            }
        }
        float a[CH];
        const int gi = i / p.g;
        int gj = j0 / p.g, rem = j0 - gj * p.g;                     // group of column j0 + q, walked without divisions
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            const int j = j0 + q;
            const bool keep = i < c_real && j < c_real && gi == gj;
            if (++rem == p.g) {
                rem = 0;
                ++gj;
            }
            const float d = keep ? e[q] * p.inv_hw : 0.f;
            acc += (double)d * (double)d;
            a[q] = p.dcoef * d;
        }
        // canonical K-major layout without swizzle: 8-row x 16-byte core matrices; row groups 128 bytes apart
        // (SBO), 16-byte K chunks 2048 bytes apart (LBO); one such operand per 128-row tile
        const int m = i >> 7, r = i & 127;
        const size_t off = (size_t)m * tile_bytes + (size_t)(j0 / CH) * 2048 + (size_t)(r >> 3) * 128 + (size_t)(r & 7) * 16;
        if (BF16) {
            uint4 w;
            w.x = Elem<__nv_bfloat16>::pack2(a[0], a[1]);
            w.y = Elem<__nv_bfloat16>::pack2(a[2], a[3]);
            w.z = Elem<__nv_bfloat16>::pack2(a[CH - 4], a[CH - 3]);
            w.w = Elem<__nv_bfloat16>::pack2(a[CH - 2], a[CH - 1]);
            *reinterpret_cast<uint4*>(dm + off) = w;
        } else {
            *reinterpret_cast<float4*>(dm + off) = make_float4(a[0], a[1], a[2], a[3]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned ticket = 0;
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (int w = 0; w < 8; ++w) a += red[w];
            __stcg(&p.blk_loss[blockIdx.x], (float)a);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double t = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += (double)__ldcg(&p.blk_loss[i]);   // fixed order
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) {
                *p.loss = (float)((double)p.loss_scale * t);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

// ====================================================================================================
// K3: dX_S = D X_S for one block over one HW range
// ====================================================================================================
constexpr int kGradBoxes = 2;     // 128-byte HW boxes per tile (1 when shared memory is short): N = 64 fp32 / 128 bf16 positions

template <bool BF16>
__global__ void __launch_bounds__(kGThreads, 1)
corr_grad_kernel(const __grid_constant__ CUtensorMap mapS0, const __grid_constant__ CUtensorMap mapS1, const CorrParams p) {
    using T = typename std::conditional<BF16, __nv_bfloat16, float>::type;
    constexpr int ES = BF16 ? 2 : 4;
    constexpr int KSTEP = BF16 ? 16 : 8;                            // channels per MMA
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // [A = the CTA's 128 rows of D][stages: gboxes x MT tiles of X_S][barriers]
    const int a_tile_bytes = (p.CBp * ES / 16) * 2048;
    const int a_bytes = (a_tile_bytes + 1023) & ~1023;
    const int stage_bytes = p.gboxes * p.MT * kGTileBytes;
    unsigned char* stages = smem + a_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + (size_t)p.nstages * stage_bytes);
    uint64_t* empty = full + p.nstages;
    uint64_t* acc_full = empty + p.nstages;       // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA = (block, 128-row tile of the output, HW range)
    const int per_blk = p.MT * p.nsplit;
    const int blk = blockIdx.x / per_blk;
    const int mt = (blockIdx.x - blk * per_blk) / p.nsplit, split = blockIdx.x - blk * per_blk - mt * p.nsplit;
    const int b = blk / p.blocks_per_sample, bi = blk - b * p.blocks_per_sample;
    const int row0 = b * p.C + bi * p.CB;
    const int c_real = min(p.CB, p.C - bi * p.CB);
    const int BN = p.gboxes * p.kbox;                              // HW positions per tile
    const int nt_total = (p.HW + BN - 1) / BN;
    const int t_lo = (int)((long long)nt_total * split / p.nsplit), t_hi = (int)((long long)nt_total * (split + 1) / p.nsplit);
    const int ncols = BN < 32 ? 32 : BN;                           // TMEM columns of one accumulator buffer
    const uint32_t tmem_cols = 2 * ncols < 32 ? 32 : 2 * ncols;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 0) tc_alloc(tmem_slot, tmem_cols);
    // A operand: rows [128 mt, 128 mt + 128) of the block's masked, scaled Gram difference (canonical layout)
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.dmask + (size_t)blk * p.dmask_bytes + (size_t)mt * a_tile_bytes);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < a_tile_bytes / 16; i += kGThreads) dst[i] = __ldg(src + i);
    }
    fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&mapS0);
            const uint32_t bytes = (uint32_t)(p.gboxes * (p.rows0 + (p.MT == 2 ? p.rows1 : 0))) * 128u;
            int s = 0;
            uint32_t ph = 0;
            for (int t = t_lo; t < t_hi; ++t) {
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full[s], bytes);
                unsigned char* st = stages + (size_t)s * stage_bytes;
                for (int x = 0; x < p.gboxes; ++x) {
                    // box x of row tile kt at (x*MT + kt) * 16 KB: the HW chunks of one 128-row tile are MT*16 KB apart
                    tma_tile2d_g2s(st + (size_t)(x * p.MT) * kGTileBytes, &mapS0, t * BN + x * p.kbox, row0, &full[s]);
                    if (p.MT == 2)
                        tma_tile2d_g2s(st + (size_t)(x * p.MT + 1) * kGTileBytes, &mapS1, t * BN + x * p.kbox, row0 + 128,
                                       &full[s]);
                }
                if (++s == p.nstages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc(BF16, 128, BN, false, false, true);      // A K-major, B MN-major
            const uint32_t a_base = smem_u32(smem);
            int s = 0;
            uint32_t ph = 0;
            for (int t = t_lo; t < t_hi; ++t) {
                const int buf = (t - t_lo) & 1;
                const uint32_t aph = (uint32_t)((t - t_lo) >> 1) & 1u;
                mbar_wait(&acc_empty[buf], aph ^ 1u);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(stages + (size_t)s * stage_bytes);
                for (int k = 0; k < p.CBp; k += KSTEP) {
                    // A: K chunk k (KSTEP*ES = 32 bytes = 2 core matrices along K, 2048 bytes apart; row groups 128)
                    const uint64_t ad = smem_desc(a_base + (uint32_t)((k * ES / 16) * 2048), 2048, 128, 0);
                    // B: channels k.. of the X_S tile, read MN-major (HW contiguous): row groups along K (SBO), the
                    // 128-byte HW chunks MT*16 KB apart (LBO); channel k lives in row tile k/128.  tf32 operands read
                    // MN-major use the 32-byte-atom swizzle (4-row groups, layout type 1)
                    const int kt = k >> 7, kr = k & 127;
                    const uint64_t bd = smem_desc(sb + (uint32_t)(kt * kGTileBytes + kr * 128), (uint32_t)(p.MT * kGTileBytes),
                                                  BF16 ? 1024 : 512, BF16 ? 2 : 1);
                    tc_mma<BF16>(tmem + (uint32_t)(buf * ncols), ad, bd, idesc, k > 0 ? 1u : 0u);
                }
                tc_commit(&empty[s]);
                tc_commit(&acc_full[buf]);
                if (++s == p.nstages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        const int q = warp & 3;
        T* out = static_cast<T*>(p.dS);
        const int ch = mt * 128 + q * 32 + lane;                     // channel of the block = TMEM lane
        for (int t = t_lo; t < t_hi; ++t) {
            const int buf = (t - t_lo) & 1;
            const uint32_t aph = (uint32_t)((t - t_lo) >> 1) & 1u;
            mbar_wait(&acc_full[buf], aph);
            tc_fence_after();
            T* orow = out + (size_t)(row0 + ch) * p.HW + (size_t)t * BN;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ncols + c0), v);
                tc_wait_ld(v);
                if (ch < c_real) {
                    const int pos = t * BN + c0;
                    if (pos + 16 <= p.HW) {
                        if (BF16) {
                            uint4 w0, w1;
                            w0.x = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[0]), __uint_as_float(v[1]));
                            w0.y = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[2]), __uint_as_float(v[3]));
                            w0.z = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[4]), __uint_as_float(v[5]));
                            w0.w = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[6]), __uint_as_float(v[7]));
                            w1.x = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[8]), __uint_as_float(v[9]));
                            w1.y = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[10]), __uint_as_float(v[11]));
                            w1.z = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[12]), __uint_as_float(v[13]));
                            w1.w = Elem<__nv_bfloat16>::pack2(__uint_as_float(v[14]), __uint_as_float(v[15]));
                            uint4* o = reinterpret_cast<uint4*>(orow + c0);
                            o[0] = w0;
                            o[1] = w1;
                        } else {
                            float4* o = reinterpret_cast<float4*>(orow + c0);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                        }
                    } else {
                        for (int i = 0; i < 16; ++i)
                            if (pos + i < p.HW) Elem<T>::store(orow + c0 + i, __uint_as_float(v[i]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, tmem_cols);
}

// ====================================================================================================
// host side
// ====================================================================================================
struct CorrPlan {
    CorrParams p;
    size_t off_ctrl, off_blk, off_partial, off_dmask, bytes;
    size_t smem1, smem3;
    bool ok;
};

constexpr int kPlanSMs = 148;     // B200; the decomposition must not depend on where the workspace size was asked

static CorrPlan corr_plan(int B, int C, int HW, int group, int dtype) {
    CorrPlan pl;
    memset(&pl, 0, sizeof(pl));
    CorrParams& p = pl.p;
    const int es = dtype == SD_BF16 ? 2 : 4;
    int g = group > C ? C : group;
    const int G = (C + g - 1) / g;
    p.B = B; p.C = C; p.HW = HW; p.g = g;
    pl.ok = g <= 256;
    const int per_tile = g <= 128 ? 128 / g : 1;             // groups that fit 128 (or one group of up to 256) channels
    p.blocks_per_sample = (G + per_tile - 1) / per_tile;
    p.nb = (G + p.blocks_per_sample - 1) / p.blocks_per_sample;   // balanced
    p.blocks_per_sample = (G + p.nb - 1) / p.nb;
    p.CB = p.nb * g;
    p.CBp = (p.CB + 15) & ~15;
    if (p.CBp > 256) pl.ok = false;
    p.MT = p.CBp > 128 ? 2 : 1;
    p.nblocks = B * p.blocks_per_sample;
    p.kbox = 128 / es;
    p.rows0 = p.CBp < 128 ? p.CBp : 128;
    p.rows1 = p.MT == 2 ? p.CBp - 128 : 0;
    const int nk = (HW + p.kbox - 1) / p.kbox;
    // K1: ONE wave of CTAs, never a second, partly filled one (two 128-row tiles: 198 KB of stages, one CTA per SM; one
    // tile: two CTAs per SM); every split costs a partial Gram in the workspace; at least 4 K steps each
    int want = kPlanSMs * (p.MT == 1 ? 2 : 1) / p.nblocks;
    p.ksplit = want < 1 ? 1 : (want > nk / 4 ? (nk / 4 > 0 ? nk / 4 : 1) : want);
    if (p.ksplit > 16) p.ksplit = 16;
    p.dmask_bytes = p.MT * (p.CBp * es / 16) * 2048;
    // K3: the A operand and at least two stages of X_S tiles must fit 200 KB
    p.gboxes = kGradBoxes;
    const int a_tile = ((p.CBp * es / 16) * 2048 + 1023) & ~1023;
    while (p.gboxes > 1 && a_tile + 2 * p.gboxes * p.MT * kGTileBytes > 200 * 1024) --p.gboxes;
    const int nt = (HW + p.gboxes * p.kbox - 1) / (p.gboxes * p.kbox);
    int want3 = 2 * kPlanSMs / (p.nblocks * p.MT);                            // two waves (one CTA per SM), not a third
    p.nsplit = want3 < 1 ? 1 : (want3 > nt ? nt : want3);
    if (p.nsplit > 64) p.nsplit = 64;
    size_t o = kArenaBytes;
    pl.off_ctrl = 0;
    pl.off_blk = o;       o += sizeof(float) * (size_t)p.nblocks * (size_t)((p.MT * 128 * p.CBp + kMaskPerCta - 1) / kMaskPerCta);
    o = (o + 255) & ~(size_t)255;
    pl.off_partial = o;   o += sizeof(float) * (size_t)p.nblocks * p.ksplit * (size_t)(p.MT * 128) * p.CBp;
    o = (o + 255) & ~(size_t)255;
    pl.off_dmask = o;     o += (size_t)p.nblocks * p.dmask_bytes;
    pl.bytes = (o + 255) & ~(size_t)255;
    return pl;
}

// upper bound over both dtypes (the C ABI's workspace query does not take one)
size_t cgd_corr_workspace_bytes(int B, int C, int HW, int group) {
    const size_t a = corr_plan(B, C, HW, group, SD_F32).bytes, b = corr_plan(B, C, HW, group, SD_BF16).bytes;
    return a > b ? a : b;
}

bool cgd_corr_geometry(int B, int C, int HW, int group, int dtype, int* rows0, int* rows1, int* kbox) {
    const CorrPlan pl = corr_plan(B, C, HW, group, dtype);
    *rows0 = pl.p.rows0;
    *rows1 = pl.p.rows1;
    *kbox = pl.p.kbox;
    return pl.ok;
}

template <bool BF16>
static cudaError_t corr_launch_t(CorrPlan& pl, const CUtensorMap* maps, cudaStream_t stream) {
    CorrParams& p = pl.p;
    // K1: as many 2*MT*16 KB stages as fit
    const int st1 = 2 * p.MT * kGTileBytes;
    int ns1 = p.MT == 1 ? 3 : (int)((200 * 1024) / st1);   // one 128-row tile: 98 KB and 256 TMEM columns, two CTAs per SM
    const size_t smem1 = (size_t)ns1 * st1 + 2048;
    auto k1 = corr_gram_kernel<BF16>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) return e;
    p.nstages = ns1;
    k1<<<p.nblocks * p.ksplit, kGThreads, smem1, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    const int chunks = (p.MT * 128 * p.CBp + kMaskPerCta - 1) / kMaskPerCta;
    corr_mask_kernel<BF16><<<p.nblocks * chunks, 256, 0, stream>>>(p);
    const int a_bytes = (((p.CBp * (BF16 ? 2 : 4)) / 16) * 2048 + 1023) & ~1023;
    const int st3 = p.gboxes * p.MT * kGTileBytes;
    int ns3 = (int)((200 * 1024 - a_bytes) / st3);
    if (ns3 > 4) ns3 = 4;
    if (ns3 < 1) return cudaErrorInvalidValue;
    const size_t smem3 = (size_t)a_bytes + (size_t)ns3 * st3 + 2048;
    auto k3 = corr_grad_kernel<BF16>;
    e = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
    if (e != cudaSuccess) return e;
    p.nstages = ns3;
    k3<<<p.nblocks * p.MT * p.nsplit, kGThreads, smem3, stream>>>(maps[4], maps[5], p);
    return cudaGetLastError();
}

// maps: CUtensorMap[6] over the (B*C, HW) views: S, T with box rows0; S, T with box rows1 (Gram); S rows0, S rows1 for
// the gradient GEMM (fp32: 128-byte swizzle with 32-byte atoms)
cudaError_t launch_cgd_corr(void* dS, float* loss, int B, int C, int HW, int group, int dtype, float alpha, float grad_scale,
                            void* workspace, const void* maps, cudaStream_t stream) {
    CorrPlan pl = corr_plan(B, C, HW, group, dtype);
    CorrParams& p = pl.p;
    char* ws = static_cast<char*>(workspace);
    const int G = (C + p.g - 1) / p.g;
    const double N = (double)B * G * p.g * p.g;
    p.inv_hw = (float)(1.0 / (double)HW);
    p.dcoef = (float)((double)grad_scale * 4.0 * (double)alpha / (N * (double)HW));
    p.loss_scale = (float)((double)alpha / N);
    p.ctrl = reinterpret_cast<unsigned*>(ws + pl.off_ctrl);
    p.blk_loss = reinterpret_cast<float*>(ws + pl.off_blk);
    p.partial = reinterpret_cast<float*>(ws + pl.off_partial);
    p.dmask = reinterpret_cast<unsigned char*>(ws + pl.off_dmask);
    p.loss = loss;
    p.dS = dS;
    const CUtensorMap* m = static_cast<const CUtensorMap*>(maps);
    return dtype == SD_BF16 ? corr_launch_t<true>(pl, m, stream) : corr_launch_t<false>(pl, m, stream);
}

}  // namespace sd
