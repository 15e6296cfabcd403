// Row-wise softmax-KL for rows that fit a THREAD-BLOCK CLUSTER: one pass over HBM, no L2 re-read, the row
// parked in TENSOR MEMORY between the statistics pass and the gradient pass.  One or two losses (CD + CGD on
// the same logits) per launch.
//
// Same mathematics as kl_rows.cu (mmseg/models/distillation/losses.py:50-58,:108-112 + backward).
// A CGD row (g = 10 channels of 128x128 logits) is 1.3 MB of S and T: no SM holds it, a cluster of 8
// does.  The cluster owns one "super-row" (a row of the loss with the larger group) at a time; CTA c
// owns slice c of it:
//
//   phase 1   the slice streams through a 6 x 32 KB shared-memory ring (1-D TMA bulk copies, mbarrier
//             full/empty pairs, a dedicated TMA warp).  Each of the 16 consumer warps pulls its vectors of
//             a chunk into registers ONCE, parks the raw bits in TMEM (tcgen05.st, 16 columns per chunk and
//             thread - the 256 KB of tensor memory hold a 32768-element fp32 slice of S and T), hands the
//             ring slot back at once - the next super-row keeps streaming in during everything below - and
//             updates ONLINE softmax statistics (running thread-local maximum, sums rescaled when it
//             moves) per PIECE, the part of one row of the smaller-group loss inside this slice.
//             Warp shuffles -> 16 warp records per piece -> one summary per piece.
//   exchange  every CTA PUSHES its piece summaries into the shared memory of all CTAs of the cluster
//             (st.async through DSMEM, completing bytes on the receiver's mbarrier): no cluster barrier, no
//             release fence behind the gradient stores; one warp per row merges what arrived.
//   phase 2   the slice comes back from TMEM (tcgen05.ld), the gradient is written once.
//
// HBM traffic is the algorithmic read S + read T + write dS.  With two losses whose temperatures are
// tau and 2*tau (CD tau = 1 next to CGD tau = 2, the reference's defaults) exp(x/tau) = exp(x/2tau)^2:
// 2 ex2 per element and phase instead of 4.
#include "rows_common.cuh"

namespace sd {

constexpr int kCCons = 512;
constexpr int kCConsWarps = kCCons / 32;
constexpr int kCThreads = kCCons + 32;
constexpr int kCChunkRows = 2;                         // 4-element vectors per consumer thread, chunk and tensor
constexpr int kCChunkVecs = kCChunkRows * kCCons;      // 1024 vectors = 4096 elements per tensor
constexpr int kCChunkBytes = kCChunkVecs * 16;         // fp32; bf16 chunks fill half a slot
constexpr int kCRing = 6;                              // ring slots of 32 KB (S chunk + T chunk)
constexpr int kCMaxChunks = kClusterMaxChunks;         // chunks per slice: 6 x 20 TMEM columns per thread
constexpr int kCRowCols = 10;                          // TMEM columns of one parked vector-row: 4 + 4 values, 2 references
constexpr int kCMaxPieces = kClusterMaxPieces;
constexpr int kCRecFloats = 8;                         // ms, mt, {zs, zt, a} x 2
constexpr int kCTmemCols = 512;
static_assert(kCMaxChunks * kCChunkRows * kCRowCols * (kCConsWarps / 4) <= kCTmemCols, "TMEM columns");
static_assert(kCChunkVecs == kClusterChunkVecs, "chunk size");

struct ClusterSmem {
    unsigned char ring[kCRing][2][kCChunkBytes];
    uint64_t full[kCRing], empty[kCRing];
    float rec[2][kCMaxPieces][kCConsWarps][kCRecFloats];  // warp records of the pieces, by iteration parity
    float summ[2][kClusterMaxSize][kCMaxPieces][kCRecFloats];  // summaries of every CTA of the cluster (pushed)
    uint64_t xch[2];                                      // ... their bytes complete on these
    float fin[kCMaxPieces + 1][4];                        // row statistics for phase 2: {Ms, Mt, coef/Zs, coef/Zt}
    float klpart[kCConsWarps][2];
    uint32_t tmem_base;
};
constexpr size_t kClusterSmemBytes = sizeof(ClusterSmem);

// ---------------------------------------------------------------- cluster / tensor-memory primitives
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of my shared-memory location `p` in the CTA of rank `cta` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(cta));
    return r;
}
// 16 bytes into a peer's shared memory; the bytes complete on the peer's mbarrier (both cluster addresses)
__device__ __forceinline__ void st_async_f4(uint32_t dst, const float4& v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                 "r"(__float_as_uint(v.w)), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {  // one warp, all lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {  // the allocating warp, all lanes
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// warp-collective: thread i of the warp owns TMEM lane (lane quarter of the warp) + i, 8 consecutive columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint4& a, uint4& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
// the loaded registers are operands of the wait: nothing may read (or copy) them before it
__device__ __forceinline__ void tmem_wait_ld(uint4& a, uint4& b, uint32_t& c, uint32_t& d) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w), "+r"(b.x), "+r"(b.y), "+r"(b.z), "+r"(b.w), "+r"(c), "+r"(d)
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <typename V>
__device__ __forceinline__ uint4 as_bits(const V& v) {
    return *reinterpret_cast<const uint4*>(&v);
}

// ---------------------------------------------------------------- statistics of a part of a row
// NL losses share the raw maxima; sums are relative to them
template <int NL>
struct PStat {
    float ms, mt;
    float zs[NL], zt[NL], a[NL];
};
template <int NL>
__device__ __forceinline__ PStat<NL> pstat_empty() {
    PStat<NL> r;
    r.ms = kMaxFloor;
    r.mt = kMaxFloor;
#pragma unroll
    for (int k = 0; k < NL; ++k) r.zs[k] = r.zt[k] = r.a[k] = 0.f;
    return r;
}
// exp2((x - ref) * c2[k]) for every loss; R == 2: c2[0] == 2*c2[1], so e[0] = e[1]^2 (one ex2 for both)
template <int NL, int R>
__device__ __forceinline__ void exps(float x, float ref, const float (&c2)[NL], float (&e)[NL]) {
    if (NL == 1) {
        e[0] = fast_exp2((x - ref) * c2[0]);
    } else if (R == 2) {
        e[NL - 1] = fast_exp2((x - ref) * c2[NL - 1]);
        e[0] = e[NL - 1] * e[NL - 1];
    } else {
#pragma unroll
        for (int k = 0; k < NL; ++k) e[k] = fast_exp2((x - ref) * c2[k]);
    }
}
// the same with the references pre-multiplied (ref2[k] = ref * c2[k]): one FFMA per exponent
template <int NL, int R>
__device__ __forceinline__ void exps(float x, const float (&ref2)[NL], const float (&c2)[NL], float (&e)[NL]) {
    if (NL == 1) {
        e[0] = fast_exp2(fmaf(x, c2[0], -ref2[0]));
    } else if (R == 2) {
        e[NL - 1] = fast_exp2(fmaf(x, c2[NL - 1], -ref2[NL - 1]));
        e[0] = e[NL - 1] * e[NL - 1];
    } else {
#pragma unroll
        for (int k = 0; k < NL; ++k) e[k] = fast_exp2(fmaf(x, c2[k], -ref2[k]));
    }
}
// reduce over `width` lanes (xor butterfly; every lane ends with the same bits): maxima first, then the
// sums rescaled to them - one exponential stage instead of one per butterfly step
template <int NL, int R>
__device__ __forceinline__ PStat<NL> pstat_reduce(const PStat<NL>& x, const float (&c2)[NL], int width) {
    PStat<NL> r;
    r.ms = x.ms;
    r.mt = x.mt;
    for (int o = width >> 1; o > 0; o >>= 1) {
        r.ms = fmaxf(r.ms, __shfl_xor_sync(0xffffffffu, r.ms, o));
        r.mt = fmaxf(r.mt, __shfl_xor_sync(0xffffffffu, r.mt, o));
    }
    float fs[NL], ft[NL];
    exps<NL, R>(x.ms, r.ms, c2, fs);
    exps<NL, R>(x.mt, r.mt, c2, ft);
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        r.zs[k] = x.zs[k] * fs[k];
        r.zt[k] = x.zt[k] * ft[k];
        r.a[k] = x.a[k] * ft[k];
    }
    for (int o = width >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            r.zs[k] += __shfl_xor_sync(0xffffffffu, r.zs[k], o);
            r.zt[k] += __shfl_xor_sync(0xffffffffu, r.zt[k], o);
            r.a[k] += __shfl_xor_sync(0xffffffffu, r.a[k], o);
        }
    }
    return r;
}
// y folded into x (both parts of the same row)
template <int NL, int R>
__device__ __forceinline__ PStat<NL> pstat_merge(const PStat<NL>& x, const PStat<NL>& y, const float (&c2)[NL]) {
    PStat<NL> r;
    r.ms = fmaxf(x.ms, y.ms);
    r.mt = fmaxf(x.mt, y.mt);
    float fxs[NL], fys[NL], fxt[NL], fyt[NL];
    exps<NL, R>(x.ms, r.ms, c2, fxs);
    exps<NL, R>(y.ms, r.ms, c2, fys);
    exps<NL, R>(x.mt, r.mt, c2, fxt);
    exps<NL, R>(y.mt, r.mt, c2, fyt);
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        r.zs[k] = __fadd_rn(__fmul_rn(x.zs[k], fxs[k]), __fmul_rn(y.zs[k], fys[k]));
        r.zt[k] = __fadd_rn(__fmul_rn(x.zt[k], fxt[k]), __fmul_rn(y.zt[k], fyt[k]));
        r.a[k] = __fadd_rn(__fmul_rn(x.a[k], fxt[k]), __fmul_rn(y.a[k], fyt[k]));
    }
    return r;
}
template <int NL>
__device__ __forceinline__ void pstat_store(float* rec, const PStat<NL>& x) {
    float v[kCRecFloats] = {x.ms, x.mt, x.zs[0], x.zt[0], x.a[0], 0.f, 0.f, 0.f};
    if (NL == 2) {
        v[5] = x.zs[NL - 1];
        v[6] = x.zt[NL - 1];
        v[7] = x.a[NL - 1];
    }
    reinterpret_cast<float4*>(rec)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(rec)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <int NL>
__device__ __forceinline__ PStat<NL> pstat_from(const float4& r0, const float4& r1) {
    PStat<NL> x;
    x.ms = r0.x;
    x.mt = r0.y;
    x.zs[0] = r0.z;
    x.zt[0] = r0.w;
    x.a[0] = r1.x;
    if (NL == 2) {
        x.zs[NL - 1] = r1.y;
        x.zt[NL - 1] = r1.z;
        x.a[NL - 1] = r1.w;
    }
    return x;
}
__device__ __forceinline__ float kl_of_row(float inv_tau, float ms, float mt, float zs, float zt, float a) {
    // KL(p||q) = sum p (t - s)/tau - lse_t + lse_s
    return inv_tau * a / zt - ((mt - ms) * inv_tau + (logf(zt) - logf(zs)));
}

// geometry of one super-row, identical on every thread of the cluster
struct SuperRow {
    int b, grp;        // sample, index of the larger-group row inside it
    int lv;            // 4-element vectors of the super-row
    size_t base;       // element offset of its first element
};
__device__ __forceinline__ SuperRow super_row(const RowsParams& p, const ClusterGeom& g, int sr) {
    SuperRow x;
    x.b = sr / g.G_big;
    x.grp = sr - x.b * g.G_big;
    const int ch = min(g.g_big, p.C - x.grp * g.g_big);
    x.lv = ch * g.hwv;
    x.base = ((size_t)x.b * p.C + (size_t)x.grp * g.g_big) * p.HW;
    return x;
}

// a thread's unit of work is a 4-element vector whatever the dtype (16 bytes of fp32, 8 bytes of bf16): the
// TMEM footprint per element and the slice capacity in elements are then the same for both
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    using type = float4;
    static __device__ __forceinline__ void unpack(const type& v, float* f) {
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    static __device__ __forceinline__ void store(float* p, const float* f) {
        st_streaming(reinterpret_cast<float4*>(p), make_float4(f[0], f[1], f[2], f[3]));
    }
};
template <>
struct Vec4<__nv_bfloat16> {
    using type = uint2;
    static __device__ __forceinline__ void unpack(const type& v, float* f) {
        Elem<__nv_bfloat16>::unpack2(v.x, f[0], f[1]);
        Elem<__nv_bfloat16>::unpack2(v.y, f[2], f[3]);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
        const uint32_t a = Elem<__nv_bfloat16>::pack2(f[0], f[1]), b = Elem<__nv_bfloat16>::pack2(f[2], f[3]);
        asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
    }
};

// R: 0 = independent exponentials per loss, 2 = l[1].tau == 2 * l[0].tau.
// With one loss, or with R == 2, phase 1 parks the EXPONENTIALS (relative to the thread's running maximum at
// that moment, parked next to them): phase 2 is then a multiply-add per element, no ex2.  Otherwise the raw
// values are parked and phase 2 recomputes.
template <typename T, int NL, int R>
__global__ void __launch_bounds__(kCThreads, 1) kl_rows_cluster_kernel(const RowsParams p, const ClusterGeom g) {
    using V = Vec4<T>;
    using vec_t = typename V::type;
    constexpr int VE = 4;
    constexpr int NE = kCChunkRows * VE;        // elements per thread, chunk and tensor
    constexpr bool kParkExp = NL == 1 || R == 2;
    constexpr int K = NL - 1;                   // the loss whose exponential comes out of the MUFU

    extern __shared__ __align__(128) unsigned char smem_raw[];
    ClusterSmem& sm = *reinterpret_cast<ClusterSmem*>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int NC = g.nc;
    const int slv = g.slv;
    const int cluster_id = blockIdx.x / NC;
    const int n_clusters = gridDim.x / NC;

    if (p.run_if != nullptr && *p.run_if == 0u) return;  // cancelled backward re-run (uniform over the grid)

    if (tid == 0) {
        for (int c = 0; c < kCRing; ++c) {
            mbar_init(&sm.full[c], 1);
            mbar_init(&sm.empty[c], kCConsWarps);
        }
        mbar_init(&sm.xch[0], 1);
        mbar_init(&sm.xch[1], 1);
        fence_barrier_init();
    }
    if (warp == kCConsWarps) tmem_alloc(&sm.tmem_base, kCTmemCols);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;
    // the only cluster-wide barrier: every CTA's mbarriers exist before a peer pushes bytes at them
    cluster_arrive_release();
    cluster_wait_acquire();

    const int n_iter = cluster_id < g.total_sr ? (g.total_sr - cluster_id + n_clusters - 1) / n_clusters : 0;

    if (warp == kCConsWarps) {
        // =====================================================================================
        // TMA warp: lane 0 streams this CTA's slices, chunk by chunk, into the ring as slots drain - up to
        // a whole super-row ahead of the consumers
        // =====================================================================================
        const uint64_t pol = l2_policy_evict_first();
        int slot = 0;
        uint32_t phase = 0;
        if (lane == 0) {
            for (int it = 0; it < n_iter; ++it) {
                const SuperRow x = super_row(p, g, cluster_id + it * n_clusters);
                const int v0 = min(x.lv, (int)rank * slv);
                const int v1 = min(x.lv, v0 + slv);
                for (int c = 0; c * kCChunkVecs < v1 - v0; ++c) {
                    mbar_wait(&sm.empty[slot], phase ^ 1u);
                    const int nv = min(kCChunkVecs, v1 - v0 - c * kCChunkVecs);
                    const uint32_t bytes = (uint32_t)nv * (uint32_t)sizeof(vec_t);
                    mbar_arrive_expect_tx(&sm.full[slot], 2u * bytes);
                    const size_t off = (x.base + (size_t)(v0 + c * kCChunkVecs) * VE) * sizeof(T);
                    tma_bulk_g2s(sm.ring[slot][0], static_cast<const char*>(p.S) + off, bytes, &sm.full[slot], pol);
                    tma_bulk_g2s(sm.ring[slot][1], static_cast<const char*>(p.T) + off, bytes, &sm.full[slot], pol);
                    if (++slot == kCRing) {
                        slot = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
        // TMEM is released after every consumer of this CTA is through with it
        bar_sync(3, kCThreads);
        tmem_dealloc(tmem_base, kCTmemCols);
        return;
    }

    // =========================================================================================
    // consumer warps
    // =========================================================================================
    float c2[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) c2[k] = p.l[k].c2;
    const int rv0 = g.rv0;             // vectors of a complete row of l[0]
    // my TMEM window: lane quarter of the warp, 128 columns per warp of that quarter; a vector-row of a
    // chunk takes kCRowCols columns: 4 + 4 parked values of S and T, and the two references
    const uint32_t tmem_mine = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 128);
    int slot = 0;
    uint32_t phase = 0;
    float kl_acc[NL];                  // lane 0 of the warps that finish rows
#pragma unroll
    for (int k = 0; k < NL; ++k) kl_acc[k] = 0.f;

    for (int it = 0; it < n_iter; ++it) {
        const int sr = cluster_id + it * n_clusters;
        const SuperRow x = super_row(p, g, sr);
        const int par = it & 1;
        const int v0 = min(x.lv, (int)rank * slv);     // my slice, in vectors of the super-row
        const int v1 = min(x.lv, v0 + slv);
        const int nvs = v1 - v0;
        const int r_first = v0 / rv0;                  // first row of l[0] (within the super-row) in my slice
        const int n_pieces = nvs > 0 ? (v1 - 1) / rv0 - r_first + 1 : 0;
        if (tid == kCCons - 1) {
            // this iteration's summaries: 32 bytes per piece of every CTA of the cluster will land in summ[par]
            int total = 0;
            for (int c = 0; c < NC; ++c) {
                const int cv0 = min(x.lv, c * slv), cv1 = min(x.lv, cv0 + slv);
                total += cv1 > cv0 ? (cv1 - 1) / rv0 - cv0 / rv0 + 1 : 0;
            }
            mbar_arrive_expect_tx(&sm.xch[par], (uint32_t)total * 32u);
        }

        // ------------------------------------------------ phase 1: piece statistics, park the slice
        // chunk-major: a chunk is pulled from the ring once; the piece it belongs to (rarely: the two or three
        // pieces it straddles) gets its running statistics updated; every vector-row is parked with the
        // reference of ITS piece
        {
            const int nchunks = (nvs + kCChunkVecs - 1) / kCChunkVecs;
            int pc = 0;                                                  // current piece
            int pv1 = min(v1, (r_first + 1) * rv0) - v0;                // ... ends here (vectors of my slice)
            PStat<NL> st = pstat_empty<NL>();
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&sm.full[slot], phase);
                const vec_t* bs = reinterpret_cast<const vec_t*>(sm.ring[slot][0]);
                const vec_t* bt = reinterpret_cast<const vec_t*>(sm.ring[slot][1]);
                float fs[NE], ft[NE];
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    V::unpack(bs[r * kCCons + tid], &fs[r * VE]);
                    V::unpack(bt[r * kCCons + tid], &ft[r * VE]);
                }
                const int cbeg = c * kCChunkVecs, cend = cbeg + kCChunkVecs;
                float ps[NE], pt[NE], pref[kCChunkRows][2];            // what gets parked
                if (cend <= pv1) {
                    // ---- the whole chunk lies in the current piece: no masks
                    float nms = st.ms, nmt = st.mt;
#pragma unroll
                    for (int i = 0; i < NE; ++i) {
                        nms = fmaxf(nms, fs[i]);
                        nmt = fmaxf(nmt, ft[i]);
                    }
                    // the maxima consumed every shared-memory read: the slot may go back to the TMA warp (an
                    // mbarrier arrival is not ordered behind LDS that are still in flight)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);
                    float refs2[NL], reft2[NL], rs[NL], rt[NL];
#pragma unroll
                    for (int k = 0; k < NL; ++k) {
                        refs2[k] = nms * c2[k];
                        reft2[k] = nmt * c2[k];
                    }
                    // exact differences for the rescale: both references may still be the -1e29 floor
                    exps<NL, R>(st.ms, nms, c2, rs);
                    exps<NL, R>(st.mt, nmt, c2, rt);
#pragma unroll
                    for (int k = 0; k < NL; ++k) {
                        st.zs[k] *= rs[k];
                        st.zt[k] *= rt[k];
                        st.a[k] *= rt[k];
                    }
                    st.ms = nms;
                    st.mt = nmt;
#pragma unroll
                    for (int i = 0; i < NE; ++i) {
                        const float d = ft[i] - fs[i];
                        float es[NL], et[NL];
                        exps<NL, R>(fs[i], refs2, c2, es);
                        exps<NL, R>(ft[i], reft2, c2, et);
#pragma unroll
                        for (int k = 0; k < NL; ++k) {
                            st.zs[k] += es[k];
                            st.zt[k] += et[k];
                            st.a[k] = fmaf(et[k], d, st.a[k]);
                        }
                        ps[i] = kParkExp ? es[K] : fs[i];
                        pt[i] = kParkExp ? et[K] : ft[i];
                    }
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) {
                        pref[r][0] = nms;
                        pref[r][1] = nmt;
                    }
                } else {
                    // ---- the chunk runs past the end of the piece (or of the slice): piece by piece, masked
                    float raw_s[NE], raw_t[NE];
#pragma unroll
                    for (int i = 0; i < NE; ++i) {
                        raw_s[i] = fs[i];
                        raw_t[i] = ft[i];
                        ps[i] = pt[i] = 0.f;
                    }
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) pref[r][0] = pref[r][1] = 0.f;
                    for (;;) {
                        // the part of piece pc inside this chunk: [lo, hi)
                        const int lo = max(cbeg, max(v0, (r_first + pc) * rv0) - v0), hi = min(min(cend, nvs), pv1);
                        bool mine[kCChunkRows];
#pragma unroll
                        for (int r = 0; r < kCChunkRows; ++r) {
                            const int v = cbeg + r * kCCons + tid;
                            mine[r] = v >= lo && v < hi;
#pragma unroll
                            for (int q = 0; q < VE; ++q) {
                                fs[r * VE + q] = mine[r] ? raw_s[r * VE + q] : kPadValue;
                                ft[r * VE + q] = mine[r] ? raw_t[r * VE + q] : kPadValue;
                            }
                        }
                        float nms = st.ms, nmt = st.mt;
#pragma unroll
                        for (int i = 0; i < NE; ++i) {
                            nms = fmaxf(nms, fs[i]);
                            nmt = fmaxf(nmt, ft[i]);
                        }
                        float refs2[NL], reft2[NL], rs[NL], rt[NL];
#pragma unroll
                        for (int k = 0; k < NL; ++k) {
                            refs2[k] = nms * c2[k];
                            reft2[k] = nmt * c2[k];
                        }
                        exps<NL, R>(st.ms, nms, c2, rs);
                        exps<NL, R>(st.mt, nmt, c2, rt);
#pragma unroll
                        for (int k = 0; k < NL; ++k) {
                            st.zs[k] *= rs[k];
                            st.zt[k] *= rt[k];
                            st.a[k] *= rt[k];
                        }
                        st.ms = nms;
                        st.mt = nmt;
#pragma unroll
                        for (int i = 0; i < NE; ++i) {
                            const float d = ft[i] - fs[i];
                            float es[NL], et[NL];
                            exps<NL, R>(fs[i], refs2, c2, es);
                            exps<NL, R>(ft[i], reft2, c2, et);
#pragma unroll
                            for (int k = 0; k < NL; ++k) {
                                st.zs[k] += es[k];
                                st.zt[k] += et[k];
                                st.a[k] = fmaf(et[k], d, st.a[k]);
                            }
                            if (mine[i / VE]) {
                                ps[i] = kParkExp ? es[K] : raw_s[i];
                                pt[i] = kParkExp ? et[K] : raw_t[i];
                            }
                        }
#pragma unroll
                        for (int r = 0; r < kCChunkRows; ++r) {
                            if (mine[r]) {
                                pref[r][0] = nms;
                                pref[r][1] = nmt;
                            }
                        }
                        if (pv1 > min(cend, nvs)) break;            // the piece continues in the next chunk
                        // the piece ends in this chunk: its warp record; on to the next piece, if any
                        st = pstat_reduce<NL, R>(st, c2, 32);
                        if (lane == 0) pstat_store<NL>(sm.rec[par][pc][warp], st);
                        st = pstat_empty<NL>();
                        ++pc;
                        if (pc >= n_pieces) break;
                        pv1 = min(v1, (r_first + pc + 1) * rv0) - v0;
                        if (max(v0, (r_first + pc) * rv0) - v0 >= min(cend, nvs)) break;   // it starts in the next chunk
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);   // (the arithmetic above consumed the reads)
                }
                // ---- park: registers -> TMEM
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    const uint32_t ta = tmem_mine + (uint32_t)((c * kCChunkRows + r) * kCRowCols);
                    tmem_st8(ta,
                             make_uint4(__float_as_uint(ps[r * VE]), __float_as_uint(ps[r * VE + 1]),
                                        __float_as_uint(ps[r * VE + 2]), __float_as_uint(ps[r * VE + 3])),
                             make_uint4(__float_as_uint(pt[r * VE]), __float_as_uint(pt[r * VE + 1]),
                                        __float_as_uint(pt[r * VE + 2]), __float_as_uint(pt[r * VE + 3])));
                    if (kParkExp) tmem_st2(ta + 8, __float_as_uint(pref[r][0]), __float_as_uint(pref[r][1]));
                }
                if (++slot == kCRing) {
                    slot = 0;
                    phase ^= 1u;
                }
                // a piece that ends exactly with this chunk (the common case) closes here
                if (pc < n_pieces && cend == pv1) {
                    st = pstat_reduce<NL, R>(st, c2, 32);
                    if (lane == 0) pstat_store<NL>(sm.rec[par][pc][warp], st);
                    st = pstat_empty<NL>();
                    ++pc;
                    pv1 = min(v1, (r_first + pc + 1) * rv0) - v0;
                }
            }
        }

        // ------------------------------------------------ 16 warp records -> one summary per piece
        bar_sync(1, kCCons);
        if (warp < n_pieces) {
            const float4* q = reinterpret_cast<const float4*>(sm.rec[par][warp][lane & 15]);
            const PStat<NL> st = pstat_reduce<NL, R>(pstat_from<NL>(q[0], q[1]), c2, 16);
            // ---- exchange: lane c pushes the summary into CTA c (this CTA included)
            if (lane < NC) {
                const uint32_t dst = map_to_cta(sm.summ[par][rank][warp], (uint32_t)lane);
                const uint32_t bar = map_to_cta(&sm.xch[par], (uint32_t)lane);
                st_async_f4(dst, make_float4(st.ms, st.mt, st.zs[0], st.zt[0]), bar);
                st_async_f4(dst + 16, make_float4(st.a[0], st.zs[NL - 1], st.zt[NL - 1], st.a[NL - 1]), bar);
            }
        }

        // one warp per row: warp 0 the super-row (two losses), warp 1 + pc the row of l[0] of piece pc
        if (NL == 2 && warp == 0) {
            mbar_wait(&sm.xch[par], (uint32_t)(it >> 1) & 1u);
            PStat<NL> acc = pstat_empty<NL>();
            for (int i = lane; i < NC * kCMaxPieces; i += 32) {
                const int c = i / kCMaxPieces, pc = i - c * kCMaxPieces;
                const int cv0 = min(x.lv, c * slv), cv1 = min(x.lv, cv0 + slv);
                const int np = cv1 > cv0 ? (cv1 - 1) / rv0 - cv0 / rv0 + 1 : 0;
                if (pc < np) {
                    const float4* q = reinterpret_cast<const float4*>(sm.summ[par][c][pc]);
                    acc = pstat_merge<NL, R>(acc, pstat_from<NL>(q[0], q[1]), c2);
                }
            }
            acc = pstat_reduce<NL, R>(acc, c2, 32);
            if (lane == 0) {
                float coef = p.l[K].coef;
                if (p.grad_out[K] != nullptr) coef *= __ldg(p.grad_out[K]);
                *reinterpret_cast<float4*>(sm.fin[kCMaxPieces]) =
                    make_float4(acc.ms, acc.mt, coef / acc.zs[K], coef / acc.zt[K]);
                if (rank == 0) {
                    const float kl = kl_of_row(p.l[K].inv_tau, acc.ms, acc.mt, acc.zs[K], acc.zt[K], acc.a[K]);
                    if (p.l[K].row_kl) p.l[K].row_kl[x.b * p.l[K].G + x.grp] = kl;
                    kl_acc[K] += kl;
                }
            }
        } else if (warp >= 1 && warp <= n_pieces) {
            const int pc = warp - 1;
            const int row = r_first + pc;                       // row of l[0] within the super-row
            const int rv_lo = row * rv0, rv_hi = min(x.lv, rv_lo + rv0);
            const int ca = rv_lo / slv, cb = (rv_hi - 1) / slv;   // its pieces live in CTAs ca..cb, one each
            mbar_wait(&sm.xch[par], (uint32_t)(it >> 1) & 1u);
            PStat<NL> acc = pstat_empty<NL>();
            if (lane <= cb - ca) {
                const int c = ca + lane;
                const int cpc = row - min(x.lv, c * slv) / rv0;
                const float4* q = reinterpret_cast<const float4*>(sm.summ[par][c][cpc]);
                acc = pstat_from<NL>(q[0], q[1]);
            }
            acc = pstat_reduce<NL, R>(acc, c2, 32);
            if (lane == 0) {
                float coef = p.l[0].coef;
                if (p.grad_out[0] != nullptr) coef *= __ldg(p.grad_out[0]);
                *reinterpret_cast<float4*>(sm.fin[pc]) = make_float4(acc.ms, acc.mt, coef / acc.zs[0], coef / acc.zt[0]);
                if ((int)rank == ca) {
                    const float kl = kl_of_row(p.l[0].inv_tau, acc.ms, acc.mt, acc.zs[0], acc.zt[0], acc.a[0]);
                    const int rowi = NL == 2 ? x.b * p.l[0].G + x.grp * p.l[NL - 1].m + row : x.b * p.l[0].G + x.grp;
                    if (p.l[0].row_kl) p.l[0].row_kl[rowi] = kl;
                    kl_acc[0] += kl;
                }
            }
        }
        // the thread that armed the exchange always sees it complete: a CTA whose slice is empty (ragged last
        // super-row) must neither re-arm the barrier early nor exit while peers still push at it
        if (tid == kCCons - 1) mbar_wait(&sm.xch[par], (uint32_t)(it >> 1) & 1u);
        bar_sync(2, kCCons);

        // ------------------------------------------------ phase 2: gradient from the parked slice
        tmem_wait_st();
        {
            T* out = static_cast<T*>(p.dS) + x.base + (size_t)v0 * VE;
            const int nchunks = (nvs + kCChunkVecs - 1) / kCChunkVecs;
            const float4 fb = *reinterpret_cast<const float4*>(sm.fin[kCMaxPieces]);   // {Ms, Mt, coef/Zs, coef/Zt} of the super-row
            // one vector-row: parked values + references -> gradient, given the statistics `fa` of its l[0] row
            auto grad_row = [&](const uint4& bs, const uint4& bt, float ref_s, float ref_t, const float4& fa, int v) {
                const float fs[VE] = {__uint_as_float(bs.x), __uint_as_float(bs.y), __uint_as_float(bs.z), __uint_as_float(bs.w)};
                const float ft[VE] = {__uint_as_float(bt.x), __uint_as_float(bt.y), __uint_as_float(bt.z), __uint_as_float(bt.w)};
                float o[VE];
                if (kParkExp) {
                    // parked: e = exp2((x - ref) c2[K]) against the thread's reference of that moment;
                    // softmax_k = e^(c2[k]/c2[K]) * exp2((ref - M_k) c2[k]) / Z_k, and ref <= M_k
                    const float gsK = (NL == 2 ? fb.z : fa.z) * fast_exp2((ref_s - (NL == 2 ? fb.x : fa.x)) * c2[K]);
                    const float gtK = (NL == 2 ? fb.w : fa.w) * fast_exp2((ref_t - (NL == 2 ? fb.y : fa.y)) * c2[K]);
                    if (NL == 2) {
                        const float gs0 = fa.z * fast_exp2((ref_s - fa.x) * c2[0]);
                        const float gt0 = fa.w * fast_exp2((ref_t - fa.y) * c2[0]);
#pragma unroll
                        for (int q = 0; q < VE; ++q) o[q] = fs[q] * fmaf(fs[q], gs0, gsK) - ft[q] * fmaf(ft[q], gt0, gtK);
                    } else {
#pragma unroll
                        for (int q = 0; q < VE; ++q) o[q] = fmaf(fs[q], gsK, -ft[q] * gtK);
                    }
                } else {
                    // raw values parked: recompute against the row maxima
                    const float rs0 = fa.x * c2[0], rt0 = fa.y * c2[0];
                    const float rs1 = fb.x * c2[K], rt1 = fb.y * c2[K];
#pragma unroll
                    for (int q = 0; q < VE; ++q) {
                        const float es0 = fast_exp2(fmaf(fs[q], c2[0], -rs0));
                        const float et0 = fast_exp2(fmaf(ft[q], c2[0], -rt0));
                        const float es1 = fast_exp2(fmaf(fs[q], c2[K], -rs1));
                        const float et1 = fast_exp2(fmaf(ft[q], c2[K], -rt1));
                        o[q] = fmaf(es0, fa.z, es1 * fb.z) - fmaf(et0, fa.w, et1 * fb.w);
                    }
                }
                V::store(out + (size_t)v * VE, o);
            };
            // parked rows in flight from TMEM, one chunk ahead of the arithmetic
            uint4 ls[kCChunkRows], lt[kCChunkRows];
            uint32_t lr[kCChunkRows][2];
            auto fetch = [&](int c) {
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    const uint32_t ta = tmem_mine + (uint32_t)((c * kCChunkRows + r) * kCRowCols);
                    tmem_ld8(ta, ls[r], lt[r]);
                    if (kParkExp) tmem_ld2(ta + 8, lr[r][0], lr[r][1]);
                }
            };
            int pc = 0;
            int pv1 = min(v1, (r_first + 1) * rv0) - v0;
            float4 fa = *reinterpret_cast<const float4*>(sm.fin[0]);
            if (nchunks > 0) fetch(0);
            for (int c = 0; c < nchunks; ++c) {
                static_assert(kCChunkRows == 2, "the waits below name the registers of two rows");
                tmem_wait_ld(ls[0], lt[0], lr[0][0], lr[0][1]);
                tmem_wait_ld(ls[1], lt[1], lr[1][0], lr[1][1]);
                uint4 bs[kCChunkRows], bt[kCChunkRows];
                float ref[kCChunkRows][2];
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    bs[r] = ls[r];
                    bt[r] = lt[r];
                    ref[r][0] = __uint_as_float(lr[r][0]);
                    ref[r][1] = __uint_as_float(lr[r][1]);
                }
                if (c + 1 < nchunks) fetch(c + 1);     // the next chunk comes in underneath the arithmetic of this one
                const int cbeg = c * kCChunkVecs, cend = cbeg + kCChunkVecs;
                if (cend <= pv1) {
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) grad_row(bs[r], bt[r], ref[r][0], ref[r][1], fa, cbeg + r * kCCons + tid);
                } else {
                    // the chunk straddles pieces or the end of the slice: look the row's piece up
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) {
                        const int v = cbeg + r * kCCons + tid;
                        if (v < nvs) {
                            const int q = (v0 + v) / rv0 - r_first;
                            grad_row(bs[r], bt[r], ref[r][0], ref[r][1], *reinterpret_cast<const float4*>(sm.fin[q]), v);
                        }
                    }
                }
                while (pc < n_pieces && pv1 <= min(cend, nvs)) {    // pieces that ended with this chunk
                    ++pc;
                    pv1 = min(v1, (r_first + pc + 1) * rv0) - v0;
                    if (pc < n_pieces) fa = *reinterpret_cast<const float4*>(sm.fin[pc]);
                }
            }
        }
    }

    // ================================ loss: warp partials -> CTA partial -> the last CTA sums in a fixed order
    if (lane == 0) {
        sm.klpart[warp][0] = kl_acc[0];
        sm.klpart[warp][1] = NL == 2 ? kl_acc[NL - 1] : 0.f;
    }
    // every consumer is through with TMEM (the TMA warp frees it); nobody pushes at this CTA any more: its
    // last exchange completed before its last gradient pass
    bar_sync(3, kCThreads);
    if (warp == 0) {
        unsigned ticket = 0;
        if (lane == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int w = 0; w < kCConsWarps; ++w) {
                s0 += sm.klpart[w][0];
                s1 += sm.klpart[w][1];
            }
            __stcg(&p.cta_part[blockIdx.x], s0);
            if (NL == 2) __stcg(&p.cta_part[kMaxGrid + blockIdx.x], s1);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double acc[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) acc[k] = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
#pragma unroll
                for (int k = 0; k < NL; ++k) acc[k] += (double)__ldcg(&p.cta_part[k * kMaxGrid + i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < NL; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NL; ++k) *p.l[k].loss = (float)((double)p.l[k].loss_scale * acc[k]);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

// ====================================================================================================
template <typename T, int NL, int R>
static cudaError_t launch_cluster_t(const RowsParams& p, ClusterGeom g, int sms, cudaStream_t stream, bool probe_only) {
    auto kern = kl_rows_cluster_kernel<T, NL, R>;
    static bool configured = false;      // per instantiation
    static int max_clusters[kClusterMaxSize + 1] = {0};
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kCThreads);
    cfg.dynamicSmemBytes = kClusterSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)g.nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters[g.nc] == 0) {
        cfg.gridDim = dim3((unsigned)(sms / g.nc * g.nc));
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) {
            cudaGetLastError();
            n = -1;
        }
        max_clusters[g.nc] = n > 0 ? n : -1;
    }
    if (max_clusters[g.nc] < 1) return cudaErrorLaunchOutOfResources;
    if (probe_only) return cudaSuccess;
    int n_clusters = max_clusters[g.nc];
    if (n_clusters > g.total_sr) n_clusters = g.total_sr;
    if (n_clusters * g.nc > kMaxGrid) n_clusters = kMaxGrid / g.nc;
    cfg.gridDim = dim3((unsigned)(n_clusters * g.nc));
    return cudaLaunchKernelEx(&cfg, kern, p, g);
}

cudaError_t launch_kl_rows_cluster(const RowsParams& p, const ClusterGeom& g, bool bf16, int sms, cudaStream_t stream,
                                   bool probe_only) {
    if (p.nl == 2) {
        // tau[1] == 2 * tau[0]: one exponential serves both losses
        const bool sq = p.l[0].c2 == 2.f * p.l[1].c2;
        if (sq)
            return bf16 ? launch_cluster_t<__nv_bfloat16, 2, 2>(p, g, sms, stream, probe_only)
                        : launch_cluster_t<float, 2, 2>(p, g, sms, stream, probe_only);
        return bf16 ? launch_cluster_t<__nv_bfloat16, 2, 0>(p, g, sms, stream, probe_only)
                    : launch_cluster_t<float, 2, 0>(p, g, sms, stream, probe_only);
    }
    return bf16 ? launch_cluster_t<__nv_bfloat16, 1, 0>(p, g, sms, stream, probe_only)
                : launch_cluster_t<float, 1, 0>(p, g, sms, stream, probe_only);
}

}  // namespace sd
