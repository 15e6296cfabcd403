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
//   exchange  ONE cluster barrier (release/acquire); one warp per row merges the summaries it needs
//             straight out of the other CTAs' shared memory (DSMEM, ld.shared::cluster).
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
constexpr int kCChunkRows = 2;                         // 16-byte vectors per consumer thread, chunk and tensor
constexpr int kCChunkVecs = kCChunkRows * kCCons;      // 1024 vectors = 16 KB per tensor
constexpr int kCChunkBytes = kCChunkVecs * 16;
constexpr int kCRing = 6;                              // ring slots of 32 KB (S chunk + T chunk)
constexpr int kCMaxChunks = kClusterMaxChunks;         // chunks per slice: 8 x 16 TMEM columns per thread
constexpr int kCMaxPieces = kClusterMaxPieces;
constexpr int kCRecFloats = 8;                         // ms, mt, {zs, zt, a} x 2
constexpr int kCTmemCols = 512;
static_assert(kCMaxChunks * kCChunkRows * 8 * (kCConsWarps / 4) <= kCTmemCols, "TMEM columns");
static_assert(kCChunkVecs == kClusterChunkVecs, "chunk size");

struct ClusterSmem {
    unsigned char ring[kCRing][2][kCChunkBytes];
    uint64_t full[kCRing], empty[kCRing];
    float rec[2][kCMaxPieces][kCConsWarps][kCRecFloats];  // warp records of the pieces, by iteration parity
    float summ[2][kCMaxPieces][kCRecFloats];              // CTA summaries, read by the peers through DSMEM
    float fin[kCMaxPieces + 1][4];                        // row statistics for phase 2: {Ms, Mt, coef/Zs, coef/Zt}
    float klpart[kCConsWarps][2];
    uint32_t tmem_base;
};
constexpr size_t kClusterSmemBytes = sizeof(ClusterSmem) + 128;

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
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
    int lv;            // 16-byte vectors of the super-row
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

// R: 0 = independent exponentials per loss, 2 = l[1].tau == 2 * l[0].tau
template <typename T, int NL, int R>
__global__ void __launch_bounds__(kCThreads, 1) kl_rows_cluster_kernel(const RowsParams p, const ClusterGeom g) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    constexpr int NE = kCChunkRows * VE;   // elements per thread, chunk and tensor

    extern __shared__ unsigned char smem_raw[];
    ClusterSmem& sm = *reinterpret_cast<ClusterSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);

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
        fence_barrier_init();
    }
    if (warp == kCConsWarps) tmem_alloc(&sm.tmem_base, kCTmemCols);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;

    const int n_iter = cluster_id < g.total_sr ? (g.total_sr - cluster_id + n_clusters - 1) / n_clusters : 0;

    if (warp == kCConsWarps) {
        // =====================================================================================
        // TMA warp: lane 0 streams this CTA's slices, chunk by chunk, into the ring as slots drain - one
        // super-row ahead of the consumers.  It joins every cluster barrier (split arrive / wait).
        // =====================================================================================
        const uint64_t pol = l2_policy_evict_first();
        int slot = 0;
        uint32_t phase = 0;
        auto load_slice = [&](int it) {
            const SuperRow x = super_row(p, g, cluster_id + it * n_clusters);
            const int v0 = min(x.lv, (int)rank * slv);
            const int v1 = min(x.lv, v0 + slv);
            for (int c = 0; c * kCChunkVecs < v1 - v0; ++c) {
                mbar_wait(&sm.empty[slot], phase ^ 1u);
                const int nv = min(kCChunkVecs, v1 - v0 - c * kCChunkVecs);
                const uint32_t bytes = (uint32_t)nv * 16u;
                mbar_arrive_expect_tx(&sm.full[slot], 2u * bytes);
                const size_t off = (x.base + (size_t)(v0 + c * kCChunkVecs) * VE) * sizeof(T);
                tma_bulk_g2s(sm.ring[slot][0], static_cast<const char*>(p.S) + off, bytes, &sm.full[slot], pol);
                tma_bulk_g2s(sm.ring[slot][1], static_cast<const char*>(p.T) + off, bytes, &sm.full[slot], pol);
                if (++slot == kCRing) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
        };
        if (lane == 0 && n_iter > 0) load_slice(0);
        for (int it = 0; it < n_iter; ++it) {
            __syncwarp();
            cluster_arrive_release();
            if (lane == 0 && it + 1 < n_iter) load_slice(it + 1);
            __syncwarp();
            cluster_wait_acquire();
        }
        // nobody may leave while a peer could still read its summaries; TMEM is released after every
        // consumer of this CTA is through with it
        cluster_arrive_release();
        cluster_wait_acquire();
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
    // my TMEM window: lane quarter of the warp, 128 columns per warp of that quarter
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

        // ------------------------------------------------ phase 1: park the slice, piece statistics
        int cur_c = -1;                                // last chunk pulled from the ring
        vec_t vs[kCChunkRows], vt[kCChunkRows];        // its vectors of this thread
        for (int pc = 0; pc < n_pieces; ++pc) {
            const int pv0 = max(v0, (r_first + pc) * rv0) - v0;       // piece, in vectors of my slice
            const int pv1 = min(v1, (r_first + pc + 1) * rv0) - v0;
            const int c_lo = pv0 / kCChunkVecs, c_hi = (pv1 - 1) / kCChunkVecs;
            PStat<NL> st = pstat_empty<NL>();
            for (int c = c_lo; c <= c_hi; ++c) {
                if (c != cur_c) {
                    // ring -> registers -> TMEM; the slot goes back to the TMA warp right away
                    mbar_wait(&sm.full[slot], phase);
                    const vec_t* bs = reinterpret_cast<const vec_t*>(sm.ring[slot][0]);
                    const vec_t* bt = reinterpret_cast<const vec_t*>(sm.ring[slot][1]);
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) {
                        vs[r] = bs[r * kCCons + tid];
                        vt[r] = bt[r * kCCons + tid];
                    }
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r)
                        tmem_st8(tmem_mine + (uint32_t)((c * kCChunkRows + r) * 8), as_bits(vs[r]), as_bits(vt[r]));
                    // the arrival must not overtake the shared-memory reads: the TMEM stores consumed them
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);
                    if (++slot == kCRing) {
                        slot = 0;
                        phase ^= 1u;
                    }
                    cur_c = c;
                }
                float fs[NE], ft[NE];
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    E::unpack(vs[r], &fs[r * VE]);
                    E::unpack(vt[r], &ft[r * VE]);
                }
                if (c * kCChunkVecs < pv0 || (c + 1) * kCChunkVecs > pv1) {
                    // the chunk sticks out of the piece (or of the slice): blank what is not ours
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) {
                        const int v = c * kCChunkVecs + r * kCCons + tid;
                        if (v < pv0 || v >= pv1) {
#pragma unroll
                            for (int q = 0; q < VE; ++q) {
                                fs[r * VE + q] = kPadValue;
                                ft[r * VE + q] = kPadValue;
                            }
                        }
                    }
                }
                // online update: new running maxima, old sums rescaled to them
                float nms = st.ms, nmt = st.mt;
#pragma unroll
                for (int i = 0; i < NE; ++i) {
                    nms = fmaxf(nms, fs[i]);
                    nmt = fmaxf(nmt, ft[i]);
                }
                float rs[NL], rt[NL];
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
                    exps<NL, R>(fs[i], nms, c2, es);
                    exps<NL, R>(ft[i], nmt, c2, et);
#pragma unroll
                    for (int k = 0; k < NL; ++k) {
                        st.zs[k] += es[k];
                        st.zt[k] += et[k];
                        st.a[k] = fmaf(et[k], d, st.a[k]);
                    }
                }
            }
            st = pstat_reduce<NL, R>(st, c2, 32);
            if (lane == 0) pstat_store<NL>(sm.rec[par][pc][warp], st);
        }

        // ------------------------------------------------ 16 warp records -> one summary per piece
        bar_sync(1, kCCons);
        if (warp < n_pieces) {
            const float4* q = reinterpret_cast<const float4*>(sm.rec[par][warp][lane & 15]);
            PStat<NL> st = pstat_reduce<NL, R>(pstat_from<NL>(q[0], q[1]), c2, 16);
            if (lane == 0) pstat_store<NL>(sm.summ[par][warp], st);
        }

        // ------------------------------------------------ exchange
        cluster_arrive_release();
        cluster_wait_acquire();

        // one warp per row: warp 0 the super-row (two losses), warp 1 + pc the row of l[0] of piece pc
        if (NL == 2 && warp == 0) {
            constexpr int K = NL - 1;
            PStat<NL> acc = pstat_empty<NL>();
            for (int i = lane; i < NC * kCMaxPieces; i += 32) {
                const int c = i / kCMaxPieces, pc = i - c * kCMaxPieces;
                const int cv0 = min(x.lv, c * slv), cv1 = min(x.lv, cv0 + slv);
                const int np = cv1 > cv0 ? (cv1 - 1) / rv0 - cv0 / rv0 + 1 : 0;
                if (pc < np) {
                    const uint32_t ra = map_to_cta(sm.summ[par][pc], (uint32_t)c);
                    acc = pstat_merge<NL, R>(acc, pstat_from<NL>(ld_dsmem_f4(ra), ld_dsmem_f4(ra + 16)), c2);
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
            PStat<NL> acc = pstat_empty<NL>();
            if (lane <= cb - ca) {
                const int c = ca + lane;
                const int cpc = row - min(x.lv, c * slv) / rv0;
                const uint32_t ra = map_to_cta(sm.summ[par][cpc], (uint32_t)c);
                acc = pstat_from<NL>(ld_dsmem_f4(ra), ld_dsmem_f4(ra + 16));
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
        bar_sync(2, kCCons);

        // ------------------------------------------------ phase 2: gradient from the parked slice
        tmem_wait_st();
        T* out = static_cast<T*>(p.dS) + x.base + (size_t)v0 * VE;
        for (int pc = 0; pc < n_pieces; ++pc) {
            const int pv0 = max(v0, (r_first + pc) * rv0) - v0;
            const int pv1 = min(v1, (r_first + pc + 1) * rv0) - v0;
            const int c_lo = pv0 / kCChunkVecs, c_hi = (pv1 - 1) / kCChunkVecs;
            // exponentials are taken against the maxima of the l[0] row (elements never exceed them); the
            // larger-group softmax absorbs the difference of the maxima in its coefficient
            const float4 fa = *reinterpret_cast<const float4*>(sm.fin[pc]);
            float refs[NL], reft[NL], ks[NL], kt[NL];
            refs[0] = fa.x;
            reft[0] = fa.y;
            ks[0] = fa.z;
            kt[0] = -fa.w;
            if (NL == 2) {
                constexpr int K = NL - 1;
                const float4 fb = *reinterpret_cast<const float4*>(sm.fin[kCMaxPieces]);
                if (R == 2) {
                    refs[K] = fa.x;
                    reft[K] = fa.y;
                    ks[K] = fb.z * fast_exp2((fa.x - fb.x) * c2[K]);
                    kt[K] = -fb.w * fast_exp2((fa.y - fb.y) * c2[K]);
                } else {
                    refs[K] = fb.x;
                    reft[K] = fb.y;
                    ks[K] = fb.z;
                    kt[K] = -fb.w;
                }
            }
            for (int c = c_lo; c <= c_hi; ++c) {
                uint4 bs[kCChunkRows], bt[kCChunkRows];
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r)
                    tmem_ld8(tmem_mine + (uint32_t)((c * kCChunkRows + r) * 8), bs[r], bt[r]);
                tmem_wait_ld();
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    const int v = c * kCChunkVecs + r * kCCons + tid;
                    if (v >= pv0 && v < pv1) {
                        float fs[VE], ft[VE], o[VE];
                        E::unpack(*reinterpret_cast<const vec_t*>(&bs[r]), fs);
                        E::unpack(*reinterpret_cast<const vec_t*>(&bt[r]), ft);
#pragma unroll
                        for (int q = 0; q < VE; ++q) {
                            float es[NL], et[NL];
                            if (NL == 2 && R != 2) {
#pragma unroll
                                for (int k = 0; k < NL; ++k) {
                                    es[k] = fast_exp2((fs[q] - refs[k]) * c2[k]);
                                    et[k] = fast_exp2((ft[q] - reft[k]) * c2[k]);
                                }
                            } else {
                                exps<NL, R>(fs[q], refs[0], c2, es);
                                exps<NL, R>(ft[q], reft[0], c2, et);
                            }
                            float acc_o = 0.f;
#pragma unroll
                            for (int k = 0; k < NL; ++k) {
                                acc_o = fmaf(es[k], ks[k], acc_o);
                                acc_o = fmaf(et[k], kt[k], acc_o);
                            }
                            o[q] = acc_o;
                        }
                        st_streaming(reinterpret_cast<vec_t*>(out) + v, E::pack(o));
                    }
                }
            }
        }
    }

    // ================================ loss: warp partials -> CTA partial -> the last CTA sums in a fixed order
    if (lane == 0) {
        sm.klpart[warp][0] = kl_acc[0];
        sm.klpart[warp][1] = NL == 2 ? kl_acc[NL - 1] : 0.f;
    }
    bar_sync(1, kCCons);
    // nobody may leave while a peer could still read its summaries
    cluster_arrive_release();
    cluster_wait_acquire();
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
