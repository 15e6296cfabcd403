// Row-wise softmax-KL for rows that fit a THREAD-BLOCK CLUSTER: one pass over HBM, no L2 re-read, the row
// parked in TENSOR MEMORY between the statistics pass and the gradient pass.  One or two losses (CD + CGD on
// the same logits) per launch.
//
// Same mathematics as kl_rows.cu (mmseg/models/distillation/losses.py:50-58,:108-112 + backward).
// A CGD row (g = 10 channels of 128x128 logits) is 1.3 MB of S and T: no SM holds it, a cluster of 8
// does.  The cluster owns one "super-row" (a row of the loss with the larger group) at a time; CTA c
// owns slice c of it.  Three kinds of warps, coupled by mbarriers only (no CTA-wide barrier in the loop):
//
//   TMA warp     streams the CTA's slices, chunk by chunk (4096 elements of S and of T), through a 6 x 32 KB
//                shared-memory ring (1-D bulk copies, full/empty mbarriers) - up to a super-row ahead.
//   8 PARK warps (phase 1) pull a chunk into registers once (16 elements of S and of T per thread), hand the ring
//                slot back, update softmax statistics of the PIECE (part of one row of the smaller-group loss
//                inside this slice) against a WARP-UNIFORM running maximum (redux.sync.max.f32), and park the
//                exponentials in tensor memory (tcgen05.st; 32 columns per thread and chunk, 8 chunk slots per
//                warp: all 512 columns; the references of a chunk go to shared memory, once per warp).  A piece
//                that ends -> one warp record (transposed butterfly: 9 shuffles for 6 sums).
//   8 GRADIENT   (phase 2) warp 8 + i reads what park warp i parked (same TMEM lane quarter, same columns):
//     warps      tcgen05.ld brings a chunk back; one multiply-add per element and loss, no ex2 per element;
//                written once; the TMEM slot goes back to the park warp (one mbarrier per pair and slot).
//   stats warp   8 warp records -> one summary per piece; PUSHES the summaries into the shared memory of every
//                CTA of the cluster (st.async through DSMEM, completing bytes on the receiver's mbarrier - no
//                cluster barrier); merges what arrived into the row statistics the gradient warps wait for.
//
// Park and gradient run CONCURRENTLY on different warps of the same SM sub-partitions: the park side is bound by
// the MUFU/FMA pipes and shared-memory reads, the gradient side by the tensor-memory read path (64 B/clk) and
// the global stores.  The park warps run up to 8 chunks ahead of the gradient warps (a slice is 5 chunks at
// 16x150x128x128): the statistics of super-row i travel through the cluster while super-row i + 1 is parked.
//
// HBM traffic is the algorithmic read S + read T + write dS.  With two losses whose temperatures are
// tau and 2*tau (CD tau = 1 next to CGD tau = 2, the reference's defaults) exp(x/tau) = exp(x/2tau)^2:
// 2 ex2 per element for both losses.
//
// Geometry contract (cabi.cu): HW % 128 == 0, so every 32-vector row of a warp lies inside one piece of one
// slice; slices are whole chunks.
#include "rows_common.cuh"
#include "park_common.cuh"
#include "launch.h"
#ifndef SD_WAIT_NS
#define SD_WAIT_NS 20000
#endif
#define mbar_wait mbar_wait_sleep<SD_WAIT_NS>

namespace sd {

constexpr int kCPark = 256;                            // park threads (warps 0..7); gradient threads: warps 8..15
constexpr int kCParkWarps = kCPark / 32;
constexpr int kCTmaWarp = 2 * kCParkWarps;             // warp 16
constexpr int kCStatWarp = 2 * kCParkWarps + 1;        // warp 17
constexpr int kCThreads = 2 * kCPark + 64;
constexpr int kCChunkRows = 4;                         // 4-element vectors per park thread, chunk and tensor
constexpr int kCChunkVecs = kCChunkRows * kCPark;      // 1024 vectors = 4096 elements per tensor
constexpr int kCChunkBytes = kCChunkVecs * 16;         // fp32; bf16 chunks fill half a slot
constexpr int kCRing = 6;                              // ring slots of 32 KB (S chunk + T chunk)
constexpr int kCSlots = 8;                             // TMEM chunk slots per park warp
constexpr int kCMaxPieces = kClusterMaxPieces;
constexpr int kCRecFloats = 12;                        // ms, mt, {zs, zt, a, dd} x 2, 2 x pad (three 16-byte words)
constexpr int kCTmemCols = 512;
static_assert(kCSlots * kCSlotCols * (kCParkWarps / 4) == kCTmemCols, "TMEM columns");
static_assert(kClusterMaxChunks <= kCSlots, "a slice must fit the TMEM slots");
static_assert(kCChunkVecs == kClusterChunkVecs, "chunk size");

struct ClusterSmem {
    unsigned char ring[kCRing][2][kCChunkBytes];
    uint64_t full[kCRing], empty[kCRing];
    uint64_t xch[2];                                      // bytes of the pushed summaries complete here
    uint64_t recbar[2];                                   // 8 park warps: "my records of this row are written"
    uint64_t finbar[2];                                   // stats warp: "the row statistics are written"
    uint64_t tfree[kCParkWarps][kCSlots];                 // gradient warp -> its park warp: "this TMEM slot is read"
    float rec[2][kCMaxPieces][kCParkWarps][kCRecFloats];  // warp records of the pieces, by row parity
    float summ[2][kClusterMaxSize][kCMaxPieces][kCRecFloats];  // summaries of every CTA of the cluster (pushed)
    float fin[2][kCMaxPieces + 1][4];                     // row statistics for phase 2: {Ms, Mt, coef/Zs, coef/Zt}
    float refs[kCParkWarps][kCSlots][2 * kCChunkRows];    // references of a parked chunk: {ms, mt} of its vector-rows
    uint32_t tmem_base;
};
constexpr size_t kClusterSmemBytes = sizeof(ClusterSmem);



// ---------------------------------------------------------------- geometry
// position of a super-row; walks sr += n_clusters without a division per row
struct RowCursor {
    int b, grp;
    __device__ __forceinline__ void init(const ClusterGeom& g, int sr) {
        b = sr / g.G_big;
        grp = sr - b * g.G_big;
    }
    __device__ __forceinline__ void advance(const ClusterGeom& g, int step) {
        grp += step;
        while (grp >= g.G_big) {
            grp -= g.G_big;
            ++b;
        }
    }
    // 4-element vectors of this super-row (the last group of a sample may be ragged)
    __device__ __forceinline__ int lv(const RowsParams& p, const ClusterGeom& g) const {
        return min(g.g_big, p.C - grp * g.g_big) * g.hwv;
    }
    __device__ __forceinline__ size_t base(const RowsParams& p, const ClusterGeom& g) const {
        return ((size_t)b * p.C + (size_t)grp * g.g_big) * p.HW;
    }
};
// my slice of a super-row of lv vectors
struct SliceGeo {
    int v0, v1, nvs;       // the slice, in vectors of the super-row; its length
    int r_first;           // first row of l[0] (within the super-row) that intersects it
    int n_pieces;          // rows of l[0] that intersect it
    int nchunks;
    uint32_t xch_pieces;   // summaries all CTAs of the cluster push for this super-row
};
__device__ __forceinline__ int pieces_of(const ClusterGeom& g, int lv, int c) {
    const int cv0 = min(lv, c * g.slv), cv1 = min(lv, cv0 + g.slv);
    return cv1 > cv0 ? (cv1 - 1) / g.rv0 - cv0 / g.rv0 + 1 : 0;
}
// total_pieces: pieces of all slices of this super-row when the host knows them (complete super-rows), else -1
__device__ __forceinline__ SliceGeo slice_geo(const ClusterGeom& g, int lv, int rank, int total_pieces = -1) {
    SliceGeo s;
    s.v0 = min(lv, rank * g.slv);
    s.v1 = min(lv, s.v0 + g.slv);
    s.nvs = s.v1 - s.v0;
    s.r_first = s.v0 / g.rv0;
    s.n_pieces = s.nvs > 0 ? (s.v1 - 1) / g.rv0 - s.r_first + 1 : 0;
    s.nchunks = (s.nvs + kCChunkVecs - 1) / kCChunkVecs;
    int total = total_pieces;
    if (total < 0) {
        total = 0;
        for (int c = 0; c < g.nc; ++c) total += pieces_of(g, lv, c);
    }
    s.xch_pieces = (uint32_t)total;
    return s;
}
// end of piece pc, in vectors of the slice
__device__ __forceinline__ int piece_end(const ClusterGeom& g, const SliceGeo& s, int pc) {
    return min(s.v1, (s.r_first + pc + 1) * g.rv0) - s.v0;
}


// R: 0 = independent exponentials per loss, 2 = l[1].tau == 2 * l[0].tau.
// With one loss, or with R == 2, phase 1 parks the EXPONENTIALS (relative to the warp's running maximum at
// that moment, kept in shared memory): phase 2 is then a multiply-add per element, no ex2.  Otherwise the raw
// values are parked and phase 2 recomputes.
#ifdef SD_CLUSTER_TIMING
#define SD_TICK(var) const long long var = clock64()
#define SD_TACC(slot, a, b) tacc[slot] += (b) - (a)
#else
#define SD_TICK(var)
#define SD_TACC(slot, a, b)
#endif

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
    const int cluster_id = blockIdx.x / NC;
    const int n_clusters = gridDim.x / NC;

    if (p.run_if != nullptr && *p.run_if == 0u) return;  // cancelled backward re-run (uniform over the grid)
#ifdef SD_CLUSTER_TIMING
    const long long tk0 = clock64();
#define SD_STAMP(i) p.pkt[blockIdx.x * 16 + (i)] = (unsigned long long)(clock64() - tk0)
#else
#define SD_STAMP(i)
#endif

    if (tid == 0) {
        for (int c = 0; c < kCRing; ++c) {
            mbar_init(&sm.full[c], 1);
            mbar_init(&sm.empty[c], kCParkWarps);
        }
        for (int q = 0; q < 2; ++q) {
            mbar_init(&sm.xch[q], 1);
            mbar_init(&sm.recbar[q], kCParkWarps);
            mbar_init(&sm.finbar[q], 1);
        }
        for (int w = 0; w < kCParkWarps; ++w)
            for (int q = 0; q < kCSlots; ++q) mbar_init(&sm.tfree[w][q], 1);
        fence_barrier_init();
    }
    if (warp == kCTmaWarp) tmem_alloc(&sm.tmem_base, kCTmemCols);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;
    // the only cluster-wide barrier: every CTA's mbarriers exist before a peer pushes bytes at them
    cluster_arrive_release();
    cluster_wait_acquire();

#ifdef SD_CLUSTER_TIMING
    if (tid == 0) SD_STAMP(10);      // prologue done (barriers, TMEM, cluster sync)
#endif
    const int n_iter = cluster_id < g.total_sr ? (g.total_sr - cluster_id + n_clusters - 1) / n_clusters : 0;
    // every complete super-row is cut the same way; only a ragged last group needs its own geometry
    const int lv_full = g.g_big * g.hwv;
    const SliceGeo geo_full = slice_geo(g, lv_full, (int)rank, g.pieces_full);
    auto geo_of = [&](int lv) { return lv == lv_full ? geo_full : slice_geo(g, lv, (int)rank); };

    if (warp == kCTmaWarp) {
        // =====================================================================================
        // TMA warp: lane 0 streams this CTA's slices, chunk by chunk, into the ring as slots drain
        // =====================================================================================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            int slot = 0;
            uint32_t phase = 0;
            RowCursor rc;
            rc.init(g, cluster_id);
            for (int it = 0; it < n_iter; ++it) {
                const int lv = rc.lv(p, g);
                const size_t base = rc.base(p, g);
                const int v0 = min(lv, (int)rank * g.slv);
                const int v1 = min(lv, v0 + g.slv);
                for (int c = 0; c * kCChunkVecs < v1 - v0; ++c) {
                    mbar_wait(&sm.empty[slot], phase ^ 1u);
                    const int nv = min(kCChunkVecs, v1 - v0 - c * kCChunkVecs);
                    const uint32_t bytes = (uint32_t)nv * (uint32_t)sizeof(vec_t);
                    mbar_arrive_expect_tx(&sm.full[slot], 2u * bytes);
                    const size_t off = (base + (size_t)(v0 + c * kCChunkVecs) * VE) * sizeof(T);
                    tma_bulk_g2s(sm.ring[slot][0], static_cast<const char*>(p.S) + off, bytes, &sm.full[slot], pol);
                    tma_bulk_g2s(sm.ring[slot][1], static_cast<const char*>(p.T) + off, bytes, &sm.full[slot], pol);
                    if (++slot == kCRing) {
                        slot = 0;
                        phase ^= 1u;
                    }
                }
                rc.advance(g, n_clusters);
            }
        }
        __syncwarp();
        // TMEM is released after every consumer of this CTA is through with it
        bar_sync(3, kCThreads);
        tmem_dealloc(tmem_base, kCTmemCols);
        return;
    }

    float c2[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) c2[k] = p.l[k].c2;
    const int rv0 = g.rv0;             // vectors of a complete row of l[0]

    if (warp == kCStatWarp) {
        // =====================================================================================
        // stats warp: records -> summaries -> (cluster) -> row statistics, one super-row after the other
        // =====================================================================================
        float kl_acc[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) kl_acc[k] = 0.f;
        RowCursor rc;
        rc.init(g, cluster_id);
        // Which summaries this lane merges depends on the geometry only: planned once for the complete
        // super-rows (integer divisions stay out of the per-row chain), again for a ragged one.
        constexpr int kSrRecs = (kClusterMaxSize * kCMaxPieces + 31) / 32;   // summaries of a super-row per lane
        constexpr int kPrPasses = (kCMaxPieces + 3) / 4;                     // rows of l[0]: four per pass, 8 lanes each
        struct StatPlan {
            int sr_off[kSrRecs];     // float offset (in summ[par]) of the summaries this lane folds into the super-row, or -1
            int pr_off[kPrPasses];   // ... of the summary (CTA ca + j, its piece of my row) this lane contributes, or -1
            int pr_row[kPrPasses];   // row of l[0] within the super-row (-1: no such piece)
            int pr_ca[kPrPasses];    // first CTA that holds a piece of it (the one that accounts for its KL term)
        };
        auto make_plan = [&](int lv, const SliceGeo& s) {
            StatPlan pl;
#pragma unroll
            for (int q = 0; q < kSrRecs; ++q) {
                const int i = lane + 32 * q;
                const int c = i / kCMaxPieces, pc = i - c * kCMaxPieces;
                pl.sr_off[q] = (c < NC && pc < pieces_of(g, lv, c)) ? (c * kCMaxPieces + pc) * kCRecFloats : -1;
            }
#pragma unroll
            for (int q = 0; q < kPrPasses; ++q) {
                const int pc = 4 * q + (lane >> 3), j = lane & 7;
                pl.pr_off[q] = -1;
                pl.pr_row[q] = -1;
                pl.pr_ca[q] = 0;
                if (pc < s.n_pieces) {
                    const int row = s.r_first + pc;                           // row of l[0] within the super-row
                    const int rv_lo = row * rv0, rv_hi = min(lv, rv_lo + rv0);
                    const int ca = rv_lo / g.slv, cb = (rv_hi - 1) / g.slv;   // its pieces live in CTAs ca..cb, one each
                    pl.pr_row[q] = row;
                    pl.pr_ca[q] = ca;
                    if (j <= cb - ca) {
                        const int c = ca + j;
                        const int cpc = row - min(lv, c * g.slv) / rv0;
                        pl.pr_off[q] = (c * kCMaxPieces + cpc) * kCRecFloats;
                    }
                }
            }
            return pl;
        };
        const StatPlan plan_full = make_plan(lv_full, geo_full);
        // where lane w (< NC) of a group of 8 pushes: summ[0][rank][0] and xch[0] of CTA w
        const uint32_t push_to = (uint32_t)(lane & 7) < (uint32_t)NC ? (uint32_t)(lane & 7) : rank;
        const uint32_t push_dst = map_to_cta(sm.summ[0][rank][0], push_to);
        const uint32_t push_bar = map_to_cta(&sm.xch[0], push_to);
        float coef[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            coef[k] = p.l[k].coef;
            if (p.grad_out[k] != nullptr) coef[k] *= __ldg(p.grad_out[k]);
        }
        auto load_stat = [&](const float* base, int off) {
            PStat<NL> x = pstat_empty<NL>();
            if (off >= 0) {
                const float4* q = reinterpret_cast<const float4*>(base + off);
                x = pstat_from<NL>(q[0], q[1], NL == 2 ? q[2] : make_float4(0.f, 0.f, 0.f, 0.f));
            }
            return x;
        };
#ifdef SD_CLUSTER_TIMING
        long long tacc[4] = {0, 0, 0, 0};
#endif
        for (int it = 0; it < n_iter; ++it) {
            const int par = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            SD_TICK(t0);
            const int lv = rc.lv(p, g);
            const bool full = lv == lv_full;
            const SliceGeo s = full ? geo_full : slice_geo(g, lv, (int)rank);
            const StatPlan pl = full ? plan_full : make_plan(lv, s);
            // this row's summaries: 32 (one loss) / 48 bytes per piece of every CTA of the cluster will land in summ[par]
            if (lane == 0) mbar_arrive_expect_tx(&sm.xch[par], s.xch_pieces * (NL == 2 ? 48u : 32u));
            mbar_wait(&sm.recbar[par], ph);
            SD_TICK(t1);
            // ---- 8 warp records -> one summary per piece, four pieces per pass (8 lanes each);
            //      lane c of the group pushes the summary into CTA c (this CTA included)
            for (int pc0 = 0; pc0 < s.n_pieces; pc0 += 4) {
                const int pc = pc0 + (lane >> 3), w = lane & 7;
                PStat<NL> st = load_stat(sm.rec[par][0][0], pc < s.n_pieces ? (pc * kCParkWarps + w) * kCRecFloats : -1);
                st = pstat_reduce<NL, R, 8>(st, c2);
                if (pc < s.n_pieces && w < NC) {
                    const uint32_t off = (uint32_t)(((par * kClusterMaxSize) * kCMaxPieces + pc) * kCRecFloats * 4);
                    const uint32_t dst = push_dst + off, bar = push_bar + (uint32_t)par * 8u;
                    st_async_f4(dst, make_float4(st.ms, st.mt, st.zs[0], st.zt[0]), bar);
                    st_async_f4(dst + 16, make_float4(st.a[0], st.dd[0], st.zs[NL - 1], st.zt[NL - 1]), bar);
                    if (NL == 2) st_async_f4(dst + 32, make_float4(st.a[NL - 1], st.dd[NL - 1], 0.f, 0.f), bar);
                }
            }
            SD_TICK(t2);
            mbar_wait(&sm.xch[par], ph);
            SD_TICK(t3);
            // ---- merge what arrived: the super-row (two losses: every summary of every CTA) and the rows of
            //      l[0] my pieces belong to (four per pass, 8 lanes each: a row lives in <= 8 CTAs).  Straight-line
            //      code: the two butterflies interleave.
            const float* sbase = sm.summ[par][0][0];
            PStat<NL> sr = pstat_empty<NL>();
            if (NL == 2) {
                sr = load_stat(sbase, pl.sr_off[0]);
#pragma unroll
                for (int q = 1; q < kSrRecs; ++q)
                    if (__any_sync(0xffffffffu, pl.sr_off[q] >= 0)) sr = pstat_merge<NL, R>(sr, load_stat(sbase, pl.sr_off[q]), c2);
            }
            PStat<NL> pr[kPrPasses];
            pr[0] = load_stat(sbase, pl.pr_off[0]);
            if (NL == 2) sr = pstat_reduce<NL, R, 32>(sr, c2);
            pr[0] = pstat_reduce<NL, R, 8>(pr[0], c2);
            if (NL == 2 && lane == 0)
                *reinterpret_cast<float4*>(sm.fin[par][kCMaxPieces]) =
                    make_float4(sr.ms, sr.mt, __fdividef(coef[K], sr.zs[K]), __fdividef(coef[K], sr.zt[K]));
            if ((lane & 7) == 0 && pl.pr_row[0] >= 0)
                *reinterpret_cast<float4*>(sm.fin[par][lane >> 3]) =
                    make_float4(pr[0].ms, pr[0].mt, __fdividef(coef[0], pr[0].zs[0]), __fdividef(coef[0], pr[0].zt[0]));
#pragma unroll
            for (int q = 1; q < kPrPasses; ++q) {
                pr[q] = pstat_empty<NL>();
                if (s.n_pieces > 4 * q) {
                    pr[q] = pstat_reduce<NL, R, 8>(load_stat(sbase, pl.pr_off[q]), c2);
                    if ((lane & 7) == 0 && pl.pr_row[q] >= 0)
                        *reinterpret_cast<float4*>(sm.fin[par][4 * q + (lane >> 3)]) =
                            make_float4(pr[q].ms, pr[q].mt, __fdividef(coef[0], pr[q].zs[0]), __fdividef(coef[0], pr[q].zt[0]));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.finbar[par]);
            SD_TICK(t4);
            // ---- off the critical path: the KL terms of the rows this CTA accounts for
            if (NL == 2 && lane == 0 && rank == 0) {
                const float kl = kl_of_row(1.f, sr.zs[K], sr.zt[K], sr.a[K], sr.dd[K]);
                if (p.l[K].row_kl) p.l[K].row_kl[rc.b * p.l[K].G + rc.grp] = kl;
                kl_acc[K] += kl;
            }
#pragma unroll
            for (int q = 0; q < kPrPasses; ++q) {
                if ((lane & 7) == 0 && pl.pr_row[q] >= 0 && (int)rank == pl.pr_ca[q]) {
                    const float kl = kl_of_row(NL == 2 && R == 2 ? 2.f : 1.f, pr[q].zs[0], pr[q].zt[0], pr[q].a[0], pr[q].dd[0]);
                    const int rowi = NL == 2 ? rc.b * p.l[0].G + rc.grp * p.l[NL - 1].m + pl.pr_row[q] : rc.b * p.l[0].G + rc.grp;
                    if (p.l[0].row_kl) p.l[0].row_kl[rowi] = kl;
                    kl_acc[0] += kl;
                }
            }
            SD_TACC(0, t0, t1);
            SD_TACC(1, t1, t2);
            SD_TACC(2, t2, t3);
            SD_TACC(3, t3, t4);
            rc.advance(g, n_clusters);
        }
#ifdef SD_CLUSTER_TIMING
        if (lane == 0)
            for (int q = 0; q < 4; ++q) p.pkt[blockIdx.x * 16 + q] = (unsigned long long)tacc[q];
#endif
        // ---- loss: lanes 0, 8, 16, 24 hold the row terms of this CTA, summed in a fixed order
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const float a8 = __shfl_sync(0xffffffffu, kl_acc[k], 8), a16 = __shfl_sync(0xffffffffu, kl_acc[k], 16),
                        a24 = __shfl_sync(0xffffffffu, kl_acc[k], 24);
            kl_acc[k] = (kl_acc[k] + a8) + (a16 + a24);
        }
        // every consumer is through with TMEM (the TMA warp frees it); nobody pushes at this CTA any more: its
        // last exchange completed above
        bar_sync(3, kCThreads);
        unsigned ticket = 0;
        if (lane == 0) {
            __stcg(&p.cta_part[blockIdx.x], kl_acc[0]);
            if (NL == 2) __stcg(&p.cta_part[kMaxGrid + blockIdx.x], kl_acc[NL - 1]);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            // the last CTA sums the partials in a fixed order
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
        return;
    }

    // =========================================================================================
    // park warps 0..7 and gradient warps 8..15: warp 8 + i retrieves what warp i parked
    // =========================================================================================
    const int pw = warp & (kCParkWarps - 1);
    const int ptid = tid & (kCPark - 1);
    // the pair's TMEM window: lane quarter of both warps, 256 columns, 8 chunk slots of 32
    const uint32_t tmem_mine = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((pw >> 2) * 256);
    const int wv = pw * 32;            // the pair's vector-rows of a chunk start here, + 256 r (r < 4)

    if (warp < kCParkWarps) {
        // ------------------------------------------------ phase 1: statistics, park the slice
        RowCursor rcA;
        rcA.init(g, cluster_id);
        SliceGeo gA = geo_full;
        int itA = 0, cA = 0;               // super-row, chunk
        int pcA = 0, pv1A = 0;             // current piece and its end
        PStat<NL> st = pstat_empty<NL>();  // its statistics so far: maxima identical in every lane, sums per lane
        uint32_t qA = 0;                   // chunks parked so far -> TMEM slot
        int slot = 0;
        uint32_t phase = 0;
#ifdef SD_CLUSTER_TIMING
        long long tacc[4] = {0, 0, 0, 0};
        const long long tstart = clock64();
#endif

        // a piece ends: the warp's record (maxima + 3 sums per loss)
        auto close_piece = [&]() {
            float v[8] = {st.zs[0], st.zt[0], st.a[0], st.dd[0], 0.f, 0.f, 0.f, 0.f};
            if (NL == 2) {
                v[4] = st.zs[NL - 1];
                v[5] = st.zt[NL - 1];
                v[6] = st.a[NL - 1];
                v[7] = st.dd[NL - 1];
            }
            const float tot = warp_sum8_transposed(v, lane);
            float* rec = sm.rec[itA & 1][pcA][pw];
            if ((lane & 3) == 0 && lane < 4 * 4 * NL) rec[2 + (lane >> 2)] = tot;
            if (lane == 1) rec[0] = st.ms;
            if (lane == 2) rec[1] = st.mt;
            st = pstat_empty<NL>();
            ++pcA;
            pv1A = piece_end(g, gA, pcA);
        };
        // the running maxima move to (at least) wms, wmt: rescale the sums (rarely needed after the first chunks)
        auto raise_refs = [&](float wms, float wmt) {
            const float nms = fmaxf(st.ms, wms), nmt = fmaxf(st.mt, wmt);
            if (nms != st.ms || nmt != st.mt) {      // warp-uniform
                float rs[NL], rt[NL], df[NL], sh[NL];
                exps<NL, R>(st.ms, nms, c2, rs);
                exps<NL, R>(st.mt, nmt, c2, rt);
                shifts<NL, R>(st.ms, st.mt, nms, nmt, c2, sh);
                factor_diffs<NL, R>(sh, rs, rt, df);
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    st.dd[k] = fmaf(st.zs[k], df[k], st.dd[k] * rt[k]);
                    st.zs[k] *= rs[k];
                    st.zt[k] *= rt[k];
                    st.a[k] = fmaf(st.zt[k], sh[k], st.a[k] * rt[k]);
                }
                st.ms = nms;
                st.mt = nmt;
            }
        };
        // N elements into the statistics of the current piece; what gets parked replaces them in fs / ft
        auto accumulate = [&](float* fs, float* ft, int n) {
            float refs2[NL], reft2[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                refs2[k] = __fmul_rn(st.ms, c2[k]);
                reft2[k] = __fmul_rn(st.mt, c2[k]);
            }
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                if (i < n) {
                    float as[NL], at[NL], es[NL], et[NL];
                    exps_args<NL, R>(fs[i], refs2, c2, as, es);
                    exps_args<NL, R>(ft[i], reft2, c2, at, et);
                    if (NL == 2 && R == 2) {
                        // e0 = eK^2: the sums of loss 0 straight from the loss-K exponentials (one FFMA each)
                        const float da = at[K] - as[K], dk = et[K] - es[K];
                        st.zs[K] += es[K];
                        st.zt[K] += et[K];
                        st.zs[0] = fmaf(es[K], es[K], st.zs[0]);
                        st.zt[0] = fmaf(et[K], et[K], st.zt[0]);
                        const float w = et[K] * da;
                        st.a[K] += w;
                        st.a[0] = fmaf(et[K], w, st.a[0]);
                        st.dd[K] += dk;
                        st.dd[0] = fmaf(dk, et[K] + es[K], st.dd[0]);      // et0 - es0 = (eK_t - eK_s)(eK_t + eK_s)
                    } else {
#pragma unroll
                        for (int k = 0; k < NL; ++k) {
                            st.zs[k] += es[k];
                            st.zt[k] += et[k];
                            st.a[k] = fmaf(et[k], at[k] - as[k], st.a[k]);
                        }
#pragma unroll
                        for (int k = 0; k < NL; ++k) st.dd[k] += et[k] - es[k];
                    }
                    if (kParkExp) {
                        fs[i] = es[K];
                        ft[i] = et[K];
                    }
                }
            }
        };

        for (; itA < n_iter; ++itA, rcA.advance(g, n_clusters)) {
            gA = geo_of(rcA.lv(p, g));
            pcA = 0;
            pv1A = piece_end(g, gA, 0);
            st = pstat_empty<NL>();
            SD_TICK(t0);
            for (cA = 0; cA < gA.nchunks; ++cA, ++qA) {
                // ---- one chunk: ring -> registers -> statistics -> TMEM
                const uint32_t ts = qA & (uint32_t)(kCSlots - 1);
                if (qA >= (uint32_t)kCSlots) {
                    // the slot's previous chunk has been retrieved by my gradient warp
                    SD_TICK(w0);
                    mbar_wait(&sm.tfree[pw][ts], ((qA >> 3) - 1u) & 1u);
                    SD_TICK(w1);
                    SD_TACC(1, w0, w1);
                    tmem_fence_after_sync();
                }
                SD_TICK(w2);
                mbar_wait(&sm.full[slot], phase);
                SD_TICK(w3);
                SD_TACC(2, w2, w3);
#ifdef SD_CLUSTER_TIMING
                if (tid == 0 && qA == 0) SD_STAMP(11);   // first chunk arrived
#endif
                const vec_t* bs = reinterpret_cast<const vec_t*>(sm.ring[slot][0]);
                const vec_t* bt = reinterpret_cast<const vec_t*>(sm.ring[slot][1]);
                float fs[NE], ft[NE];
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    V::unpack(bs[r * kCPark + ptid], &fs[r * VE]);
                    V::unpack(bt[r * kCPark + ptid], &ft[r * VE]);
                }
                float mxs[kCChunkRows], mxt[kCChunkRows];
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
                    mxs[r] = fmaxf(fmaxf(fs[r * VE], fs[r * VE + 1]), fmaxf(fs[r * VE + 2], fs[r * VE + 3]));
                    mxt[r] = fmaxf(fmaxf(ft[r * VE], ft[r * VE + 1]), fmaxf(ft[r * VE + 2], ft[r * VE + 3]));
                }
                const int vA = cA * kCChunkVecs + wv;      // my first vector-row, in vectors of the slice; + 256 r
                float rf[2 * kCChunkRows];
                while (pcA < gA.n_pieces && vA >= pv1A) close_piece();
                if (vA + (kCChunkRows - 1) * kCPark < pv1A) {
                    // ---- all my vector-rows lie in the current piece (the common case)
                    const float wms = warp_max_uniform(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])));
                    const float wmt = warp_max_uniform(fmaxf(fmaxf(mxt[0], mxt[1]), fmaxf(mxt[2], mxt[3])));
                    // every lane's shared-memory reads went into the maxima: the slot may go back to the TMA warp
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);
                    raise_refs(wms, wmt);
                    accumulate(fs, ft, NE);
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) {
                        rf[2 * r] = st.ms;
                        rf[2 * r + 1] = st.mt;
                    }
                } else {
                    // ---- a piece (or the slice) ends in this chunk: vector-row by vector-row
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) {
                        const int v = vA + r * kCPark;
                        while (pcA < gA.n_pieces && v >= pv1A) close_piece();
                        rf[2 * r] = rf[2 * r + 1] = 0.f;
                        if (v < gA.nvs) {
                            const float wms = warp_max_uniform(mxs[r]);
                            const float wmt = warp_max_uniform(mxt[r]);
                            raise_refs(wms, wmt);
                            accumulate(&fs[r * VE], &ft[r * VE], VE);
                            rf[2 * r] = st.ms;
                            rf[2 * r + 1] = st.mt;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);   // (the maxima above consumed the reads)
                }
                // ---- park: registers -> TMEM; the references once per warp
                Parked pk;
#pragma unroll
                for (int r = 0; r < kCChunkRows; ++r) {
#pragma unroll
                    for (int q = 0; q < VE; ++q) {
                        pk.w[r * 8 + q] = __float_as_uint(fs[r * VE + q]);
                        pk.w[r * 8 + 4 + q] = __float_as_uint(ft[r * VE + q]);
                    }
                }
                tmem_st32(tmem_mine + ts * kCSlotCols, pk);
                if (kParkExp && lane == 0) {
                    float4* d = reinterpret_cast<float4*>(sm.refs[pw][ts]);
                    d[0] = make_float4(rf[0], rf[1], rf[2], rf[3]);
                    d[1] = make_float4(rf[4], rf[5], rf[6], rf[7]);
                }
                if (++slot == kCRing) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
            // the slice is parked (or empty): the records of the pieces still open; what I stored in TMEM is
            // complete and ordered before the arrival the gradient warps will (transitively) wait on
            while (pcA < gA.n_pieces) close_piece();
            tmem_wait_st();
            tmem_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.recbar[itA & 1]);
            SD_TICK(t1);
            SD_TACC(0, t0, t1);
#ifdef SD_CLUSTER_TIMING
            if (tid == 0 && itA == 0) SD_STAMP(12);      // first row parked
#endif
        }
#ifdef SD_CLUSTER_TIMING
        if (tid == 0) {
            for (int q = 0; q < 3; ++q) p.pkt[blockIdx.x * 16 + 4 + q] = (unsigned long long)tacc[q];
            p.pkt[blockIdx.x * 16 + 7] = (unsigned long long)(clock64() - tstart);
        }
#endif
    } else {
        // ------------------------------------------------ phase 2: gradient from the parked slice
        RowCursor rcB;
        rcB.init(g, cluster_id);
        uint32_t qB = 0;                   // chunks retrieved so far -> TMEM slot
#ifdef SD_CLUSTER_TIMING
        long long tacc[4] = {0, 0, 0, 0};
#endif
        for (int itB = 0; itB < n_iter; ++itB, rcB.advance(g, n_clusters)) {
            const int par = itB & 1;
            const SliceGeo gB = geo_of(rcB.lv(p, g));
            if (gB.nchunks == 0) continue;
            SD_TICK(t0);
            mbar_wait(&sm.finbar[par], (uint32_t)(itB >> 1) & 1u);
            tmem_fence_after_sync();
            SD_TICK(t1);
#ifdef SD_CLUSTER_TIMING
            if (tid == kCPark && itB == 0) SD_STAMP(13);     // first row statistics arrived
            if (tid == kCPark && itB == n_iter - 1) SD_STAMP(14);   // last row statistics arrived
#endif
            T* out = static_cast<T*>(p.dS) + rcB.base(p, g) + (size_t)gB.v0 * VE;
            const float4 fb = *reinterpret_cast<const float4*>(sm.fin[par][kCMaxPieces]);   // {Ms, Mt, coef/Zs, coef/Zt} of the super-row
            int pc = 0;
            int pv1 = piece_end(g, gB, 0);
            float4 fa = *reinterpret_cast<const float4*>(sm.fin[par][0]);
            // one vector-row: parked values (+ the references they were taken against) -> gradient, given the
            // statistics `fa` of its l[0] row
            auto grad = [&](const uint32_t* w, float ref_s, float ref_t, int v) {
                float fs[VE], ft[VE], o[VE];
#pragma unroll
                for (int q = 0; q < VE; ++q) {
                    fs[q] = __uint_as_float(w[q]);
                    ft[q] = __uint_as_float(w[4 + q]);
                }
                if (kParkExp) {
                    // parked: e = exp2((x - ref) c2[K]); softmax_k = e^(c2[k]/c2[K]) * exp2((ref - M_k) c2[k]) / Z_k, ref <= M_k
                    const float gsK = (NL == 2 ? fb.z : fa.z) * ref_factor(ref_s, NL == 2 ? fb.x : fa.x, c2[K]);
                    const float gtK = (NL == 2 ? fb.w : fa.w) * ref_factor(ref_t, NL == 2 ? fb.y : fa.y, c2[K]);
                    if (NL == 2) {
                        const float gs0 = fa.z * ref_factor(ref_s, fa.x, c2[0]);
                        const float gt0 = fa.w * ref_factor(ref_t, fa.y, c2[0]);
#pragma unroll
                        for (int q = 0; q < VE; ++q) o[q] = fs[q] * fmaf(fs[q], gs0, gsK) - ft[q] * fmaf(ft[q], gt0, gtK);
                    } else {
#pragma unroll
                        for (int q = 0; q < VE; ++q) o[q] = fmaf(fs[q], gsK, -ft[q] * gtK);
                    }
                } else {
                    // raw values parked: recompute against the row maxima
                    const float rs0 = __fmul_rn(fa.x, c2[0]), rt0 = __fmul_rn(fa.y, c2[0]);
                    const float rs1 = __fmul_rn(fb.x, c2[K]), rt1 = __fmul_rn(fb.y, c2[K]);
#pragma unroll
                    for (int q = 0; q < VE; ++q) {
                        const float es0 = fast_exp2(fmaf(fs[q], c2[0], -rs0));
                        const float et0 = fast_exp2(fmaf(ft[q], c2[0], -rt0));
                        const float es1 = fast_exp2(fmaf(fs[q], c2[K], -rs1));
                        const float et1 = fast_exp2(fmaf(ft[q], c2[K], -rt1));
                        o[q] = fmaf(es0, fa.z, es1 * fb.z) - fmaf(et0, fa.w, et1 * fb.w);
                    }
                }
                V::store(out + (size_t)(v + lane) * VE, o);
            };
            auto grad_chunk = [&](int c, const Parked& pk, uint32_t ts) {
                const int vA = c * kCChunkVecs + wv;
                float rf[2 * kCChunkRows];
#pragma unroll
                for (int q = 0; q < 2 * kCChunkRows; ++q) rf[q] = 0.f;
                if (kParkExp) {
                    const float4* d = reinterpret_cast<const float4*>(sm.refs[pw][ts]);
                    const float4 d0 = d[0], d1 = d[1];
                    rf[0] = d0.x; rf[1] = d0.y; rf[2] = d0.z; rf[3] = d0.w;
                    rf[4] = d1.x; rf[5] = d1.y; rf[6] = d1.z; rf[7] = d1.w;
                }
                // the parked values are in registers, the references too: the slot may be parked into again
                __syncwarp();
                tmem_fence_before_sync();
                if (lane == 0) mbar_arrive(&sm.tfree[pw][ts]);
                auto next_piece = [&]() {
                    ++pc;
                    pv1 = piece_end(g, gB, pc);
                    if (pc < gB.n_pieces) fa = *reinterpret_cast<const float4*>(sm.fin[par][pc]);
                };
                while (pc < gB.n_pieces && vA >= pv1) next_piece();
                if (vA + (kCChunkRows - 1) * kCPark < pv1) {
                    if (kParkExp && NL == 2) {
                        // all vector-rows in one piece, parked against the same references: the four factors once
                        const float gsK = fb.z * ref_factor(rf[0], fb.x, c2[K]);
                        const float gtK = fb.w * ref_factor(rf[1], fb.y, c2[K]);
                        const float gs0 = fa.z * ref_factor(rf[0], fa.x, c2[0]);
                        const float gt0 = fa.w * ref_factor(rf[1], fa.y, c2[0]);
#pragma unroll
                        for (int r = 0; r < kCChunkRows; ++r) {
                            float o[VE];
#pragma unroll
                            for (int q = 0; q < VE; ++q) {
                                const float es = __uint_as_float(pk.w[r * 8 + q]), et = __uint_as_float(pk.w[r * 8 + 4 + q]);
                                o[q] = es * fmaf(es, gs0, gsK) - et * fmaf(et, gt0, gtK);
                            }
                            V::store(out + (size_t)(vA + r * kCPark + lane) * VE, o);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < kCChunkRows; ++r) grad(&pk.w[r * 8], rf[2 * r], rf[2 * r + 1], vA + r * kCPark);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kCChunkRows; ++r) {
                        const int v = vA + r * kCPark;
                        while (pc < gB.n_pieces && v >= pv1) next_piece();
                        if (v < gB.nvs) grad(&pk.w[r * 8], rf[2 * r], rf[2 * r + 1], v);
                    }
                }
            };
            // parked chunks in flight from TMEM, one ahead of the arithmetic, in two register sets
            Parked pa, pb;
            const int n = gB.nchunks;
            tmem_ld32(tmem_mine + (qB & (uint32_t)(kCSlots - 1)) * kCSlotCols, pa);
            for (int c = 0; c < n; c += 2) {
                const uint32_t t0_ = (qB + (uint32_t)c) & (uint32_t)(kCSlots - 1), t1_ = (t0_ + 1) & (uint32_t)(kCSlots - 1),
                               t2_ = (t0_ + 2) & (uint32_t)(kCSlots - 1);
                tmem_wait_ld(pa);
                if (c + 1 < n) tmem_ld32(tmem_mine + t1_ * kCSlotCols, pb);
                grad_chunk(c, pa, t0_);
                if (c + 1 < n) {
                    tmem_wait_ld(pb);
                    if (c + 2 < n) tmem_ld32(tmem_mine + t2_ * kCSlotCols, pa);
                    grad_chunk(c + 1, pb, t1_);
                }
            }
            qB += (uint32_t)n;
            SD_TICK(t2);
            SD_TACC(0, t0, t1);
            SD_TACC(1, t1, t2);
        }
#ifdef SD_CLUSTER_TIMING
        if (tid == kCPark)
            for (int q = 0; q < 2; ++q) p.pkt[blockIdx.x * 16 + 8 + q] = (unsigned long long)tacc[q];
#endif
    }
#ifdef SD_CLUSTER_TIMING
    if (tid == kCPark) SD_STAMP(15);                         // last gradient written
#endif
    // every consumer is through with TMEM (the TMA warp frees it)
    bar_sync(3, kCThreads);
}

// ====================================================================================================
template <typename T, int NL, int R>
static cudaError_t launch_cluster_t(const RowsParams& p, ClusterGeom g, int sms, cudaStream_t stream, bool probe_only) {
    auto kern = kl_rows_cluster_kernel<T, NL, R>;
    static std::atomic<bool> configured[kMaxDevices];      // per instantiation and device
    static std::atomic<int> max_clusters_dev[kMaxDevices][kClusterMaxSize + 1];
    const int dev = device_slot();
    if (!configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    std::atomic<int>* max_clusters = max_clusters_dev[dev];
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
