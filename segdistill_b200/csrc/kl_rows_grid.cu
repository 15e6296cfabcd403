// Row-wise softmax-KL for rows longer than one SM can hold, resident in the WHOLE GRID: one pass over HBM, no L2
// re-read, every SM of the GPU busy.  One or two losses (CD + CGD on the same logits) per launch.
//
// Same mathematics as kl_rows.cu (mmseg/models/distillation/losses.py:50-58,:108-112 + backward).
// kl_rows_cluster.cu keeps a long row (CGD, g = 10 channels of 128x128 logits: 1.3 MB of S and T) resident in a
// thread-block cluster of 8 - but only 15 such clusters fit the GPU's GPCs: 120 of 148 SMs work.  Here the row is
// spread over CTAs that need not be neighbours: the grid is launched cooperatively (one CTA per SM, all resident),
// a UNIT is a run of 1 .. 4 chunks (4096 elements of S and of T each) of one row of the smaller-group loss (4; for
// short work lists 1 for the tail: see "units" below), unit u belongs to CTA u % grid - the units of one row are
// worked on at the same time by different SMs - and the
// softmax statistics of a unit travel as an epoch-tagged packet through global memory (L2), like in
// kl_rows_stream.cu, instead of through distributed shared memory.  Between the statistics and the gradient a unit
// is PARKED IN TENSOR MEMORY, as in the cluster kernel: 8 chunk slots per SM, so up to eight chunks are in flight
// while the packets of the row-mates arrive - nobody waits for a neighbour in the steady state.
//
//   TMA warp     streams the CTA's units, chunk by chunk, through a 6 x 32 KB shared-memory ring; the same warp is the
//                PUBLISHER: 8 warp records of a parked unit -> the unit's packet in global memory (both jobs polled).
//   8 PARK warps (phase 1) chunk -> registers (16 elements of S and of T per thread), statistics of the unit against a
//                warp-uniform running maximum, exponentials -> tensor memory (tcgen05.st); a unit that ends -> one
//                warp record.
//   8 GRADIENT   (phase 2) warp 8 + i retrieves what park warp i parked (tcgen05.ld) once the statistics of the
//     warps      unit's rows are known: one multiply-add per element and loss, no second ex2; dS written once.
//   3 gather     the packets of the row-mates of a unit (warp i: the CTA's units j % 3 == i), merged (row of the
//     warps      smaller-group loss: the mates of that row only; row of the larger-group loss: all of them) -> the row
//                statistics the gradient warps wait for; the KL terms of the rows whose first unit is the CTA's.
//
// The same kernel walks the work list of SEVERAL (student, teacher) pairs (template parameter MULTI,
// sd_kl_rows_group_fwd_bwd: one loss per pair, temperature / weight / tensors per unit).  The sums of the park warps
// and the gradient use two fp32 per instruction (FFMA2 / FADD2 / FMUL2, common.cuh).
//
// Geometry contract (cabi.cu): HW % 128 == 0 (a warp's 32 consecutive 4-element vectors are all inside or all outside
// a unit), units of whole chunks (the last unit of a row may be shorter), at most 64 units per row of any fused loss
// and no more than the grid holds, no channel gather, no fused MSE.
#include "rows_common.cuh"
#include "park_common.cuh"
#include "launch.h"

#include <type_traits>

namespace sd {

constexpr int kGGradWarps = 8;                         // gradient warps (two per TMEM lane quarter)
constexpr int kGGatherWarps = 3;                       // packets -> row statistics (warp i: units j % 3 == i)
constexpr int kGChunkVecs = 1024;                      // 4-element vectors per chunk and tensor = 4096 elements
constexpr int kGChunkBytes = kGChunkVecs * 16;         // fp32; bf16 chunks fill half a slot
constexpr int kGRing = 6;                              // ring slots of 32 KB (S chunk + T chunk)
constexpr int kGSlots = 8;                             // TMEM chunk slots per park warp
constexpr int kGDepth = 8;                             // units between park and gradient (a unit is >= 1 chunk)
constexpr int kGRecFloats = 12;                        // ms, mt, {zs, zt, a, dd} x 2, 2 x pad (three 16-byte words)
constexpr int kGTmemCols = 512;
constexpr int kGPktWords = 2 + 4 * kMaxLosses;         // words of a packet in use (same order as a warp record)
static_assert(kGridUnitMaxChunks * 2 <= kGSlots, "two units must fit the TMEM slots (deadlock freedom)");
static_assert(kGPktWords <= kPktWords, "packet size");
static_assert(kGChunkVecs * 4 == kGridChunkElems, "chunk size");

// PW park warps (8: 16 element pairs per thread and chunk, 96 registers per thread; 16: 8 pairs, 72 registers - four
// park warps per scheduler instead of two), 8 gradient warps (each serves PW / 8 park warps of its TMEM lane quarter),
// the TMA + publisher warp, 3 gather warps
template <int PW>
struct GridCfg {
    static constexpr int kParkWarps = PW;
    static constexpr int kPark = 32 * PW;                          // park threads (warps 0 .. PW - 1)
    static constexpr int kChunkRows = kGChunkVecs / kPark;         // 4-element vectors per park thread, chunk and tensor
    static constexpr int kSlotCols = 8 * kChunkRows;               // TMEM columns of a parked chunk per thread
    static constexpr int kWarpCols = kGSlots * kSlotCols;          // ... of a park warp
    static constexpr int kTmaWarp = PW + kGGradWarps;
    static constexpr int kThreads = 32 * (PW + kGGradWarps + 1 + kGGatherWarps);
    static_assert(kWarpCols * (PW / 4) == kGTmemCols, "TMEM columns");
};

template <int PW>
struct GridSmem {
    unsigned char ring[kGRing][2][kGChunkBytes];
    uint64_t full[kGRing], empty[kGRing];
    uint64_t recbar[kGDepth];                             // park warps: "my record of this unit is written"
    uint64_t finbar[kGDepth];                             // gather warp: "the row statistics of this unit are written"
    uint64_t tfree[PW][kGSlots];                          // gradient warp -> park warp: "this TMEM slot is read"
    float rec[kGDepth][PW][kGRecFloats];                  // warp records of the units in flight
    float fin[kGDepth][2][4];                             // {Ms, Mt, coef/Zs, coef/Zt} of the unit's row of l[0], of l[1]
    float refs[PW][kGSlots][2];                           // references {ms, mt} a parked chunk was taken against
    float klpart[kGGatherWarps][kMaxLosses];              // KL sums of the gather warps
    float klseg[kGGatherWarps][kMaxSegs];                 // ... per pair (launches over several pairs)
#ifdef SD_GRID_TIMING
    long long stamp[kGDepth];                             // clock at which park warp 0 finished the unit
#endif
    uint32_t tmem_base;
};

// waits that normally last a microsecond or more: poll, then sleep - a polling warp takes issue slots from the park warps
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned sleep_ns) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(sleep_ns);
}
__device__ __forceinline__ void ld_relaxed_v2u64(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}


// ---------------------------------------------------------------- units: a coarse region, then a fine one
// The work list is cut into units of p.chunk_elems elements (up to 4 chunks) - few unit boundaries, few packets.
// Where it is short (fewer than 8 rounds of the grid; cabi.cu) its tail - everything after the last whole round - is
// cut into units of p.f_chunk_elems (one chunk): the grid's last round is then shared by all SMs instead of leaving
// most of them idle.  (At 16 rounds the fine tail measured 1 us slower: its units wait for 40 row-mates each.)  The
// regions meet at a row boundary of the larger-group loss (sample p.split_b, row p.split_row of l[0]).
struct GridRegion {
    int chunk_elems, nch_full, nch_last, ups;
};
__device__ __forceinline__ GridRegion grid_region(const RowsParams& p, bool fine) {
    GridRegion g;
    g.chunk_elems = fine ? p.f_chunk_elems : p.chunk_elems;
    g.nch_full = fine ? p.f_nch_full : p.nch_full;
    g.nch_last = fine ? p.f_nch_last : p.nch_last;
    g.ups = fine ? p.f_units_per_sample : p.units_per_sample;
    return g;
}
// first unit (within the sample) of l[0] row j; j == number of rows gives the end
__device__ __forceinline__ int grid_unit_start(const RowsParams& p, const GridRegion& g, int j) {
    return j <= p.G_full ? j * g.nch_full : g.ups;
}
struct GUnit {
    int b, grp, ck, nch;   // sample, row of l[0], unit within the row, units of the row
    int r;                 // unit within the sample
    int e0, len;           // first element within the row, elements
    bool fine;
};
__device__ __forceinline__ GUnit grid_decode(const RowsParams& p, long long u64) {
    // (cabi.cu: fewer than 2^31 units - 32-bit divisions)
    GUnit x;
    const unsigned u = (unsigned)u64, uc = (unsigned)p.units_coarse;
    x.fine = u >= uc;
    const GridRegion g = grid_region(p, x.fine);
    int r;
    if (!x.fine) {
        x.b = (int)(u / (unsigned)g.ups);
        r = (int)(u - (unsigned)x.b * (unsigned)g.ups);
    } else {
        unsigned v = u - uc;
        const int first = grid_unit_start(p, g, p.split_row);   // the fine units of sample split_b start here
        const unsigned part = (unsigned)(g.ups - first);
        if (v < part) {
            x.b = p.split_b;
            r = first + (int)v;
        } else {
            v -= part;
            const unsigned q = v / (unsigned)g.ups;
            x.b = p.split_b + 1 + (int)q;
            r = (int)(v - q * (unsigned)g.ups);
        }
    }
    x.r = r;
    const int full_units = p.G_full * g.nch_full;
    int g_real;
    if (r < full_units) {
        x.grp = g.nch_full == 1 ? r : r / g.nch_full;
        x.ck = r - x.grp * g.nch_full;
        x.nch = g.nch_full;
        g_real = p.l[0].g;
    } else {
        x.grp = p.G_full;
        x.ck = r - full_units;
        x.nch = g.nch_last;
        g_real = p.g_last;
    }
    x.e0 = x.ck * g.chunk_elems;
    x.len = min(g.chunk_elems, g_real * p.HW - x.e0);
    return x;
}
// index in the work list of unit r of sample b, in the region `fine`
__device__ __forceinline__ long long grid_unit_index(const RowsParams& p, const GridRegion& g, bool fine, int b, int r) {
    if (!fine) return (long long)b * g.ups + r;
    const int first = grid_unit_start(p, g, p.split_row);
    if (b == p.split_b) return p.units_coarse + (r - first);
    return p.units_coarse + (g.ups - first) + (long long)(b - p.split_b - 1) * g.ups + r;
}


// ---------------------------------------------------------------- several (student, teacher) pairs, one work list
// MULTI launches (sd_kl_rows_group_fwd_bwd): one loss per pair, every pair cut into coarse units only; the units of
// pair k are [seg[k].unit0, seg[k + 1].unit0).  `seg` = index of the pair.
template <typename Pairs>
__device__ __forceinline__ GUnit grid_decode_multi(const Pairs& gp, long long u64, int& seg) {
    const unsigned u = (unsigned)u64;
    int k = 0;
#pragma unroll
    for (int i = 1; i < kMaxSegs; ++i)
        if (i < gp.nseg && u >= (unsigned)gp.seg[i].unit0) k = i;
    seg = k;
    const GroupSeg& sg = gp.seg[k];
    const unsigned v = u - (unsigned)sg.unit0;
    GUnit x;
    x.fine = false;
    x.b = (int)(v / (unsigned)sg.units_per_sample);
    const int r = (int)(v - (unsigned)x.b * (unsigned)sg.units_per_sample);
    x.r = r;
    const int full_units = sg.G_full * sg.nch_full;
    int g_real;
    if (r < full_units) {
        x.grp = sg.nch_full == 1 ? r : r / sg.nch_full;
        x.ck = r - x.grp * sg.nch_full;
        x.nch = sg.nch_full;
        g_real = sg.g;
    } else {
        x.grp = sg.G_full;
        x.ck = r - full_units;
        x.nch = sg.nch_last;
        g_real = sg.g_last;
    }
    x.e0 = x.ck * sg.chunk_elems;
    x.len = min(sg.chunk_elems, g_real * sg.HW - x.e0);
    return x;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// cycle counters per CTA (scripts/grid_timing.py; build with SD_NVCC_EXTRA=-DSD_GRID_TIMING)
#ifdef SD_GRID_TIMING
#define GT_DECL(n) long long gt_acc[n] = {}
#define GT_TICK(var) const long long var = clock64()
#define GT_ACC(slot, a, b) gt_acc[slot] += (b) - (a)
#define GT_OUT(first, n) do { for (int gi = 0; gi < (n); ++gi) p.dbg[blockIdx.x * 16 + (first) + gi] = (unsigned long long)gt_acc[gi]; } while (0)
#else
#define GT_DECL(n)
#define GT_TICK(var)
#define GT_ACC(slot, a, b)
#define GT_OUT(first, n)
#endif

// R: 0 = independent exponentials per loss, 2 = l[1].tau == 2 * l[0].tau (one ex2 serves both).
// With one loss, or with R == 2, phase 1 parks the EXPONENTIALS (relative to the warp's running maximum at that
// moment, kept in shared memory): phase 2 is then a multiply-add per element, no ex2.  Otherwise the raw values are
// parked and phase 2 recomputes.
// what a launch over one pair passes instead of the pairs' table (never read: every use is behind `if (MULTI)`)
struct NoPairs {
    int nseg;
    const float* grad_out;
    GroupSeg seg[1];
};
template <bool MULTI>
using PairTable = typename std::conditional<MULTI, GroupParams, NoPairs>::type;

template <typename T, int NL, int R, bool MULTI, int PW>
__global__ void __launch_bounds__(GridCfg<PW>::kThreads, 1) kl_rows_grid_kernel(const RowsParams p, const PairTable<MULTI> gp) {
    static_assert(!MULTI || (NL == 1 && R == 0), "several pairs: one loss each");
    using Cfg = GridCfg<PW>;
    constexpr int kGParkWarps = Cfg::kParkWarps, kGPark = Cfg::kPark, kGChunkRows = Cfg::kChunkRows;
    constexpr int kSlotCols = Cfg::kSlotCols, kGTmaWarp = Cfg::kTmaWarp, kGThreads = Cfg::kThreads;
    constexpr int NSRC = PW / kGGradWarps;      // park warps a gradient warp serves
    using ParkedT = typename std::conditional<PW == 8, Parked, Parked16>::type;
    using V = Vec4<T>;
    using vec_t = typename V::type;
    constexpr int VE = 4;
    constexpr int NE = kGChunkRows * VE;        // elements per thread, chunk and tensor
    constexpr bool kParkExp = NL == 1 || R == 2;
    constexpr int K = NL - 1;                   // the loss whose exponential comes out of the MUFU

    extern __shared__ __align__(128) unsigned char smem_raw[];
    GridSmem<PW>& sm = *reinterpret_cast<GridSmem<PW>*>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (p.run_if != nullptr && *p.run_if == 0u) return;  // cancelled backward re-run (uniform over the grid)
#ifdef SD_GRID_TIMING
    if (tid == 0) p.dbg[blockIdx.x * 16 + 11] = global_ns();
#endif

    const int grid = (int)gridDim.x;
    const int n_units = blockIdx.x < p.total_units ? (int)((p.total_units - blockIdx.x + grid - 1) / grid) : 0;

    // a unit of the work list, the pair it belongs to, where it starts in that pair's tensors
    auto unit_at = [&](long long u, int& seg) {
        if (MULTI) return grid_decode_multi(gp, u, seg);
        seg = 0;
        return grid_decode(p, u);
    };
    auto elem_base = [&](const GUnit& x, int seg) {
        if (MULTI) return ((size_t)x.b * gp.seg[seg].C + (size_t)x.grp * gp.seg[seg].g) * gp.seg[seg].HW + (size_t)x.e0;
        return ((size_t)x.b * p.C + (size_t)x.grp * p.l[0].g) * p.HW + (size_t)x.e0;
    };

    // ---- prologue.  The TMA warp sets up the ring's barriers itself and has the first chunks on their way before the
    //      rest of the CTA is set up (barriers of the other hand-offs, tensor-memory allocation).
    struct Stream {                  // the TMA warp's position in the CTA's work list (identical in every lane)
        int slot;
        uint32_t phase;
        int jt, c;                   // unit being streamed and its chunk
        size_t base;                 // first element of that unit
        int nvs;                     // its length in vectors
        const char* pS;              // the pair's tensors
        const char* pT;
    } ts_ = {0, 0u, 0, 0, 0, 0, static_cast<const char*>(p.S), static_cast<const char*>(p.T)};
    const uint64_t pol = l2_policy_evict_first();
    auto stream_unit = [&](int j) {
        int seg;
        const GUnit x = unit_at((long long)blockIdx.x + (long long)j * grid, seg);
        ts_.base = elem_base(x, seg);
        ts_.nvs = x.len / VE;
        if (MULTI) {
            ts_.pS = static_cast<const char*>(gp.seg[seg].S);
            ts_.pT = static_cast<const char*>(gp.seg[seg].T);
        }
    };
    // the next chunk of the work list into ring slot ts_.slot (the caller knows it is free)
    auto stream_chunk = [&]() {
        if (lane == 0) {
            const int nv = min(kGChunkVecs, ts_.nvs - ts_.c * kGChunkVecs);
            const uint32_t bytes = (uint32_t)nv * (uint32_t)sizeof(vec_t);
            mbar_arrive_expect_tx(&sm.full[ts_.slot], 2u * bytes);
            const size_t off = (ts_.base + (size_t)ts_.c * kGChunkVecs * VE) * sizeof(T);
            tma_bulk_g2s(sm.ring[ts_.slot][0], ts_.pS + off, bytes, &sm.full[ts_.slot], pol);
            tma_bulk_g2s(sm.ring[ts_.slot][1], ts_.pT + off, bytes, &sm.full[ts_.slot], pol);
        }
        if (++ts_.slot == kGRing) {
            ts_.slot = 0;
            ts_.phase ^= 1u;
        }
        if (++ts_.c * kGChunkVecs >= ts_.nvs) {
            ts_.c = 0;
            if (++ts_.jt < n_units) stream_unit(ts_.jt);
        }
    };
    if (warp == kGTmaWarp) {
        if (lane == 0) {
            for (int c = 0; c < kGRing; ++c) {
                mbar_init(&sm.full[c], 1);
                mbar_init(&sm.empty[c], kGParkWarps);
            }
            fence_barrier_init();
        }
        __syncwarp();
        if (n_units > 0) stream_unit(0);
        for (int i = 0; i < kGRing && ts_.jt < n_units; ++i) stream_chunk();     // every slot is free
    } else if (tid == 0) {
        for (int q = 0; q < kGDepth; ++q) {
            mbar_init(&sm.recbar[q], kGParkWarps);
            mbar_init(&sm.finbar[q], 1);
        }
        for (int w = 0; w < kGParkWarps; ++w)
            for (int q = 0; q < kGSlots; ++q) mbar_init(&sm.tfree[w][q], 1);
        fence_barrier_init();
    } else if (warp == kGTmaWarp + 1) {
        tmem_alloc(&sm.tmem_base, kGTmemCols);       // (gather warp 0 also frees it)
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;

    float c2[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) c2[k] = MULTI ? gp.seg[0].c2 : p.l[k].c2;      // (MULTI: set per unit)
    const int n_gather = p.grid_knobs[0];          // active gather warps (1 .. kGGatherWarps)
    const unsigned ns_fin = (unsigned)p.grid_knobs[1], ns_tma = (unsigned)p.grid_knobs[2], ns_stat = (unsigned)p.grid_knobs[3];

    if (warp == kGTmaWarp) {
        // =====================================================================================
        // TMA + publisher warp.  Lane 0 streams this CTA's units, chunk by chunk, into the ring as slots drain; the
        // warp turns the 8 warp records of a parked unit into its packet in global memory as soon as they are
        // written.  Neither job ever waits for the other (or for another CTA): both are polled.
        // =====================================================================================
        // tag of this launch's packets: the workspace's launch counter (bumped by the last CTA to finish, i.e. after
        // every CTA has read it) + 1, so a packet left by any earlier launch never validates
        const unsigned long long tag = (unsigned long long)(__ldcg(&p.ctrl[2]) + 1u) << 32;
        int jp = 0;                 // unit to publish next
        while (jp < n_units) {
            int ev = 0;
            if (lane == 0) {
                if (ts_.jt < n_units && mbar_test_wait(&sm.empty[ts_.slot], ts_.phase ^ 1u)) ev |= 1;
                if (mbar_test_wait(&sm.recbar[jp & (kGDepth - 1)], (uint32_t)(jp >> 3) & 1u)) ev |= 2;
            }
            ev = __shfl_sync(0xffffffffu, ev, 0);
            if (ev & 1) stream_chunk();
            if (ev & 2) {
                if (MULTI) {
                    int seg;
                    unit_at((long long)blockIdx.x + (long long)jp * grid, seg);
                    c2[0] = gp.seg[seg].c2;
                }
                const float4* q = reinterpret_cast<const float4*>(sm.rec[jp & (kGDepth - 1)][lane & (PW - 1)]);
                PStat<NL> st = pstat_from<NL>(q[0], q[1], NL == 2 ? q[2] : make_float4(0.f, 0.f, 0.f, 0.f));
                st = pstat_reduce<NL, R, PW>(st, c2);
                float val = st.ms;
                if (lane == 1) val = st.mt;
                if (lane == 2) val = st.zs[0];
                if (lane == 3) val = st.zt[0];
                if (lane == 4) val = st.a[0];
                if (lane == 5) val = st.dd[0];
                if (NL == 2) {
                    if (lane == 6) val = st.zs[K];
                    if (lane == 7) val = st.zt[K];
                    if (lane == 8) val = st.a[K];
                    if (lane == 9) val = st.dd[K];
                }
                if (lane < 2 + 4 * NL)
                    st_relaxed_u64(p.pkt + ((size_t)blockIdx.x + (size_t)jp * grid) * kPktWords + lane,
                                   tag | (unsigned long long)__float_as_uint(val));
                ++jp;
            }
            if (ev == 0) __nanosleep(ns_tma);
        }
        bar_sync(3, kGThreads);
        return;
    }

    if (warp > kGTmaWarp) {
        // =====================================================================================
        // gather warps: packets of a unit's row-mates (all CTAs) -> row statistics for the gradient warps; warp i takes
        // the units j % kGGatherWarps == i of this CTA
        // =====================================================================================
        const int gw = warp - kGTmaWarp - 1;
        const unsigned epoch = __ldcg(&p.ctrl[2]) + 1u;
        float kl_acc[NL];
        float coef[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            kl_acc[k] = 0.f;
            coef[k] = p.l[k].coef;
            if (p.grad_out[k] != nullptr) coef[k] *= __ldg(p.grad_out[k]);
        }
        float gscale = 1.f;        // MULTI: d(total)/d(losses), one scalar for the launch
        if (MULTI) {
            if (gp.grad_out != nullptr) gscale = __ldg(gp.grad_out);
            if (lane < kMaxSegs) sm.klseg[gw][lane] = 0.f;
            __syncwarp();
        }
        GT_DECL(6);
        for (int j = gw < n_gather ? gw : n_units; j < n_units; j += n_gather) {
            // ---- the packets of the row-mates of unit j: units [rowu, rowu + rown); its row of l[0] is the mates
            //      [i0, i0 + n0)
            const long long u = (long long)blockIdx.x + (long long)j * grid;
            int seg;
            const GUnit x = unit_at(u, seg);
            const GridRegion rg = grid_region(p, x.fine);
            if (MULTI) {
                c2[0] = gp.seg[seg].c2;
                coef[0] = gp.seg[seg].coef * gscale;
            }
            long long rowu;
            int rown, i0, n0, rk = 0;
            if (NL == 2) {
                const int m = p.l[K].m;
                rk = x.grp / m;
                const int j0 = rk * m, j1 = min(j0 + m, p.l[0].G);
                const int us = grid_unit_start(p, rg, j0);
                rown = grid_unit_start(p, rg, j1) - us;
                rowu = grid_unit_index(p, rg, x.fine, x.b, us);
                i0 = (int)(u - x.ck - rowu);
                n0 = x.nch;
            } else {
                rowu = u - x.ck;
                rown = x.nch;
                i0 = 0;
                n0 = rown;
            }
            // (my own packet is among them: nothing can be complete before my CTA's park warps are through the unit)
            GT_TICK(g0);
            mbar_wait_backoff(&sm.recbar[j & (kGDepth - 1)], (uint32_t)(j >> 3) & 1u, ns_stat);
            GT_TICK(g1);
            GT_ACC(0, g0, g1);
            PStat<NL> mate[2];
            unsigned spins = 0;
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int idx = lane + 32 * h;
                    mate[h] = pstat_empty<NL>();
                    if (idx < rown) {
                        const unsigned long long* q = p.pkt + (size_t)(rowu + idx) * kPktWords;
                        unsigned long long w[2 + 4 * NL];
#pragma unroll
                        for (int i = 0; i < 1 + 2 * NL; ++i) ld_relaxed_v2u64(q + 2 * i, w[2 * i], w[2 * i + 1]);
#pragma unroll
                        for (int i = 0; i < 2 + 4 * NL; ++i) ok = ok && (unsigned)(w[i] >> 32) == epoch;
                        mate[h].ms = __uint_as_float((unsigned)w[0]);
                        mate[h].mt = __uint_as_float((unsigned)w[1]);
                        mate[h].zs[0] = __uint_as_float((unsigned)w[2]);
                        mate[h].zt[0] = __uint_as_float((unsigned)w[3]);
                        mate[h].a[0] = __uint_as_float((unsigned)w[4]);
                        mate[h].dd[0] = __uint_as_float((unsigned)w[5]);
                        if (NL == 2) {
                            mate[h].zs[K] = __uint_as_float((unsigned)w[2 + 4 * K]);
                            mate[h].zt[K] = __uint_as_float((unsigned)w[3 + 4 * K]);
                            mate[h].a[K] = __uint_as_float((unsigned)w[4 + 4 * K]);
                            mate[h].dd[K] = __uint_as_float((unsigned)w[5 + 4 * K]);
                        }
                    }
                }
                if (__all_sync(0xffffffffu, ok)) break;
                if (++spins > kSpinLimit) {
                    // never expected (the launch is cooperative: every CTA is resident).  The flag makes the last CTA
                    // report NaN losses instead of numbers built on a stale packet.
                    if (lane == 0) atomicExch(&p.ctrl[1], 1u);
                    break;
                }
                __nanosleep(ns_stat);
            }
            GT_TICK(g2);
            GT_ACC(1, g1, g2);
            const int d = j & (kGDepth - 1);
            // ---- on the critical path (the gradient warps wait for it): maxima and the sums Zs, Zt of the unit's rows,
            //      nothing else - plain (max, sum exp) merges, one redux / butterfly each
            const bool in0[2] = {NL == 1 || (lane >= i0 && lane < i0 + n0), NL == 1 || (lane + 32 >= i0 && lane + 32 < i0 + n0)};
            {
                const float MsK = warp_max_uniform(fmaxf(mate[0].ms, mate[1].ms));
                const float MtK = warp_max_uniform(fmaxf(mate[0].mt, mate[1].mt));
                float zsK = mate[0].zs[K] * ref_factor(mate[0].ms, MsK, c2[K]);
                float ztK = mate[0].zt[K] * ref_factor(mate[0].mt, MtK, c2[K]);
                if (rown > 32) {
                    zsK = fmaf(mate[1].zs[K], ref_factor(mate[1].ms, MsK, c2[K]), zsK);
                    ztK = fmaf(mate[1].zt[K], ref_factor(mate[1].mt, MtK, c2[K]), ztK);
                }
                float Ms0 = MsK, Mt0 = MtK, zs0 = 0.f, zt0 = 0.f;
                if (NL == 2) {
                    Ms0 = warp_max_uniform(fmaxf(in0[0] ? mate[0].ms : kMaxFloor, in0[1] ? mate[1].ms : kMaxFloor));
                    Mt0 = warp_max_uniform(fmaxf(in0[0] ? mate[0].mt : kMaxFloor, in0[1] ? mate[1].mt : kMaxFloor));
                    zs0 = in0[0] ? mate[0].zs[0] * ref_factor(mate[0].ms, Ms0, c2[0]) : 0.f;
                    zt0 = in0[0] ? mate[0].zt[0] * ref_factor(mate[0].mt, Mt0, c2[0]) : 0.f;
                    if (rown > 32 && in0[1]) {
                        zs0 = fmaf(mate[1].zs[0], ref_factor(mate[1].ms, Ms0, c2[0]), zs0);
                        zt0 = fmaf(mate[1].zt[0], ref_factor(mate[1].mt, Mt0, c2[0]), zt0);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    zsK += __shfl_xor_sync(0xffffffffu, zsK, o);
                    ztK += __shfl_xor_sync(0xffffffffu, ztK, o);
                    if (NL == 2) {
                        zs0 += __shfl_xor_sync(0xffffffffu, zs0, o);
                        zt0 += __shfl_xor_sync(0xffffffffu, zt0, o);
                    }
                }
                if (lane == 0) {
                    if (NL == 2) {
                        *reinterpret_cast<float4*>(sm.fin[d][0]) = make_float4(Ms0, Mt0, __fdividef(coef[0], zs0), __fdividef(coef[0], zt0));
                        *reinterpret_cast<float4*>(sm.fin[d][1]) = make_float4(MsK, MtK, __fdividef(coef[K], zsK), __fdividef(coef[K], ztK));
                    } else {
                        *reinterpret_cast<float4*>(sm.fin[d][0]) = make_float4(MsK, MtK, __fdividef(coef[0], zsK), __fdividef(coef[0], ztK));
                    }
                }
                __syncwarp();
#ifdef SD_GRID_TIMING
                gt_acc[4] += clock64() - sm.stamp[d];     // park warp 0 through the unit -> row statistics ready
#endif
                if (lane == 0) mbar_arrive(&sm.finbar[d]);
            }
            GT_TICK(g3);
            GT_ACC(2, g2, g3);
            // ---- off the critical path: the KL terms of the rows whose first unit is mine, from the complete statistics
            //      (compensated merges of the a and dd sums: common.cuh)
            const bool own0 = x.ck == 0, ownK = NL == 2 && u == rowu;
            if (own0 || ownK) {
                PStat<NL> pr, sr = pstat_empty<NL>();
                if (NL == 2) {
                    PStat<NL> all = mate[0], mine[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) mine[h] = in0[h] ? mate[h] : pstat_empty<NL>();
                    if (rown > 32) {
                        all = pstat_merge<NL, R>(mate[0], mate[1], c2);
                        mine[0] = pstat_merge<NL, R>(mine[0], mine[1], c2);
                    }
                    if (ownK) sr = pstat_reduce<NL, R, 32>(all, c2);
                    pr = pstat_reduce<NL, R, 32>(mine[0], c2);
                } else {
                    PStat<NL> all = mate[0];
                    if (rown > 32) all = pstat_merge<NL, R>(mate[0], mate[1], c2);
                    pr = pstat_reduce<NL, R, 32>(all, c2);
                }
                if (lane == 0) {
                    if (own0) {
                        const float kl = kl_of_row(NL == 2 && R == 2 ? 2.f : 1.f, pr.zs[0], pr.zt[0], pr.a[0], pr.dd[0]);
                        if (MULTI) {
                            if (gp.seg[seg].row_kl) gp.seg[seg].row_kl[x.b * gp.seg[seg].G + x.grp] = kl;
                            sm.klseg[gw][seg] += kl;
                        } else {
                            if (p.l[0].row_kl) p.l[0].row_kl[x.b * p.l[0].G + x.grp] = kl;
                            kl_acc[0] += kl;
                        }
                    }
                    if (ownK) {
                        const float kl = kl_of_row(1.f, sr.zs[K], sr.zt[K], sr.a[K], sr.dd[K]);
                        if (p.l[K].row_kl) p.l[K].row_kl[x.b * p.l[K].G + rk] = kl;
                        kl_acc[K] += kl;
                    }
                }
            }
            GT_TICK(g4);
            GT_ACC(3, g3, g4);
#ifdef SD_GRID_TIMING
            gt_acc[5] += 1;
#endif
        }
#ifdef SD_GRID_TIMING
        if (gw == 0 && lane == 0) GT_OUT(5, 6);
#endif
        // ---- loss: this CTA's partial = the gather warps' sums in warp order; the last CTA sums the partials in a fixed
        //      order.  All of it happens while the gradient warps are still busy with the last unit.
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NL; ++k) sm.klpart[gw][k] = kl_acc[k];
        }
        bar_sync(4, 32 * kGGatherWarps);        // every gather warp's partial is written
        if (gw != 0) {
            bar_sync(3, kGThreads);
            return;
        }
        unsigned ticket = 0;
        if (MULTI && lane < gp.nseg) {
            float acc = sm.klseg[0][lane];
#pragma unroll
            for (int w = 1; w < kGGatherWarps; ++w) acc += sm.klseg[w][lane];
            __stcg(&p.cta_part[lane * kMaxGrid + blockIdx.x], acc);
        }
        if (MULTI) {
            __threadfence();
            __syncwarp();
        }
        if (lane == 0) {
            if (!MULTI) {
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    float acc = sm.klpart[0][k];
#pragma unroll
                    for (int w = 1; w < kGGatherWarps; ++w) acc += sm.klpart[w][k];
                    __stcg(&p.cta_part[k * kMaxGrid + blockIdx.x], acc);
                }
            }
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (MULTI && ticket == gridDim.x - 1) {
            // the last CTA sums every pair's partials in a fixed order
            __threadfence();
            const bool timed_out = __ldcg(&p.ctrl[1]) != 0u;
            for (int k = 0; k < gp.nseg; ++k) {
                double acc = 0.0;
                for (int i = lane; i < (int)gridDim.x; i += 32) acc += (double)__ldcg(&p.cta_part[k * kMaxGrid + i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
                if (lane == 0)
                    *gp.seg[k].loss = timed_out ? __int_as_float(0x7fc00000) : (float)((double)gp.seg[k].loss_scale * acc);
            }
            if (lane == 0) {
                atomicAdd(&p.ctrl[2], 1u);
                atomicExch(&p.ctrl[0], 0u);
            }
        } else if (ticket == gridDim.x - 1) {
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
                const bool timed_out = __ldcg(&p.ctrl[1]) != 0u;
#pragma unroll
                for (int k = 0; k < NL; ++k)
                    *p.l[k].loss = timed_out ? __int_as_float(0x7fc00000) : (float)((double)p.l[k].loss_scale * acc[k]);
                atomicAdd(&p.ctrl[2], 1u);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
        // every consumer is through with TMEM: this warp allocated it, this warp frees it
        bar_sync(3, kGThreads);
        tmem_dealloc(tmem_base, kGTmemCols);
#ifdef SD_GRID_TIMING
        if (lane == 0) p.dbg[blockIdx.x * 16 + 15] = global_ns();
#endif
        return;
    }

    // =========================================================================================
    // park warps 0..7 and gradient warps 8..15: warp 8 + i retrieves what warp i parked
    // =========================================================================================
    // TMEM window of park warp w: the lane quarter of the warp (w & 3 - a warp can only reach its own quarter),
    // Cfg::kWarpCols columns from (w >> 2) * Cfg::kWarpCols: 8 chunk slots of kSlotCols.  A gradient warp reads the
    // windows of the NSRC park warps of its quarter it is paired with.
    auto tmem_of = [&](int w) {
        return tmem_base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * Cfg::kWarpCols);
    };
    const int pw = warp;               // (park warps) my index
    const int ptid = tid;              // (park warps) my index among the park threads
    const uint32_t tmem_mine = tmem_of(warp);
    const int wv = pw * 32;            // (park warps) my vector-rows of a chunk start here, + kGPark r

    if (warp < kGParkWarps) {
        // ------------------------------------------------ phase 1: statistics, park the unit
        PStat<NL> st = pstat_empty<NL>();  // statistics of the unit so far: maxima identical in every lane, sums per lane
        uint32_t q = 0;                    // chunks parked so far -> TMEM slot
        GT_DECL(8);
        GT_TICK(gt_start);
        int slot = 0;
        uint32_t phase = 0;

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
        // N (even) elements into the statistics of the unit; what gets parked replaces them in fs / ft.  Two elements
        // per instruction (FFMA2 / FADD2 / FMUL2, common.cuh): every sum runs as two partial sums per lane.
        auto accumulate = [&](float* fs, float* ft, int n) {
            const F2 neg1 = f2_dup(-1.f);
            F2 C2[NL], NRS[NL], NRT[NL], ZS[NL], ZT[NL], A[NL], DD[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                C2[k] = f2_dup(c2[k]);
                NRS[k] = f2_dup(-__fmul_rn(st.ms, c2[k]));
                NRT[k] = f2_dup(-__fmul_rn(st.mt, c2[k]));
                ZS[k] = f2_make(st.zs[k], 0.f);
                ZT[k] = f2_make(st.zt[k], 0.f);
                A[k] = f2_make(st.a[k], 0.f);
                DD[k] = f2_make(st.dd[k], 0.f);
            }
            auto ex2x2 = [](const F2& x) {
                float lo, hi;
                f2_split(x, lo, hi);
                return f2_make(fast_exp2(lo), fast_exp2(hi));
            };
#pragma unroll
            for (int i = 0; i < NE; i += 2) {
                if (i < n) {
                    const F2 s2 = f2_make(fs[i], fs[i + 1]), t2 = f2_make(ft[i], ft[i + 1]);
                    if (NL == 2 && R == 2) {
                        // e0 = eK^2: the sums of loss 0 straight from the loss-K exponentials (one FFMA2 each)
                        const F2 as2 = f2_fma(s2, C2[K], NRS[K]), at2 = f2_fma(t2, C2[K], NRT[K]);
                        const F2 es2 = ex2x2(as2), et2 = ex2x2(at2);
                        const F2 da = f2_fma(as2, neg1, at2), dk = f2_fma(es2, neg1, et2);
                        ZS[K] = f2_add(ZS[K], es2);
                        ZT[K] = f2_add(ZT[K], et2);
                        ZS[0] = f2_fma(es2, es2, ZS[0]);
                        ZT[0] = f2_fma(et2, et2, ZT[0]);
                        const F2 w = f2_mul(et2, da);
                        A[K] = f2_add(A[K], w);
                        A[0] = f2_fma(et2, w, A[0]);
                        DD[K] = f2_add(DD[K], dk);
                        DD[0] = f2_fma(dk, f2_add(et2, es2), DD[0]);     // et0 - es0 = (eK_t - eK_s)(eK_t + eK_s)
                        f2_split(es2, fs[i], fs[i + 1]);
                        f2_split(et2, ft[i], ft[i + 1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < NL; ++k) {
                            const F2 as2 = f2_fma(s2, C2[k], NRS[k]), at2 = f2_fma(t2, C2[k], NRT[k]);
                            const F2 es2 = ex2x2(as2), et2 = ex2x2(at2);
                            ZS[k] = f2_add(ZS[k], es2);
                            ZT[k] = f2_add(ZT[k], et2);
                            A[k] = f2_fma(et2, f2_fma(as2, neg1, at2), A[k]);
                            DD[k] = f2_add(DD[k], f2_fma(es2, neg1, et2));
                            if (kParkExp && k == K) {
                                f2_split(es2, fs[i], fs[i + 1]);
                                f2_split(et2, ft[i], ft[i + 1]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                st.zs[k] = f2_sum(ZS[k]);
                st.zt[k] = f2_sum(ZT[k]);
                st.a[k] = f2_sum(A[k]);
                st.dd[k] = f2_sum(DD[k]);
            }
        };

        for (int j = 0; j < n_units; ++j) {
            GT_TICK(d0);
            int seg;
            const GUnit x = unit_at((long long)blockIdx.x + (long long)j * grid, seg);
            if (MULTI) c2[0] = gp.seg[seg].c2;
            const int nvs = x.len / VE;
            const int nchunks = (nvs + kGChunkVecs - 1) / kGChunkVecs;
            st = pstat_empty<NL>();
            GT_TICK(d1);
            GT_ACC(3, d0, d1);
            for (int c = 0; c < nchunks; ++c, ++q) {
                // ---- one chunk: ring -> registers -> statistics -> TMEM
                const uint32_t ts = q & (uint32_t)(kGSlots - 1);
                if (q >= (uint32_t)kGSlots) {
                    // the slot's previous chunk has been retrieved by my gradient warp
                    GT_TICK(w0);
                    mbar_wait(&sm.tfree[pw][ts], ((q >> 3) - 1u) & 1u);
                    GT_TICK(w1);
                    GT_ACC(0, w0, w1);
                    tmem_fence_after_sync();
                }
                GT_TICK(w2);
                mbar_wait(&sm.full[slot], phase);
                GT_TICK(w3);
                GT_ACC(1, w2, w3);
#ifdef SD_GRID_TIMING
                if (tid == 0 && q == 0) p.dbg[blockIdx.x * 16 + 12] = global_ns();
#endif
                const vec_t* bs = reinterpret_cast<const vec_t*>(sm.ring[slot][0]);
                const vec_t* bt = reinterpret_cast<const vec_t*>(sm.ring[slot][1]);
                float fs[NE], ft[NE];
#pragma unroll
                for (int r = 0; r < kGChunkRows; ++r) {
                    V::unpack(bs[r * kGPark + ptid], &fs[r * VE]);
                    V::unpack(bt[r * kGPark + ptid], &ft[r * VE]);
                }
                float mxs[kGChunkRows], mxt[kGChunkRows];
#pragma unroll
                for (int r = 0; r < kGChunkRows; ++r) {
                    mxs[r] = fmaxf(fmaxf(fs[r * VE], fs[r * VE + 1]), fmaxf(fs[r * VE + 2], fs[r * VE + 3]));
                    mxt[r] = fmaxf(fmaxf(ft[r * VE], ft[r * VE + 1]), fmaxf(ft[r * VE + 2], ft[r * VE + 3]));
                }
                const int vA = c * kGChunkVecs + wv;      // my first vector-row, in vectors of the unit; + 256 r
                if (vA + (kGChunkRows - 1) * kGPark < nvs) {
                    // ---- all my vector-rows lie inside the unit (the common case)
                    float ms_ = mxs[0], mt_ = mxt[0];
#pragma unroll
                    for (int r = 1; r < kGChunkRows; ++r) {
                        ms_ = fmaxf(ms_, mxs[r]);
                        mt_ = fmaxf(mt_, mxt[r]);
                    }
                    const float wms = warp_max_uniform(ms_);
                    const float wmt = warp_max_uniform(mt_);
                    // every lane's shared-memory reads went into the maxima: the slot may go back to the TMA warp
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);
                    GT_TICK(r0);
                    raise_refs(wms, wmt);
                    GT_TICK(r1);
                    GT_ACC(4, r0, r1);
                    accumulate(fs, ft, NE);
                    GT_TICK(r2);
                    GT_ACC(5, r1, r2);
                } else {
                    // ---- the unit ends in this chunk: only the vector-rows inside it count
                    float ms_ = kMaxFloor, mt_ = kMaxFloor;
#pragma unroll
                    for (int r = 0; r < kGChunkRows; ++r) {
                        if (vA + r * kGPark < nvs) {
                            ms_ = fmaxf(ms_, mxs[r]);
                            mt_ = fmaxf(mt_, mxt[r]);
                        }
                    }
                    const float wms = warp_max_uniform(ms_);
                    const float wmt = warp_max_uniform(mt_);
                    if (lane == 0) mbar_arrive(&sm.empty[slot]);
                    raise_refs(wms, wmt);
#pragma unroll
                    for (int r = 0; r < kGChunkRows; ++r)
                        if (vA + r * kGPark < nvs) accumulate(&fs[r * VE], &ft[r * VE], VE);
                }
                // ---- park: registers -> TMEM; the references once per warp
                ParkedT pk;
#pragma unroll
                for (int r = 0; r < kGChunkRows; ++r) {
#pragma unroll
                    for (int e = 0; e < VE; ++e) {
                        pk.w[r * 8 + e] = __float_as_uint(fs[r * VE + e]);
                        pk.w[r * 8 + 4 + e] = __float_as_uint(ft[r * VE + e]);
                    }
                }
                GT_TICK(s0);
                if constexpr (PW == 8) tmem_st32(tmem_mine + ts * kSlotCols, pk);
                else tmem_st16(tmem_mine + ts * kSlotCols, pk);
                if (kParkExp && lane == 0) *reinterpret_cast<float2*>(sm.refs[pw][ts]) = make_float2(st.ms, st.mt);
                GT_TICK(s1);
                GT_ACC(6, s0, s1);
                if (++slot == kGRing) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
            // ---- the unit is parked: the warp's record (maxima + 4 sums per loss, transposed butterfly); what I
            //      stored in TMEM is complete and ordered before the arrival the gradient warps (transitively) wait on
            GT_TICK(c0);
            {
                float v[8] = {st.zs[0], st.zt[0], st.a[0], st.dd[0], 0.f, 0.f, 0.f, 0.f};
                if (NL == 2) {
                    v[4] = st.zs[NL - 1];
                    v[5] = st.zt[NL - 1];
                    v[6] = st.a[NL - 1];
                    v[7] = st.dd[NL - 1];
                }
                const float tot = warp_sum8_transposed(v, lane);
                float* rec = sm.rec[j & (kGDepth - 1)][pw];
                if ((lane & 3) == 0 && lane < 4 * 4 * NL) rec[2 + (lane >> 2)] = tot;
                if (lane == 1) rec[0] = st.ms;
                if (lane == 2) rec[1] = st.mt;
            }
            tmem_wait_st();
            tmem_fence_before_sync();
            __syncwarp();
#ifdef SD_GRID_TIMING
            if (tid == 0) sm.stamp[j & (kGDepth - 1)] = clock64();
            gt_acc[7] += clock64() - c0;
#endif
            if (lane == 0) mbar_arrive(&sm.recbar[j & (kGDepth - 1)]);
        }
#ifdef SD_GRID_TIMING
        if (tid == 0) {
            gt_acc[2] = clock64() - gt_start;
            GT_OUT(0, 3);
            p.dbg[blockIdx.x * 16 + 13] = global_ns();
        }
#endif
    } else {
        // ------------------------------------------------ phase 2: gradient from the parked unit
        // gradient warp g reads the park warps src0 + 4 s (s < NSRC) of its lane quarter: consecutive TMEM windows
        const int gq = warp - kGParkWarps;
        const int src0 = (gq & 3) + 4 * (NSRC * (gq >> 2));
        const uint32_t tmem_src0 = tmem_of(src0);
        uint32_t q = 0;                    // chunks retrieved so far -> TMEM slot
        GT_DECL(2);
        GT_TICK(gt_start);
        for (int j = 0; j < n_units; ++j) {
            int seg;
            const GUnit x = unit_at((long long)blockIdx.x + (long long)j * grid, seg);
            if (MULTI) c2[0] = gp.seg[seg].c2;
            const int nvs = x.len / VE;
            const int n = (nvs + kGChunkVecs - 1) / kGChunkVecs;
            const int d = j & (kGDepth - 1);
            GT_TICK(w0);
            mbar_wait_backoff(&sm.finbar[d], (uint32_t)(j >> 3) & 1u, ns_fin);
            GT_TICK(w1);
            GT_ACC(0, w0, w1);
            tmem_fence_after_sync();
            T* out = static_cast<T*>(MULTI ? gp.seg[seg].dS : p.dS) + elem_base(x, seg);
            const float4 fa = *reinterpret_cast<const float4*>(sm.fin[d][0]);                 // row of l[0]
            const float4 fb = NL == 2 ? *reinterpret_cast<const float4*>(sm.fin[d][1]) : fa;   // row of l[1]
            // one parked chunk of park warp `src`: values (+ the references they were taken against) -> gradient
            auto grad_chunk = [&](int c, const ParkedT& pk, uint32_t ts, int src) {
                float2 rf = make_float2(0.f, 0.f);
                if (kParkExp) rf = *reinterpret_cast<const float2*>(sm.refs[src][ts]);
                // the parked values are in registers, the references too: the slot may be parked into again
                __syncwarp();
                tmem_fence_before_sync();
                if (lane == 0) mbar_arrive(&sm.tfree[src][ts]);
                const int vA = c * kGChunkVecs + src * 32;
                float gsK = 0.f, gtK = 0.f, gs0 = 0.f, gt0 = 0.f, rs0 = 0.f, rt0 = 0.f, rs1 = 0.f, rt1 = 0.f;
                if (kParkExp) {
                    // parked: e = exp2((x - ref) c2[K]); softmax_k = e^(c2[k]/c2[K]) * exp2((ref - M_k) c2[k]) / Z_k, ref <= M_k
                    gsK = fb.z * ref_factor(rf.x, fb.x, c2[K]);
                    gtK = fb.w * ref_factor(rf.y, fb.y, c2[K]);
                    if (NL == 2) {
                        gs0 = fa.z * ref_factor(rf.x, fa.x, c2[0]);
                        gt0 = fa.w * ref_factor(rf.y, fa.y, c2[0]);
                    }
                } else {
                    // raw values parked: recompute against the row maxima
                    rs0 = __fmul_rn(fa.x, c2[0]);
                    rt0 = __fmul_rn(fa.y, c2[0]);
                    rs1 = __fmul_rn(fb.x, c2[K]);
                    rt1 = __fmul_rn(fb.y, c2[K]);
                }
                const F2 GSK = f2_dup(gsK), NGTK = f2_dup(-gtK), GS0 = f2_dup(gs0), NGT0 = f2_dup(-gt0);
#pragma unroll
                for (int r = 0; r < kGChunkRows; ++r) {
                    if (vA + r * kGPark < nvs) {
                        float o[VE];
#pragma unroll
                        for (int e = 0; e < VE; e += 2) {
                            const float vs0 = __uint_as_float(pk.w[r * 8 + e]), vs1 = __uint_as_float(pk.w[r * 8 + e + 1]);
                            const float vt0 = __uint_as_float(pk.w[r * 8 + 4 + e]), vt1 = __uint_as_float(pk.w[r * 8 + 4 + e + 1]);
                            if (kParkExp && NL == 2) {
                                // vs (vs gs0 + gsK) - vt (vt gt0 + gtK), two elements per instruction
                                const F2 vs = f2_make(vs0, vs1), vt = f2_make(vt0, vt1);
                                const F2 u = f2_mul(vt, f2_fma(vt, NGT0, NGTK));
                                f2_split(f2_fma(vs, f2_fma(vs, GS0, GSK), u), o[e], o[e + 1]);
                            } else if (kParkExp) {
                                const F2 vs = f2_make(vs0, vs1), vt = f2_make(vt0, vt1);
                                f2_split(f2_fma(vs, GSK, f2_mul(vt, NGTK)), o[e], o[e + 1]);
                            } else {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const float vs = h ? vs1 : vs0, vt = h ? vt1 : vt0;
                                    const float es0 = fast_exp2(fmaf(vs, c2[0], -rs0));
                                    const float et0 = fast_exp2(fmaf(vt, c2[0], -rt0));
                                    const float es1 = fast_exp2(fmaf(vs, c2[K], -rs1));
                                    const float et1 = fast_exp2(fmaf(vt, c2[K], -rt1));
                                    o[e + h] = fmaf(es0, fa.z, es1 * fb.z) - fmaf(et0, fa.w, et1 * fb.w);
                                }
                            }
                        }
                        V::store(out + (size_t)(vA + r * kGPark + lane) * VE, o);
                    }
                }
            };
            if constexpr (PW == 8) {
                // parked chunks in flight from TMEM, one ahead of the arithmetic, in two register sets
                Parked pa, pb;
                tmem_ld32(tmem_src0 + (q & (uint32_t)(kGSlots - 1)) * kSlotCols, pa);
                for (int c = 0; c < n; c += 2) {
                    const uint32_t t0_ = (q + (uint32_t)c) & (uint32_t)(kGSlots - 1), t1_ = (t0_ + 1) & (uint32_t)(kGSlots - 1),
                                   t2_ = (t0_ + 2) & (uint32_t)(kGSlots - 1);
                    tmem_wait_ld(pa);
                    if (c + 1 < n) tmem_ld32(tmem_src0 + t1_ * kSlotCols, pb);
                    grad_chunk(c, pa, t0_, src0);
                    if (c + 1 < n) {
                        tmem_wait_ld(pb);
                        if (c + 2 < n) tmem_ld32(tmem_src0 + t2_ * kSlotCols, pa);
                        grad_chunk(c + 1, pb, t1_, src0);
                    }
                }
            } else {
                // two park warps per gradient warp: both halves of a chunk in one round trip to TMEM
                for (int c = 0; c < n; ++c) {
                    const uint32_t ts = (q + (uint32_t)c) & (uint32_t)(kGSlots - 1);
                    Parked16 pa, pb;
                    tmem_ld16(tmem_src0 + ts * kSlotCols, pa);
                    tmem_ld16(tmem_src0 + Cfg::kWarpCols + ts * kSlotCols, pb);
                    tmem_wait_ld(pa, pb);
                    grad_chunk(c, pa, ts, src0);
                    grad_chunk(c, pb, ts, src0 + 4);
                }
            }
            q += (uint32_t)n;
        }
#ifdef SD_GRID_TIMING
        if (tid == kGPark) {
            gt_acc[1] = clock64() - gt_start;
            GT_OUT(3, 2);
            p.dbg[blockIdx.x * 16 + 14] = global_ns();
        }
#endif
    }
    // every consumer is through with TMEM (gather warp 0 frees it)
    bar_sync(3, kGThreads);
}

// ====================================================================================================
template <typename T, int NL, int R, bool MULTI, int PW>
static cudaError_t launch_grid_pw(const RowsParams& p, const PairTable<MULTI>& gp, int sms, cudaStream_t stream, bool probe_only) {
    auto kern = kl_rows_grid_kernel<T, NL, R, MULTI, PW>;
    constexpr int kGThreads = GridCfg<PW>::kThreads;
    constexpr size_t kGridSmemBytes = sizeof(GridSmem<PW>);
    static std::atomic<int> ctas_per_sm_dev[kMaxDevices];  // per instantiation and device; -1: cannot run
    std::atomic<int>& ctas_per_sm = ctas_per_sm_dev[device_slot()];
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGridSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kGThreads, kGridSmemBytes);
        if (e != cudaSuccess) return e;
        ctas_per_sm = n >= 1 ? 1 : -1;
    }
    if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (probe_only) return cudaSuccess;
    long long grid = sms;                      // one CTA per SM: each allocates all of its SM's tensor memory
    if (grid > p.total_units) grid = p.total_units;
    if (grid > kMaxGrid) grid = kMaxGrid;
    if (p.max_row_units > grid) return cudaErrorInvalidConfiguration;   // (cabi.cu checks against the SM count first)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGThreads);
    cfg.dynamicSmemBytes = kGridSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // every CTA must be resident: they read each other's packets
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p, gp);
}

// Park warps per CTA: 8.  The 16-warp variant (8 element pairs per thread and chunk, 72 registers, four park warps per
// scheduler) is correct and measured no faster - bf16 fused 83.5 vs 80.3 us, fp32 the same: with 8 park warps busy 92 %
// of the time the SM is bound by its instruction throughput, not by latency.  -DSD_GRID_PW16 builds it in
// (SEGDISTILL_GRID_KNOBS=4099,... selects it: bits 8.. of the first knob).
template <typename T, int NL, int R, bool MULTI>
static cudaError_t launch_grid_t(const RowsParams& p, const PairTable<MULTI>& gp, int sms, cudaStream_t stream, bool probe_only) {
    RowsParams q = p;
    q.grid_knobs[0] &= 255;
#ifdef SD_GRID_PW16
    if ((p.grid_knobs[0] >> 8) == 16) return launch_grid_pw<T, NL, R, MULTI, 16>(q, gp, sms, stream, probe_only);
#endif
    return launch_grid_pw<T, NL, R, MULTI, 8>(q, gp, sms, stream, probe_only);
}

cudaError_t launch_kl_rows_grid(const RowsParams& p, bool bf16, int sms, cudaStream_t stream, bool probe_only) {
    static const NoPairs none = {};
    if (p.nl == 2) {
        // tau[1] == 2 * tau[0]: one exponential serves both losses
        const bool sq = p.l[0].c2 == 2.f * p.l[1].c2;
        if (sq)
            return bf16 ? launch_grid_t<__nv_bfloat16, 2, 2, false>(p, none, sms, stream, probe_only)
                        : launch_grid_t<float, 2, 2, false>(p, none, sms, stream, probe_only);
        return bf16 ? launch_grid_t<__nv_bfloat16, 2, 0, false>(p, none, sms, stream, probe_only)
                    : launch_grid_t<float, 2, 0, false>(p, none, sms, stream, probe_only);
    }
    return bf16 ? launch_grid_t<__nv_bfloat16, 1, 0, false>(p, none, sms, stream, probe_only)
                : launch_grid_t<float, 1, 0, false>(p, none, sms, stream, probe_only);
}

// several pairs, one loss each: gp.seg[] cut into units of whole chunks (chunk_elems a multiple of 4096, at most 4 chunks),
// gp.total_units / ctrl / cta_part / pkt / grad_out filled; knobs: grid_knobs as in RowsParams
cudaError_t launch_kl_rows_grid_group(const GroupParams& gp, int max_row_units, const int* knobs, bool bf16, int sms,
                                      cudaStream_t stream, bool probe_only) {
    RowsParams p = {};
    p.nl = 1;
    p.total_units = gp.total_units;
    p.units_coarse = gp.total_units;
    p.max_row_units = max_row_units;
    p.ctrl = gp.ctrl;
    p.cta_part = gp.cta_part;
    p.pkt = gp.pkt;
    for (int i = 0; i < 4; ++i) p.grid_knobs[i] = knobs[i];
    return bf16 ? launch_grid_t<__nv_bfloat16, 1, 0, true>(p, gp, sms, stream, probe_only)
                : launch_grid_t<float, 1, 0, true>(p, gp, sms, stream, probe_only);
}

}  // namespace sd
