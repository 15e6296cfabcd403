// Row-wise softmax-KL for rows that do not fit one CTA's registers, and for two losses fused over one
// (student, teacher) pair: a warp-specialised streaming TWO-PHASE kernel with an L2-resident re-read.
//
// Same mathematics as kl_rows.cu (mmseg/models/distillation/losses.py:50-58,:108-112 + backward).
// A row of CGD with g = 10 on 128x128 logits is 163 840 elements (1.3 MB for S and T): no SM can hold
// it, so its softmax statistics need every chunk before any gradient can be written.  Instead of
// stalling a CTA until the other chunks' partials arrive, each persistent CTA runs its units
// (7168-element chunks, u = blockIdx.x + j*gridDim.x) through two phases that are `delay` units apart:
//
//   phase 1 (unit j)          chunk -> shared (TMA) -> thread-local max, exponentials, per-warp partial
//                             (max, sum, sum p*(t-s)) per loss.  Nothing is kept.
//   phase 2 (unit j - delay)  the row statistics are known by now; the chunk streams in AGAIN - it is
//                             still in the 126 MB L2, the re-read costs no HBM traffic - and dS is
//                             written from the stream.
//
// Warp roles (no CTA-wide barrier anywhere):
//   14 consumer warps   do the arithmetic; they wait only on mbarriers (a ring slot is full, the row
//                       statistics of the phase-2 unit are there) and never on each other.
//   1 TMA warp          one lane issues the 1-D TMA bulk copies of the tasks, in consumption order, as
//                       ring slots drain.
//   1 control warp      merges the consumers' partials and publishes the unit's packet (epoch-tagged
//                       8-byte words in global memory: no atomics, no counters to reset), and - ahead
//                       of the consumers - gathers the packets of the phase-2 unit's row-mates from the
//                       other CTAs and broadcasts the row statistics through shared memory.
//
// HBM traffic stays the algorithmic read S + read T + write dS; the price is a second pass of
// exponentials (4 ex2 per element and loss instead of 2).  Two CTAs per SM (<= 64 registers, a 3-stage
// 84 KB ring each).  NL = 2 serves two losses with nested rows (CD + CGD on the same logits).
#include "rows_common.cuh"
#include "launch.h"

namespace sd {

constexpr int kSThreads = 512;
constexpr int kSCons = 448;                        // consumer threads (14 warps); warp 14 issues TMA, warp 15 is the control warp
constexpr int kSConsWarps = kSCons / 32;
constexpr int kSEPT = 16;                          // elements per consumer thread and tensor of one unit
constexpr int kSSlotVecRows = 2;
constexpr int kSSlotVecs = kSSlotVecRows * kSCons; // 896 vectors = 14 KB per tensor
constexpr int kSSlotBytes = kSSlotVecs * 16;
constexpr int kSStageBytes = 2 * kSSlotBytes;      // S + T
constexpr int kSStages = 3;
constexpr int kSBars = 2 * kSStages + 8;           // full, empty, part_ready[2], part_free[2], stat_ready[2], stat_free[2]
constexpr int kSRec = 12;                          // warp record: ms, mt, {zs, zt, a, dd} x NL; [6]: sq (MSE, NL == 1)
constexpr size_t kStreamSmemBytes = (size_t)kSStages * kSStageBytes + kSBars * sizeof(uint64_t) +
                                    2 * 16 * kSRec * sizeof(float) + 2 * kMaxLosses * 8 * sizeof(float);

template <typename T, int NL, bool MSE>
__global__ void __launch_bounds__(kSThreads, 2) kl_rows_stream_kernel(const RowsParams p) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    constexpr int NV = kSEPT / VE;                // vectors per thread and tensor: 4 (fp32) / 2 (bf16)
    constexpr int NJ = NV / kSSlotVecRows;        // ring slots of a whole unit: 2 (fp32) / 1 (bf16)
    static_assert(NV % kSSlotVecRows == 0, "layout");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kSStages * kSStageBytes);
    uint64_t* empty = full + kSStages;
    uint64_t* part_ready = empty + kSStages;   // [2] consumers' warp records of a phase-1 unit are written
    uint64_t* part_free = part_ready + 2;      // [2] the control warp has read them
    uint64_t* stat_ready = part_free + 2;      // [2] row statistics of a phase-2 unit are in bcast
    uint64_t* stat_free = stat_ready + 2;      // [2] every consumer warp has read them
    float* red = reinterpret_cast<float*>(full + kSBars);  // [2][16][kSRec]
    float* bcast = red + 2 * 16 * kSRec;                   // [2][kMaxLosses][8]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (p.run_if != nullptr && *p.run_if == 0u) return;  // cancelled backward re-run (uniform over the grid)

    if (tid == 0) {
        for (int s = 0; s < kSStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kSConsWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&part_ready[i], kSConsWarps);
            mbar_init(&part_free[i], 1);
            mbar_init(&stat_ready[i], 1);
            mbar_init(&stat_free[i], kSConsWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // tag of this launch's unit packets: the workspace's launch counter (bumped by the last CTA to finish,
    // i.e. after every CTA has read it) + 1, so a packet left by any earlier launch never validates
    const unsigned epoch = __ldcg(&p.ctrl[2]) + 1u;
    const int grid = (int)gridDim.x;
    const int n_units = blockIdx.x < p.total_units ? (int)((p.total_units - blockIdx.x + grid - 1) / grid) : 0;
    const int D = p.delay;
    const int n_steps = n_units > 0 ? n_units + D : 0;
    // step t: phase 1 of this CTA's unit t (t < n_units), then phase 2 of its unit t - D (t >= D)

    if (warp == kSConsWarps) {
        // =====================================================================================
        // TMA warp: one lane streams the chunks of the tasks, in consumption order, into ring slots as
        // they drain
        // =====================================================================================
        if (lane != 0) return;
        int pstage = 0;
        uint32_t pphase = 0;
        const uint64_t pol_keep = l2_policy_evict_last();
        const uint64_t pol_stream = l2_policy_evict_first();
        auto load_unit = [&](const Unit& x, uint64_t pol) {
            const int nvec = x.len / VE;
            for (int v0 = 0; v0 < nvec; v0 += kSSlotVecs) {
                mbar_wait(&empty[pstage], pphase ^ 1u);
                const int nv = min(kSSlotVecs, nvec - v0);
                const uint32_t bytes = (uint32_t)nv * 16u;
                mbar_arrive_expect_tx(&full[pstage], 2u * bytes);
                unsigned char* dst_s = smem + (size_t)pstage * kSStageBytes;
                unsigned char* dst_t = dst_s + kSSlotBytes;
                const int e = x.e0 + v0 * VE;
                if (p.perm == nullptr) {
                    const size_t off = (((size_t)x.b * p.C + (size_t)x.grp * p.l[0].g) * p.HW + e) * sizeof(T);
                    tma_bulk_g2s(dst_s, static_cast<const char*>(p.S) + off, bytes, &full[pstage], pol);
                    tma_bulk_g2s(dst_t, static_cast<const char*>(p.T) + off, bytes, &full[pstage], pol);
                } else {
                    int remaining = nv * VE;
                    int cur = e;
                    uint32_t doff = 0;
                    while (remaining > 0) {  // gathered channels: one copy per channel segment
                        const int j = cur / p.HW;
                        const int pos = cur - j * p.HW;
                        const int n = min(remaining, p.HW - pos);
                        const size_t off = perm_elem_offset(p, x, cur) * sizeof(T);
                        const uint32_t nb = (uint32_t)n * (uint32_t)sizeof(T);
                        tma_bulk_g2s(dst_s + doff, static_cast<const char*>(p.S) + off, nb, &full[pstage], pol);
                        tma_bulk_g2s(dst_t + doff, static_cast<const char*>(p.T) + off, nb, &full[pstage], pol);
                        doff += nb;
                        cur += n;
                        remaining -= n;
                    }
                }
                if (++pstage == kSStages) {
                    pstage = 0;
                    pphase ^= 1u;
                }
            }
        };
        UnitCursor c1, c2;
        c1.init(p, blockIdx.x);
        c2.init(p, blockIdx.x);
        for (int step = 0; step < n_steps; ++step) {
            if (step < n_units) {
                load_unit(decode_unit(p, c1.b, c1.r), pol_keep);
                c1.advance(p, grid);
            }
            if (step >= D) {
                load_unit(decode_unit(p, c2.b, c2.r), pol_stream);
                c2.advance(p, grid);
            }
        }
        return;
    }

    if (warp == kSConsWarps + 1) {
        // =====================================================================================
        // control warp: unit packets out, row statistics in
        // =====================================================================================
        float cta_kl[NL], cta_sq = 0.f;  // lane 0, in unit order (deterministic)
#pragma unroll
        for (int k = 0; k < NL; ++k) cta_kl[k] = 0.f;
        UnitCursor c1, c2;
        c1.init(p, blockIdx.x);
        c2.init(p, blockIdx.x);
        for (int step = 0; step < n_steps; ++step) {
            const bool has1 = step < n_units, has2 = step >= D;
            const int par1 = step & 1, par2 = (step - D) & 1;
            const uint32_t ph1 = (uint32_t)(step >> 1) & 1u, ph2 = (uint32_t)((step - D) >> 1) & 1u;
            if (has2) {
                // ---- row statistics of the phase-2 unit from the packets of its row-mates (all CTAs)
                const long long u = c2.u;
                const Unit x2 = decode_unit(p, c2.b, c2.r);
                mbar_wait(&stat_free[par2], ph2 ^ 1u);
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    int rown, rowi;
                    long long rowu;
                    if (k == 0) {
                        rown = x2.nch;
                        rowu = u - x2.ck;
                        rowi = x2.b * p.l[0].G + x2.grp;
                    } else {
                        const int m = p.l[k].m;
                        const int rk = x2.grp / m;
                        const int j0 = rk * m;
                        const int j1 = min(j0 + m, p.l[0].G);
                        const int us = unit_start(p, j0);
                        rown = unit_start(p, j1) - us;
                        rowu = (long long)x2.b * p.units_per_sample + us;
                        rowi = x2.b * p.l[k].G + rk;
                    }
                    RowStat acc = rowstat_empty();
                    for (int j = lane; j < rown; j += 32) {
                        const unsigned long long* q = p.pkt + (size_t)(rowu + j) * kPktWords + 6 * k;
                        unsigned long long w[6];
                        unsigned spins = 0;
                        for (;;) {
                            bool ok = true;
#pragma unroll
                            for (int i = 0; i < 6; ++i) {
                                w[i] = ld_relaxed_u64(q + i);
                                ok = ok && (unsigned)(w[i] >> 32) == epoch;
                            }
                            if (ok) break;
                            if (++spins > kSpinLimit) {
                                // never expected (the launch is cooperative: every CTA is resident).  The flag makes
                                // the last CTA report NaN losses instead of numbers built on a stale packet.
                                atomicExch(&p.ctrl[1], 1u);
                                break;
                            }
                            __nanosleep(64);
                        }
                        RowStat r;
                        r.ms = __uint_as_float((unsigned)w[0]);
                        r.zs = __uint_as_float((unsigned)w[1]);
                        r.mt = __uint_as_float((unsigned)w[2]);
                        r.zt = __uint_as_float((unsigned)w[3]);
                        r.a = __uint_as_float((unsigned)w[4]);
                        r.dd = __uint_as_float((unsigned)w[5]);
                        acc = rowstat_merge(acc, r, p.l[k].c2);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        RowStat other;
                        other.ms = __shfl_xor_sync(0xffffffffu, acc.ms, o);
                        other.zs = __shfl_xor_sync(0xffffffffu, acc.zs, o);
                        other.mt = __shfl_xor_sync(0xffffffffu, acc.mt, o);
                        other.zt = __shfl_xor_sync(0xffffffffu, acc.zt, o);
                        other.a = __shfl_xor_sync(0xffffffffu, acc.a, o);
                        other.dd = __shfl_xor_sync(0xffffffffu, acc.dd, o);
                        acc = rowstat_merge(acc, other, p.l[k].c2);
                    }
                    if (lane == 0) {
                        float coef = p.l[k].coef;
                        if (p.grad_out[k] != nullptr) coef *= __ldg(p.grad_out[k]);
                        float* b = bcast + (par2 * kMaxLosses + k) * 8;
                        b[0] = __fmul_rn(acc.ms, p.l[k].c2);
                        b[1] = coef / acc.zs;
                        b[2] = __fmul_rn(acc.mt, p.l[k].c2);
                        b[3] = coef / acc.zt;
                        if (u == rowu) {
                            const float kl = kl_from_stats(acc.zs, acc.zt, acc.a, acc.dd);
                            if (p.l[k].row_kl) p.l[k].row_kl[rowi] = kl;
                            cta_kl[k] += kl;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&stat_ready[par2]);
                c2.advance(p, grid);
            }
            if (has1) {
                // ---- merge the consumer warps' records of the phase-1 unit, publish its packet
                mbar_wait(&part_ready[par1], ph1);
                const float* q = red + (par1 * 16 + (lane < kSConsWarps ? lane : 0)) * kSRec;
                const float4 r0 = reinterpret_cast<const float4*>(q)[0];
                const float4 r1 = reinterpret_cast<const float4*>(q)[1];
                const float4 r2 = reinterpret_cast<const float4*>(q)[2];
                __syncwarp();
                if (lane == 0) mbar_arrive(&part_free[par1]);
                const bool live = lane < kSConsWarps;
                const float rms = live ? r0.x : -INFINITY, rmt = live ? r0.y : -INFINITY;
                // {zs, zt, a, dd} of loss k: r0.zw r1.xy, r1.zw r2.xy
                const float rz[kMaxLosses][4] = {{r0.z, r0.w, r1.x, r1.y}, {r1.z, r1.w, r2.x, r2.y}};
                const float Ms = warp_max(rms);
                const float Mt = warp_max(rmt);
                float val = 0.f;
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    const float c2k = p.l[k].c2;
                    const float fs = live ? ref_factor(rms, Ms, c2k) : 0.f;
                    const float ft = live ? ref_factor(rmt, Mt, c2k) : 0.f;
                    const float gx = live ? merge_shift(rms, rmt, Ms, Mt, c2k) : 0.f;
                    const float zsk = live ? rz[k][0] : 0.f, ztk = (live ? rz[k][1] : 0.f) * ft;
                    const float Zs = warp_sum(zsk * fs);
                    const float Zt = warp_sum(ztk);
                    const float A = warp_sum(fmaf(ztk, gx, (live ? rz[k][2] : 0.f) * ft));
                    const float DD = warp_sum(fmaf(zsk, factor_diff(fs, ft, gx), (live ? rz[k][3] : 0.f) * ft));
                    if (lane == 6 * k + 0) val = Ms;
                    if (lane == 6 * k + 1) val = Zs;
                    if (lane == 6 * k + 2) val = Mt;
                    if (lane == 6 * k + 3) val = Zt;
                    if (lane == 6 * k + 4) val = A;
                    if (lane == 6 * k + 5) val = DD;
                }
                if (MSE) cta_sq += warp_sum(live ? r1.z : 0.f);
                if (lane < 6 * NL)
                    st_relaxed_u64(p.pkt + (size_t)c1.u * kPktWords + lane,
                                   ((unsigned long long)epoch << 32) | __float_as_uint(val));
                c1.advance(p, grid);
            }
        }

        // ---- loss: per-CTA partials, the last CTA sums them in a fixed order
        unsigned ticket = 0;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NL; ++k) __stcg(&p.cta_part[k * kMaxGrid + blockIdx.x], cta_kl[k]);
            __stcg(&p.cta_part[kMaxLosses * kMaxGrid + blockIdx.x], cta_sq);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double acc[NL + 1];
#pragma unroll
            for (int k = 0; k <= NL; ++k) acc[k] = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
#pragma unroll
                for (int k = 0; k < NL; ++k) acc[k] += (double)__ldcg(&p.cta_part[k * kMaxGrid + i]);
                acc[NL] += (double)__ldcg(&p.cta_part[kMaxLosses * kMaxGrid + i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k <= NL; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
            }
            if (lane == 0) {
                // a control warp that gave up waiting for a packet (ctrl[1]) built statistics on stale data: the
                // losses of this launch are reported as NaN - visible to the caller without a synchronisation -
                // and the flag stays set (the Python binding's workspace_error_flag() reads it)
                const bool timed_out = __ldcg(&p.ctrl[1]) != 0u;
#pragma unroll
                for (int k = 0; k < NL; ++k)
                    *p.l[k].loss = timed_out ? __int_as_float(0x7fc00000) : (float)((double)p.l[k].loss_scale * acc[k]);
                if (MSE && p.mse_loss) *p.mse_loss = (float)((double)p.mse_scale * acc[NL]);
                atomicAdd(&p.ctrl[2], 1u);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
        return;
    }

    // =========================================================================================
    // consumer warps
    // =========================================================================================
    int stage = 0;
    uint32_t phase = 0;
    auto release_slot = [&]() {  // hand a drained slot back: one arrival per consumer warp
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kSStages) {
            stage = 0;
            phase ^= 1u;
        }
    };

    UnitCursor c1, c2;
    c1.init(p, blockIdx.x);
    c2.init(p, blockIdx.x);
    for (int step = 0; step < n_steps; ++step) {
        // ------------------------------------------------ phase 1: partial statistics of unit `step`
        if (step < n_units) {
            const Unit x = decode_unit(p, c1.b, c1.r);
            const int nvec = x.len / VE;
            const bool whole = nvec == NV * kSCons;
            // pass A over the unit's slots: thread-local maxima (the slots stay put)
            float ms = kMaxFloor, mt = kMaxFloor;
            {
                int st = stage;
                uint32_t ph = phase;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (whole || j * kSSlotVecs < nvec) {
                        mbar_wait(&full[st], ph);
                        const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)st * kSStageBytes);
                        const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)st * kSStageBytes + kSSlotBytes);
#pragma unroll
                        for (int r = 0; r < kSSlotVecRows; ++r) {
                            if (whole || (j * kSSlotVecRows + r) * kSCons + tid < nvec) {
                                float fs[VE], ft[VE];
                                E::unpack(bs[r * kSCons + tid], fs);
                                E::unpack(bt[r * kSCons + tid], ft);
#pragma unroll
                                for (int q = 0; q < VE; ++q) {
                                    ms = fmaxf(ms, fs[q]);
                                    mt = fmaxf(mt, ft[q]);
                                }
                            }
                        }
                        if (++st == kSStages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
            // pass B: exponentials against the local maxima, partial sums; the slots are handed back
            // (dd = sum (et - es) term by term, common.cuh; against thread-local maxima zs = zt - dd is exact enough)
            float zt[NL], dd[NL], a[NL], sq = 0.f;
            float ms2[NL], mt2[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                zt[k] = 0.f;
                dd[k] = 0.f;
                a[k] = 0.f;
                ms2[k] = __fmul_rn(ms, p.l[k].c2);
                mt2[k] = __fmul_rn(mt, p.l[k].c2);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (whole || j * kSSlotVecs < nvec) {
                    const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kSStageBytes);
                    const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kSStageBytes + kSSlotBytes);
                    vec_t vs[kSSlotVecRows], vt[kSSlotVecRows];
#pragma unroll
                    for (int r = 0; r < kSSlotVecRows; ++r) {
                        if (whole || (j * kSSlotVecRows + r) * kSCons + tid < nvec) {
                            vs[r] = bs[r * kSCons + tid];
                            vt[r] = bt[r * kSCons + tid];
                        }
                    }
#pragma unroll
                    for (int r = 0; r < kSSlotVecRows; ++r) {
                        if (whole || (j * kSSlotVecRows + r) * kSCons + tid < nvec) {
                            float fs[VE], ft[VE];
                            E::unpack(vs[r], fs);
                            E::unpack(vt[r], ft);
#pragma unroll
                            for (int q = 0; q < VE; ++q) {
                                if (MSE) {
                                    const float d = ft[q] - fs[q];
                                    sq = fmaf(d, d, sq);
                                }
#pragma unroll
                                for (int k = 0; k < NL; ++k) {
                                    const float as = fmaf(fs[q], p.l[k].c2, -ms2[k]), at = fmaf(ft[q], p.l[k].c2, -mt2[k]);
                                    const float es = fast_exp2(as);
                                    const float et = fast_exp2(at);
                                    zt[k] += et;
                                    dd[k] += et - es;
                                    a[k] = fmaf(et, at - as, a[k]);
                                }
                            }
                        }
                    }
                    // only now: the arrival must not overtake the shared-memory reads above (an mbarrier
                    // arrive is not ordered behind LDS that are still in flight; the arithmetic is)
                    release_slot();
                }
            }
            // warp record: raw maxima are common to all losses, sums are rescaled to them
            const float msw = warp_max(ms), mtw = warp_max(mt);
            float rec[kSRec];
#pragma unroll
            for (int i = 0; i < kSRec; ++i) rec[i] = 0.f;
            rec[0] = msw;
            rec[1] = mtw;
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const float c2k = p.l[k].c2;
                const float fs = ref_factor(ms, msw, c2k);
                const float ft = ref_factor(mt, mtw, c2k);
                const float zs = zt[k] - dd[k];
                const float gx = merge_shift2(ms2[k], mt2[k], __fmul_rn(msw, c2k), __fmul_rn(mtw, c2k));
                rec[2 + 4 * k] = warp_sum(zs * fs);
                rec[3 + 4 * k] = warp_sum(zt[k] * ft);
                rec[4 + 4 * k] = warp_sum(fmaf(zt[k] * ft, gx, a[k] * ft));
                rec[5 + 4 * k] = warp_sum(fmaf(zs, factor_diff(fs, ft, gx), dd[k] * ft));
            }
            if (MSE) rec[6] = warp_sum(sq);
            const int par1 = step & 1;
            if (lane == 0) {
                mbar_wait(&part_free[par1], ((uint32_t)(step >> 1) & 1u) ^ 1u);
                float* my_red = red + (par1 * 16 + warp) * kSRec;
                reinterpret_cast<float4*>(my_red)[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
                reinterpret_cast<float4*>(my_red)[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
                reinterpret_cast<float4*>(my_red)[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
                mbar_arrive(&part_ready[par1]);
            }
            c1.advance(p, grid);
        }

        // ------------------------------------------------ phase 2: gradient of unit `step - D`
        if (step >= D) {
            const Unit x = decode_unit(p, c2.b, c2.r);
            const int nvec = x.len / VE;
            const bool whole = nvec == NV * kSCons;
            const int par2 = (step - D) & 1;
            mbar_wait(&stat_ready[par2], (uint32_t)((step - D) >> 1) & 1u);
            float ms2[NL], mt2[NL], ks[NL], kt[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const float4 b = *reinterpret_cast<const float4*>(bcast + (par2 * kMaxLosses + k) * 8);
                ms2[k] = b.x;
                ks[k] = b.y;
                mt2[k] = b.z;
                kt[k] = b.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&stat_free[par2]);
            T* out = static_cast<T*>(p.dS);
            vec_t* dst = reinterpret_cast<vec_t*>(out + ((size_t)x.b * p.C + (size_t)x.grp * p.l[0].g) * p.HW + x.e0) + tid;
            const bool gathered = p.perm != nullptr;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (whole || j * kSSlotVecs < nvec) {
                    mbar_wait(&full[stage], phase);
                    const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kSStageBytes);
                    const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kSStageBytes + kSSlotBytes);
                    vec_t vs[kSSlotVecRows], vt[kSSlotVecRows];
#pragma unroll
                    for (int r = 0; r < kSSlotVecRows; ++r) {
                        if (whole || (j * kSSlotVecRows + r) * kSCons + tid < nvec) {
                            vs[r] = bs[r * kSCons + tid];
                            vt[r] = bt[r * kSCons + tid];
                        }
                    }
#pragma unroll
                    for (int r = 0; r < kSSlotVecRows; ++r) {
                        const int v = j * kSSlotVecRows + r;
                        const int vi = v * kSCons + tid;
                        if (whole || vi < nvec) {
                            float fs[VE], ft[VE], o[VE];
                            E::unpack(vs[r], fs);
                            E::unpack(vt[r], ft);
#pragma unroll
                            for (int q = 0; q < VE; ++q) {
                                float acc = MSE ? p.mse_gcoef * (fs[q] - ft[q]) : 0.f;
#pragma unroll
                                for (int k = 0; k < NL; ++k) {
                                    const float es = fast_exp2(fmaf(fs[q], p.l[k].c2, -ms2[k]));
                                    const float et = fast_exp2(fmaf(ft[q], p.l[k].c2, -mt2[k]));
                                    acc = fmaf(es, ks[k], acc);
                                    acc = fmaf(-et, kt[k], acc);
                                }
                                o[q] = acc;
                            }
                            if (!gathered) dst[v * kSCons] = E::pack(o);
                            else *reinterpret_cast<vec_t*>(out + perm_elem_offset(p, x, x.e0 + vi * VE)) = E::pack(o);
                        }
                    }
                    release_slot();  // after the arithmetic that consumed the shared-memory reads (see phase 1)
                }
            }
            c2.advance(p, grid);
        }
    }
}

// ====================================================================================================
template <typename T, int NL, bool MSE>
static cudaError_t launch_stream_t(const RowsParams& p, int sms, cudaStream_t stream) {
    auto kern = kl_rows_stream_kernel<T, NL, MSE>;
    static std::atomic<int> ctas_per_sm_dev[kMaxDevices];  // per instantiation and device
    std::atomic<int>& ctas_per_sm = ctas_per_sm_dev[device_slot()];
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kSThreads, kStreamSmemBytes);
        if (e != cudaSuccess) return e;
        if (n < 1) return cudaErrorLaunchOutOfResources;
        ctas_per_sm = n > 2 ? 2 : n;
    }
    long long grid = (long long)sms * ctas_per_sm;
    if (grid > p.total_units) grid = p.total_units;
    if (grid > kMaxGrid) grid = kMaxGrid;
    RowsParams q = p;
    // phase 2 trails phase 1 by `delay` units: MORE than the row length in grid rounds, or two control warps
    // would wait for each other's packets; 2 by default so that the packets are normally there when read
    long long need = (q.max_row_units - 1 + grid - 1) / grid + 1;
    if (q.delay < need) q.delay = (int)need;
    if (q.delay < 1) q.delay = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kSThreads);
    cfg.dynamicSmemBytes = kStreamSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // every CTA must be resident: they read each other's packets
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, q);
}

cudaError_t launch_kl_rows_stream(const RowsParams& p, bool bf16, int sms, cudaStream_t stream) {
    const bool mse = p.mse_gcoef != 0.f || p.mse_loss != nullptr;
    if (p.nl == 2) {
        return bf16 ? launch_stream_t<__nv_bfloat16, 2, false>(p, sms, stream)
                    : launch_stream_t<float, 2, false>(p, sms, stream);
    }
    if (bf16) {
        return mse ? launch_stream_t<__nv_bfloat16, 1, true>(p, sms, stream)
                   : launch_stream_t<__nv_bfloat16, 1, false>(p, sms, stream);
    }
    return mse ? launch_stream_t<float, 1, true>(p, sms, stream) : launch_stream_t<float, 1, false>(p, sms, stream);
}

int kl_rows_stream_chunk_capacity() { return kSCons * kSEPT; }

}  // namespace sd
