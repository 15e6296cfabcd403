// One launch for the channel-mode softmax-KL of SEVERAL (student, teacher) pairs: the dispatcher's step when a config
// distils more than one layer (mmseg/models/distillation/opts.py:87-112 loops over the `distillation` entries and the
// reference launches its whole ATen chain once per entry; BASELINE config 2: CGD on the four MiT stage maps).
//
// Separate launches pay a fixed ~10 us each (launch, prologue, first TMA round trip, last-CTA loss reduction) - more
// than the streaming time of the small stage maps.  Here the units (chunks of <= 7168 row elements) of all pairs form
// ONE work list that a persistent cooperative grid walks with the two-phase scheme of kl_rows_stream.cu:
//
//   phase 1 (unit j)          chunk -> shared (TMA) -> thread-local max, exponentials, per-warp partial sums; the unit's
//                             packet (epoch-tagged words in global memory) is published by the control warp.
//   phase 2 (unit j - delay)  the row statistics are merged from the packets of the row's units (any CTA), the chunk
//                             streams in again (L2-resident) and dS is written.
//
// Every pair has its own shape, group size, temperature, weight and loss scalar; rows may be ragged (C % g != 0) and of
// any length.  Same mathematics and the same cancellation-free KL as the other row kernels (common.cuh).
#include "rows_common.cuh"
#include "launch.h"

namespace sd {

constexpr int kGThreads = 512;
constexpr int kGCons = 448;                        // consumer threads (14 warps); warp 14 issues TMA, warp 15 is the control warp
constexpr int kGConsWarps = kGCons / 32;
constexpr int kGEPT = 16;                          // elements per consumer thread and tensor of one unit
constexpr int kGSlotVecRows = 2;
constexpr int kGSlotVecs = kGSlotVecRows * kGCons; // 896 vectors = 14 KB per tensor
constexpr int kGSlotBytes = kGSlotVecs * 16;
constexpr int kGStageBytes = 2 * kGSlotBytes;      // S + T
constexpr int kGStages = 3;
constexpr int kGBars = 2 * kGStages + 8;
constexpr int kGRec = 8;                           // warp record: ms, mt, zs, zt, a, dd
constexpr size_t kGroupSmemBytes = (size_t)kGStages * kGStageBytes + kGBars * sizeof(uint64_t) +
                                   2 * 16 * kGRec * sizeof(float) + 2 * 8 * sizeof(float);

struct GUnit {
    int si, b, grp, ck, nch;
    int e0, len;
};
__device__ __forceinline__ GUnit g_decode(const GroupParams& gp, int si, int b, int r) {
    const GroupSeg& s = gp.seg[si];
    GUnit x;
    x.si = si;
    x.b = b;
    const int full_units = s.G_full * s.nch_full;
    int g_real;
    if (r < full_units) {
        x.grp = r / s.nch_full;
        x.ck = r - x.grp * s.nch_full;
        x.nch = s.nch_full;
        g_real = s.g;
    } else {
        x.grp = s.G_full;
        x.ck = r - full_units;
        x.nch = s.nch_last;
        g_real = s.g_last;
    }
    const int L = g_real * s.HW;
    x.e0 = x.ck * s.chunk_elems;
    x.len = min(s.chunk_elems, L - x.e0);
    return x;
}
// walks the units of one CTA (u = blockIdx.x, += gridDim.x) across the pairs
struct GCursor {
    long long u;
    int si, b, r;
    __device__ __forceinline__ void init(const GroupParams& gp, long long u0) {
        u = u0;
        si = 0;
        while (si + 1 < gp.nseg && u0 >= gp.seg[si + 1].unit0) ++si;
        const long long local = u0 - gp.seg[si].unit0;
        b = (int)(local / gp.seg[si].units_per_sample);
        r = (int)(local - (long long)b * gp.seg[si].units_per_sample);
    }
    __device__ __forceinline__ void advance(const GroupParams& gp, int step) {
        u += step;
        r += step;
        while (si < gp.nseg) {
            const int ups = gp.seg[si].units_per_sample;
            if (r < ups) break;
            r -= ups;
            if (++b == gp.seg[si].B) {
                b = 0;
                ++si;
            }
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(kGThreads, 2) kl_rows_group_kernel(const __grid_constant__ GroupParams gp) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    constexpr int NV = kGEPT / VE;                // vectors per thread and tensor: 4 (fp32) / 2 (bf16)
    constexpr int NJ = NV / kGSlotVecRows;        // ring slots of a whole unit: 2 (fp32) / 1 (bf16)

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kGStages * kGStageBytes);
    uint64_t* empty = full + kGStages;
    uint64_t* part_ready = empty + kGStages;   // [2] consumers' warp records of a phase-1 unit are written
    uint64_t* part_free = part_ready + 2;      // [2] the control warp has read them
    uint64_t* stat_ready = part_free + 2;      // [2] row statistics of a phase-2 unit are in bcast
    uint64_t* stat_free = stat_ready + 2;      // [2] every consumer warp has read them
    float* red = reinterpret_cast<float*>(full + kGBars);  // [2][16][kGRec]
    float* bcast = red + 2 * 16 * kGRec;                   // [2][8]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < kGStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kGConsWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&part_ready[i], kGConsWarps);
            mbar_init(&part_free[i], 1);
            mbar_init(&stat_ready[i], 1);
            mbar_init(&stat_free[i], kGConsWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const unsigned epoch = __ldcg(&gp.ctrl[2]) + 1u;      // tag of this launch's packets (see kl_rows_stream.cu)
    const int grid = (int)gridDim.x;
    const int n_units = blockIdx.x < gp.total_units ? (int)((gp.total_units - blockIdx.x + grid - 1) / grid) : 0;
    const int D = gp.delay;
    const int n_steps = n_units > 0 ? n_units + D : 0;
    float gscale = 1.f;
    if (gp.grad_out != nullptr) gscale = __ldg(gp.grad_out);

    if (warp == kGConsWarps) {
        // ===================================================================================== TMA warp
        if (lane != 0) return;
        int pstage = 0;
        uint32_t pphase = 0;
        const uint64_t pol_keep = l2_policy_evict_last();
        const uint64_t pol_stream = l2_policy_evict_first();
        auto load_unit = [&](const GUnit& x, uint64_t pol) {
            const GroupSeg& s = gp.seg[x.si];
            const int nvec = x.len / VE;
            for (int v0 = 0; v0 < nvec; v0 += kGSlotVecs) {
                mbar_wait(&empty[pstage], pphase ^ 1u);
                const int nv = min(kGSlotVecs, nvec - v0);
                const uint32_t bytes = (uint32_t)nv * 16u;
                mbar_arrive_expect_tx(&full[pstage], 2u * bytes);
                unsigned char* dst_s = smem + (size_t)pstage * kGStageBytes;
                const int e = x.e0 + v0 * VE;
                const size_t off = (((size_t)x.b * s.C + (size_t)x.grp * s.g) * s.HW + e) * sizeof(T);
                tma_bulk_g2s(dst_s, static_cast<const char*>(s.S) + off, bytes, &full[pstage], pol);
                tma_bulk_g2s(dst_s + kGSlotBytes, static_cast<const char*>(s.T) + off, bytes, &full[pstage], pol);
                if (++pstage == kGStages) {
                    pstage = 0;
                    pphase ^= 1u;
                }
            }
        };
        GCursor c1, c2;
        c1.init(gp, blockIdx.x);
        c2.init(gp, blockIdx.x);
        for (int step = 0; step < n_steps; ++step) {
            if (step < n_units) {
                load_unit(g_decode(gp, c1.si, c1.b, c1.r), pol_keep);
                c1.advance(gp, grid);
            }
            if (step >= D) {
                load_unit(g_decode(gp, c2.si, c2.b, c2.r), pol_stream);
                c2.advance(gp, grid);
            }
        }
        return;
    }

    if (warp == kGConsWarps + 1) {
        // ===================================================================================== control warp
        float cta_kl[kMaxSegs];          // lane 0, in unit order (deterministic)
#pragma unroll
        for (int k = 0; k < kMaxSegs; ++k) cta_kl[k] = 0.f;
        GCursor c1, c2;
        c1.init(gp, blockIdx.x);
        c2.init(gp, blockIdx.x);
        for (int step = 0; step < n_steps; ++step) {
            const bool has1 = step < n_units, has2 = step >= D;
            const int par1 = step & 1, par2 = (step - D) & 1;
            const uint32_t ph1 = (uint32_t)(step >> 1) & 1u, ph2 = (uint32_t)((step - D) >> 1) & 1u;
            if (has2) {
                // ---- row statistics of the phase-2 unit from the packets of its row-mates (all CTAs)
                const long long u = c2.u;
                const GUnit x2 = g_decode(gp, c2.si, c2.b, c2.r);
                const GroupSeg& s = gp.seg[x2.si];
                mbar_wait(&stat_free[par2], ph2 ^ 1u);
                const int rown = x2.nch;
                const long long rowu = u - x2.ck;
                RowStat acc = rowstat_empty();
                for (int j = lane; j < rown; j += 32) {
                    const unsigned long long* q = gp.pkt + (size_t)(rowu + j) * kPktWords;
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
                            atomicExch(&gp.ctrl[1], 1u);      // never expected (cooperative launch); losses become NaN
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
                    acc = rowstat_merge(acc, r, s.c2);
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
                    acc = rowstat_merge(acc, other, s.c2);
                }
                if (lane == 0) {
                    const float coef = s.coef * gscale;
                    float* b = bcast + par2 * 8;
                    b[0] = __fmul_rn(acc.ms, s.c2);
                    b[1] = coef / acc.zs;
                    b[2] = __fmul_rn(acc.mt, s.c2);
                    b[3] = coef / acc.zt;
                    if (u == rowu) {
                        const float kl = kl_from_stats(acc.zs, acc.zt, acc.a, acc.dd);
                        if (s.row_kl) s.row_kl[x2.b * s.G + x2.grp] = kl;
#pragma unroll
                        for (int k = 0; k < kMaxSegs; ++k)
                            if (k == x2.si) cta_kl[k] += kl;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&stat_ready[par2]);
                c2.advance(gp, grid);
            }
            if (has1) {
                // ---- merge the consumer warps' records of the phase-1 unit, publish its packet
                const float c2k = gp.seg[c1.si].c2;
                mbar_wait(&part_ready[par1], ph1);
                const float* q = red + (par1 * 16 + (lane < kGConsWarps ? lane : 0)) * kGRec;
                const float4 r0 = reinterpret_cast<const float4*>(q)[0];
                const float4 r1 = reinterpret_cast<const float4*>(q)[1];
                __syncwarp();
                if (lane == 0) mbar_arrive(&part_free[par1]);
                const bool live = lane < kGConsWarps;
                const float rms = live ? r0.x : -INFINITY, rmt = live ? r0.y : -INFINITY;
                const float Ms = warp_max(rms);
                const float Mt = warp_max(rmt);
                const float fs = live ? ref_factor(rms, Ms, c2k) : 0.f;
                const float ft = live ? ref_factor(rmt, Mt, c2k) : 0.f;
                const float gx = live ? merge_shift(rms, rmt, Ms, Mt, c2k) : 0.f;
                const float zsk = live ? r0.z : 0.f, ztk = (live ? r0.w : 0.f) * ft;
                const float Zs = warp_sum(zsk * fs);
                const float Zt = warp_sum(ztk);
                const float A = warp_sum(fmaf(ztk, gx, (live ? r1.x : 0.f) * ft));
                const float DD = warp_sum(fmaf(zsk, factor_diff(fs, ft, gx), (live ? r1.y : 0.f) * ft));
                float val = 0.f;
                if (lane == 0) val = Ms;
                if (lane == 1) val = Zs;
                if (lane == 2) val = Mt;
                if (lane == 3) val = Zt;
                if (lane == 4) val = A;
                if (lane == 5) val = DD;
                if (lane < 6)
                    st_relaxed_u64(gp.pkt + (size_t)c1.u * kPktWords + lane,
                                   ((unsigned long long)epoch << 32) | __float_as_uint(val));
                c1.advance(gp, grid);
            }
        }

        // ---- losses: per-CTA partials per pair, the last CTA sums them in a fixed order
        unsigned ticket = 0;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kMaxSegs; ++k)
                if (k < gp.nseg) __stcg(&gp.cta_part[k * kMaxGrid + blockIdx.x], cta_kl[k]);
            __threadfence();
            ticket = atomicAdd(&gp.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            const bool timed_out = __ldcg(&gp.ctrl[1]) != 0u;
            for (int k = 0; k < gp.nseg; ++k) {
                double acc = 0.0;
                for (int i = lane; i < (int)gridDim.x; i += 32) acc += (double)__ldcg(&gp.cta_part[k * kMaxGrid + i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
                if (lane == 0)
                    *gp.seg[k].loss = timed_out ? __int_as_float(0x7fc00000) : (float)((double)gp.seg[k].loss_scale * acc);
            }
            if (lane == 0) {
                atomicAdd(&gp.ctrl[2], 1u);
                atomicExch(&gp.ctrl[0], 0u);
            }
        }
        return;
    }

    // ========================================================================================= consumer warps
    int stage = 0;
    uint32_t phase = 0;
    auto release_slot = [&]() {  // hand a drained slot back: one arrival per consumer warp
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kGStages) {
            stage = 0;
            phase ^= 1u;
        }
    };

    GCursor c1, c2;
    c1.init(gp, blockIdx.x);
    c2.init(gp, blockIdx.x);
    for (int step = 0; step < n_steps; ++step) {
        // ------------------------------------------------ phase 1: partial statistics of unit `step`
        if (step < n_units) {
            const GUnit x = g_decode(gp, c1.si, c1.b, c1.r);
            const float c2k = gp.seg[x.si].c2;
            const int nvec = x.len / VE;
            const bool whole = nvec == NV * kGCons;
            // pass A over the unit's slots: thread-local maxima (the slots stay put)
            float ms = kMaxFloor, mt = kMaxFloor;
            {
                int st = stage;
                uint32_t ph = phase;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (whole || j * kGSlotVecs < nvec) {
                        mbar_wait(&full[st], ph);
                        const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)st * kGStageBytes);
                        const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)st * kGStageBytes + kGSlotBytes);
#pragma unroll
                        for (int r = 0; r < kGSlotVecRows; ++r) {
                            if (whole || (j * kGSlotVecRows + r) * kGCons + tid < nvec) {
                                float fs[VE], ft[VE];
                                E::unpack(bs[r * kGCons + tid], fs);
                                E::unpack(bt[r * kGCons + tid], ft);
#pragma unroll
                                for (int q = 0; q < VE; ++q) {
                                    ms = fmaxf(ms, fs[q]);
                                    mt = fmaxf(mt, ft[q]);
                                }
                            }
                        }
                        if (++st == kGStages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
            // pass B: exponentials against the local maxima, partial sums; the slots are handed back
            // (a = sum et (at - as), dd = sum (et - es) term by term, common.cuh; zs = zt - dd against thread-local maxima)
            float zt = 0.f, dd = 0.f, a = 0.f;
            const float ms2 = __fmul_rn(ms, c2k), mt2 = __fmul_rn(mt, c2k);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (whole || j * kGSlotVecs < nvec) {
                    const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kGStageBytes);
                    const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kGStageBytes + kGSlotBytes);
                    vec_t vs[kGSlotVecRows], vt[kGSlotVecRows];
#pragma unroll
                    for (int r = 0; r < kGSlotVecRows; ++r) {
                        if (whole || (j * kGSlotVecRows + r) * kGCons + tid < nvec) {
                            vs[r] = bs[r * kGCons + tid];
                            vt[r] = bt[r * kGCons + tid];
                        }
                    }
#pragma unroll
                    for (int r = 0; r < kGSlotVecRows; ++r) {
                        if (whole || (j * kGSlotVecRows + r) * kGCons + tid < nvec) {
                            float fs[VE], ft[VE];
                            E::unpack(vs[r], fs);
                            E::unpack(vt[r], ft);
#pragma unroll
                            for (int q = 0; q < VE; ++q) {
                                const float as = fmaf(fs[q], c2k, -ms2), at = fmaf(ft[q], c2k, -mt2);
                                const float es = fast_exp2(as);
                                const float et = fast_exp2(at);
                                zt += et;
                                dd += et - es;
                                a = fmaf(et, at - as, a);
                            }
                        }
                    }
                    release_slot();          // (after the arithmetic that consumed the shared-memory reads)
                }
            }
            const float msw = warp_max(ms), mtw = warp_max(mt);
            const float fs = ref_factor(ms, msw, c2k);
            const float ft = ref_factor(mt, mtw, c2k);
            const float zs = zt - dd;
            const float gx = merge_shift2(ms2, mt2, __fmul_rn(msw, c2k), __fmul_rn(mtw, c2k));
            const float r_zs = warp_sum(zs * fs), r_zt = warp_sum(zt * ft);
            const float r_a = warp_sum(fmaf(zt * ft, gx, a * ft));
            const float r_dd = warp_sum(fmaf(zs, factor_diff(fs, ft, gx), dd * ft));
            const int par1 = step & 1;
            if (lane == 0) {
                mbar_wait(&part_free[par1], ((uint32_t)(step >> 1) & 1u) ^ 1u);
                float* my_red = red + (par1 * 16 + warp) * kGRec;
                reinterpret_cast<float4*>(my_red)[0] = make_float4(msw, mtw, r_zs, r_zt);
                reinterpret_cast<float4*>(my_red)[1] = make_float4(r_a, r_dd, 0.f, 0.f);
                mbar_arrive(&part_ready[par1]);
            }
            c1.advance(gp, grid);
        }

        // ------------------------------------------------ phase 2: gradient of unit `step - D`
        if (step >= D) {
            const GUnit x = g_decode(gp, c2.si, c2.b, c2.r);
            const GroupSeg& s = gp.seg[x.si];
            const float c2k = s.c2;
            const int nvec = x.len / VE;
            const bool whole = nvec == NV * kGCons;
            const int par2 = (step - D) & 1;
            mbar_wait(&stat_ready[par2], (uint32_t)((step - D) >> 1) & 1u);
            const float4 b = *reinterpret_cast<const float4*>(bcast + par2 * 8);
            const float ms2 = b.x, ks = b.y, mt2 = b.z, kt = b.w;
            __syncwarp();
            if (lane == 0) mbar_arrive(&stat_free[par2]);
            T* out = static_cast<T*>(s.dS);
            vec_t* dst = reinterpret_cast<vec_t*>(out + ((size_t)x.b * s.C + (size_t)x.grp * s.g) * s.HW + x.e0) + tid;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (whole || j * kGSlotVecs < nvec) {
                    mbar_wait(&full[stage], phase);
                    const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kGStageBytes);
                    const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kGStageBytes + kGSlotBytes);
                    vec_t vs[kGSlotVecRows], vt[kGSlotVecRows];
#pragma unroll
                    for (int r = 0; r < kGSlotVecRows; ++r) {
                        if (whole || (j * kGSlotVecRows + r) * kGCons + tid < nvec) {
                            vs[r] = bs[r * kGCons + tid];
                            vt[r] = bt[r * kGCons + tid];
                        }
                    }
#pragma unroll
                    for (int r = 0; r < kGSlotVecRows; ++r) {
                        const int v = j * kGSlotVecRows + r;
                        if (whole || v * kGCons + tid < nvec) {
                            float fs[VE], ft[VE], o[VE];
                            E::unpack(vs[r], fs);
                            E::unpack(vt[r], ft);
#pragma unroll
                            for (int q = 0; q < VE; ++q) {
                                const float es = fast_exp2(fmaf(fs[q], c2k, -ms2));
                                const float et = fast_exp2(fmaf(ft[q], c2k, -mt2));
                                o[q] = fmaf(es, ks, -et * kt);
                            }
                            dst[v * kGCons] = E::pack(o);
                        }
                    }
                    release_slot();
                }
            }
            c2.advance(gp, grid);
        }
    }
}

// ====================================================================================================
template <typename T>
static cudaError_t launch_group_t(const GroupParams& gp, int max_row_units, int sms, cudaStream_t stream) {
    auto kern = kl_rows_group_kernel<T>;
    static std::atomic<int> ctas_per_sm_dev[kMaxDevices];
    std::atomic<int>& ctas_per_sm = ctas_per_sm_dev[device_slot()];
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGroupSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kGThreads, kGroupSmemBytes);
        if (e != cudaSuccess) return e;
        if (n < 1) return cudaErrorLaunchOutOfResources;
        ctas_per_sm = n > 2 ? 2 : n;
    }
    long long grid = (long long)sms * ctas_per_sm;
    if (grid > gp.total_units) grid = gp.total_units;
    if (grid > kMaxGrid) grid = kMaxGrid;
    GroupParams q = gp;
    // phase 2 trails phase 1 by MORE than the longest row in grid rounds (see kl_rows_stream.cu)
    long long need = (max_row_units - 1 + grid - 1) / grid + 1;
    if (q.delay < need) q.delay = (int)need;
    if (q.delay < 1) q.delay = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGThreads);
    cfg.dynamicSmemBytes = kGroupSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // every CTA must be resident: they read each other's packets
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, q);
}

cudaError_t launch_kl_rows_group(const GroupParams& gp, int max_row_units, bool bf16, int sms, cudaStream_t stream) {
    return bf16 ? launch_group_t<__nv_bfloat16>(gp, max_row_units, sms, stream)
                : launch_group_t<float>(gp, max_row_units, sms, stream);
}

int kl_rows_group_chunk_capacity() { return kGCons * kGEPT; }

}  // namespace sd
