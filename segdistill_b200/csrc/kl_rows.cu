// Row-wise softmax-KL, forward + backward fused (CD / CGD / plain KLDLoss): rows of up to 16384
// elements, one row per CTA pass.  (Longer rows and two fused losses: kl_rows_stream.cu.)
//
// Replaces the ATen chain of mmseg/models/distillation/losses.py:35-42 (channel gather),
// :50-58 (group reshape, -1e9 pad) and :108-112 (div, log_softmax, softmax, kl_div, mul)
// plus its autograd backward.  Row = `g` consecutive (gathered) channels x HW.
//
//   kl_rows_tma_kernel      persistent, one CTA per SM.  Rows stream into a 7-stage shared-memory ring
//                           with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) issued by
//                           one elected thread as soon as a slot has been drained; the 16 warps pull a
//                           row into REGISTERS (32 elements of S and of T per thread), exponentiate
//                           against a thread-local maximum (no block barrier before the
//                           exponentials), combine (max, sum) partials with warp shuffles and one CTA
//                           barrier, and write dS straight from registers.  HBM traffic is the
//                           algorithmic 12 B/elem (fp32) / 6 B/elem (bf16): S and T are read once, dS
//                           written once; 2 ex2 per element.
//   kl_rows_generic_*       any alignment / any row length: three plain passes.
#include "rows_common.cuh"
#include "launch.h"

namespace sd {

// ------------------------------------------------------------------ configuration
constexpr int kThreads = 512;                     // 16 warps x 128 registers
constexpr int kWarps = kThreads / 32;
constexpr int kDataRegs = 32;                     // register-resident elements per thread and tensor (NL = 1)
constexpr int kSlotVecRows = 2;                   // 16-byte vectors per thread and ring slot
constexpr int kSlotVecs = kSlotVecRows * kThreads;  // 1024 vectors
constexpr int kSlotBytes = kSlotVecs * 16;        // 16 KB per tensor
constexpr int kStageBytes = 2 * kSlotBytes;       // S + T
constexpr int kStages = 7;                        // 224 KB ring
// thread 0's TMA-issue state and loss accumulators (shared memory, see the kernel)
struct ProducerState {
    UnitCursor cur;
    int v0, stage, free_slots;
    uint64_t pol;
    float kl, sq;
    float q[32][6];        // warp 0: {zs, zt, a2, dd, row} of up to 32 rows awaiting their KL, the lane's KL sum
};
constexpr size_t kRowsSmemBytes = (size_t)kStages * kStageBytes + (kStages + 1) * sizeof(uint64_t) +
                                  2 * kWarps * kRedFloats * sizeof(float) + sizeof(ProducerState);
// running maximum of packed bf16 pairs over the four words of a 16-byte vector (fp32 vectors: not used)
__device__ __forceinline__ uint32_t packed_max4(uint32_t m, const uint4& v) {
    uint32_t a, b;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(a) : "r"(v.x), "r"(v.y));
    asm("max.bf16x2 %0, %1, %2;" : "=r"(b) : "r"(v.z), "r"(v.w));
    asm("max.bf16x2 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b));
    asm("max.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(m), "r"(a));
    return m;
}
__device__ __forceinline__ uint32_t packed_max4(uint32_t m, const float4&) { return m; }

// ================================ TMA issue (thread 0 only) ================================
// Slots are refilled in ring order.  `free_slots` counts slots every thread has drained: a unit's
// slots are released by the CTA barrier that follows its ring->register copy.  The state lives in
// shared memory: only thread 0 touches it, and the 16 compute warps need all 128 registers.
template <typename T>
__device__ __forceinline__ void rows_issue_loads(const RowsParams& p, ProducerState& ps, unsigned char* smem, uint64_t* full,
                                                 int newly_free) {
    constexpr int VE = Elem<T>::kVec;
    int free_slots = ps.free_slots + newly_free, v0 = ps.v0, pstage = ps.stage;
    while (free_slots > 0 && ps.cur.u < p.total_units) {
        const Unit x = decode_unit(p, ps.cur.b, ps.cur.r);
        const int nvec = x.len / VE;
        const int nv = min(kSlotVecs, nvec - v0);
        const uint32_t bytes = (uint32_t)nv * 16u;
        mbar_arrive_expect_tx(&full[pstage], 2u * bytes);
        unsigned char* dst_s = smem + (size_t)pstage * kStageBytes;
        unsigned char* dst_t = dst_s + kSlotBytes;
        const int e = x.e0 + v0 * VE;
        if (p.perm == nullptr) {
            const size_t off = (((size_t)x.b * p.C + (size_t)x.grp * p.l[0].g) * p.HW + e) * sizeof(T);
            tma_bulk_g2s(dst_s, static_cast<const char*>(p.S) + off, bytes, &full[pstage], ps.pol);
            tma_bulk_g2s(dst_t, static_cast<const char*>(p.T) + off, bytes, &full[pstage], ps.pol);
        } else {
            // gathered channels: one copy per channel segment
            int remaining = nv * VE;
            int cur = e;
            uint32_t doff = 0;
            while (remaining > 0) {
                const int j = cur / p.HW;
                const int pos = cur - j * p.HW;
                const int n = min(remaining, p.HW - pos);
                const size_t off = perm_elem_offset(p, x, cur) * sizeof(T);
                const uint32_t nb = (uint32_t)n * (uint32_t)sizeof(T);
                tma_bulk_g2s(dst_s + doff, static_cast<const char*>(p.S) + off, nb, &full[pstage], ps.pol);
                tma_bulk_g2s(dst_t + doff, static_cast<const char*>(p.T) + off, nb, &full[pstage], ps.pol);
                doff += nb;
                cur += n;
                remaining -= n;
            }
        }
        v0 += kSlotVecs;
        if (v0 >= nvec) {
            v0 = 0;
            ps.cur.advance(p, (int)gridDim.x);
        }
        if (++pstage == kStages) pstage = 0;
        --free_slots;
    }
    ps.free_slots = free_slots;
    ps.v0 = v0;
    ps.stage = pstage;
}

// ================================ loss: per-CTA partials, last CTA sums them in a fixed order ================================
template <bool MSE>
__device__ __forceinline__ void rows_finish_loss(const RowsParams& p, const ProducerState& ps, int warp, int lane) {
    if (warp == 0) {
        unsigned ticket = 0;
        if (lane == 0) {
            __stcg(&p.cta_part[blockIdx.x], ps.kl);
            __stcg(&p.cta_part[kMaxLosses * kMaxGrid + blockIdx.x], ps.sq);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double kl = 0.0, sq = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
                kl += (double)__ldcg(&p.cta_part[i]);
                sq += (double)__ldcg(&p.cta_part[kMaxLosses * kMaxGrid + i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kl += __shfl_down_sync(0xffffffffu, kl, o);
                sq += __shfl_down_sync(0xffffffffu, sq, o);
            }
            if (lane == 0) {
                *p.l[0].loss = (float)((double)p.l[0].loss_scale * kl);
                if (MSE && p.mse_loss) *p.mse_loss = (float)((double)p.mse_scale * sq);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

// per-phase cycle counters of threads 0 and 480 (scripts/rows_timing.py; build with SD_NVCC_EXTRA=-DSD_ROWS_TIMING)
#ifdef SD_ROWS_TIMING
#define SD_RT_DECL long long rt_acc[6] = {0, 0, 0, 0, 0, 0}, rt_last = 0, rt_rows = 0, rt_begin = clock64();
#define SD_RT_START rt_last = clock64(); ++rt_rows;
#define SD_RT_MARK(k) { const long long rt_now = clock64(); rt_acc[k] += rt_now - rt_last; rt_last = rt_now; }
#define SD_RT_WRITE if (tid == 0 || tid == 480) { unsigned long long* d = p.dbg + blockIdx.x * 16 + (tid ? 8 : 0); \
    for (int k = 0; k < 6; ++k) d[k] = rt_acc[k]; d[6] = rt_rows; d[7] = clock64() - rt_begin; }
#else
#define SD_RT_DECL
#define SD_RT_START
#define SD_RT_MARK(k)
#define SD_RT_WRITE
#endif

template <typename T, bool MSE>
__global__ void __launch_bounds__(kThreads, 1) kl_rows_tma_kernel(const RowsParams p) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    constexpr int EPT = kDataRegs;       // elements per thread and tensor
    constexpr int NV = EPT / VE;         // 16-byte vectors per thread and tensor
    constexpr int NJ = NV / kSlotVecRows;  // ring slots of a whole chunk
    static_assert(NV % kSlotVecRows == 0, "layout");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
    float* red = reinterpret_cast<float*>(full + kStages + 1);      // [2][kWarps][kRedFloats]
    ProducerState& ps = *reinterpret_cast<ProducerState*>(red + 2 * kWarps * kRedFloats);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        ps.cur.init(p, blockIdx.x);
        ps.v0 = 0;
        ps.stage = 0;
        ps.free_slots = kStages;
        ps.pol = l2_policy_evict_first();
        ps.kl = 0.f;
        ps.sq = 0.f;
    }
    __syncthreads();

    auto issue_loads = [&](int newly_free) { rows_issue_loads<T>(p, ps, smem, full, newly_free); };
    if (tid == 0) issue_loads(0);

    // ================================ 16 warps, chunk lives in registers ================================
    const float c2 = p.l[0].c2;
    float s[EPT], t[EPT];  // raw values, then (unless MSE) their exponentials
    int stage = 0;
    uint32_t phase = 0;
    int par = 0;

    UnitCursor cur;
    int nrow = 0;
    if (warp == 0) ps.q[lane][5] = 0.f;
    SD_RT_DECL
    for (cur.init(p, blockIdx.x); cur.u < p.total_units; cur.advance(p, (int)gridDim.x)) {
        SD_RT_START
        const Unit x = decode_unit(p, cur.b, cur.r);
        const int nvec = x.len / VE;
        const bool whole = nvec == NV * kThreads;  // every thread holds NV vectors

        // ---- ring -> registers
        const int nslots = (nvec + kSlotVecs - 1) / kSlotVecs;
        // (bf16: the thread's maxima from the packed words - max.bf16x2, one instruction per two elements)
        uint32_t pms = 0xff80ff80u, pmt = 0xff80ff80u;      // (-inf, -inf)
        if (whole) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                mbar_wait(&full[stage], phase);
                const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes);
                const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes + kSlotBytes);
#pragma unroll
                for (int r = 0; r < kSlotVecRows; ++r) {
                    const int v = j * kSlotVecRows + r;
                    const vec_t vs = bs[r * kThreads + tid], vt = bt[r * kThreads + tid];
                    if constexpr (sizeof(T) == 2) {
                        pms = packed_max4(pms, vs);
                        pmt = packed_max4(pmt, vt);
                    }
                    E::unpack(vs, &s[v * VE]);
                    E::unpack(vt, &t[v * VE]);
                }
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                s[i] = kPadValue;
                t[i] = kPadValue;
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (j * kSlotVecs < nvec) {
                    mbar_wait(&full[stage], phase);
                    const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes);
                    const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes + kSlotBytes);
#pragma unroll
                    for (int r = 0; r < kSlotVecRows; ++r) {
                        const int v = j * kSlotVecRows + r;
                        if (v * kThreads + tid < nvec) {
                            E::unpack(bs[r * kThreads + tid], &s[v * VE]);
                            E::unpack(bt[r * kThreads + tid], &t[v * VE]);
                        }
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }

        // ---- thread-local maxima of the raw values: no barrier before the exponentials
        float ms = kMaxFloor, mt = kMaxFloor;
        if (sizeof(T) == 2 && whole) {
            float lo, hi;
            Elem<__nv_bfloat16>::unpack2(pms, lo, hi);
            ms = fmaxf(ms, fmaxf(lo, hi));
            Elem<__nv_bfloat16>::unpack2(pmt, lo, hi);
            mt = fmaxf(mt, fmaxf(lo, hi));
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                ms = fmaxf(ms, s[i]);
                mt = fmaxf(mt, t[i]);
            }
        }

        // every thread holds its elements in registers (the maxima consumed every shared-memory read):
        // hand the unit's slots back to the TMA thread now, so the next loads fly during the exponentials
        SD_RT_MARK(0)
        __syncthreads();
        if (tid == 0) issue_loads(nslots);
        SD_RT_MARK(1)

        // ---- exponentials (kept in registers), thread-partial sums relative to (ms, mt); a = sum et (at - as) and
        //      dd = sum (et - es) term by term (common.cuh: KL without cancellation).  Against thread-local maxima zs
        //      and zt lie in [1, 32], so zs = zt - dd is as accurate as a sum of its own: one accumulator less
        // (bf16 inputs: two elements per instruction - FFMA2 / FADD2, common.cuh; that kernel is bound by instruction
        //  issue.  The fp32 kernel is HBM-bound and measured 1.6 % slower with the packed forms: it keeps the scalar ones)
        constexpr bool kPacked = sizeof(T) == 2;
        float zt = 0.f, dd = 0.f, a = 0.f, sq = 0.f;
        const float ms2 = __fmul_rn(ms, c2), mt2 = __fmul_rn(mt, c2);
        if constexpr (kPacked) {
            const F2 C2 = f2_dup(c2), NMS = f2_dup(-ms2), NMT = f2_dup(-mt2), neg1 = f2_dup(-1.f);
            F2 ZT = f2_dup(0.f), DD = f2_dup(0.f), A = f2_dup(0.f), SQ = f2_dup(0.f);
#pragma unroll
            for (int i = 0; i < EPT; i += 2) {
                const F2 s2 = f2_make(s[i], s[i + 1]), t2 = f2_make(t[i], t[i + 1]);
                if (MSE) {
                    const F2 d = f2_fma(s2, neg1, t2);
                    SQ = f2_fma(d, d, SQ);
                }
                const F2 as2 = f2_fma(s2, C2, NMS), at2 = f2_fma(t2, C2, NMT);
                float as0, as1, at0, at1;
                f2_split(as2, as0, as1);
                f2_split(at2, at0, at1);
                const float es0 = fast_exp2(as0), es1 = fast_exp2(as1), et0 = fast_exp2(at0), et1 = fast_exp2(at1);
                const F2 es2 = f2_make(es0, es1), et2 = f2_make(et0, et1);
                ZT = f2_add(ZT, et2);
                DD = f2_add(DD, f2_fma(es2, neg1, et2));
                A = f2_fma(et2, f2_fma(as2, neg1, at2), A);
                if (!MSE) {
                    s[i] = es0;
                    s[i + 1] = es1;
                    t[i] = et0;
                    t[i + 1] = et1;
                }
            }
            zt = f2_sum(ZT);
            dd = f2_sum(DD);
            a = f2_sum(A);
            if (MSE) sq = f2_sum(SQ);
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                if (MSE) {
                    const float d = t[i] - s[i];
                    sq = fmaf(d, d, sq);
                }
                const float as = fmaf(s[i], c2, -ms2), at = fmaf(t[i], c2, -mt2);
                const float es = fast_exp2(as);
                const float et = fast_exp2(at);
                zt += et;
                dd += et - es;
                a = fmaf(et, at - as, a);
                if (!MSE) {
                    s[i] = es;
                    t[i] = et;
                }
            }
        }
        const float zs = zt - dd;
        SD_RT_MARK(2)

        // ---- warp: sums rescaled to the warp maxima (redux.sync for the maxima, transposed butterflies for the sums:
        //      the reductions were 4 of the 27 instructions per element)
        const float msw = warp_max_uniform(ms), mtw = warp_max_uniform(mt);
        {
            const float fs = ref_factor(ms, msw, c2);
            const float ft = ref_factor(mt, mtw, c2);
            const float gx = merge_shift2(ms2, mt2, __fmul_rn(msw, c2), __fmul_rn(mtw, c2));
            const float ztf = zt * ft;
            float* my_red = red + (par * kWarps + warp) * kRedFloats;       // ms, mt, zs, zt, a, dd, sq, -
            if (MSE) {
                const float v[8] = {zs * fs, ztf, fmaf(ztf, gx, a * ft), fmaf(zs, factor_diff(fs, ft, gx), dd * ft), sq, 0.f, 0.f, 0.f};
                const float tot = warp_sum8_transposed(v, lane);
                if ((lane & 3) == 0 && lane < 20) my_red[2 + (lane >> 2)] = tot;
            } else {
                const float v[4] = {zs * fs, ztf, fmaf(ztf, gx, a * ft), fmaf(zs, factor_diff(fs, ft, gx), dd * ft)};
                const float tot = warp_sum4_transposed(v, lane);
                if ((lane & 7) == 0) my_red[2 + (lane >> 3)] = tot;
            }
            if (lane == 1) my_red[0] = msw;
            if (lane == 2) my_red[1] = mtw;
        }
        SD_RT_MARK(3)
        __syncthreads();

        // ---- CTA = row: every warp merges the 16 warp records (lanes l and l+16 mirror each other); the sums only
        //      thread 0 needs (KL, MSE) are merged by warp 0 alone
        float Ms, Mt, Zs, Zt, A = 0.f, DD = 0.f, SQ = 0.f;
        {
            const float* q = red + (par * kWarps + (lane & 15)) * kRedFloats;
            const float4 r0 = reinterpret_cast<const float4*>(q)[0];
            const float4 r1 = reinterpret_cast<const float4*>(q)[1];
            Ms = warp_max_uniform(r0.x);
            Mt = warp_max_uniform(r0.y);
            const float fs = ref_factor(r0.x, Ms, c2);
            const float ft = ref_factor(r0.y, Mt, c2);
            Zs = sum16(r0.z * fs);
            Zt = sum16(r0.w * ft);
            if (warp == 0) {
                const float gx = merge_shift(r0.x, r0.y, Ms, Mt, c2);
                A = sum16(fmaf(r0.w * ft, gx, r1.x * ft));
                DD = sum16(fmaf(r0.z, factor_diff(fs, ft, gx), r1.y * ft));
                if (MSE) SQ = sum16(r1.z);
            }
        }
        par ^= 1;

        if (warp == 0) {
            // lane (row count & 31) of warp 0 keeps the row's statistics: the logarithms of kl_from_stats run once per
            // 32 rows with all lanes busy (~70 instructions that the other 15 warps would wait for behind the barrier)
            if (lane == (nrow & 31)) {
                ps.q[lane][0] = Zs;
                ps.q[lane][1] = Zt;
                ps.q[lane][2] = A;
                ps.q[lane][3] = DD;
                ps.q[lane][4] = __int_as_float(x.b * p.l[0].G + x.grp);
            }
            if (MSE && lane == 0) ps.sq += SQ;
            if ((nrow & 31) == 31 || cur.u + (long long)gridDim.x >= p.total_units) {
                __syncwarp();
                if (lane <= (nrow & 31)) {
                    const float kl = kl_from_stats(ps.q[lane][0], ps.q[lane][1], ps.q[lane][2], ps.q[lane][3]);
                    if (p.l[0].row_kl) p.l[0].row_kl[__float_as_int(ps.q[lane][4])] = kl;
                    ps.q[lane][5] += kl;
                }
                __syncwarp();
            }
        }
        ++nrow;

        SD_RT_MARK(4)
        // ---- gradient straight from registers
        float coef = p.l[0].coef;
        if (p.grad_out[0] != nullptr) coef *= __ldg(p.grad_out[0]);
        const float ks = coef * ref_factor(ms, Ms, c2) / Zs;
        const float kt = coef * ref_factor(mt, Mt, c2) / Zt;
        const F2 KS = f2_dup(ks), NKT = f2_dup(-kt);
        // (MSE: the exponents are computed again; as common subexpressions with the statistics loop they were all kept
        //  alive across the reductions, i.e. spilled - see kl_rows_pack_kernel)
        float ms2g = ms2, mt2g = mt2;
        if (MSE) asm volatile("" : "+f"(ms2g), "+f"(mt2g));
        auto grad_vec = [&](int v, float* o) {
#pragma unroll
            for (int q = 0; q < VE; q += 2) {
                const int i = v * VE + q;
                if (MSE) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float es = fast_exp2(fmaf(s[i + h], c2, -ms2g));
                        const float et = fast_exp2(fmaf(t[i + h], c2, -mt2g));
                        o[q + h] = fmaf(es, ks, -et * kt) + p.mse_gcoef * (s[i + h] - t[i + h]);
                    }
                } else if constexpr (kPacked) {
                    f2_split(f2_fma(f2_make(s[i], s[i + 1]), KS, f2_mul(f2_make(t[i], t[i + 1]), NKT)), o[q], o[q + 1]);
                } else {
                    o[q] = fmaf(s[i], ks, -t[i] * kt);
                    o[q + 1] = fmaf(s[i + 1], ks, -t[i + 1] * kt);
                }
            }
        };
        T* out = static_cast<T*>(p.dS);
        if (p.perm == nullptr) {
            vec_t* dst = reinterpret_cast<vec_t*>(out + ((size_t)x.b * p.C + (size_t)x.grp * p.l[0].g) * p.HW + x.e0) + tid;
            if (whole) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float o[VE];
                    grad_vec(v, o);
                    dst[v * kThreads] = E::pack(o);
                }
            } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if (v * kThreads + tid < nvec) {
                        float o[VE];
                        grad_vec(v, o);
                        dst[v * kThreads] = E::pack(o);
                    }
                }
            }
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int vi = v * kThreads + tid;
                if (vi < nvec) {
                    float o[VE];
                    grad_vec(v, o);
                    *reinterpret_cast<vec_t*>(out + perm_elem_offset(p, x, x.e0 + vi * VE)) = E::pack(o);
                }
            }
        }
        SD_RT_MARK(5)
    }
    SD_RT_WRITE

    if (warp == 0) {
        const float cta_kl = warp_sum(ps.q[lane][5]);
        if (lane == 0) ps.kl = cta_kl;
        __syncwarp();
    }
    rows_finish_loss<MSE>(p, ps, warp, lane);
}

// ====================================================================================================
// bf16 rows of exactly 16384 elements: references = the ROW maxima, ONE CTA barrier per row
// ====================================================================================================
// The bf16 launch of kl_rows_tma_kernel is not HBM-bound: its phases (ring -> registers with the bf16 -> fp32 unpack on
// the ALU pipe, exponentials on the MUFU pipe, the reductions' shuffle chains, the gradient) run one after the other
// between two CTA barriers, each on a different pipe (scripts/rows_timing.py: 1300 + 700 + 2100 + 1200 + 750 cycles per
// row).  Here
//   * a row stays PACKED (16 + 16 registers) from the ring until the exponentials, so the unpack runs next to the MUFU
//     instructions instead of before them;
//   * row r+1 comes out of the ring (packed, with its packed maxima) DURING the exponentials of row r - shared-memory
//     loads under the MUFU instructions - and its warp maxima are published together with row r's warp sums: the one
//     barrier of row r gives every thread the sums of row r and the maxima of row r+1.  Every thread exponentiates
//     against the ROW maxima, so the sums of threads and warps simply add - no rescale factors (ex2), no merge shifts;
//   * ring stages are whole rows (3 x 64 KB): one mbarrier wait and two bulk copies per row;
//   * the gradient leaves through shared memory and one 2 KB bulk store per warp (cp.async.bulk shared -> global):
//     128-bit global stores from 512 threads back up the queue that shared-memory loads, shuffles and MUFU
//     instructions share, and every phase behind the gradient sweep waited for them.  A warp holds 2 KB of
//     consecutive row elements (vector = warp * 128 + v * 32 + lane), so no CTA barrier is involved.
// Same statistics as everywhere (common.cuh, KL without cancellation): zs, zt, a2 = sum et (at - as), dd = sum (et - es).
constexpr int kRmStages = 3;
constexpr int kRmRowBytes = kThreads * kDataRegs * 2;            // one tensor's row: 32 KB
constexpr int kRmStageBytes = 2 * kRmRowBytes;                   // S + T
struct RmProducer {
    long long u_next;      // next unit this CTA loads
    int stage;             // ... into this stage
    uint64_t pol;
    float q[32][6];        // warp 0: {zs, zt, a2, dd, row} of up to 32 rows awaiting their KL, the lane's KL sum
};
constexpr int kRmWarpVecs = kDataRegs / 8 * 32;                  // 16-byte vectors of a row one warp holds (contiguous: 2 KB)
constexpr size_t kRmSmemBytes = (size_t)kRmStages * kRmStageBytes + kRmRowBytes + (kRmStages + 1) * sizeof(uint64_t) +
                                2 * kWarps * kRedFloats * sizeof(float) + sizeof(RmProducer);

// thread 0: one row into the stage at the producer's head (a no-op behind the CTA's last row)
__device__ __forceinline__ void rm_issue_row(const RowsParams& p, RmProducer& ps, unsigned char* smem, uint64_t* full) {
    const long long u = ps.u_next;
    if (u >= p.total_units) return;
    const int st = ps.stage;
    unsigned char* dst_s = smem + (size_t)st * kRmStageBytes;
    unsigned char* dst_t = dst_s + kRmRowBytes;
    mbar_arrive_expect_tx(&full[st], (uint32_t)kRmStageBytes);
    if (p.perm == nullptr) {
        // whole rows, C % g == 0: row u starts u rows into the tensor
        const size_t off = (size_t)u * kRmRowBytes;
        tma_bulk_g2s(dst_s, static_cast<const char*>(p.S) + off, kRmRowBytes, &full[st], ps.pol);
        tma_bulk_g2s(dst_t, static_cast<const char*>(p.T) + off, kRmRowBytes, &full[st], ps.pol);
    } else {
        // gathered channels: one copy per channel
        const int G = p.l[0].G, g = p.l[0].g;
        const int b = (int)(u / G), grp = (int)(u - (long long)b * G);
        const uint32_t nb = (uint32_t)p.HW * 2u;
        for (int j = 0; j < g; ++j) {
            const size_t off = (((size_t)b * p.C + p.perm[grp * g + j]) * p.HW) * 2u;
            tma_bulk_g2s(dst_s + (size_t)j * nb, static_cast<const char*>(p.S) + off, nb, &full[st], ps.pol);
            tma_bulk_g2s(dst_t + (size_t)j * nb, static_cast<const char*>(p.T) + off, nb, &full[st], ps.pol);
        }
    }
    ps.u_next = u + gridDim.x;
    ps.stage = st + 1 == kRmStages ? 0 : st + 1;
}

__global__ void __launch_bounds__(kThreads, 1) kl_rows_rm_kernel(const RowsParams p) {
    using T = __nv_bfloat16;
    using E = Elem<T>;
    constexpr int VE = E::kVec;
    constexpr int EPT = kDataRegs;
    constexpr int NV = EPT / VE;                 // 4 vectors per thread and tensor
    constexpr int NW = EPT / 2;                  // 16 packed words per thread and tensor
    constexpr uint32_t kNegInf2 = 0xff80ff80u;   // (-inf, -inf)

    extern __shared__ __align__(128) unsigned char smem[];
    uint4* out_stage = reinterpret_cast<uint4*>(smem + (size_t)kRmStages * kRmStageBytes);       // the row's gradient
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kRmStages * kRmStageBytes + kRmRowBytes);
    float* red = reinterpret_cast<float*>(full + kRmStages + 1);    // [2][kWarps][kRedFloats]: zs, zt, a, dd, ms', mt', -, -
    RmProducer& ps = *reinterpret_cast<RmProducer*>(red + 2 * kWarps * kRedFloats);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < kRmStages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        ps.u_next = blockIdx.x;
        ps.stage = 0;
        ps.pol = l2_policy_evict_first();
        for (int s = 0; s < kRmStages; ++s) rm_issue_row(p, ps, smem, full);
    }
    __syncthreads();

    const float c2 = p.l[0].c2;
    const int G = p.l[0].G;
    uint32_t rs[NW], rt[NW];                     // the row that is exponentiated next, packed
    uint32_t pms = kNegInf2, pmt = kNegInf2;     // its packed maxima
    int stage = 0;
    uint32_t phase = 0;
    int par = 0;

    // vector v of the row at the head of the ring -> registers (v is a compile-time constant after unrolling)
    auto load_vec = [&](int v) {
        const uint4* bs = reinterpret_cast<const uint4*>(smem + (size_t)stage * kRmStageBytes);
        const uint4* bt = reinterpret_cast<const uint4*>(smem + (size_t)stage * kRmStageBytes + kRmRowBytes);
        const uint4 vs = bs[warp * kRmWarpVecs + v * 32 + lane], vt = bt[warp * kRmWarpVecs + v * 32 + lane];
        pms = packed_max4(pms, vs);
        pmt = packed_max4(pmt, vt);
        rs[4 * v + 0] = vs.x; rs[4 * v + 1] = vs.y; rs[4 * v + 2] = vs.z; rs[4 * v + 3] = vs.w;
        rt[4 * v + 0] = vt.x; rt[4 * v + 1] = vt.y; rt[4 * v + 2] = vt.z; rt[4 * v + 3] = vt.w;
    };
    auto next_stage = [&]() {
        if (++stage == kRmStages) {
            stage = 0;
            phase ^= 1u;
        }
    };
    // the warp's maxima of the row in (pms, pmt) -> its record
    auto publish_max = [&](float* rec) {
        float lo, hi;
        E::unpack2(pms, lo, hi);
        const float msw = warp_max_uniform(fmaxf(lo, hi));
        E::unpack2(pmt, lo, hi);
        const float mtw = warp_max_uniform(fmaxf(lo, hi));
        if (lane == 0) *reinterpret_cast<float2*>(rec + 4) = make_float2(msw, mtw);
    };
    // the row maxima from the 16 records, as references of the exponents
    auto read_refs = [&](const float* recs, float& sig, float& th) {
        const float2 wm = *reinterpret_cast<const float2*>(recs + (lane & 15) * kRedFloats + 4);
        sig = __fmul_rn(warp_max_uniform(wm.x), c2);
        th = __fmul_rn(warp_max_uniform(wm.y), c2);
    };

    const unsigned total = (unsigned)p.total_units;     // rows of the launch (< 2^31: checked by the caller)
    unsigned u = blockIdx.x;
    bool have = u < total;
    float sig = 0.f, th = 0.f;
    if (have) {                                  // (uniform over the CTA)
        mbar_wait(&full[stage], phase);
#pragma unroll
        for (int v = 0; v < NV; ++v) load_vec(v);
        next_stage();
        publish_max(red + (par * kWarps + warp) * kRedFloats);
        __syncthreads();
        if (tid == 0) rm_issue_row(p, ps, smem, full);
        read_refs(red + par * kWarps * kRedFloats, sig, th);
        par ^= 1;
    }
    // warp 0: statistics of up to 32 rows, one per lane (shared memory: the compute warps need all their registers)
    int nrow = 0;
    if (warp == 0) ps.q[lane][5] = 0.f;
    auto flush_kl = [&](int last) {              // rows (last & ~31) .. last of this CTA
        __syncwarp();
        if (lane <= (last & 31)) {
            const float kl = kl_from_stats(ps.q[lane][0], ps.q[lane][1], ps.q[lane][2], ps.q[lane][3]);
            if (p.l[0].row_kl) p.l[0].row_kl[__float_as_uint(ps.q[lane][4])] = kl;
            ps.q[lane][5] += kl;
        }
        __syncwarp();
    };
    SD_RT_DECL
    while (have) {
        SD_RT_START
        const unsigned un = u + gridDim.x;
        const bool next = un < total;
        if (next) mbar_wait(&full[stage], phase);
        SD_RT_MARK(0)

        // ---- exponentials against the row maxima (kept in registers), the thread's four sums; behind every
        //      vector of this row the same vector of the next row comes out of the ring
        float es[EPT], et[EPT];
        float v4[4];
        pms = kNegInf2;
        pmt = kNegInf2;
        {
            const F2 C2 = f2_dup(c2), NS = f2_dup(-sig), NT = f2_dup(-th), neg1 = f2_dup(-1.f);
            F2 ZS = f2_dup(0.f), ZT = f2_dup(0.f), DD = f2_dup(0.f), A = f2_dup(0.f);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                float s0, s1, t0, t1;
                E::unpack2(rs[w], s0, s1);
                E::unpack2(rt[w], t0, t1);
                const F2 as2 = f2_fma(f2_make(s0, s1), C2, NS), at2 = f2_fma(f2_make(t0, t1), C2, NT);
                float as0, as1, at0, at1;
                f2_split(as2, as0, as1);
                f2_split(at2, at0, at1);
                const float es0 = fast_exp2(as0), es1 = fast_exp2(as1), et0 = fast_exp2(at0), et1 = fast_exp2(at1);
                const F2 es2 = f2_make(es0, es1), et2 = f2_make(et0, et1);
                ZS = f2_add(ZS, es2);
                ZT = f2_add(ZT, et2);
                DD = f2_add(DD, f2_fma(es2, neg1, et2));
                A = f2_fma(et2, f2_fma(as2, neg1, at2), A);
                es[2 * w] = es0;
                es[2 * w + 1] = es1;
                et[2 * w] = et0;
                et[2 * w + 1] = et1;
                if ((w & 3) == 3 && next) load_vec(w >> 2);
            }
            v4[0] = f2_sum(ZS);
            v4[1] = f2_sum(ZT);
            v4[2] = f2_sum(A);
            v4[3] = f2_sum(DD);
        }
        SD_RT_MARK(1)
        float* my_red = red + (par * kWarps + warp) * kRedFloats;
        {
            const float tot = warp_sum4_transposed(v4, lane);
            if ((lane & 7) == 0) my_red[lane >> 3] = tot;
        }
        if (next) {
            next_stage();
            publish_max(my_red);
        }
        SD_RT_MARK(2)
        __syncthreads();          // row u: sums complete; row un: maxima complete, its stage drained
        if (tid == 0 && next) rm_issue_row(p, ps, smem, full);
        SD_RT_MARK(3)

        // ---- CTA = row: the 16 warp records add up (lanes l and l+16 mirror each other)
        const float* recs = red + par * kWarps * kRedFloats;
        const float4 r4 = *reinterpret_cast<const float4*>(recs + (lane & 15) * kRedFloats);
        const float Zs = sum16(r4.x), Zt = sum16(r4.y);
        if (warp == 0) {
            // lane (row count & 31) of warp 0 keeps the row's statistics: the logarithms of kl_from_stats run once per
            // 32 rows with all lanes busy (~70 instructions that the other 15 warps would wait for behind every barrier)
            const float A = sum16(r4.z), DD = sum16(r4.w);
            if (lane == (nrow & 31)) {
                ps.q[lane][0] = Zs;
                ps.q[lane][1] = Zt;
                ps.q[lane][2] = A;
                ps.q[lane][3] = DD;
                ps.q[lane][4] = __uint_as_float(u);
            }
            if ((nrow & 31) == 31 || !next) flush_kl(nrow);
        }
        ++nrow;
        SD_RT_MARK(4)

        // ---- gradient from registers into the warp's 2 KB of the staging row, from there by one bulk store
        float coef = p.l[0].coef;
        if (p.grad_out[0] != nullptr) coef *= __ldg(p.grad_out[0]);
        const F2 KS = f2_dup(__fdividef(coef, Zs)), NKT = f2_dup(-__fdividef(coef, Zt));     // (bf16 results)
        uint4* my_out = out_stage + warp * kRmWarpVecs;
        if (lane == 0) tma_bulk_wait_read0();        // the previous row's store has read the staging row
        __syncwarp();
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float o[VE];
#pragma unroll
            for (int q = 0; q < VE; q += 2) {
                const int i = v * VE + q;
                f2_split(f2_fma(f2_make(es[i], es[i + 1]), KS, f2_mul(f2_make(et[i], et[i + 1]), NKT)), o[q], o[q + 1]);
            }
            my_out[v * 32 + lane] = E::pack(o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            T* out = static_cast<T*>(p.dS);
            constexpr int kWarpElems = kRmWarpVecs * VE;             // 1024 consecutive row elements
            if (p.perm == nullptr) {
                tma_bulk_s2g(out + (size_t)u * (kThreads * EPT) + warp * kWarpElems, my_out, kWarpElems * 2);
            } else {
                // gathered channels: one store per channel piece
                const unsigned b = u / (unsigned)G, grp = u - b * (unsigned)G;
                int e = warp * kWarpElems, left = kWarpElems;
                const unsigned char* src = reinterpret_cast<const unsigned char*>(my_out);
                while (left > 0) {
                    const int j = e / p.HW, pos = e - j * p.HW, n = min(left, p.HW - pos);
                    tma_bulk_s2g(out + ((size_t)b * p.C + p.perm[grp * p.l[0].g + j]) * p.HW + pos, src, (uint32_t)n * 2u);
                    src += (size_t)n * 2;
                    e += n;
                    left -= n;
                }
            }
            tma_bulk_commit();
        }
        if (next) read_refs(recs, sig, th);
        par ^= 1;
        u = un;
        have = next;
        SD_RT_MARK(5)
    }
    SD_RT_WRITE
    if (lane == 0) tma_bulk_wait0();             // the warp's last store is complete

    // ---- loss: per-CTA partials, last CTA sums them in a fixed order
    if (warp == 0) {
        const float cta_kl = warp_sum(ps.q[lane][5]);
        unsigned ticket = 0;
        if (lane == 0) {
            __stcg(&p.cta_part[blockIdx.x], cta_kl);
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
                *p.l[0].loss = (float)((double)p.l[0].loss_scale * kl);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

// ====================================================================================================
// packed short rows: several WHOLE rows per CTA pass
// ====================================================================================================
// Rows of 256 ... 8192 elements (16x16 ... 64x64 maps, CD on 512-channel features) would leave most of a 512-thread
// CTA idle or pay two CTA barriers per tiny row.  Here a unit is 512/TPR consecutive rows (they are contiguous in
// memory: no shuffle, C % g == 0), TPR threads x 32 elements each; every thread still keeps its elements in
// registers, reductions stay inside the row's team (sub-warp shuffles for TPR < 32; for TPR > 32 one exchange
// through shared memory), the ring / TMA machinery is the one of kl_rows_tma_kernel.
template <typename T, bool MSE>
__global__ void __launch_bounds__(kThreads, 1) kl_rows_pack_kernel(const RowsParams p) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    constexpr int EPT = kDataRegs;
    constexpr int NV = EPT / VE;

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
    float* red = reinterpret_cast<float*>(full + kStages + 1);      // [2][kWarps][kRedFloats]
    ProducerState& ps = *reinterpret_cast<ProducerState*>(red + 2 * kWarps * kRedFloats);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int TPR = p.pack_tpr;
    const int RPU = p.pack_rows;
    const int nvec_row = TPR * NV;                       // 16-byte vectors per row
    const int R = p.l[0].R;
    const int seg = TPR < 32 ? TPR : 32;                 // lanes of a row inside one warp
    const int tw = TPR >> 5;                             // warps per row (0: several rows per warp)
    const int q = tid / TPR, lt = tid - q * TPR;         // my row within the unit, my position in its team

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        ps.cur.u = blockIdx.x;
        ps.stage = 0;
        ps.free_slots = kStages;
        ps.pol = l2_policy_evict_first();
    }
    __syncthreads();

    auto unit_vecs = [&](long long u) {                  // vectors of unit u (the last one may hold fewer rows)
        const long long r0 = u * RPU;
        const int nrows = (int)(R - r0 < RPU ? R - r0 : RPU);
        return nrows * nvec_row;
    };
    // TMA issue (thread 0): whole units only, as many as there are free slots for
    auto issue_loads = [&](int newly_free) {
        int free_slots = ps.free_slots + newly_free, pstage = ps.stage;
        while (ps.cur.u < p.pack_units) {
            const int nvec = unit_vecs(ps.cur.u);
            const int nslots = (nvec + kSlotVecs - 1) / kSlotVecs;
            if (nslots > free_slots) break;
            const size_t base = (size_t)ps.cur.u * RPU * nvec_row * 16;     // bytes
            for (int sl = 0; sl < nslots; ++sl) {
                const int nv = min(kSlotVecs, nvec - sl * kSlotVecs);
                const uint32_t bytes = (uint32_t)nv * 16u;
                mbar_arrive_expect_tx(&full[pstage], 2u * bytes);
                unsigned char* dst_s = smem + (size_t)pstage * kStageBytes;
                const size_t off = base + (size_t)sl * kSlotBytes;
                tma_bulk_g2s(dst_s, static_cast<const char*>(p.S) + off, bytes, &full[pstage], ps.pol);
                tma_bulk_g2s(dst_s + kSlotBytes, static_cast<const char*>(p.T) + off, bytes, &full[pstage], ps.pol);
                if (++pstage == kStages) pstage = 0;
            }
            free_slots -= nslots;
            ps.cur.u += gridDim.x;
        }
        ps.free_slots = free_slots;
        ps.stage = pstage;
    };
    if (tid == 0) issue_loads(0);

    const float c2 = p.l[0].c2;
    float s[EPT], t[EPT];
    float my_kl = 0.f, my_sq = 0.f;
    int stage = 0;
    uint32_t phase = 0;
    int par = 0;

    for (long long u = blockIdx.x; u < p.pack_units; u += gridDim.x) {
        const int nvec = unit_vecs(u);
        const int nslots = (nvec + kSlotVecs - 1) / kSlotVecs;
        const int nrows = nvec / nvec_row;
        const bool active = q < nrows;
        const int cv0 = q * nvec_row + lt;               // my first vector of the unit; the others follow at stride TPR
        if (active) {
            // a row never straddles slots (nvec_row <= 1024 divides the slot): one barrier per thread
            const int sl = cv0 / kSlotVecs;
            int st = stage + sl;
            uint32_t ph = phase;
            if (st >= kStages) {
                st -= kStages;
                ph ^= 1u;
            }
            mbar_wait(&full[st], ph);
            const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)st * kStageBytes) + (cv0 - sl * kSlotVecs);
            const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)st * kStageBytes + kSlotBytes) + (cv0 - sl * kSlotVecs);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                E::unpack(bs[j * TPR], &s[j * VE]);
                E::unpack(bt[j * TPR], &t[j * VE]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                s[i] = kPadValue;
                t[i] = kPadValue;
            }
        }
        stage += nslots;
        if (stage >= kStages) {
            stage -= kStages;
            phase ^= 1u;
        }

        float ms = fmaxf(s[0], kMaxFloor), mt = fmaxf(t[0], kMaxFloor);
#pragma unroll
        for (int i = 1; i < EPT; ++i) {
            ms = fmaxf(ms, s[i]);
            mt = fmaxf(mt, t[i]);
        }
        // the maxima consumed every shared-memory read: hand the unit's slots back
        __syncthreads();
        if (tid == 0) issue_loads(nslots);

        float zt = 0.f, dd = 0.f, a = 0.f;       // zs = zt - dd: see kl_rows_tma_kernel
        const float ms2 = __fmul_rn(ms, c2), mt2 = __fmul_rn(mt, c2);
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            if (MSE) {
                const float d = t[i] - s[i];
                my_sq = fmaf(d, d, my_sq);
            }
            const float as = fmaf(s[i], c2, -ms2), at = fmaf(t[i], c2, -mt2);
            const float es = fast_exp2(as);
            const float et = fast_exp2(at);
            zt += et;
            dd += et - es;
            a = fmaf(et, at - as, a);
            if (!MSE) {
                s[i] = es;
                t[i] = et;
            }
        }
        const float zs = zt - dd;
        // ---- the row's lanes of this warp
        float Ms = ms, Mt = mt;
        for (int o = seg >> 1; o > 0; o >>= 1) {
            Ms = fmaxf(Ms, __shfl_xor_sync(0xffffffffu, Ms, o));
            Mt = fmaxf(Mt, __shfl_xor_sync(0xffffffffu, Mt, o));
        }
        float Zs, Zt, A, DD;
        {
            const float fs = ref_factor(ms, Ms, c2), ft = ref_factor(mt, Mt, c2);
            const float gx = merge_shift2(ms2, mt2, __fmul_rn(Ms, c2), __fmul_rn(Mt, c2));
            Zs = zs * fs;
            Zt = zt * ft;
            A = fmaf(Zt, gx, a * ft);
            DD = fmaf(zs, factor_diff(fs, ft, gx), dd * ft);
        }
        for (int o = seg >> 1; o > 0; o >>= 1) {
            Zs += __shfl_xor_sync(0xffffffffu, Zs, o);
            Zt += __shfl_xor_sync(0xffffffffu, Zt, o);
            A += __shfl_xor_sync(0xffffffffu, A, o);
            DD += __shfl_xor_sync(0xffffffffu, DD, o);
        }
        // ---- the row's warps (rows wider than a warp): one exchange through shared memory
        if (tw > 1) {
            if (lane == 0) {
                float* my_red = red + (par * kWarps + warp) * kRedFloats;
                reinterpret_cast<float4*>(my_red)[0] = make_float4(Ms, Mt, Zs, Zt);
                my_red[4] = A;
                my_red[5] = DD;
            }
            __syncthreads();
            const float* rq = red + (par * kWarps + (warp / tw) * tw + (lane & (tw - 1))) * kRedFloats;
            const float4 r0 = reinterpret_cast<const float4*>(rq)[0];
            const float r1 = rq[4], r2 = rq[5];
            float M2s = r0.x, M2t = r0.y;
            for (int o = tw >> 1; o > 0; o >>= 1) {
                M2s = fmaxf(M2s, __shfl_xor_sync(0xffffffffu, M2s, o));
                M2t = fmaxf(M2t, __shfl_xor_sync(0xffffffffu, M2t, o));
            }
            const float fs = ref_factor(r0.x, M2s, c2), ft = ref_factor(r0.y, M2t, c2);
            const float gx = merge_shift(r0.x, r0.y, M2s, M2t, c2);
            float z2s = r0.z * fs, z2t = r0.w * ft, a2 = fmaf(r0.w * ft, gx, r1 * ft);
            float d2 = fmaf(r0.z, factor_diff(fs, ft, gx), r2 * ft);
            for (int o = tw >> 1; o > 0; o >>= 1) {
                z2s += __shfl_xor_sync(0xffffffffu, z2s, o);
                z2t += __shfl_xor_sync(0xffffffffu, z2t, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                d2 += __shfl_xor_sync(0xffffffffu, d2, o);
            }
            Ms = M2s;
            Mt = M2t;
            Zs = z2s;
            Zt = z2t;
            A = a2;
            DD = d2;
            par ^= 1;
        }
        if (active && lt == 0) {
            const float kl = kl_from_stats(Zs, Zt, A, DD);
            if (p.l[0].row_kl) p.l[0].row_kl[u * RPU + q] = kl;
            my_kl += kl;
        }
        // ---- gradient straight from registers
        if (active) {
            float coef = p.l[0].coef;
            if (p.grad_out[0] != nullptr) coef *= __ldg(p.grad_out[0]);
            const float ks = coef * ref_factor(ms, Ms, c2) / Zs;
            const float kt = coef * ref_factor(mt, Mt, c2) / Zt;
            // (MSE: the exponents are computed again below; seen as the same expressions as in the statistics loop they
            //  were all kept alive across the reductions - spilled to local memory, which this kernel's shared-memory
            //  carve-out leaves no L1 for)
            float ms2g = ms2, mt2g = mt2;
            if (MSE) asm volatile("" : "+f"(ms2g), "+f"(mt2g));
            vec_t* dst = reinterpret_cast<vec_t*>(static_cast<T*>(p.dS)) + (size_t)u * RPU * nvec_row + cv0;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                float o[VE];
#pragma unroll
                for (int k = 0; k < VE; ++k) {
                    const int i = j * VE + k;
                    if (MSE) {
                        const float es = fast_exp2(fmaf(s[i], c2, -ms2g));
                        const float et = fast_exp2(fmaf(t[i], c2, -mt2g));
                        o[k] = fmaf(es, ks, -et * kt) + p.mse_gcoef * (s[i] - t[i]);
                    } else {
                        o[k] = fmaf(s[i], ks, -t[i] * kt);
                    }
                }
                dst[j * TPR] = E::pack(o);
            }
        }
    }

    // ================================ loss: thread partials -> CTA partial (fixed order) -> last CTA sums
    __syncthreads();
    {
        const float wk = warp_sum(my_kl), wq = MSE ? warp_sum(my_sq) : 0.f;
        if (lane == 0) {
            red[warp] = wk;
            red[kWarps + warp] = wq;
        }
    }
    __syncthreads();
    if (warp == 0) {
        unsigned ticket = 0;
        if (lane == 0) {
            float ck = 0.f, cq = 0.f;
            for (int w = 0; w < kWarps; ++w) {
                ck += red[w];
                cq += red[kWarps + w];
            }
            __stcg(&p.cta_part[blockIdx.x], ck);
            __stcg(&p.cta_part[kMaxLosses * kMaxGrid + blockIdx.x], cq);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double kl = 0.0, sq = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
                kl += (double)__ldcg(&p.cta_part[i]);
                sq += (double)__ldcg(&p.cta_part[kMaxLosses * kMaxGrid + i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kl += __shfl_down_sync(0xffffffffu, kl, o);
                sq += __shfl_down_sync(0xffffffffu, sq, o);
            }
            if (lane == 0) {
                *p.l[0].loss = (float)((double)p.l[0].loss_scale * kl);
                if (MSE && p.mse_loss) *p.mse_loss = (float)((double)p.mse_scale * sq);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

// ====================================================================================================
// generic path (single loss): unit = (sample, logical channel, plane chunk); no alignment requirement
// ====================================================================================================
constexpr int kGenThreads = 256;

struct GenUnit {
    int b, cl, k, n;
    size_t off;  // element offset of the chunk
};
__device__ __forceinline__ GenUnit decode_gen(const RowsParams& p, long long u) {
    GenUnit x;
    x.k = (int)(u % p.KC);
    const long long bc = u / p.KC;
    x.cl = (int)(bc % p.C);
    x.b = (int)(bc / p.C);
    const int ch = p.perm ? p.perm[x.cl] : x.cl;
    x.off = ((size_t)x.b * p.C + ch) * p.HW + (size_t)x.k * kGenericChunk;
    x.n = min(kGenericChunk, p.HW - x.k * kGenericChunk);
    return x;
}

template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* scratch /* [N][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[i * 8 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < kGenThreads / 32; ++w) acc += scratch[i * 8 + w];
        v[i] = acc;
    }
}
__device__ __forceinline__ void block_max2(float& a, float& b, float* scratch /* [2][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_max(a);
    b = warp_max(b);
    __syncthreads();
    if (lane == 0) {
        scratch[warp] = a;
        scratch[8 + warp] = b;
    }
    __syncthreads();
    a = -INFINITY;
    b = -INFINITY;
#pragma unroll
    for (int w = 0; w < kGenThreads / 32; ++w) {
        a = fmaxf(a, scratch[w]);
        b = fmaxf(b, scratch[8 + w]);
    }
}

template <typename T>
__global__ void __launch_bounds__(kGenThreads) kl_rows_generic_stats(const RowsParams p) {
    using E = Elem<T>;
    __shared__ float scratch[5 * 8];
    const long long u = blockIdx.x;
    const GenUnit x = decode_gen(p, u);
    const float c2 = p.l[0].c2;
    const T* s = static_cast<const T*>(p.S) + x.off;
    const T* t = static_cast<const T*>(p.T) + x.off;
    float ms = -INFINITY, mt = -INFINITY;
    for (int i = threadIdx.x; i < x.n; i += kGenThreads) {
        ms = fmaxf(ms, E::load(s + i));
        mt = fmaxf(mt, E::load(t + i));
    }
    block_max2(ms, mt, scratch);
    const float ms2 = __fmul_rn(ms, c2), mt2 = __fmul_rn(mt, c2);
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // zs, zt, a, sq, dd
    for (int i = threadIdx.x; i < x.n; i += kGenThreads) {
        const float a = E::load(s + i), b = E::load(t + i);
        const float d = b - a;
        const float as = fmaf(a, c2, -ms2), at = fmaf(b, c2, -mt2);
        const float es = fast_exp2(as);
        const float et = fast_exp2(at);
        acc[0] += es;
        acc[1] += et;
        acc[2] = fmaf(et, at - as, acc[2]);
        acc[3] = fmaf(d, d, acc[3]);
        acc[4] += et - es;
    }
    block_sum<5>(acc, scratch);
    if (threadIdx.x == 0) {
        float* slot = p.unit_part + (size_t)u * kPartWords;
        slot[0] = ms;
        slot[1] = acc[0];
        slot[2] = mt;
        slot[3] = acc[1];
        slot[4] = acc[2];
        slot[5] = acc[3];
        slot[6] = acc[4];
    }
}

template <typename T>
__global__ void __launch_bounds__(kGenThreads) kl_rows_generic_grad(const RowsParams p) {
    using E = Elem<T>;
    __shared__ float row_stat[8];
    const long long u = blockIdx.x;
    const GenUnit x = decode_gen(p, u);
    const float c2 = p.l[0].c2;
    const int g = p.l[0].g;
    const int grp = x.cl / g;
    const int c0 = grp * g;
    const int g_real = min(g, p.C - c0);
    const long long u0 = ((long long)x.b * p.C + c0) * p.KC;
    const int nparts = g_real * p.KC;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        RowStat acc = rowstat_empty();
        for (int k = lane; k < nparts; k += 32) {
            const float* q = p.unit_part + (size_t)(u0 + k) * kPartWords;
            acc = rowstat_merge(acc, RowStat{q[0], q[1], q[2], q[3], q[4], q[6]}, c2);
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
            acc = rowstat_merge(acc, other, c2);
        }
        if (lane == 0) {
            row_stat[0] = acc.ms;
            row_stat[1] = acc.zs;
            row_stat[2] = acc.mt;
            row_stat[3] = acc.zt;
            row_stat[4] = acc.a;
            row_stat[5] = acc.dd;
        }
    }
    __syncthreads();
    const float Ms = row_stat[0], Zs = row_stat[1], Mt = row_stat[2], Zt = row_stat[3], A = row_stat[4];
    if (threadIdx.x == 0 && u == u0) {
        const float kl = kl_from_stats(Zs, Zt, A, row_stat[5]);
        p.l[0].row_kl[x.b * p.l[0].G + grp] = kl;
    }
    const float ms2 = __fmul_rn(Ms, c2), mt2 = __fmul_rn(Mt, c2);
    const float ks = p.l[0].coef / Zs, kt = p.l[0].coef / Zt;
    const T* s = static_cast<const T*>(p.S) + x.off;
    const T* t = static_cast<const T*>(p.T) + x.off;
    T* o = static_cast<T*>(p.dS) + x.off;
    for (int i = threadIdx.x; i < x.n; i += kGenThreads) {
        const float a = E::load(s + i), b = E::load(t + i);
        const float es = fast_exp2(fmaf(a, c2, -ms2));
        const float et = fast_exp2(fmaf(b, c2, -mt2));
        E::store(o + i, fmaf(es, ks, -et * kt) + p.mse_gcoef * (a - b));
    }
}

// one CTA: loss = loss_scale * sum(row_kl) (fixed order); mse_loss = mse_scale * sum(unit sq partials)
__global__ void __launch_bounds__(1024) kl_rows_generic_finalize(const RowsParams p, long long n_units) {
    __shared__ double sh[2][32];
    double kl = 0.0, sq = 0.0;
    for (int i = threadIdx.x; i < p.l[0].R; i += 1024) kl += (double)p.l[0].row_kl[i];
    if (p.mse_loss)
        for (long long i = threadIdx.x; i < n_units; i += 1024) sq += (double)p.unit_part[(size_t)i * kPartWords + 5];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kl += __shfl_down_sync(0xffffffffu, kl, o);
        sq += __shfl_down_sync(0xffffffffu, sq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = kl;
        sh[1][threadIdx.x >> 5] = sq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; ++w) {
            a += sh[0][w];
            b += sh[1][w];
        }
        *p.l[0].loss = (float)((double)p.l[0].loss_scale * a);
        if (p.mse_loss) *p.mse_loss = (float)((double)p.mse_scale * b);
    }
}

// ====================================================================================================
// host launchers
// ====================================================================================================
template <typename T, bool MSE>
static cudaError_t launch_tma_t(const RowsParams& p, int grid, cudaStream_t stream) {
    auto kern = kl_rows_tma_kernel<T, MSE>;
    static std::atomic<bool> configured[kMaxDevices];  // per instantiation and device
    const int dev = device_slot();
    if (!configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    kern<<<grid, kThreads, kRowsSmemBytes, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_kl_rows_tma(const RowsParams& p, bool bf16, int grid, cudaStream_t stream) {
    const bool mse = p.mse_gcoef != 0.f || p.mse_loss != nullptr;
    if (bf16) {
        return mse ? launch_tma_t<__nv_bfloat16, true>(p, grid, stream)
                   : launch_tma_t<__nv_bfloat16, false>(p, grid, stream);
    }
    return mse ? launch_tma_t<float, true>(p, grid, stream) : launch_tma_t<float, false>(p, grid, stream);
}

// bf16 rows of exactly kl_rows_tma_chunk_capacity() elements (the caller checks: every unit is a whole row)
cudaError_t launch_kl_rows_rm(const RowsParams& p, int grid, cudaStream_t stream) {
    auto kern = kl_rows_rm_kernel;
    static std::atomic<bool> configured[kMaxDevices];
    const int dev = device_slot();
    if (!configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRmSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    kern<<<grid, kThreads, kRmSmemBytes, stream>>>(p);
    return cudaGetLastError();
}
int kl_rows_tma_chunk_capacity() { return kThreads * kDataRegs; }

template <typename T, bool MSE>
static cudaError_t launch_pack_t(const RowsParams& p, int grid, cudaStream_t stream) {
    auto kern = kl_rows_pack_kernel<T, MSE>;
    static std::atomic<bool> configured[kMaxDevices];  // per instantiation and device
    const int dev = device_slot();
    if (!configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    kern<<<grid, kThreads, kRowsSmemBytes, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_kl_rows_pack(const RowsParams& p, bool bf16, int grid, cudaStream_t stream) {
    const bool mse = p.mse_gcoef != 0.f || p.mse_loss != nullptr;
    if (bf16) {
        return mse ? launch_pack_t<__nv_bfloat16, true>(p, grid, stream) : launch_pack_t<__nv_bfloat16, false>(p, grid, stream);
    }
    return mse ? launch_pack_t<float, true>(p, grid, stream) : launch_pack_t<float, false>(p, grid, stream);
}

cudaError_t launch_kl_rows_generic(const RowsParams& p, bool bf16, cudaStream_t stream) {
    const long long units = (long long)p.B * p.C * p.KC;
    if (bf16) {
        kl_rows_generic_stats<__nv_bfloat16><<<(unsigned)units, kGenThreads, 0, stream>>>(p);
        kl_rows_generic_grad<__nv_bfloat16><<<(unsigned)units, kGenThreads, 0, stream>>>(p);
    } else {
        kl_rows_generic_stats<float><<<(unsigned)units, kGenThreads, 0, stream>>>(p);
        kl_rows_generic_grad<float><<<(unsigned)units, kGenThreads, 0, stream>>>(p);
    }
    kl_rows_generic_finalize<<<1, 1024, 0, stream>>>(p, units);
    return cudaGetLastError();
}

}  // namespace sd
