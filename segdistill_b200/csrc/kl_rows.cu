// Row-wise softmax-KL, forward + backward fused (CD / CGD / plain KLDLoss), one or two losses
// over the same (student, teacher) pair per pass.
//
// Replaces the ATen chain of mmseg/models/distillation/losses.py:35-42 (channel gather),
// :50-58 (group reshape, -1e9 pad) and :108-112 (div, log_softmax, softmax, kl_div, mul)
// plus its autograd backward.  Row = `g` consecutive (gathered) channels x HW.
//
//   kl_rows_tma_kernel      persistent, one CTA per SM.  Row chunks stream into a 7-stage
//                           shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//                           complete_tx) issued by one elected thread as soon as a slot has been
//                           drained; the 16 warps pull a chunk into REGISTERS (32 elements of S and
//                           of T per thread), exponentiate against a thread-local maximum (no block
//                           barrier before the exponentials), combine (max, sum) partials with
//                           warp shuffles and ONE CTA barrier, and write dS straight from
//                           registers.  HBM traffic is the algorithmic 12 B/elem (fp32) /
//                           6 B/elem (bf16): S and T are read once, dS written once.
//                           NL = 2 fuses two losses with nested rows (e.g. CD + CGD on the same
//                           logits): still one read of S and T and one write of the summed
//                           gradient.  Rows longer than one chunk are split over several CTAs which
//                           exchange partials through epoch-tagged 8-byte packets in global memory
//                           (no atomics, no counters to reset; all CTAs co-resident: cooperative
//                           launch).
//   kl_rows_generic_*       any alignment / any row length: three plain passes.
#include "rows_common.cuh"

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
constexpr size_t kRowsSmemBytes = (size_t)kStages * kStageBytes + (kStages + 1) * sizeof(uint64_t) +
                                  2 * kWarps * kRedFloats * sizeof(float) + kMaxLosses * 8 * sizeof(float);
template <typename T, int NL, bool MSE>
__global__ void __launch_bounds__(kThreads, 1) kl_rows_tma_kernel(const RowsParams p) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    constexpr int EPT = kDataRegs / NL;  // elements per thread and tensor
    constexpr int NV = EPT / VE;         // 16-byte vectors per thread and tensor
    constexpr int NJ = (NV + kSlotVecRows - 1) / kSlotVecRows;  // ring slots of a whole chunk
    static_assert(NV >= 1 && EPT % VE == 0, "layout");
    static_assert(!(MSE && NL > 1), "the fused MSE term is a single-loss feature");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
    float* red = reinterpret_cast<float*>(full + kStages + 1);      // [2][kWarps][kRedFloats]
    float* bcast = red + 2 * kWarps * kRedFloats;                   // [kMaxLosses][8]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (p.run_if != nullptr && *p.run_if == 0u) return;  // cancelled backward re-run (uniform over the grid)

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // ================================ TMA issue (thread 0 only) ================================
    // Slots are refilled in ring order.  `free_slots` counts slots every thread has drained: a
    // unit's slots are released by the CTA barrier that follows its ring->register copy.
    UnitCursor prod;
    prod.init(p, blockIdx.x);
    int prod_v0 = 0, prod_stage = 0, free_slots = kStages;
    uint64_t pol = 0;
    if (tid == 0) pol = l2_policy_evict_first();
    auto issue_loads = [&]() {
        while (free_slots > 0 && prod.u < p.total_units) {
            const Unit x = decode_unit(p, prod.b, prod.r);
            const int nvec = x.len / VE;
            const int nv = min(kSlotVecRows * kThreads, nvec - prod_v0);
            const uint32_t bytes = (uint32_t)nv * 16u;
            mbar_arrive_expect_tx(&full[prod_stage], 2u * bytes);
            unsigned char* dst_s = smem + (size_t)prod_stage * kStageBytes;
            unsigned char* dst_t = dst_s + kSlotBytes;
            const int e = x.e0 + prod_v0 * VE;
            if (p.perm == nullptr) {
                const size_t off = (((size_t)x.b * p.C + (size_t)x.grp * p.l[0].g) * p.HW + e) * sizeof(T);
                tma_bulk_g2s(dst_s, static_cast<const char*>(p.S) + off, bytes, &full[prod_stage], pol);
                tma_bulk_g2s(dst_t, static_cast<const char*>(p.T) + off, bytes, &full[prod_stage], pol);
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
                    tma_bulk_g2s(dst_s + doff, static_cast<const char*>(p.S) + off, nb, &full[prod_stage], pol);
                    tma_bulk_g2s(dst_t + doff, static_cast<const char*>(p.T) + off, nb, &full[prod_stage], pol);
                    doff += nb;
                    cur += n;
                    remaining -= n;
                }
            }
            prod_v0 += kSlotVecs;
            if (prod_v0 >= nvec) {
                prod_v0 = 0;
                prod.advance(p, (int)gridDim.x);
            }
            if (++prod_stage == kStages) prod_stage = 0;
            --free_slots;
        }
    };
    if (tid == 0) issue_loads();

    // ================================ 16 warps, chunk lives in registers ================================
    float c2[NL], coef[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        c2[k] = p.l[k].c2;
        coef[k] = p.l[k].coef;
        if (p.grad_out[k] != nullptr) coef[k] *= *p.grad_out[k];
    }
    // s/t: raw values, then the exponentials of loss NL-1; xs/xt: exponentials of loss 0 when NL == 2
    float s[EPT], t[EPT];
    float xs[NL > 1 ? EPT : 1], xt[NL > 1 ? EPT : 1];
    float cta_kl[NL], cta_sq = 0.f;  // accumulated by thread 0 in unit order (deterministic)
#pragma unroll
    for (int k = 0; k < NL; ++k) cta_kl[k] = 0.f;
    int stage = 0;
    uint32_t phase = 0;
    int par = 0;

    UnitCursor cur;
    for (cur.init(p, blockIdx.x); cur.u < p.total_units; cur.advance(p, (int)gridDim.x)) {
        const long long u = cur.u;
        const Unit x = decode_unit(p, cur.b, cur.r);
        const int nvec = x.len / VE;
        const bool whole = nvec == NV * kThreads;  // every thread holds NV vectors

        // ---- ring -> registers
        const int nslots = (nvec + kSlotVecs - 1) / kSlotVecs;
        if (whole) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                mbar_wait(&full[stage], phase);
                const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes);
                const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes + kSlotBytes);
#pragma unroll
                for (int r = 0; r < kSlotVecRows; ++r) {
                    const int v = j * kSlotVecRows + r;
                    if (v < NV) {
                        E::unpack(bs[r * kThreads + tid], &s[v * VE]);
                        E::unpack(bt[r * kThreads + tid], &t[v * VE]);
                    }
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
                        if (v < NV && v * kThreads + tid < nvec) {
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
        float ms = fmaxf(s[0], kMaxFloor), mt = fmaxf(t[0], kMaxFloor);
#pragma unroll
        for (int i = 1; i < EPT; ++i) {
            ms = fmaxf(ms, s[i]);
            mt = fmaxf(mt, t[i]);
        }

        // every thread holds its elements in registers (the maxima consumed every shared-memory read):
        // hand the unit's slots back to the TMA thread now, so the next loads fly during the exponentials
        __syncthreads();
        if (tid == 0) {
            free_slots += nslots;
            issue_loads();
        }

        // ---- exponentials (kept in registers), thread-partial sums relative to (ms, mt)
        float zs[NL], zt[NL], a[NL], sq = 0.f;
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            zs[k] = 0.f;
            zt[k] = 0.f;
            a[k] = 0.f;
        }
        {
            float ms2[NL], mt2[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                ms2[k] = ms * c2[k];
                mt2[k] = mt * c2[k];
            }
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const float d = t[i] - s[i];
                if (MSE) sq = fmaf(d, d, sq);
                if (NL > 1) {
                    const float es0 = fast_exp2(fmaf(s[i], c2[0], -ms2[0]));
                    const float et0 = fast_exp2(fmaf(t[i], c2[0], -mt2[0]));
                    zs[0] += es0;
                    zt[0] += et0;
                    a[0] = fmaf(et0, d, a[0]);
                    xs[i] = es0;
                    xt[i] = et0;
                }
                const float es = fast_exp2(fmaf(s[i], c2[NL - 1], -ms2[NL - 1]));
                const float et = fast_exp2(fmaf(t[i], c2[NL - 1], -mt2[NL - 1]));
                zs[NL - 1] += es;
                zt[NL - 1] += et;
                a[NL - 1] = fmaf(et, d, a[NL - 1]);
                if (!MSE) {
                    s[i] = es;
                    t[i] = et;
                }
            }
        }

        // ---- warp: raw maxima are common to all losses, sums are rescaled to them
        const float msw = warp_max(ms), mtw = warp_max(mt);
        float rec[kRedFloats];
#pragma unroll
        for (int i = 0; i < kRedFloats; ++i) rec[i] = 0.f;
        rec[0] = msw;
        rec[1] = mtw;
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const float fs = fast_exp2((ms - msw) * c2[k]);
            const float ft = fast_exp2((mt - mtw) * c2[k]);
            rec[2 + 3 * k] = warp_sum(zs[k] * fs);
            rec[3 + 3 * k] = warp_sum(zt[k] * ft);
            rec[4 + 3 * k] = warp_sum(a[k] * ft);
        }
        if (MSE) rec[5] = warp_sum(sq);
        float* my_red = red + (par * kWarps + warp) * kRedFloats;
        if (lane == 0) {
            reinterpret_cast<float4*>(my_red)[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
            reinterpret_cast<float4*>(my_red)[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
        }
        __syncthreads();

        // ---- CTA: every warp merges the 16 warp records (lanes l and l+16 mirror each other)
        float Ms, Mt, Zs[NL], Zt[NL], A[NL], SQ = 0.f;
        {
            const float* q = red + (par * kWarps + (lane & 15)) * kRedFloats;
            const float4 r0 = reinterpret_cast<const float4*>(q)[0];
            const float4 r1 = reinterpret_cast<const float4*>(q)[1];
            const float rr[kRedFloats] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            Ms = max16(rr[0]);
            Mt = max16(rr[1]);
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const float fs = fast_exp2((rr[0] - Ms) * c2[k]);
                const float ft = fast_exp2((rr[1] - Mt) * c2[k]);
                Zs[k] = sum16(rr[2 + 3 * k] * fs);
                Zt[k] = sum16(rr[3 + 3 * k] * ft);
                A[k] = sum16(rr[4 + 3 * k] * ft);
            }
            if (MSE) SQ = sum16(rr[5]);
        }
        par ^= 1;

        // ---- rows split over several CTAs: exchange partials through epoch-tagged packets
        int rown[NL];          // units in this unit's row of loss k
        long long rowu[NL];    // first unit of that row
        int rowi[NL];          // row index (for row_kl)
        rown[0] = x.nch;
        rowu[0] = u - x.ck;
        rowi[0] = x.b * p.l[0].G + x.grp;
#pragma unroll
        for (int k = 1; k < NL; ++k) {
            const int m = p.l[k].m;
            const int rk = x.grp / m;
            const int j0 = rk * m;
            const int j1 = min(j0 + m, p.l[0].G);
            const int us = unit_start(p, j0);
            rown[k] = unit_start(p, j1) - us;
            rowu[k] = (long long)x.b * p.units_per_sample + us;
            rowi[k] = x.b * p.l[k].G + rk;
        }
        bool split = false;
#pragma unroll
        for (int k = 0; k < NL; ++k) split = split || rown[k] > 1;

        float Msr[NL], Mtr[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            Msr[k] = Ms;
            Mtr[k] = Mt;
        }
        if (split) {
            if (warp == 0) {
                // publish this unit's partials: one 8-byte {value, epoch} word per lane
                {
                    float val = 0.f;
#pragma unroll
                    for (int k = 0; k < NL; ++k) {
                        if (lane == 6 * k + 0) val = Ms;
                        if (lane == 6 * k + 1) val = Zs[k];
                        if (lane == 6 * k + 2) val = Mt;
                        if (lane == 6 * k + 3) val = Zt[k];
                        if (lane == 6 * k + 4) val = A[k];
                    }
                    if (lane < 6 * NL)
                        st_relaxed_u64(p.pkt + (size_t)u * kPktWords + lane,
                                       ((unsigned long long)p.epoch << 32) | __float_as_uint(val));
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    if (rown[k] > 1) {
                        RowStat acc = rowstat_empty();
                        for (int j = lane; j < rown[k]; j += 32) {
                            const unsigned long long* q = p.pkt + (size_t)(rowu[k] + j) * kPktWords + 6 * k;
                            unsigned long long w[5];
                            unsigned spins = 0;
                            for (;;) {
                                bool ok = true;
#pragma unroll
                                for (int i = 0; i < 5; ++i) {
                                    w[i] = ld_relaxed_u64(q + i);
                                    ok = ok && (unsigned)(w[i] >> 32) == p.epoch;
                                }
                                if (ok) break;
                                if (++spins > kSpinLimit) {
                                    atomicExch(&p.ctrl[1], 1u);  // never expected: reported by the host wrapper
                                    break;
                                }
                                __nanosleep(40);
                            }
                            RowStat r;
                            r.ms = __uint_as_float((unsigned)w[0]);
                            r.zs = __uint_as_float((unsigned)w[1]);
                            r.mt = __uint_as_float((unsigned)w[2]);
                            r.zt = __uint_as_float((unsigned)w[3]);
                            r.a = __uint_as_float((unsigned)w[4]);
                            acc = rowstat_merge(acc, r, c2[k]);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            RowStat other;
                            other.ms = __shfl_xor_sync(0xffffffffu, acc.ms, o);
                            other.zs = __shfl_xor_sync(0xffffffffu, acc.zs, o);
                            other.mt = __shfl_xor_sync(0xffffffffu, acc.mt, o);
                            other.zt = __shfl_xor_sync(0xffffffffu, acc.zt, o);
                            other.a = __shfl_xor_sync(0xffffffffu, acc.a, o);
                            acc = rowstat_merge(acc, other, c2[k]);
                        }
                        if (lane == 0) {
                            bcast[8 * k + 0] = acc.ms;
                            bcast[8 * k + 1] = acc.zs;
                            bcast[8 * k + 2] = acc.mt;
                            bcast[8 * k + 3] = acc.zt;
                            bcast[8 * k + 4] = acc.a;
                        }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                if (rown[k] > 1) {
                    Msr[k] = bcast[8 * k + 0];
                    Zs[k] = bcast[8 * k + 1];
                    Mtr[k] = bcast[8 * k + 2];
                    Zt[k] = bcast[8 * k + 3];
                    A[k] = bcast[8 * k + 4];
                }
            }
        }

        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                if (u == rowu[k]) {
                    // KL(p||q) = sum p (t - s)/tau - lse_t + lse_s
                    const float kl = p.l[k].inv_tau * A[k] / Zt[k] -
                                     ((Mtr[k] - Msr[k]) * p.l[k].inv_tau + (logf(Zt[k]) - logf(Zs[k])));
                    if (p.l[k].row_kl) p.l[k].row_kl[rowi[k]] = kl;
                    cta_kl[k] += kl;
                }
            }
            if (MSE) cta_sq += SQ;
        }

        // ---- gradient straight from registers
        float ks[NL], kt[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            ks[k] = coef[k] * fast_exp2((ms - Msr[k]) * c2[k]) / Zs[k];
            kt[k] = coef[k] * fast_exp2((mt - Mtr[k]) * c2[k]) / Zt[k];
        }
        const float ms2 = ms * c2[0], mt2 = mt * c2[0];
        auto grad_vec = [&](int v, float* o) {
#pragma unroll
            for (int q = 0; q < VE; ++q) {
                const int i = v * VE + q;
                if (MSE) {
                    const float es = fast_exp2(fmaf(s[i], c2[0], -ms2));
                    const float et = fast_exp2(fmaf(t[i], c2[0], -mt2));
                    o[q] = fmaf(es, ks[0], -et * kt[0]) + p.mse_gcoef * (s[i] - t[i]);
                } else if (NL > 1) {
                    o[q] = fmaf(xs[i], ks[0], -xt[i] * kt[0]) + fmaf(s[i], ks[NL - 1], -t[i] * kt[NL - 1]);
                } else {
                    o[q] = fmaf(s[i], ks[0], -t[i] * kt[0]);
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
    }

    // ================================ loss: per-CTA partials, last CTA sums them in a fixed order ================================
    if (warp == 0) {
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
#pragma unroll
                for (int k = 0; k < NL; ++k) *p.l[k].loss = (float)((double)p.l[k].loss_scale * acc[k]);
                if (MSE && p.mse_loss) *p.mse_loss = (float)((double)p.mse_scale * acc[NL]);
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
    __shared__ float scratch[4 * 8];
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
    const float ms2 = ms * c2, mt2 = mt * c2;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // zs, zt, a, sq
    for (int i = threadIdx.x; i < x.n; i += kGenThreads) {
        const float a = E::load(s + i), b = E::load(t + i);
        const float d = b - a;
        const float et = fast_exp2(fmaf(b, c2, -mt2));
        acc[0] += fast_exp2(fmaf(a, c2, -ms2));
        acc[1] += et;
        acc[2] = fmaf(et, d, acc[2]);
        acc[3] = fmaf(d, d, acc[3]);
    }
    block_sum<4>(acc, scratch);
    if (threadIdx.x == 0) {
        float* slot = p.unit_part + (size_t)u * kPartWords;
        slot[0] = ms;
        slot[1] = acc[0];
        slot[2] = mt;
        slot[3] = acc[1];
        slot[4] = acc[2];
        slot[5] = acc[3];
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
            acc = rowstat_merge(acc, RowStat{q[0], q[1], q[2], q[3], q[4]}, c2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            RowStat other;
            other.ms = __shfl_xor_sync(0xffffffffu, acc.ms, o);
            other.zs = __shfl_xor_sync(0xffffffffu, acc.zs, o);
            other.mt = __shfl_xor_sync(0xffffffffu, acc.mt, o);
            other.zt = __shfl_xor_sync(0xffffffffu, acc.zt, o);
            other.a = __shfl_xor_sync(0xffffffffu, acc.a, o);
            acc = rowstat_merge(acc, other, c2);
        }
        if (lane == 0) {
            row_stat[0] = acc.ms;
            row_stat[1] = acc.zs;
            row_stat[2] = acc.mt;
            row_stat[3] = acc.zt;
            row_stat[4] = acc.a;
        }
    }
    __syncthreads();
    const float Ms = row_stat[0], Zs = row_stat[1], Mt = row_stat[2], Zt = row_stat[3], A = row_stat[4];
    if (threadIdx.x == 0 && u == u0) {
        const float kl = p.l[0].inv_tau * A / Zt - ((Mt - Ms) * p.l[0].inv_tau + (logf(Zt) - logf(Zs)));
        p.l[0].row_kl[x.b * p.l[0].G + grp] = kl;
    }
    const float ms2 = Ms * c2, mt2 = Mt * c2;
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
template <typename T, int NL, bool MSE>
static cudaError_t launch_tma_t(const RowsParams& p, int grid, bool cooperative, cudaStream_t stream) {
    auto kern = kl_rows_tma_kernel<T, NL, MSE>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kRowsSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = cooperative ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_kl_rows_tma(const RowsParams& p, bool bf16, int grid, bool cooperative, cudaStream_t stream) {
    const bool mse = p.mse_gcoef != 0.f || p.mse_loss != nullptr;
    if (p.nl == 2) {
        return bf16 ? launch_tma_t<__nv_bfloat16, 2, false>(p, grid, cooperative, stream)
                    : launch_tma_t<float, 2, false>(p, grid, cooperative, stream);
    }
    if (bf16) {
        return mse ? launch_tma_t<__nv_bfloat16, 1, true>(p, grid, cooperative, stream)
                   : launch_tma_t<__nv_bfloat16, 1, false>(p, grid, cooperative, stream);
    }
    return mse ? launch_tma_t<float, 1, true>(p, grid, cooperative, stream)
               : launch_tma_t<float, 1, false>(p, grid, cooperative, stream);
}

int kl_rows_tma_chunk_capacity(int nl) { return kThreads * (kDataRegs / nl); }

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
