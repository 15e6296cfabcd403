// Row-wise softmax-KL, forward + backward fused (CD / CGD / plain KLDLoss).
//
// Replaces the ATen chain of mmseg/models/distillation/losses.py:35-42 (channel gather),
// :50-58 (group reshape, -1e9 pad) and :108-112 (div, log_softmax, softmax, kl_div, mul)
// plus its autograd backward.  Row = `g` consecutive (gathered) channels x HW.
//
//   kl_rows_tma_kernel      persistent: row chunks stream into a shared-memory ring with 1-D TMA
//                           bulk copies (cp.async.bulk + mbarrier complete_tx) issued by one
//                           elected thread as soon as a slot has been drained; the 16 warps pull
//                           a chunk (<= 16384 elements) into REGISTERS, reduce max / sum-exp,
//                           and write dS straight from registers.  HBM traffic is the
//                           algorithmic 12 B/elem (fp32) / 6 B/elem (bf16): S and T are read
//                           once, dS written once.  Rows longer than one chunk are split over
//                           several CTAs which exchange (max, sum) partials through global
//                           memory flags (all CTAs co-resident: cooperative launch).
//   kl_rows_generic_*       any alignment / any row length: three plain passes.
#include "common.cuh"
#include "params.h"

namespace sd {

// ------------------------------------------------------------------ configuration
constexpr int kCons = 512;                        // threads per CTA (16 warps x 128 registers)
constexpr int kConsWarps = kCons / 32;
constexpr int kRowsThreads = kCons;
constexpr int kEPT = 32;                          // elements per consumer thread and tensor
constexpr int kChunkCap = kCons * kEPT;           // 16384 elements per chunk
constexpr int kSlotVecRows = 2;                   // 16-byte vectors per thread and ring slot
constexpr int kSlotVecs = kSlotVecRows * kCons;   // 1024 vectors
constexpr int kSlotBytes = kSlotVecs * 16;        // 16 KB per tensor
constexpr int kStageBytes = 2 * kSlotBytes;       // S + T
constexpr int kStages = 7;                        // 224 KB ring
constexpr unsigned kSpinLimit = 1u << 24;
constexpr size_t kRowsSmemBytes = (size_t)kStages * kStageBytes + 2 * kStages * sizeof(uint64_t) +  // (2nd barrier array: spare)
                                  kConsWarps * (sizeof(float2) + sizeof(float4)) + 8 * sizeof(float);

struct Unit {
    int b, grp, ck, nch, row;
    int e0;   // first logical row element of this chunk
    int len;  // elements in this chunk
};

__device__ __forceinline__ Unit decode_unit(const RowsParams& p, long long u) {
    Unit x;
    x.b = (int)(u / p.units_per_sample);
    const int r = (int)(u - (long long)x.b * p.units_per_sample);
    const int full_units = p.G_full * p.nch_full;
    int g_real;
    if (r < full_units) {
        x.grp = r / p.nch_full;
        x.ck = r - x.grp * p.nch_full;
        x.nch = p.nch_full;
        g_real = p.g;
    } else {
        x.grp = p.G_full;
        x.ck = r - full_units;
        x.nch = p.nch_last;
        g_real = p.g_last;
    }
    const int L = g_real * p.HW;
    x.e0 = x.ck * p.chunk_elems;
    x.len = min(p.chunk_elems, L - x.e0);
    x.row = x.b * p.G + x.grp;
    return x;
}

// global element offset of logical row element e (gathered channel order)
__device__ __forceinline__ size_t row_elem_offset(const RowsParams& p, const Unit& x, int e) {
    if (p.perm == nullptr) return ((size_t)x.b * p.C + (size_t)x.grp * p.g) * p.HW + e;
    const int j = e / p.HW;
    const int pos = e - j * p.HW;
    const int ch = p.perm[x.grp * p.g + j];
    return ((size_t)x.b * p.C + ch) * p.HW + pos;
}

template <typename T, bool MSE>
__global__ void __launch_bounds__(kRowsThreads, 1) kl_rows_tma_kernel(const RowsParams p) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    constexpr int NV = kEPT / VE;  // vector rows per thread

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
    float2* red_max = reinterpret_cast<float2*>(full + 2 * kStages);
    float4* red_sum = reinterpret_cast<float4*>(red_max + kConsWarps);
    float* bcast = reinterpret_cast<float*>(red_sum + kConsWarps);

    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // ================================ TMA issue (thread 0 only) ================================
    // Slots are refilled in ring order.  `free_slots` counts slots every thread has drained: a
    // unit's slots are released by the CTA barrier that follows its ring->register copy.
    long long prod_u = blockIdx.x;
    int prod_v0 = 0, prod_stage = 0, free_slots = kStages;
    uint64_t pol = 0;
    if (threadIdx.x == 0) pol = l2_policy_evict_first();
    auto issue_loads = [&]() {
        while (free_slots > 0 && prod_u < p.total_units) {
            const Unit x = decode_unit(p, prod_u);
            const int nvec = x.len / VE;
            const int nv = min(kSlotVecs, nvec - prod_v0);
            const uint32_t bytes = (uint32_t)nv * 16u;
            mbar_arrive_expect_tx(&full[prod_stage], 2u * bytes);
            unsigned char* dst_s = smem + (size_t)prod_stage * kStageBytes;
            unsigned char* dst_t = dst_s + kSlotBytes;
            const int e = x.e0 + prod_v0 * VE;
            if (p.perm == nullptr) {
                const size_t off = row_elem_offset(p, x, e) * sizeof(T);
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
                    const size_t off = row_elem_offset(p, x, cur) * sizeof(T);
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
                prod_u += gridDim.x;
            }
            if (++prod_stage == kStages) prod_stage = 0;
            --free_slots;
        }
    };
    if (threadIdx.x == 0) issue_loads();

    // ================================ 16 warps, chunk lives in registers ================================
    const int tid = threadIdx.x;
    const int cwarp = tid >> 5;
    const float c2 = p.c2;
    float s[kEPT], t[kEPT];
    float cta_kl = 0.f, cta_sq = 0.f;  // accumulated by tid 0 in unit order (deterministic)
    int stage = 0;
    uint32_t phase = 0;

    for (long long u = blockIdx.x; u < p.total_units; u += gridDim.x) {
        const Unit x = decode_unit(p, u);
        const int nvec = x.len / VE;

        // ---- ring -> registers, running max of the raw values
        const int nslots = (nvec + kSlotVecs - 1) / kSlotVecs;
        float mxs = -INFINITY, mxt = -INFINITY;
#pragma unroll
        for (int j = 0; j < NV / kSlotVecRows; ++j) {
            if (j * kSlotVecs < nvec) {
                mbar_wait(&full[stage], phase);
                const vec_t* bs = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes);
                const vec_t* bt = reinterpret_cast<const vec_t*>(smem + (size_t)stage * kStageBytes + kSlotBytes);
#pragma unroll
                for (int r = 0; r < kSlotVecRows; ++r) {
                    const int v = j * kSlotVecRows + r;
                    if (v * kCons + tid < nvec) {
                        const vec_t a = bs[r * kCons + tid];
                        const vec_t b = bt[r * kCons + tid];
                        E::unpack(a, &s[v * VE]);
                        E::unpack(b, &t[v * VE]);
#pragma unroll
                        for (int k = 0; k < VE; ++k) {
                            mxs = fmaxf(mxs, s[v * VE + k]);
                            mxt = fmaxf(mxt, t[v * VE + k]);
                        }
                    }
                }
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }

        // ---- chunk max over the CTA
        mxs = warp_max(mxs);
        mxt = warp_max(mxt);
        if (lane == 0) red_max[cwarp] = make_float2(mxs, mxt);
        __syncthreads();
        if (tid == 0) {  // every thread holds its elements in registers: this unit's slots are free
            free_slots += nslots;
            issue_loads();
        }
        float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
        for (int w = 0; w < kConsWarps; ++w) {
            const float2 r = red_max[w];
            ms = fmaxf(ms, r.x);
            mt = fmaxf(mt, r.y);
        }
        const float ms2 = ms * c2, mt2 = mt * c2;

        // ---- exponentials (kept in registers), partial sums
        float zs = 0.f, zt = 0.f, a = 0.f, sq = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            if (v * kCons + tid < nvec) {
#pragma unroll
                for (int k = 0; k < VE; ++k) {
                    const int i = v * VE + k;
                    const float d = t[i] - s[i];
                    const float es = fast_exp2(fmaf(s[i], c2, -ms2));
                    const float et = fast_exp2(fmaf(t[i], c2, -mt2));
                    zs += es;
                    zt += et;
                    a = fmaf(et, d, a);
                    if (MSE) {
                        sq = fmaf(d, d, sq);
                    } else {
                        s[i] = es;
                        t[i] = et;
                    }
                }
            }
        }
        zs = warp_sum(zs);
        zt = warp_sum(zt);
        a = warp_sum(a);
        if (MSE) sq = warp_sum(sq);
        if (lane == 0) red_sum[cwarp] = make_float4(zs, zt, a, sq);
        __syncthreads();
        float Zs = 0.f, Zt = 0.f, A = 0.f, SQ = 0.f;
#pragma unroll
        for (int w = 0; w < kConsWarps; ++w) {
            const float4 r = red_sum[w];
            Zs += r.x;
            Zt += r.y;
            A += r.z;
            SQ += r.w;
        }

        // ---- rows split over several CTAs: exchange (max, sum) partials through global memory
        float Ms = ms, Mt = mt, fs = 1.f, ft = 1.f;
        if (x.nch > 1) {
            unsigned* rcnt = &p.row_cnt[x.row & (kRowCntRing - 1)];
            if (tid == 0) {
                float* slot = p.unit_part + (size_t)u * kPartWords;
                __stcg(reinterpret_cast<float4*>(slot), make_float4(ms, Zs, mt, Zt));
                __stcg(slot + 4, A);
                __threadfence();
                atomicAdd(rcnt, 1u);
                unsigned spins = 0;
                while (ld_acquire_gpu(rcnt) < (unsigned)x.nch) {
                    if (++spins > kSpinLimit) {
                        atomicExch(&p.ctrl[1], 1u);  // never expected: reported by the host wrapper
                        break;
                    }
                    __nanosleep(32);
                }
            }
            __syncthreads();
            if (cwarp == 0) {
                const long long u0 = u - x.ck;
                Stat ss = {-INFINITY, 0.f}, st = {-INFINITY, 0.f};
                float aa = 0.f;
                for (int k = lane; k < x.nch; k += 32) {
                    const float* q = p.unit_part + (size_t)(u0 + k) * kPartWords;
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(q));
                    const float ak = __ldcg(q + 4);
                    ss = stat_merge(ss, Stat{v.x, v.y}, c2);
                    const float nm = fmaxf(st.m, v.z);
                    const float fo = (st.z > 0.f) ? fast_exp2((st.m - nm) * c2) : 0.f;
                    const float fn = fast_exp2((v.z - nm) * c2);
                    st.z = st.z * fo + v.w * fn;
                    aa = aa * fo + ak * fn;
                    st.m = nm;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    Stat os, ot;
                    os.m = __shfl_xor_sync(0xffffffffu, ss.m, o);
                    os.z = __shfl_xor_sync(0xffffffffu, ss.z, o);
                    ot.m = __shfl_xor_sync(0xffffffffu, st.m, o);
                    ot.z = __shfl_xor_sync(0xffffffffu, st.z, o);
                    const float oa = __shfl_xor_sync(0xffffffffu, aa, o);
                    ss = stat_merge(ss, os, c2);
                    const float nm = fmaxf(st.m, ot.m);
                    const float f1 = (st.z > 0.f) ? fast_exp2((st.m - nm) * c2) : 0.f;
                    const float f2 = (ot.z > 0.f) ? fast_exp2((ot.m - nm) * c2) : 0.f;
                    // symmetric form: both partners compute bit-identical results
                    const float z1 = st.z * f1, z2 = ot.z * f2;
                    const float a1 = aa * f1, a2 = oa * f2;
                    st.z = (lane & o) ? (z2 + z1) : (z1 + z2);
                    aa = (lane & o) ? (a2 + a1) : (a1 + a2);
                    st.m = nm;
                }
                if (lane == 0) {
                    bcast[0] = ss.m;
                    bcast[1] = ss.z;
                    bcast[2] = st.m;
                    bcast[3] = st.z;
                    bcast[4] = aa;
                }
            }
            __syncthreads();
            Ms = bcast[0];
            Zs = bcast[1];
            Mt = bcast[2];
            Zt = bcast[3];
            A = bcast[4];
            fs = fast_exp2((ms - Ms) * c2);
            ft = fast_exp2((mt - Mt) * c2);
            if (tid == 0) {
                const unsigned old = atomicAdd(rcnt, 1u);
                if (old == 2u * (unsigned)x.nch - 1u) atomicExch(rcnt, 0u);  // last one out resets
            }
        }

        if (tid == 0) {
            if (x.ck == 0) {
                // KL(p||q) = sum p (t - s)/tau - lse_t + lse_s
                const float kl = p.inv_tau * A / Zt - ((Mt - Ms) * p.inv_tau + (logf(Zt) - logf(Zs)));
                if (p.row_kl) p.row_kl[x.row] = kl;
                cta_kl += kl;
            }
            if (MSE) cta_sq += SQ;
        }

        // ---- gradient straight from registers
        const float ks = p.coef * fs / Zs;
        const float kt = p.coef * ft / Zt;
        T* out = static_cast<T*>(p.dS);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int vi = v * kCons + tid;
            if (vi < nvec) {
                float o[VE];
#pragma unroll
                for (int k = 0; k < VE; ++k) {
                    const int i = v * VE + k;
                    if (MSE) {
                        const float es = fast_exp2(fmaf(s[i], c2, -ms2));
                        const float et = fast_exp2(fmaf(t[i], c2, -mt2));
                        o[k] = fmaf(es, ks, -et * kt) + p.mse_gcoef * (s[i] - t[i]);
                    } else {
                        o[k] = fmaf(s[i], ks, -t[i] * kt);
                    }
                }
                const size_t off = row_elem_offset(p, x, x.e0 + vi * VE);
                *reinterpret_cast<vec_t*>(out + off) = E::pack(o);
            }
        }
    }

    // ================================ loss: per-CTA partials, last CTA sums them in a fixed order ================================
    if (cwarp == 0) {
        unsigned ticket = 0;
        if (lane == 0) {
            __stcg(&p.cta_part[blockIdx.x], cta_kl);
            __stcg(&p.cta_part[kMaxGrid + blockIdx.x], cta_sq);
            __threadfence();
            ticket = atomicAdd(&p.ctrl[0], 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            double kl = 0.0, sq = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
                kl += (double)__ldcg(&p.cta_part[i]);
                sq += (double)__ldcg(&p.cta_part[kMaxGrid + i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kl += __shfl_down_sync(0xffffffffu, kl, o);
                sq += __shfl_down_sync(0xffffffffu, sq, o);
            }
            if (lane == 0) {
                *p.loss = (float)((double)p.loss_scale * kl);
                if (MSE && p.mse_loss) *p.mse_loss = (float)((double)p.mse_scale * sq);
                atomicExch(&p.ctrl[0], 0u);
            }
        }
    }
}

// ====================================================================================================
// generic path: unit = (sample, logical channel, plane chunk); no alignment requirement
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
    const T* s = static_cast<const T*>(p.S) + x.off;
    const T* t = static_cast<const T*>(p.T) + x.off;
    float ms = -INFINITY, mt = -INFINITY;
    for (int i = threadIdx.x; i < x.n; i += kGenThreads) {
        ms = fmaxf(ms, E::load(s + i));
        mt = fmaxf(mt, E::load(t + i));
    }
    block_max2(ms, mt, scratch);
    const float ms2 = ms * p.c2, mt2 = mt * p.c2;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // zs, zt, a, sq
    for (int i = threadIdx.x; i < x.n; i += kGenThreads) {
        const float a = E::load(s + i), b = E::load(t + i);
        const float d = b - a;
        const float et = fast_exp2(fmaf(b, p.c2, -mt2));
        acc[0] += fast_exp2(fmaf(a, p.c2, -ms2));
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
    const int grp = x.cl / p.g;
    const int c0 = grp * p.g;
    const int g_real = min(p.g, p.C - c0);
    const long long u0 = ((long long)x.b * p.C + c0) * p.KC;
    const int nparts = g_real * p.KC;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        Stat ss = {-INFINITY, 0.f}, st = {-INFINITY, 0.f};
        float aa = 0.f;
        for (int k = lane; k < nparts; k += 32) {
            const float* q = p.unit_part + (size_t)(u0 + k) * kPartWords;
            ss = stat_merge(ss, Stat{q[0], q[1]}, p.c2);
            const float nm = fmaxf(st.m, q[2]);
            const float fo = (st.z > 0.f) ? fast_exp2((st.m - nm) * p.c2) : 0.f;
            const float fn = fast_exp2((q[2] - nm) * p.c2);
            st.z = st.z * fo + q[3] * fn;
            aa = aa * fo + q[4] * fn;
            st.m = nm;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Stat os, ot;
            os.m = __shfl_xor_sync(0xffffffffu, ss.m, o);
            os.z = __shfl_xor_sync(0xffffffffu, ss.z, o);
            ot.m = __shfl_xor_sync(0xffffffffu, st.m, o);
            ot.z = __shfl_xor_sync(0xffffffffu, st.z, o);
            const float oa = __shfl_xor_sync(0xffffffffu, aa, o);
            ss = stat_merge(ss, os, p.c2);
            const float nm = fmaxf(st.m, ot.m);
            const float f1 = (st.z > 0.f) ? fast_exp2((st.m - nm) * p.c2) : 0.f;
            const float f2 = (ot.z > 0.f) ? fast_exp2((ot.m - nm) * p.c2) : 0.f;
            const float z1 = st.z * f1, z2 = ot.z * f2;
            const float a1 = aa * f1, a2 = oa * f2;
            st.z = (lane & o) ? (z2 + z1) : (z1 + z2);
            aa = (lane & o) ? (a2 + a1) : (a1 + a2);
            st.m = nm;
        }
        if (lane == 0) {
            row_stat[0] = ss.m;
            row_stat[1] = ss.z;
            row_stat[2] = st.m;
            row_stat[3] = st.z;
            row_stat[4] = aa;
        }
    }
    __syncthreads();
    const float Ms = row_stat[0], Zs = row_stat[1], Mt = row_stat[2], Zt = row_stat[3], A = row_stat[4];
    if (threadIdx.x == 0 && u == u0) {
        const float kl = p.inv_tau * A / Zt - ((Mt - Ms) * p.inv_tau + (logf(Zt) - logf(Zs)));
        p.row_kl[x.b * p.G + grp] = kl;
    }
    const float ms2 = Ms * p.c2, mt2 = Mt * p.c2;
    const float ks = p.coef / Zs, kt = p.coef / Zt;
    const T* s = static_cast<const T*>(p.S) + x.off;
    const T* t = static_cast<const T*>(p.T) + x.off;
    T* o = static_cast<T*>(p.dS) + x.off;
    for (int i = threadIdx.x; i < x.n; i += kGenThreads) {
        const float a = E::load(s + i), b = E::load(t + i);
        const float es = fast_exp2(fmaf(a, p.c2, -ms2));
        const float et = fast_exp2(fmaf(b, p.c2, -mt2));
        E::store(o + i, fmaf(es, ks, -et * kt) + p.mse_gcoef * (a - b));
    }
}

// one CTA: loss = loss_scale * sum(row_kl) (fixed order); mse_loss = mse_scale * sum(unit sq partials)
__global__ void __launch_bounds__(1024) kl_rows_generic_finalize(const RowsParams p, long long n_units) {
    __shared__ double sh[2][32];
    double kl = 0.0, sq = 0.0;
    for (int i = threadIdx.x; i < p.R; i += 1024) kl += (double)p.row_kl[i];
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
        *p.loss = (float)((double)p.loss_scale * a);
        if (p.mse_loss) *p.mse_loss = (float)((double)p.mse_scale * b);
    }
}

// ====================================================================================================
// host launchers
// ====================================================================================================
template <typename T, bool MSE>
static cudaError_t launch_tma_t(const RowsParams& p, int grid, bool cooperative, cudaStream_t stream) {
    auto kern = kl_rows_tma_kernel<T, MSE>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kRowsThreads);
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
    if (bf16) {
        return mse ? launch_tma_t<__nv_bfloat16, true>(p, grid, cooperative, stream)
                   : launch_tma_t<__nv_bfloat16, false>(p, grid, cooperative, stream);
    }
    return mse ? launch_tma_t<float, true>(p, grid, cooperative, stream)
               : launch_tma_t<float, false>(p, grid, cooperative, stream);
}

int kl_rows_tma_chunk_capacity() { return kChunkCap; }

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
