// Device-side building blocks shared by the sm_100a kernels: mbarrier / TMA (bulk async
// copy) wrappers, named barriers, warp reductions, bf16 packing, fast exp2.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sd {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a while when the phase is not complete)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// the same for waits that may last microseconds: with a suspend-time hint the failed try becomes a sleep that the
// phase completion ends (SASS: NANOSLEEP.SYNCS) instead of a poll every few dozen cycles - polling warps share the
// MIO queue with the MUFU / shared-memory instructions of the warps that do the work
template <uint32_t kHintNs>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(kHintNs)
            : "memory");
    } while (ok == 0);
}

// ---------------------------------------------------------------- TMA: 1-D bulk copy global -> shared
// bytes % 16 == 0, both addresses 16-byte aligned; completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar, uint64_t l2_policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_policy)
        : "memory");
}
// TMA: 3-D tiled tensor copy global -> shared (tensor map built on the host)
__device__ __forceinline__ void tma_tile3d_g2s(void* dst_smem, const void* tmap, int c0, int c1, int c2,
                                               uint64_t* bar, uint64_t l2_policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(l2_policy)
        : "memory");
}
// ---------------------------------------------------------------- TMA: 1-D bulk copy shared -> global
// The writes of the calling thread('s warp, after __syncwarp) to `src_smem` must be made visible to the async proxy
// first (fence_proxy_async_smem).  Completion is tracked per thread in bulk groups.
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// TMA: 3-D tiled tensor copy shared -> global (elements outside the tensor are not written); bulk-group completion
__device__ __forceinline__ void tma_tile3d_s2g(const void* tmap, int c0, int c1, int c2, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src_smem)) : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of the thread's groups have READ their shared-memory source (it may be overwritten)
__device__ __forceinline__ void tma_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed
__device__ __forceinline__ void tma_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// streaming data is read exactly once: ask L2 to evict it first
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// data that will be read again soon (phase 2 of the streaming kernel): keep it in L2
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// ---------------------------------------------------------------- named barrier (subset of the CTA)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- math
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


// ---------------------------------------------------------------- two fp32 per instruction (sm_100: FFMA2 / FADD2 / FMUL2)
// The packed forms deliver the same 128 results per clock and SM as the scalar ones but take ONE issue slot for two
// results (scripts/probe/ffma2_bench.cu) - what the issue-bound kernels are short of.  IEEE round-to-nearest per
// element, like the scalar instructions.  A value pair lives in an aligned 64-bit register pair.
struct F2 {
    unsigned long long v;
};
__device__ __forceinline__ F2 f2_make(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 f2_dup(float x) { return f2_make(x, x); }
__device__ __forceinline__ void f2_split(const F2& x, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x.v));
}
__device__ __forceinline__ float f2_sum(const F2& x) {
    float lo, hi;
    f2_split(x, lo, hi);
    return lo + hi;
}
__device__ __forceinline__ F2 f2_fma(const F2& a, const F2& b, const F2& c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 f2_add(const F2& a, const F2& b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 f2_mul(const F2& a, const F2& b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// maximum over the warp, identical bits in every lane (also orders the lanes: everybody contributed)
__device__ __forceinline__ float warp_max_uniform(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// sums of 8 values over the 32 lanes in 9 shuffles (halve the values a lane carries at every step, then two
// plain steps); lane L returns the total of v[L >> 2].  Fixed order: deterministic.
__device__ __forceinline__ float warp_sum8_transposed(const float (&v)[8], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    float w[4], u[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = b4 ? v[j + 4] : v[j], send = b4 ? v[j] : v[j + 4];
        w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = b3 ? w[j + 2] : w[j], send = b3 ? w[j] : w[j + 2];
        u[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float keep = b2 ? u[1] : u[0], send = b2 ? u[0] : u[1];
    float t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}
// the same for 4 values in 6 shuffles; lane L returns the total of v[L >> 3]
__device__ __forceinline__ float warp_sum4_transposed(const float (&v)[4], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
    float w[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = b4 ? v[j + 2] : v[j], send = b4 ? v[j] : v[j + 2];
        w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const float keep = b3 ? w[1] : w[0], send = b3 ? w[0] : w[1];
    float t = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// ---------------------------------------------------------------- global memory, L2-coherent access
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_streaming(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_streaming(uint4* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------- element types
template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static constexpr int kVec = 4;  // elements per 16-byte vector
    using vec_t = float4;
    static __device__ __forceinline__ void unpack(const vec_t& v, float* f) {
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    static __device__ __forceinline__ vec_t pack(const float* f) { return make_float4(f[0], f[1], f[2], f[3]); }
    static __device__ __forceinline__ float load(const float* p) { return *p; }
    static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <>
struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    using vec_t = uint4;
    static __device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
        lo = __uint_as_float(w << 16);
        hi = __uint_as_float(w & 0xffff0000u);
    }
    static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ void unpack(const vec_t& v, float* f) {
        unpack2(v.x, f[0], f[1]); unpack2(v.y, f[2], f[3]);
        unpack2(v.z, f[4], f[5]); unpack2(v.w, f[6], f[7]);
    }
    static __device__ __forceinline__ vec_t pack(const float* f) {
        return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
    static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------- KL without cancellation when S ~ T
// With as_i = s_i c2 - sigma, at_i = t_i c2 - theta the exponents the kernels feed to ex2 (sigma = fl(ms c2),
// theta = fl(mt c2) the references, c2 = log2(e)/tau), es_i = 2^as_i, et_i = 2^at_i, zs = sum es, zt = sum et:
//     KL(p||q) = sum p_i ln(p_i / q_i) = ln2 * (sum et_i (at_i - as_i)) / zt  -  ln(zt / zs)
// exactly - the references cancel algebraically, so a constant offset between teacher and student (which a softmax
// does not see) never enters.  Both terms are of the size of the differences t - s; the KL of a nearly converged
// student is their difference, of second order.  ln(zt) - ln(zs) from two rounded logarithms would lose everything
// below ulp(ln z) ~ 1e-6 - as large as the whole KL there (the reference's own log_softmax chain, losses.py:108-111,
// is 6e-4 .. 1e-2 off on such inputs, tests/golden/kld_*near*).  Every kernel therefore accumulates, next to zs, zt:
//     a2 = sum et_i (at_i - as_i)          dd = sum (et_i - es_i)        (term by term: sums of small terms)
// and evaluates ln(zt / zs) = log1p(dd / zs).  When partial sums taken against references (sigma_k, theta_k) are
// merged into (sigma, theta), with fs_k = 2^(sigma_k - sigma), ft_k = 2^(theta_k - theta) and the shift
// x_k = (theta_k - sigma_k) - (theta - sigma) = log2(ft_k / fs_k):
//     a2 = sum_k ft_k (a2_k + zt_k x_k)     dd = sum_k dd_k ft_k + zs_k (ft_k - fs_k),   ft_k - fs_k = fs_k (2^x_k - 1)
// x_k comes from compensated differences (exact), 2^x - 1 from a polynomial - never from subtracting rounded factors.

// 2^x - 1 for |x| <= 0.25, relative error ~1e-8 (Taylor in y = x ln 2, |y| <= 0.174)
__device__ __forceinline__ float exp2m1_small(float x) {
    const float y = x * kLn2;
    float p = 1.f / 720.f;
    p = fmaf(p, y, 1.f / 120.f);
    p = fmaf(p, y, 1.f / 24.f);
    p = fmaf(p, y, 1.f / 6.f);
    p = fmaf(p, y, 0.5f);
    p = fmaf(p, y, 1.f);
    return p * y;
}
// ft - fs for rescale factors fs = 2^xs, ft = 2^xt, given x = xt - xs.  Away from x ~ 0 nothing cancels.
__device__ __forceinline__ float factor_diff(float fs, float ft, float x) {
    return fabsf(x) <= 0.25f ? fs * exp2m1_small(x) : ft - fs;
}
// a - b and its exact rounding error (Knuth's TwoSum on a + (-b))
__device__ __forceinline__ float two_diff(float a, float b, float& err) {
    const float d = __fsub_rn(a, b);
    const float bb = __fsub_rn(d, a);
    err = __fsub_rn(__fsub_rn(a, __fsub_rn(d, bb)), __fadd_rn(b, bb));
    return d;
}
// the shift x_k = (theta_k - sigma_k) - (theta - sigma) of a part merged into new references, all four in the
// exponent (log2) domain; exact up to one final rounding whatever offset separates teacher and student
__device__ __forceinline__ float merge_shift2(float sk, float tk, float s, float t) {
    float e1, e2;
    const float d1 = two_diff(tk, sk, e1), d2 = two_diff(t, s, e2);
    return __fadd_rn(__fsub_rn(d1, d2), __fsub_rn(e1, e2));
}
// the same from references in the value domain; the products are rounded exactly as the kernels round the
// references of their exponentials (fmaf(x, c2, -fl(m c2)))
__device__ __forceinline__ float merge_shift(float msk, float mtk, float ms, float mt, float c2) {
    return merge_shift2(__fmul_rn(msk, c2), __fmul_rn(mtk, c2), __fmul_rn(ms, c2), __fmul_rn(mt, c2));
}
// 2^(fl(m_k c2) - fl(m c2)): rescale factor of sums taken against m_k to the reference m >= m_k
__device__ __forceinline__ float ref_factor(float mk, float m, float c2) {
    return fast_exp2(__fmul_rn(mk, c2) - __fmul_rn(m, c2));
}
// KL of one row from its merged statistics.  log1p(dd / zs) needs zt / zs = 1 + dd / zs away from 0; where the two
// sums differ by more than a factor 2 the plain logarithms are as good (nothing cancels there).
__device__ __forceinline__ float kl_from_stats(float zs, float zt, float a2, float dd) {
    const float r = dd / zs;
    const float lse = fabsf(r) < 0.5f ? log1pf(r) : logf(zt) - logf(zs);
    return kLn2 * a2 / zt - lse;
}

// softmax statistics of a piece of a row: max m and z = sum exp2((x - m)*c2)
struct Stat {
    float m, z;
};
// merge two pieces of the same row (c2 = log2(e)/tau).  Pieces with z == 0 are empty.
__device__ __forceinline__ Stat stat_merge(const Stat& a, const Stat& b, float c2) {
    Stat r;
    r.m = fmaxf(a.m, b.m);
    const float fa = (a.z > 0.f) ? fast_exp2((a.m - r.m) * c2) : 0.f;
    const float fb = (b.z > 0.f) ? fast_exp2((b.m - r.m) * c2) : 0.f;
    r.z = a.z * fa + b.z * fb;
    return r;
}

}  // namespace sd
