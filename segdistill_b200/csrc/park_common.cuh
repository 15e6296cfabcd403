// Device helpers shared by the kernels that park a row in TENSOR MEMORY between the statistics pass and the
// gradient pass (kl_rows_cluster.cu: rows resident in a thread-block cluster; kl_rows_grid.cu: rows resident in the
// whole cooperative grid): tcgen05 alloc / st / ld wrappers, DSMEM pushes, the statistics of a part of a row for
// one or two fused losses and their merges, 4-element vectors of either dtype.
#pragma once
#include "rows_common.cuh"

namespace sd {

constexpr int kCSlotCols = 32;                         // TMEM columns of one parked chunk per thread: 4 vector-rows x (4 of S + 4 of T)

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
// one parked chunk of a thread: 32 consecutive TMEM columns of its lane.
// warp-collective: thread i of the warp owns TMEM lane (lane quarter of the warp) + i
struct Parked {
    uint32_t w[kCSlotCols];
};
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const Parked& x) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(x.w[0]), "r"(x.w[1]), "r"(x.w[2]), "r"(x.w[3]), "r"(x.w[4]), "r"(x.w[5]), "r"(x.w[6]), "r"(x.w[7]), "r"(x.w[8]), "r"(x.w[9]), "r"(x.w[10]), "r"(x.w[11]), "r"(x.w[12]), "r"(x.w[13]), "r"(x.w[14]), "r"(x.w[15]), "r"(x.w[16]), "r"(x.w[17]), "r"(x.w[18]), "r"(x.w[19]), "r"(x.w[20]), "r"(x.w[21]), "r"(x.w[22]), "r"(x.w[23]), "r"(x.w[24]), "r"(x.w[25]), "r"(x.w[26]), "r"(x.w[27]), "r"(x.w[28]), "r"(x.w[29]), "r"(x.w[30]), "r"(x.w[31])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, Parked& x) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(x.w[0]), "=r"(x.w[1]), "=r"(x.w[2]), "=r"(x.w[3]), "=r"(x.w[4]), "=r"(x.w[5]), "=r"(x.w[6]), "=r"(x.w[7]), "=r"(x.w[8]), "=r"(x.w[9]), "=r"(x.w[10]), "=r"(x.w[11]), "=r"(x.w[12]), "=r"(x.w[13]), "=r"(x.w[14]), "=r"(x.w[15]), "=r"(x.w[16]), "=r"(x.w[17]), "=r"(x.w[18]), "=r"(x.w[19]), "=r"(x.w[20]), "=r"(x.w[21]), "=r"(x.w[22]), "=r"(x.w[23]), "=r"(x.w[24]), "=r"(x.w[25]), "=r"(x.w[26]), "=r"(x.w[27]), "=r"(x.w[28]), "=r"(x.w[29]), "=r"(x.w[30]), "=r"(x.w[31])
                 : "r"(taddr)
                 : "memory");
}
// the loaded registers are operands of the wait: nothing may read (or copy) them before it
__device__ __forceinline__ void tmem_wait_ld(Parked& x) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(x.w[0]), "+r"(x.w[1]), "+r"(x.w[2]), "+r"(x.w[3]), "+r"(x.w[4]), "+r"(x.w[5]), "+r"(x.w[6]), "+r"(x.w[7]), "+r"(x.w[8]), "+r"(x.w[9]), "+r"(x.w[10]), "+r"(x.w[11]), "+r"(x.w[12]), "+r"(x.w[13]), "+r"(x.w[14]), "+r"(x.w[15]), "+r"(x.w[16]), "+r"(x.w[17]), "+r"(x.w[18]), "+r"(x.w[19]), "+r"(x.w[20]), "+r"(x.w[21]), "+r"(x.w[22]), "+r"(x.w[23]), "+r"(x.w[24]), "+r"(x.w[25]), "+r"(x.w[26]), "+r"(x.w[27]), "+r"(x.w[28]), "+r"(x.w[29]), "+r"(x.w[30]), "+r"(x.w[31])
                 :
                 : "memory");
}

// 16-column variants (kl_rows_grid.cu with 16 park warps: 8 element pairs per thread and chunk)
struct Parked16 {
    uint32_t w[16];
};
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const Parked16& x) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(x.w[0]), "r"(x.w[1]), "r"(x.w[2]), "r"(x.w[3]), "r"(x.w[4]), "r"(x.w[5]), "r"(x.w[6]), "r"(x.w[7]), "r"(x.w[8]), "r"(x.w[9]), "r"(x.w[10]), "r"(x.w[11]), "r"(x.w[12]), "r"(x.w[13]), "r"(x.w[14]), "r"(x.w[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, Parked16& x) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(x.w[0]), "=r"(x.w[1]), "=r"(x.w[2]), "=r"(x.w[3]), "=r"(x.w[4]), "=r"(x.w[5]), "=r"(x.w[6]), "=r"(x.w[7]), "=r"(x.w[8]), "=r"(x.w[9]), "=r"(x.w[10]), "=r"(x.w[11]), "=r"(x.w[12]), "=r"(x.w[13]), "=r"(x.w[14]), "=r"(x.w[15])
                 : "r"(taddr)
                 : "memory");
}
// (the loaded registers of both halves are operands of the wait: nothing may read or copy them before it)
__device__ __forceinline__ void tmem_wait_ld(Parked16& a, Parked16& b) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a.w[0]), "+r"(a.w[1]), "+r"(a.w[2]), "+r"(a.w[3]), "+r"(a.w[4]), "+r"(a.w[5]), "+r"(a.w[6]), "+r"(a.w[7]), "+r"(a.w[8]), "+r"(a.w[9]), "+r"(a.w[10]), "+r"(a.w[11]), "+r"(a.w[12]), "+r"(a.w[13]), "+r"(a.w[14]), "+r"(a.w[15]),
                   "+r"(b.w[0]), "+r"(b.w[1]), "+r"(b.w[2]), "+r"(b.w[3]), "+r"(b.w[4]), "+r"(b.w[5]), "+r"(b.w[6]), "+r"(b.w[7]), "+r"(b.w[8]), "+r"(b.w[9]), "+r"(b.w[10]), "+r"(b.w[11]), "+r"(b.w[12]), "+r"(b.w[13]), "+r"(b.w[14]), "+r"(b.w[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- statistics of a part of a row
// NL losses share the raw maxima; sums are relative to them
// (a = sum et (at - as) and dd = sum (et - es), accumulated term by term: common.cuh, "KL without cancellation".
// With R == 2 - one exponent for both losses, e[0] = e[1]^2 - a[0] is kept in units of the loss-1 exponent, i.e. it is
// half the true sum, everywhere up to kl_of_row.)
template <int NL>
struct PStat {
    float ms, mt;
    float zs[NL], zt[NL], a[NL], dd[NL];
};
template <int NL>
__device__ __forceinline__ PStat<NL> pstat_empty() {
    PStat<NL> r;
    r.ms = kMaxFloor;
    r.mt = kMaxFloor;
#pragma unroll
    for (int k = 0; k < NL; ++k) r.zs[k] = r.zt[k] = r.a[k] = r.dd[k] = 0.f;
    return r;
}
// exp2((x - ref) * c2[k]) for every loss; R == 2: c2[0] == 2*c2[1], so e[0] = e[1]^2 (one ex2 for both)
template <int NL, int R>
__device__ __forceinline__ void exps(float x, float ref, const float (&c2)[NL], float (&e)[NL]) {
    // (fl(x c2) - fl(ref c2), not (x - ref) c2: the per-element exponents are taken against fl(ref c2))
    if (NL == 1) {
        e[0] = ref_factor(x, ref, c2[0]);
    } else if (R == 2) {
        e[NL - 1] = ref_factor(x, ref, c2[NL - 1]);
        e[0] = e[NL - 1] * e[NL - 1];
    } else {
#pragma unroll
        for (int k = 0; k < NL; ++k) e[k] = ref_factor(x, ref, c2[k]);
    }
}
// the same, returning the exponents too (arg[0] in units of the loss-1 exponent when R == 2)
template <int NL, int R>
__device__ __forceinline__ void exps_args(float x, const float (&ref2)[NL], const float (&c2)[NL], float (&arg)[NL],
                                          float (&e)[NL]) {
    if (NL == 1) {
        arg[0] = fmaf(x, c2[0], -ref2[0]);
        e[0] = fast_exp2(arg[0]);
    } else if (R == 2) {
        arg[NL - 1] = fmaf(x, c2[NL - 1], -ref2[NL - 1]);
        arg[0] = arg[NL - 1];
        e[NL - 1] = fast_exp2(arg[NL - 1]);
        e[0] = e[NL - 1] * e[NL - 1];
    } else {
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            arg[k] = fmaf(x, c2[k], -ref2[k]);
            e[k] = fast_exp2(arg[k]);
        }
    }
}
// shift of the exponent gap per loss when sums taken against (ms, mt) move to (Ms, Mt) (common.cuh: merge_shift), in
// the units a[k] is kept in
template <int NL, int R>
__device__ __forceinline__ void shifts(float ms, float mt, float Ms, float Mt, const float (&c2)[NL], float (&x)[NL]) {
    if (NL == 2 && R == 2) {
        x[NL - 1] = merge_shift(ms, mt, Ms, Mt, c2[NL - 1]);
        x[0] = x[NL - 1];
    } else {
#pragma unroll
        for (int k = 0; k < NL; ++k) x[k] = merge_shift(ms, mt, Ms, Mt, c2[k]);
    }
}
// ft - fs per loss for sums taken against (ms, mt) that move to the references (Ms, Mt): see factor_diff
template <int NL, int R>
__device__ __forceinline__ void factor_diffs(const float (&x)[NL], const float (&fs)[NL], const float (&ft)[NL],
                                             float (&df)[NL]) {
    if (NL == 2 && R == 2) {
        df[NL - 1] = factor_diff(fs[NL - 1], ft[NL - 1], x[NL - 1]);
        df[0] = df[NL - 1] * (ft[NL - 1] + fs[NL - 1]);      // fs[0] = fs[1]^2, ft[0] = ft[1]^2
    } else {
#pragma unroll
        for (int k = 0; k < NL; ++k) df[k] = factor_diff(fs[k], ft[k], x[k]);
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
template <int NL, int R, int WIDTH>
__device__ __forceinline__ PStat<NL> pstat_reduce(const PStat<NL>& x, const float (&c2)[NL]) {
    PStat<NL> r;
    r.ms = x.ms;
    r.mt = x.mt;
#pragma unroll
    for (int o = WIDTH >> 1; o > 0; o >>= 1) {
        r.ms = fmaxf(r.ms, __shfl_xor_sync(0xffffffffu, r.ms, o));
        r.mt = fmaxf(r.mt, __shfl_xor_sync(0xffffffffu, r.mt, o));
    }
    float fs[NL], ft[NL], df[NL], sh[NL];
    exps<NL, R>(x.ms, r.ms, c2, fs);
    exps<NL, R>(x.mt, r.mt, c2, ft);
    shifts<NL, R>(x.ms, x.mt, r.ms, r.mt, c2, sh);
    factor_diffs<NL, R>(sh, fs, ft, df);
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        r.zs[k] = x.zs[k] * fs[k];
        r.zt[k] = x.zt[k] * ft[k];
        r.a[k] = fmaf(r.zt[k], sh[k], x.a[k] * ft[k]);
        r.dd[k] = fmaf(x.zs[k], df[k], x.dd[k] * ft[k]);
    }
#pragma unroll
    for (int o = WIDTH >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            r.zs[k] += __shfl_xor_sync(0xffffffffu, r.zs[k], o);
            r.zt[k] += __shfl_xor_sync(0xffffffffu, r.zt[k], o);
            r.a[k] += __shfl_xor_sync(0xffffffffu, r.a[k], o);
            r.dd[k] += __shfl_xor_sync(0xffffffffu, r.dd[k], o);
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
    float dfx[NL], dfy[NL], shx[NL], shy[NL];
    shifts<NL, R>(x.ms, x.mt, r.ms, r.mt, c2, shx);
    shifts<NL, R>(y.ms, y.mt, r.ms, r.mt, c2, shy);
    factor_diffs<NL, R>(shx, fxs, fxt, dfx);
    factor_diffs<NL, R>(shy, fys, fyt, dfy);
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        const float zx = __fmul_rn(x.zt[k], fxt[k]), zy = __fmul_rn(y.zt[k], fyt[k]);
        r.zs[k] = __fadd_rn(__fmul_rn(x.zs[k], fxs[k]), __fmul_rn(y.zs[k], fys[k]));
        r.zt[k] = __fadd_rn(zx, zy);
        r.a[k] = __fadd_rn(fmaf(zx, shx[k], __fmul_rn(x.a[k], fxt[k])), fmaf(zy, shy[k], __fmul_rn(y.a[k], fyt[k])));
        r.dd[k] = __fadd_rn(fmaf(x.zs[k], dfx[k], __fmul_rn(x.dd[k], fxt[k])), fmaf(y.zs[k], dfy[k], __fmul_rn(y.dd[k], fyt[k])));
    }
    return r;
}
// record layout: {ms, mt, zs0, zt0} {a0, dd0, zs1, zt1} {a1, dd1, -, -}
template <int NL>
__device__ __forceinline__ PStat<NL> pstat_from(const float4& r0, const float4& r1, const float4& r2) {
    PStat<NL> x;
    x.ms = r0.x;
    x.mt = r0.y;
    x.zs[0] = r0.z;
    x.zt[0] = r0.w;
    x.a[0] = r1.x;
    x.dd[0] = r1.y;
    if (NL == 2) {
        x.zs[NL - 1] = r1.z;
        x.zt[NL - 1] = r1.w;
        x.a[NL - 1] = r2.x;
        x.dd[NL - 1] = r2.y;
    }
    return x;
}
// a_unit: 2 for loss 0 of a launch with R == 2 (its a is kept in units of the loss-1 exponent), else 1
__device__ __forceinline__ float kl_of_row(float a_unit, float zs, float zt, float a, float dd) {
    return kl_from_stats(zs, zt, a_unit * a, dd);
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

}  // namespace sd
