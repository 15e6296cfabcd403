// Feature MSE, forward + backward fused, and the backward grad_output scaling helper.
//
// sd_mse_fwd_bwd replaces nn.MSELoss / weight*mean((s-t)**2) (mmseg/models/distillation/
// losses.py:178,190,:202,235,:829) and its backward: one streaming pass, 12 B/elem fp32.
#include "common.cuh"
#include "params.h"

namespace sd {

constexpr int kMseThreads = 256;

template <typename T, bool VEC>
__global__ void __launch_bounds__(kMseThreads) mse_kernel(const T* __restrict__ S, const T* __restrict__ Tt,
                                                          T* __restrict__ dS, float* __restrict__ partials,
                                                          long long n, float gcoef) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    __shared__ float sh[kMseThreads / 32];
    float sq = 0.f;
    const long long stride = (long long)gridDim.x * kMseThreads;
    const long long gid = (long long)blockIdx.x * kMseThreads + threadIdx.x;
    long long done = 0;
    if (VEC) {
        const long long nvec = n / VE;
        const vec_t* vs = reinterpret_cast<const vec_t*>(S);
        const vec_t* vt = reinterpret_cast<const vec_t*>(Tt);
        vec_t* vo = reinterpret_cast<vec_t*>(dS);
        for (long long i = gid; i < nvec; i += stride) {
            float a[VE], b[VE], o[VE];
            E::unpack(__ldcs(vs + i), a);
            E::unpack(__ldcs(vt + i), b);
#pragma unroll
            for (int k = 0; k < VE; ++k) {
                const float d = a[k] - b[k];
                sq = fmaf(d, d, sq);
                o[k] = gcoef * d;
            }
            vo[i] = E::pack(o);
        }
        done = nvec * VE;
    }
    for (long long i = done + gid; i < n; i += stride) {
        const float d = E::load(S + i) - E::load(Tt + i);
        sq = fmaf(d, d, sq);
        E::store(dS + i, gcoef * d);
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kMseThreads / 32; ++w) a += sh[w];
        partials[blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(1024) mse_finalize(const float* __restrict__ partials, int nparts, float scale,
                                                     float* __restrict__ loss) {
    __shared__ double sh[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 1024) a += (double)partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sh[w];
        *loss = (float)((double)scale * t);
    }
}

// x[0 .. n) *= gv by the whole grid.  The grids are small (two CTAs per SM: the usual launch finds gv == 1 and only has
// to start and end - 1184 CTAs took 3 us to do that, 296 take half), so a thread keeps four 16-byte vectors in flight.
template <typename T>
__device__ __forceinline__ void scale_span(T* __restrict__ x, long long n, float gv, bool vec) {
    using E = Elem<T>;
    using vec_t = typename E::vec_t;
    constexpr int VE = E::kVec;
    const long long stride = (long long)gridDim.x * 256;
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    long long done = 0;
    if (vec) {
        const long long nvec = n / VE;
        vec_t* vx = reinterpret_cast<vec_t*>(x);
        long long i = gid;
        for (; i + 3 * stride < nvec; i += 4 * stride) {
            vec_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = vx[i + u * stride];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float a[VE];
                E::unpack(v[u], a);
#pragma unroll
                for (int k = 0; k < VE; ++k) a[k] *= gv;
                vx[i + u * stride] = E::pack(a);
            }
        }
        for (; i < nvec; i += stride) {
            float a[VE];
            E::unpack(vx[i], a);
#pragma unroll
            for (int k = 0; k < VE; ++k) a[k] *= gv;
            vx[i] = E::pack(a);
        }
        done = nvec * VE;
    }
    for (long long i = done + gid; i < n; i += stride) E::store(x + i, E::load(x + i) * gv);
}

// the step's loss scalars into slot (*cursor mod slots) of the device-resident log ring, cursor += 1 (one warp)
struct LogSink {
    const float* values;   // null: nothing to append
    int n;
    float* ring;
    unsigned* cursor;
    unsigned slots;
};
__device__ __forceinline__ void log_store(const LogSink& k, unsigned c, float v0) {
    float* dst = k.ring + (size_t)(c % k.slots) * k.n;
    const int lane = threadIdx.x;
    if (lane < k.n) dst[lane] = v0;
    for (int i = lane + 32; i < k.n; i += 32) dst[i] = k.values[i];
    __syncwarp();
    if (lane == 0) *k.cursor = c + 1u;
}
// (the cursor and a lane's value are loaded before either is used: one memory round trip, not two)
__device__ __forceinline__ void log_append(const LogSink& k) {
    const unsigned c = *k.cursor;
    const float v0 = (int)threadIdx.x < k.n ? k.values[threadIdx.x] : 0.f;
    log_store(k, c, v0);
}

// (sink: the backward's scaling launch can carry the step's log append - sd_scale_grad_log -, so that a step under
//  DeferredLogs holds no launch of its own for it)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) scale_grad_kernel(T* __restrict__ x, long long n, const float* __restrict__ g,
                                                         const LogSink sink) {
    // the appending warp issues its loads (cursor, value) together with the load of *g: the launch stays one memory
    // round trip long (three dependent ones measured +1.5 us on the step of every rank at N > 1)
    const bool appends = sink.values != nullptr && blockIdx.x == 0 && threadIdx.x < 32;
    unsigned c = 0;
    float v0 = 0.f;
    if (appends) {
        c = *sink.cursor;
        if ((int)threadIdx.x < sink.n) v0 = sink.values[threadIdx.x];
    }
    const float gv = *g;
    if (appends) log_store(sink, c, v0);
    if (gv == 1.0f) return;  // the usual case: loss enters the total as a plain sum
    scale_span<T>(x, n, gv, VEC);
}

// The same for the tensors of a grouped launch (sd_kl_rows_group_fwd_bwd): tensor k = blockIdx.y, dS[k] *= *g[k]; ONE launch
// for all of them (a dispatcher step over several layers otherwise pays one launch per layer in its backward).
struct ScaleGroup {
    void* x[kMaxSegs];
    const float* g[kMaxSegs];
    long long n[kMaxSegs];
};
template <typename T>
__global__ void __launch_bounds__(256) scale_grad_group_kernel(const ScaleGroup a) {
    const int k = blockIdx.y;
    const float gv = *a.g[k];
    if (gv == 1.0f) return;
    T* x = static_cast<T*>(a.x[k]);
    scale_span<T>(x, a.n[k], gv, (reinterpret_cast<uintptr_t>(x) & 15u) == 0);
}
cudaError_t launch_scale_grad_group(int n_tensors, void* const* dS, const long long* numel, bool bf16, const float* const* g,
                                    int grid, cudaStream_t stream) {
    ScaleGroup a = {};
    for (int k = 0; k < n_tensors; ++k) {
        a.x[k] = dS[k];
        a.g[k] = g[k];
        a.n[k] = numel[k];
    }
    const dim3 gr((unsigned)grid, (unsigned)n_tensors);
    if (bf16) scale_grad_group_kernel<__nv_bfloat16><<<gr, 256, 0, stream>>>(a);
    else scale_grad_group_kernel<float><<<gr, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

// Backward of a fused two-loss launch: dS was computed for d(total)/d(loss_k) == 1.  When the two
// upstream gradients are equal (the loss terms enter the total as a plain sum, possibly times one
// loss scale) dS is scaled in place and *flag = 0; otherwise *flag = 1 and the caller's conditional
// re-run of the fused kernel (run_if = flag) rebuilds dS with the individual factors.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) scale_grad2_kernel(T* __restrict__ x, long long n, const float* __restrict__ g0,
                                                          const float* __restrict__ g1, unsigned* __restrict__ flag) {
    const float a = *g0, b = *g1;
    const bool uniform = a == b;
    if (blockIdx.x == 0 && threadIdx.x == 0) *flag = uniform ? 0u : 1u;
    if (!uniform || a == 1.0f) return;
    scale_span<T>(x, n, a, VEC);
}

// The step's loss scalars into slot (*cursor mod slots) of a device-resident ring, cursor += 1: one warp, capturable
// in a CUDA graph (the slot is chosen on the device).  The ring is all-reduced once per `slots` steps instead of
// one collective per step (segdistill_b200/dist.py: DeferredLogs).
__global__ void log_push_kernel(const float* __restrict__ values, int n, float* __restrict__ ring,
                                unsigned* __restrict__ cursor, unsigned slots) {
    log_append(LogSink{values, n, ring, cursor, slots});
}

cudaError_t launch_log_push(const float* values, int n, float* ring, unsigned* cursor, int slots, cudaStream_t stream) {
    log_push_kernel<<<1, 32, 0, stream>>>(values, n, ring, cursor, (unsigned)slots);
    return cudaGetLastError();
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

cudaError_t launch_mse(const void* S, const void* T, void* dS, float* loss, float* partials, long long n, bool bf16,
                       float gcoef, float scale, int grid, cudaStream_t stream) {
    const bool vec = aligned16(S) && aligned16(T) && aligned16(dS);
    if (bf16) {
        auto s = static_cast<const __nv_bfloat16*>(S);
        auto t = static_cast<const __nv_bfloat16*>(T);
        auto o = static_cast<__nv_bfloat16*>(dS);
        if (vec) mse_kernel<__nv_bfloat16, true><<<grid, kMseThreads, 0, stream>>>(s, t, o, partials, n, gcoef);
        else mse_kernel<__nv_bfloat16, false><<<grid, kMseThreads, 0, stream>>>(s, t, o, partials, n, gcoef);
    } else {
        auto s = static_cast<const float*>(S);
        auto t = static_cast<const float*>(T);
        auto o = static_cast<float*>(dS);
        if (vec) mse_kernel<float, true><<<grid, kMseThreads, 0, stream>>>(s, t, o, partials, n, gcoef);
        else mse_kernel<float, false><<<grid, kMseThreads, 0, stream>>>(s, t, o, partials, n, gcoef);
    }
    mse_finalize<<<1, 1024, 0, stream>>>(partials, grid, scale, loss);
    return cudaGetLastError();
}

cudaError_t launch_scale_grad(void* dS, long long n, bool bf16, const float* g, int grid, cudaStream_t stream,
                              const float* log_values, int log_n, float* log_ring, unsigned* log_cursor, int log_slots) {
    const LogSink sink{log_values, log_n, log_ring, log_cursor, (unsigned)(log_slots > 0 ? log_slots : 1)};
    const bool vec = aligned16(dS);
    if (bf16) {
        auto x = static_cast<__nv_bfloat16*>(dS);
        if (vec) scale_grad_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(x, n, g, sink);
        else scale_grad_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(x, n, g, sink);
    } else {
        auto x = static_cast<float*>(dS);
        if (vec) scale_grad_kernel<float, true><<<grid, 256, 0, stream>>>(x, n, g, sink);
        else scale_grad_kernel<float, false><<<grid, 256, 0, stream>>>(x, n, g, sink);
    }
    return cudaGetLastError();
}

cudaError_t launch_scale_grad2(void* dS, long long n, bool bf16, const float* g0, const float* g1, unsigned* flag,
                               int grid, cudaStream_t stream) {
    const bool vec = aligned16(dS);
    if (bf16) {
        auto x = static_cast<__nv_bfloat16*>(dS);
        if (vec) scale_grad2_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(x, n, g0, g1, flag);
        else scale_grad2_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(x, n, g0, g1, flag);
    } else {
        auto x = static_cast<float*>(dS);
        if (vec) scale_grad2_kernel<float, true><<<grid, 256, 0, stream>>>(x, n, g0, g1, flag);
        else scale_grad2_kernel<float, false><<<grid, 256, 0, stream>>>(x, n, g0, g1, flag);
    }
    return cudaGetLastError();
}

}  // namespace sd
