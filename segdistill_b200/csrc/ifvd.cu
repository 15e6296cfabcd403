// IFVDLoss similarity term, forward + backward.
//
// Replaces mmseg/models/distillation/losses.py:218-235: the Python loop over all C classes that builds the
// per-class centre maps (`for i in range(C)`, ~10 full-size ATen ops per class for S and for T), the two
// nn.CosineSimilarity(dim=1) calls, nn.MSELoss and the autograd backward of all of it.
//
//   sim_X(p) = cos(X[:, p], centre_X[:, k(p)]),  centre_X[:, k] = sum_{p' in k} X[:, p'] / (n_k + 1e-6)
//   loss     = weight * mean_p (sim_S(p) - sim_T(p))^2               (weight = 10 in the reference)
//   pixels without a class (label outside [0, C), e.g. ignore index 255) keep their own feature as centre
//   (sim = 1 for S and for T): no loss, no gradient.
//
// Four launches + a finalize, all on tensors that stay in the 126 MB L2 at the sizes the reference trains on:
//   1. ifvd_class_sums_kernel<false>   class sums of S and T and the class counts - a segmented reduction, one warp per
//                                      (sample, channel) plane, warp-private shared-memory bins; lanes holding the same
//                                      class are found with match.any and summed in lane order by the lowest of
//                                      them, so there are no float atomics and the result is deterministic.
//                                      (Cosine similarity is scale-invariant: the sums are used, the division by
//                                      n_k + 1e-6 appears only in the eps clamp and in the gradient.)
//   2. ifvd_sim_kernel                 one thread per pixel: both similarities in one sweep over the channels,
//                                      (sim_S - sim_T)^2 partials, per-pixel backward coefficients.
//   3. ifvd_class_sums_kernel<true>    the gradient reaching each centre: U_k = sum_{p in k} g_p * S[:, p]/|S[:, p]|,
//                                      V_k = sum_{p in k} g_p * sim_S(p)   (same segmented reduction, weighted).
//   4. ifvd_grad_kernel                dS[:, p] = g_p*(c^/|f| - sim*f/|f|^2) + (U_k/|c| - V_k*c/|c|^2)/(n_k + 1e-6)
//   5. ifvd_finalize_kernel            loss partials summed in a fixed order (fp64).
//
// eps handling follows ATen's cosine_similarity (x / max(|x|, 1e-8) for both arguments), including the zero
// gradient through a clamped norm.
#include "common.cuh"
#include "launch.h"
#include "params.h"

namespace sd {

constexpr int kIfvdWarps = 8;          // planes per CTA of the class-sum kernels
constexpr int kIfvdPixThreads = 128;   // pixels per CTA of the per-pixel kernels
constexpr float kCosEps = 1e-8f;       // ATen cosine_similarity eps (the reference uses the default)
constexpr float kCentreEps = 1e-6f;    // losses.py:229-230

// ------------------------------------------------------------------------------------------------
// class sums: out[(tensor, b), c, k] = sum over the pixels p of class k of value(c, p)
//   plain:    value = X[b, c, p]              (virtual channel c == C: 1 -> class counts), X = S (z = 0) or T (z = 1)
//   weighted: value = a0[p] * S[b, c, p]      (virtual channel c == C: a1[p])
// ------------------------------------------------------------------------------------------------
template <typename T, bool WEIGHTED>
__global__ void __launch_bounds__(kIfvdWarps * 32) ifvd_class_sums_kernel(const IfvdParams p) {
    extern __shared__ float bins[];  // [kIfvdWarps][K1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int c = blockIdx.x * kIfvdWarps + warp;
    const int C = p.C, HW = p.HW, K1 = C + 1;
    if (c > C) return;  // warp-uniform; no CTA barrier below

    float* my = bins + warp * K1;
    for (int k = lane; k < K1; k += 32) my[k] = 0.f;
    __syncwarp();

    const bool second = !WEIGHTED && blockIdx.z == 1;
    const T* feat = static_cast<const T*>(second ? p.T : p.S) + ((size_t)b * C + (c < C ? c : 0)) * HW;
    const int* cls = p.cls + (size_t)b * HW;
    const float* a0 = p.pix + (size_t)b * HW;
    const float* a1 = a0 + (size_t)p.B * HW;
    const bool real = c < C;

    constexpr int U = 4;
    for (int i0 = 0; i0 < HW; i0 += 32 * U) {
        int k[U];
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = i0 + u * 32 + lane;
            k[u] = C;      // lanes past the end of the plane: the "no class" bin, value 0
            v[u] = 0.f;
            if (px < HW) {
                k[u] = __ldg(cls + px);
                if (WEIGHTED) v[u] = real ? __ldg(a0 + px) * Elem<T>::load(feat + px) : __ldg(a1 + px);
                else v[u] = real ? Elem<T>::load(feat + px) : 1.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned m = __match_any_sync(0xffffffffu, k[u]);
            float acc;
            if (m == 0xffffffffu) {  // the whole warp sits in one class (the usual case on label maps)
                acc = v[u];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            } else {                 // lanes of a class summed in lane order
                acc = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float o = __shfl_sync(0xffffffffu, v[u], j);
                    if ((m >> j) & 1u) acc += o;
                }
            }
            if (lane == __ffs(m) - 1) my[k[u]] += acc;
            __syncwarp();
        }
    }

    float* out = (WEIGHTED ? p.wsum : p.sums + (second ? (size_t)p.B * K1 * K1 : 0)) + ((size_t)b * K1 + c) * K1;
    for (int k = lane; k < K1; k += 32) out[k] = my[k];
}

// ------------------------------------------------------------------------------------------------
// per pixel: sim_S, sim_T, loss partial, backward coefficients
//   pix[0] = g/|f|   pix[1] = g*sim_S   pix[2] = g*sim_S/|f|^2 (0 when |f| is clamped)   pix[3] = max(|centre|, eps)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kIfvdPixThreads) ifvd_sim_kernel(const IfvdParams p) {
    __shared__ float red[kIfvdPixThreads / 32];
    const int b = blockIdx.y;
    const int px = blockIdx.x * kIfvdPixThreads + threadIdx.x;
    const int C = p.C, HW = p.HW, K1 = C + 1;
    float d2 = 0.f;
    if (px < HW) {
        const int k = __ldg(p.cls + (size_t)b * HW + px);
        const T* s = static_cast<const T*>(p.S) + (size_t)b * C * HW + px;
        const T* t = static_cast<const T*>(p.T) + (size_t)b * C * HW + px;
        float dots = 0.f, fs2 = 0.f, cs2 = 0.f, dott = 0.f, ft2 = 0.f, ct2 = 0.f, inv_n = 1.f;
        if (k < C) {
            const float* zs = p.sums + (size_t)b * K1 * K1 + k;
            const float* zt = zs + (size_t)p.B * K1 * K1;
#pragma unroll 5
            for (int c = 0; c < C; ++c) {
                const float fs = Elem<T>::load(s + (size_t)c * HW), ft = Elem<T>::load(t + (size_t)c * HW);
                const float cs = __ldg(zs + (size_t)c * K1), ct = __ldg(zt + (size_t)c * K1);
                dots = fmaf(fs, cs, dots); fs2 = fmaf(fs, fs, fs2); cs2 = fmaf(cs, cs, cs2);
                dott = fmaf(ft, ct, dott); ft2 = fmaf(ft, ft, ft2); ct2 = fmaf(ct, ct, ct2);
            }
            inv_n = 1.f / (__ldg(zs + (size_t)C * K1) + kCentreEps);
        } else {  // no class: the pixel is its own centre
#pragma unroll 5
            for (int c = 0; c < C; ++c) {
                const float fs = Elem<T>::load(s + (size_t)c * HW), ft = Elem<T>::load(t + (size_t)c * HW);
                fs2 = fmaf(fs, fs, fs2);
                ft2 = fmaf(ft, ft, ft2);
            }
            dots = cs2 = fs2;
            dott = ct2 = ft2;
        }
        const float nfs_raw = sqrtf(fs2), nft = fmaxf(sqrtf(ft2), kCosEps);
        const float nfs = fmaxf(nfs_raw, kCosEps);
        const float ncs = fmaxf(sqrtf(cs2) * inv_n, kCosEps), nct = fmaxf(sqrtf(ct2) * inv_n, kCosEps);
        const float sim_s = dots * inv_n / (nfs * ncs);
        const float sim_t = dott * inv_n / (nft * nct);
        const float d = sim_s - sim_t;
        const float g = k < C ? p.gcoef * d : 0.f;
        d2 = d * d;
        const size_t o = (size_t)b * HW + px, n = (size_t)p.B * HW;
        const float inv_nf = 1.f / nfs;
        p.pix[o] = g * inv_nf;
        p.pix[n + o] = g * sim_s;
        p.pix[2 * n + o] = nfs_raw > kCosEps ? g * sim_s * inv_nf * inv_nf : 0.f;
        p.pix[3 * n + o] = ncs;
    }
    d2 = warp_sum(d2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kIfvdPixThreads / 32; ++w) a += red[w];
        p.part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = a;
    }
}

// ------------------------------------------------------------------------------------------------
// gradient
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kIfvdPixThreads) ifvd_grad_kernel(const IfvdParams p) {
    const int b = blockIdx.y;
    const int px = blockIdx.x * kIfvdPixThreads + threadIdx.x;
    const int C = p.C, HW = p.HW, K1 = C + 1;
    if (px >= HW) return;
    const int k = __ldg(p.cls + (size_t)b * HW + px);
    const T* s = static_cast<const T*>(p.S) + (size_t)b * C * HW + px;
    T* out = static_cast<T*>(p.dS) + (size_t)b * C * HW + px;
    if (k >= C) {
        for (int c = 0; c < C; ++c) Elem<T>::store(out + (size_t)c * HW, 0.f);
        return;
    }
    const size_t o = (size_t)b * HW + px, n = (size_t)p.B * HW;
    const float a0 = p.pix[o], a2 = p.pix[2 * n + o], nc = p.pix[3 * n + o];
    const float* zs = p.sums + (size_t)b * K1 * K1 + k;
    const float* us = p.wsum + (size_t)b * K1 * K1 + k;
    const float inv_n = 1.f / (__ldg(zs + (size_t)C * K1) + kCentreEps);
    const float cp = inv_n / nc;                                    // d centre^ / d sum
    const float dp = nc > kCosEps ? __ldg(us + (size_t)C * K1) * cp * cp : 0.f;
    const float ks = a0 * cp - dp;
#pragma unroll 5
    for (int c = 0; c < C; ++c) {
        const float f = Elem<T>::load(s + (size_t)c * HW);
        const float z = __ldg(zs + (size_t)c * K1), u = __ldg(us + (size_t)c * K1);
        Elem<T>::store(out + (size_t)c * HW, fmaf(ks, z, fmaf(cp, u, -a2 * f)));
    }
}

__global__ void __launch_bounds__(256) ifvd_finalize_kernel(const float* part, int n, float scale, float* loss) {
    __shared__ double sh[8];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) a += (double)part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        *loss = (float)((double)scale * t);
    }
}

template <typename T>
static cudaError_t launch_ifvd_t(const IfvdParams& p, float loss_scale, cudaStream_t stream) {
    const int K1 = p.C + 1;
    const size_t bins = (size_t)kIfvdWarps * K1 * sizeof(float);
    const dim3 gsum((K1 + kIfvdWarps - 1) / kIfvdWarps, p.B, 2), gwsum(gsum.x, p.B, 1);
    const dim3 gpix((p.HW + kIfvdPixThreads - 1) / kIfvdPixThreads, p.B, 1);
    ifvd_class_sums_kernel<T, false><<<gsum, kIfvdWarps * 32, bins, stream>>>(p);
    ifvd_sim_kernel<T><<<gpix, kIfvdPixThreads, 0, stream>>>(p);
    ifvd_class_sums_kernel<T, true><<<gwsum, kIfvdWarps * 32, bins, stream>>>(p);
    ifvd_grad_kernel<T><<<gpix, kIfvdPixThreads, 0, stream>>>(p);
    ifvd_finalize_kernel<<<1, 256, 0, stream>>>(p.part, (int)(gpix.x * gpix.y), loss_scale, p.loss);
    return cudaGetLastError();
}

int ifvd_pix_threads() { return kIfvdPixThreads; }

cudaError_t launch_ifvd_sim(const IfvdParams& p, bool bf16, float loss_scale, cudaStream_t stream) {
    return bf16 ? launch_ifvd_t<__nv_bfloat16>(p, loss_scale, stream) : launch_ifvd_t<float>(p, loss_scale, stream);
}

}  // namespace sd
