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
// Launches (all on tensors that stay in the 126 MB L2 at the sizes the reference trains on):
//   1. ifvd_class_sums_kernel<plain>   class sums of S and T and the class counts - a segmented reduction with
//                                      lane = channel: a warp walks its pixels one by one, the class of a pixel is
//                                      warp-uniform, a run of pixels of one class is summed in a register and then
//                                      added to the lane's private shared-memory bin.  No atomics, no shuffle
//                                      reductions, fixed summation order -> deterministic.  (Cosine similarity is
//                                      scale-invariant: the sums are used, the division by n_k + 1e-6 appears only in
//                                      the eps clamp and in the gradient.)
//      ifvd_combine_kernel             when the pixels of a sample are spread over several CTAs: their partial
//                                      sums added in CTA order.
//   2. ifvd_sim_kernel                 per pixel: both similarities in one sweep over the channels (8 channel slices
//                                      per pixel), (sim_S - sim_T)^2 partials, per-pixel backward coefficients.
//   3. ifvd_class_sums_kernel<weighted> (+ combine)  the gradient reaching each centre:
//                                      U_k = sum_{p in k} g_p * S[:, p]/|S[:, p]|,  V_k = sum_{p in k} g_p * sim_S(p).
//   4. ifvd_grad_kernel                dS[:, p] = g_p*(c^/|f| - sim*f/|f|^2) + (U_k/|c| - V_k*c/|c|^2)/(n_k + 1e-6)
//   5. ifvd_finalize_kernel            loss partials summed in a fixed order (fp64).
//
// eps handling follows ATen's cosine_similarity (x / max(|x|, 1e-8) for both arguments), including the zero
// gradient through a clamped norm.
#include <type_traits>

#include "common.cuh"
#include "launch.h"
#include "params.h"

namespace sd {

constexpr int kIfvdWarps = 8;          // planes per CTA of the class-sum kernels
constexpr int kIfvdPixThreads = 32;    // pixels per CTA of the per-pixel kernels
constexpr int kIfvdSlices = 8;         // channel slices per pixel of the per-pixel kernels
constexpr float kCosEps = 1e-8f;       // ATen cosine_similarity eps (the reference uses the default)
constexpr float kCentreEps = 1e-6f;    // losses.py:229-230

// ------------------------------------------------------------------------------------------------
// class sums: out[(tensor, b), c, k] = sum over the pixels p of class k of value(c, p)
//   plain:    value = X[b, c, p]              (virtual channel c == C: 1 -> class counts), X = S (z even) or T (z odd)
//   weighted: value = a0[p] * S[b, c, p]      (virtual channel c == C: a1[p])
//
// lane = channel.  A warp owns 32 channels and a run of pixels; it stages [32 channels x 32 pixels] through a padded
// shared-memory tile (coalesced along the pixels on the way in, one channel per lane on the way out) and then walks the
// 32 pixels one by one: the class of a pixel is the same for every lane, so the control flow is warp-uniform, the bin
// `bins[k][lane]` is private to the lane, and a run of pixels of one class is summed in a register before it touches
// the bin.  No match/shuffle reductions, no atomics; every (channel, class) sum has a fixed order: pixels in order
// inside a warp, warps in order inside a CTA, CTAs (`splits` pixel ranges) in order in ifvd_combine_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kTileMaxPitch = 65;   // 64 pixel columns (bf16, 16-byte loads) + 1

template <int N>
__device__ __forceinline__ void load_vec(const float* p, float* f) {   // N floats, 16-byte aligned
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + i);
        f[i] = v.x; f[i + 1] = v.y; f[i + 2] = v.z; f[i + 3] = v.w;
    }
}
template <int N>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float* f) {   // 8 bf16, 16-byte aligned
    static_assert(N == 8, "bf16: one 16-byte load");
    Elem<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(p), f);
}

// VEC (HW % EPL == 0, 16-byte aligned tensors): a lane fetches 16 bytes = EPL consecutive pixels of a channel per load,
// 8 lanes cover a 128-byte piece of a channel row, a warp-load covers 4 channel rows; a step is 8*EPL pixels (32 fp32,
// 64 bf16 - 64-byte pieces of a row were measured 2x slower per element).  Otherwise: scalar loads, 32 pixels per step.
template <typename T, bool WEIGHTED, bool VEC>
__global__ void __launch_bounds__(kIfvdWarps * 32, 1) ifvd_class_sums_kernel(const IfvdParams p) {
    constexpr int EPL = VEC ? 16 / (int)sizeof(T) : 1;   // pixels per lane and row-load
    constexpr int PXS = VEC ? 8 * EPL : 32;              // pixels per step
    constexpr int HALVES = PXS / 32;                     // 32-pixel walks per step
    constexpr int PITCH = PXS + 1;
    constexpr int NV = VEC ? 8 * EPL : 32;               // staged values per lane and step
    extern __shared__ float smem_f[];  // bins [kIfvdWarps][K1][32], tiles [kIfvdWarps][32][kTileMaxPitch]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, HW = p.HW, K1 = C + 1;
    const int split = blockIdx.x, c0 = blockIdx.y * 32;
    const int b = WEIGHTED ? blockIdx.z : blockIdx.z >> 1;
    const bool second = !WEIGHTED && (blockIdx.z & 1);

    float* bins = smem_f + (size_t)warp * K1 * 32;
    float* tile = smem_f + (size_t)kIfvdWarps * K1 * 32 + (size_t)warp * 32 * kTileMaxPitch;
    for (int i = lane; i < K1 * 32; i += 32) bins[i] = 0.f;

    // this warp's run of steps
    const int steps = (HW + PXS - 1) / PXS;
    const int nsplit = WEIGHTED ? p.wsplits : p.splits;
    const int per_cta = (steps + nsplit - 1) / nsplit;
    const int per_warp = (per_cta + kIfvdWarps - 1) / kIfvdWarps;
    const int s0 = split * per_cta + warp * per_warp;
    const int s1 = min(min(s0 + per_warp, (split + 1) * per_cta), steps);

    const T* feat = static_cast<const T*>(second ? p.T : p.S) + ((size_t)b * C + c0) * HW;
    const int* cls = p.cls + (size_t)b * HW;
    const float* a0 = p.pix + (size_t)b * HW;
    const float* a1 = a0 + (size_t)p.B * HW;

    int cur = C;       // class of the run being summed in `acc` (C: the bin nobody reads)
    float acc = 0.f;
    // one step = PXS pixels x 32 channels; the next step's global loads fly while this one is walked.  The loads land
    // in registers RAW (no conversion, no arithmetic on them before the next step's shared-memory stores): a use right
    // behind each load would make the in-order warp wait for every load in turn (bf16: 8 round trips per step).
    using vec_t = typename Elem<T>::vec_t;
    using raw_t = typename std::conditional<VEC, vec_t, T>::type;
    constexpr int NRAW = VEC ? 8 : 32;
    raw_t raw[NRAW];
    float w[EPL], virt[EPL];   // weights of this lane's pixels; values of the virtual channel c == C
    int kreg[HALVES];
    const int sr = lane >> 3, pxl = EPL * (lane & 7);  // VEC: sub-row and first pixel of this lane's loads
    const int rv = C - c0;                              // row of the virtual channel in this group (if 0 <= rv < 32)
    auto load_step = [&](int st) {
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
            const int px = st * PXS + h * 32 + lane;
            kreg[h] = px < HW ? __ldg(cls + px) : C;
        }
        if constexpr (VEC) {
            const int q0 = st * PXS + pxl;
            const bool in = q0 < HW;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = 4 * q + sr;
                raw[q] = vec_t{};
                if (in && c0 + r < C) raw[q] = *reinterpret_cast<const vec_t*>(feat + (size_t)r * HW + q0);
            }
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                w[e] = WEIGHTED ? 0.f : 1.f;
                virt[e] = (!WEIGHTED && in) ? 1.f : 0.f;
            }
            if (WEIGHTED && in) {
                load_vec<EPL>(a0 + q0, w);
                if (rv >= 0 && rv < 32 && (rv & 3) == sr) load_vec<EPL>(a1 + q0, virt);
            }
        } else {
            const int px = st * 32 + lane;
            const bool in = px < HW;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                raw[r] = T(0.f);
                if (in && c0 + r < C) raw[r] = feat[(size_t)r * HW + px];
            }
            w[0] = WEIGHTED ? (in ? __ldg(a0 + px) : 0.f) : 1.f;
            virt[0] = in ? (WEIGHTED ? __ldg(a1 + px) : 1.f) : 0.f;
        }
    };
    if (s0 < s1) load_step(s0);
    __syncwarp();
    for (int st = s0; st < s1; ++st) {
        if constexpr (VEC) {  // bank = (row + pixel) % 32 = (4q + sr + pxl + e) % 32: distinct over the warp for EPL = 4, 2-way for 8
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = 4 * q + sr;
                float d[EPL];
                Elem<T>::unpack(raw[q], d);
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    float val = WEIGHTED ? w[e] * d[e] : d[e];
                    if (r == rv) val = virt[e];
                    tile[r * PITCH + pxl + e] = val;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                float val = (float)raw[r];
                if (WEIGHTED) val *= w[0];
                if (r == rv) val = virt[0];
                tile[r * PITCH + lane] = val;
            }
        }
        int kstep[HALVES];
#pragma unroll
        for (int h = 0; h < HALVES; ++h) kstep[h] = kreg[h];
        __syncwarp();
        if (st + 1 < s1) load_step(st + 1);
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
            const int kcur = kstep[h];
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = tile[lane * PITCH + h * 32 + j];
            // bit j: pixel j starts a new run (its class differs from its predecessor's)
            const int kprev = __shfl_up_sync(0xffffffffu, kcur, 1);
            const unsigned chg = __ballot_sync(0xffffffffu, lane == 0 ? kcur != cur : kcur != kprev);
            if (chg == 0u) {  // the whole walk continues the current run (the usual case on label maps)
                float t[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) t[q] = (x[4 * q] + x[4 * q + 1]) + (x[4 * q + 2] + x[4 * q + 3]);
                acc += ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if ((chg >> j) & 1u) {  // warp-uniform
                        bins[cur * 32 + lane] += acc;
                        acc = 0.f;
                        cur = __shfl_sync(0xffffffffu, kcur, j);
                    }
                    acc += x[j];
                }
            }
        }
        __syncwarp();
    }
    bins[cur * 32 + lane] += acc;
    __syncthreads();

    // warps summed in order; out[(tensor, b)][k][c]: one contiguous row of channels per class
    float* out = (nsplit > 1 ? p.spart + (size_t)split * (WEIGHTED ? 1 : 2) * p.B * K1 * K1 : (WEIGHTED ? p.wsum : p.sums)) +
                 ((size_t)(second ? p.B : 0) + b) * K1 * K1;
    const float* all = smem_f;
    for (int i = threadIdx.x; i < 32 * K1; i += kIfvdWarps * 32) {
        const int k = i >> 5, ch = i & 31;
        if (c0 + ch > C) continue;
        float a = 0.f;
#pragma unroll
        for (int wv = 0; wv < kIfvdWarps; ++wv) a += all[((size_t)wv * K1 + k) * 32 + ch];
        out[(size_t)k * K1 + c0 + ch] = a;
    }
}

// pixel ranges summed in order (only launched when splits > 1)
__global__ void __launch_bounds__(256) ifvd_combine_kernel(const float* spart, float* out, long long n, int splits) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += spart[(size_t)s * n + i];
    out[i] = a;
}

// ------------------------------------------------------------------------------------------------
// per pixel: sim_S, sim_T, loss partial, backward coefficients
//   pix[0] = g/|f|   pix[1] = g*sim_S   pix[2] = g*sim_S/|f|^2 (0 when |f| is clamped)   pix[3] = max(|centre|, eps)
// CTA = 32 pixels (threadIdx.x, coalesced) x kIfvdSlices channel slices (threadIdx.y): the sweep over the channels is
// split so that the two-sample batches the reference trains on still fill the GPU; slices are merged in order.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kIfvdPixThreads* kIfvdSlices) ifvd_sim_kernel(const IfvdParams p) {
    __shared__ float red[kIfvdSlices][6][kIfvdPixThreads];
    const int b = blockIdx.y;
    const int lane = threadIdx.x, slice = threadIdx.y;
    const int px = blockIdx.x * kIfvdPixThreads + lane;
    const int C = p.C, HW = p.HW, K1 = C + 1;
    const int cps = (C + kIfvdSlices - 1) / kIfvdSlices;
    const int cbeg = slice * cps, cend = min(C, cbeg + cps);
    const bool in = px < HW;
    const int k = in ? __ldg(p.cls + (size_t)b * HW + px) : C;
    float dots = 0.f, fs2 = 0.f, cs2 = 0.f, dott = 0.f, ft2 = 0.f, ct2 = 0.f;
    if (in) {
        const T* s = static_cast<const T*>(p.S) + (size_t)b * C * HW + px;
        const T* t = static_cast<const T*>(p.T) + (size_t)b * C * HW + px;
        if (k < C) {
            const float* zs = p.sums + ((size_t)b * K1 + k) * K1;   // the class's row: C channel sums, then the count
            const float* zt = zs + (size_t)p.B * K1 * K1;
#pragma unroll 4
            for (int c = cbeg; c < cend; ++c) {
                const float fs = Elem<T>::load(s + (size_t)c * HW), ft = Elem<T>::load(t + (size_t)c * HW);
                const float cs = __ldg(zs + c), ct = __ldg(zt + c);
                dots = fmaf(fs, cs, dots); fs2 = fmaf(fs, fs, fs2); cs2 = fmaf(cs, cs, cs2);
                dott = fmaf(ft, ct, dott); ft2 = fmaf(ft, ft, ft2); ct2 = fmaf(ct, ct, ct2);
            }
        } else {  // no class: the pixel is its own centre
#pragma unroll 4
            for (int c = cbeg; c < cend; ++c) {
                const float fs = Elem<T>::load(s + (size_t)c * HW), ft = Elem<T>::load(t + (size_t)c * HW);
                fs2 = fmaf(fs, fs, fs2);
                ft2 = fmaf(ft, ft, ft2);
            }
            dots = cs2 = fs2;
            dott = ct2 = ft2;
        }
    }
    red[slice][0][lane] = dots; red[slice][1][lane] = fs2; red[slice][2][lane] = cs2;
    red[slice][3][lane] = dott; red[slice][4][lane] = ft2; red[slice][5][lane] = ct2;
    __syncthreads();
    if (slice != 0) return;
    float d2 = 0.f;
    if (in) {
        dots = fs2 = cs2 = dott = ft2 = ct2 = 0.f;
#pragma unroll
        for (int q = 0; q < kIfvdSlices; ++q) {
            dots += red[q][0][lane]; fs2 += red[q][1][lane]; cs2 += red[q][2][lane];
            dott += red[q][3][lane]; ft2 += red[q][4][lane]; ct2 += red[q][5][lane];
        }
        const float inv_n = k < C ? 1.f / (__ldg(p.sums + ((size_t)b * K1 + k) * K1 + C) + kCentreEps) : 1.f;
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
    if (lane == 0) p.part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = d2;
}

// ------------------------------------------------------------------------------------------------
// gradient (same CTA shape; every slice writes its own channels)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kIfvdPixThreads* kIfvdSlices) ifvd_grad_kernel(const IfvdParams p) {
    const int b = blockIdx.y;
    const int px = blockIdx.x * kIfvdPixThreads + threadIdx.x;
    const int C = p.C, HW = p.HW, K1 = C + 1;
    const int cps = (C + kIfvdSlices - 1) / kIfvdSlices;
    const int cbeg = threadIdx.y * cps, cend = min(C, cbeg + cps);
    if (px >= HW) return;
    const int k = __ldg(p.cls + (size_t)b * HW + px);
    const T* s = static_cast<const T*>(p.S) + (size_t)b * C * HW + px;
    T* out = static_cast<T*>(p.dS) + (size_t)b * C * HW + px;
    const bool acc = p.accumulate != 0;   // dS already holds another loss's gradient (IFVDLoss: the per-pixel KL)
    if (k >= C) {
        if (!acc)
            for (int c = cbeg; c < cend; ++c) Elem<T>::store(out + (size_t)c * HW, 0.f);
        return;
    }
    const size_t o = (size_t)b * HW + px, n = (size_t)p.B * HW;
    const float a0 = p.pix[o], a2 = p.pix[2 * n + o], nc = p.pix[3 * n + o];
    const float* zs = p.sums + ((size_t)b * K1 + k) * K1;
    const float* us = p.wsum + ((size_t)b * K1 + k) * K1;
    const float inv_n = 1.f / (__ldg(zs + C) + kCentreEps);
    const float cp = inv_n / nc;                                    // d centre^ / d sum
    const float dp = nc > kCosEps ? __ldg(us + C) * cp * cp : 0.f;
    const float ks = a0 * cp - dp;
#pragma unroll 4
    for (int c = cbeg; c < cend; ++c) {
        const float f = Elem<T>::load(s + (size_t)c * HW);
        const float z = __ldg(zs + c), u = __ldg(us + c);
        const float g = fmaf(ks, z, fmaf(cp, u, -a2 * f));
        Elem<T>::store(out + (size_t)c * HW, acc ? Elem<T>::load(out + (size_t)c * HW) + g : g);
    }
}

// ------------------------------------------------------------------------------------------------
// class map: losses.py:218-224.  nn.Upsample(size, mode='nearest') of the label map (ATen's nearest index:
// identity, >> 1 for an exact doubling, else min(floor(dst * in/out), in - 1) in fp32), then the class index where the
// label equals one of range(C), else C.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in, int out, float scale) {
    if (out == in) return dst;
    if (out == 2 * in) return dst >> 1;
    return min((int)floorf((float)dst * scale), in - 1);
}
__global__ void __launch_bounds__(256) ifvd_class_map_kernel(const long long* target, int* cls, int B, int Ht, int Wt,
                                                             int h, int w, int C) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)B * h * w) return;
    const int x = (int)(i % w), y = (int)((i / w) % h), b = (int)(i / ((long long)w * h));
    const int sy = nearest_src(y, Ht, h, (float)Ht / (float)h), sx = nearest_src(x, Wt, w, (float)Wt / (float)w);
    const long long lab = target[((size_t)b * Ht + sy) * Wt + sx];
    cls[i] = lab >= 0 && lab < C ? (int)lab : C;
}

cudaError_t launch_ifvd_class_map(const long long* target, int* cls, int B, int Ht, int Wt, int h, int w, int C,
                                  cudaStream_t stream) {
    const long long n = (long long)B * h * w;
    ifvd_class_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(target, cls, B, Ht, Wt, h, w, C);
    return cudaGetLastError();
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
    const size_t smem = (size_t)kIfvdWarps * (K1 * 32 + 32 * kTileMaxPitch) * sizeof(float);
    const bool vec = p.vec != 0;
    auto sums = vec ? ifvd_class_sums_kernel<T, false, true> : ifvd_class_sums_kernel<T, false, false>;
    auto wsums = vec ? ifvd_class_sums_kernel<T, true, true> : ifvd_class_sums_kernel<T, true, false>;
    static std::atomic<bool> configured[kMaxDevices][2];  // per instantiation (T), device and load width
    const int dev = device_slot();
    if (!configured[dev][vec].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(sums, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wsums, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        configured[dev][vec].store(true, std::memory_order_release);
    }
    const int groups = (K1 + 31) / 32;
    const dim3 gsum(p.splits, groups, 2 * p.B), gwsum(p.wsplits, groups, p.B);
    const dim3 gpix((p.HW + kIfvdPixThreads - 1) / kIfvdPixThreads, p.B, 1);
    const long long plane = (long long)p.B * K1 * K1;
    sums<<<gsum, kIfvdWarps * 32, smem, stream>>>(p);
    if (p.splits > 1)
        ifvd_combine_kernel<<<(unsigned)((2 * plane + 255) / 256), 256, 0, stream>>>(p.spart, p.sums, 2 * plane, p.splits);
    ifvd_sim_kernel<T><<<gpix, dim3(kIfvdPixThreads, kIfvdSlices), 0, stream>>>(p);
    wsums<<<gwsum, kIfvdWarps * 32, smem, stream>>>(p);
    if (p.wsplits > 1)
        ifvd_combine_kernel<<<(unsigned)((plane + 255) / 256), 256, 0, stream>>>(p.spart, p.wsum, plane, p.wsplits);
    ifvd_grad_kernel<T><<<gpix, dim3(kIfvdPixThreads, kIfvdSlices), 0, stream>>>(p);
    ifvd_finalize_kernel<<<1, 256, 0, stream>>>(p.part, (int)(gpix.x * gpix.y), loss_scale, p.loss);
    return cudaGetLastError();
}

// class-sum bins (C+1 classes x 32 channels per warp) must fit the CTA's shared memory
int ifvd_max_channels() { return (int)((227 * 1024 / sizeof(float) / kIfvdWarps - 32 * kTileMaxPitch) / 32) - 1; }

int ifvd_pix_threads() { return kIfvdPixThreads; }

cudaError_t launch_ifvd_sim(const IfvdParams& p, bool bf16, float loss_scale, cudaStream_t stream) {
    return bf16 ? launch_ifvd_t<__nv_bfloat16>(p, loss_scale, stream) : launch_ifvd_t<float>(p, loss_scale, stream);
}

}  // namespace sd
