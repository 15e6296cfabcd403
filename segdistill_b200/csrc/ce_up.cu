// The student head's supervised loss at label resolution, forward + backward fused, the up-sampled logits never
// materialised:  BaseDecodeHead.losses (mmseg/models/decode_heads/decode_head.py:217-237) =
//     resize(seg_logit, size=label.shape[2:], bilinear, align_corners=False)        (:221-225)
//  -> CrossEntropyLoss / cross_entropy (mmseg/models/losses/cross_entropy_loss.py:9-32, :138-198: F.cross_entropy with
//     reduction='none', ignore_index, class_weight; then weight_reduce_loss, losses/utils.py:25-56)
//  -> accuracy(seg_logit, seg_label) top-1 (mmseg/models/losses/accuracy.py:4-46)
// plus autograd's backward through all of it.  The reference runs this on (B, 150, 512, 512) fp32 maps (157 MB per
// sample pair of temporaries); here a thread owns one LOW-resolution cell, regenerates the s x s up-sampled pixels of
// its block per channel from the 3 x 3 cells around it (constant bilinear weights for an integer scale, up_common.cuh),
// keeps the per-pixel statistics in registers over two sweeps of the channels and sends the gradient back through the
// transposed stencil - the structure of kl_pixels_up_kernel (kl_rows_up.cu), with the teacher replaced by a label:
//
//   sweep 1   per pixel: Z = sum_c exp(v_c - ref), the largest v_c, and v at the pixel's label channel
//             -> nll = ln Z + ref - v_y (times class / pixel weight, 0 for ignored pixels), correct = (v_y is the largest)
//   sweep 2   per channel: g = coef_pixel (softmax_c - [c == y]) folded onto the 3 x 3 cells (transposed stencil, nine
//             shared-memory planes, fixed summation order: deterministic), written once at low resolution.
//
// scale 1 (labels at logit resolution) runs the same code with a one-pixel block.  fp32 arithmetic; dX has the dtype
// of the logits.
#include <type_traits>

#include "up_common.cuh"
#include "launch.h"

namespace sd {

template <int NP>
struct CeSmem {
    float st[2][kPxCh][kPxLoad * kPxLoad];   // [stage][channel][cell]
    float planes[2][9][kPxPlane];
    float red[kPxThreads / 32];
    // exact redo (a pixel's sum vanished against the cell's reference - neighbouring cells more than ~87 apart): one
    // reference per pixel, fl(max_c v * log2e); only the owning thread reads its column
    float r2[NP][kPxThreads];
};

template <typename T, int S>
__global__ void __launch_bounds__(kPxThreads, 2) ce_up_kernel(const CeParams p) {
    constexpr int SB = S > 4 ? 4 : S;              // window side
    constexpr int NWIN = (S / SB) * (S / SB);
    constexpr int NP = SB * SB;
    __shared__ CeSmem<NP> sm;
    const int tid = threadIdx.x;
    const int ty = tid / kPxTile, tx = tid % kPxTile;
    const int tiles_x = (p.Wl + kPxOwn - 1) / kPxOwn, tiles_y = (p.Hl + kPxOwn - 1) / kPxOwn;
    const long long n_tiles = (long long)p.B * tiles_y * tiles_x;
    const size_t plane_elems = (size_t)p.Hl * p.Wl;
    const int Ws = p.Wl * S;
    const size_t label_plane = (size_t)p.Hl * S * Ws;
    float loss_acc = 0.f, hit_acc = 0.f;
    for (long long unit = blockIdx.x; unit < n_tiles * NWIN; unit += gridDim.x) {
        const long long tile = unit / NWIN;
        const int my_win = (int)(unit - tile * NWIN);
        const int b = (int)(tile / (tiles_y * tiles_x));
        const int trem = (int)(tile - (long long)b * tiles_y * tiles_x);
        const int i0 = (trem / tiles_x) * kPxOwn - 1, j0 = (trem % tiles_x) * kPxOwn - 1;   // first computed cell
        const int i = i0 + ty, j = j0 + tx;                                                 // my cell
        const bool in_map = i >= 0 && i < p.Hl && j >= 0 && j < p.Wl;
        const bool owned = in_map && ty >= 1 && ty <= kPxOwn && tx >= 1 && tx <= kPxOwn;
        const T* gX = static_cast<const T*>(p.X) + (size_t)b * p.C * plane_elems;
        auto load_chunk = [&](int c0, int stage) {
            const int nch = min(kPxCh, p.C - c0);
            for (int e = tid; e < nch * kPxLoad * kPxLoad; e += kPxThreads) {
                const int ch = e / (kPxLoad * kPxLoad), cell = e - ch * (kPxLoad * kPxLoad);
                const int ly = cell / kPxLoad, lx = cell - ly * kPxLoad;
                const int yi = min(max(i0 - 1 + ly, 0), p.Hl - 1), xj = min(max(j0 - 1 + lx, 0), p.Wl - 1);
                sm.st[stage][ch][cell] = up_load<T>(gX + (size_t)(c0 + ch) * plane_elems + (size_t)yi * p.Wl + xj);
            }
        };
        auto nbhd = [&](const float* cellp, float (&a)[3][3]) {
            const float* q = cellp + ty * kPxLoad + tx;          // loaded cell (ty, tx) = my cell (-1, -1)
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int e = 0; e < 3; ++e) a[d][e] = q[d * kPxLoad + e];
        };
        const int n_chunks = (p.C + kPxCh - 1) / kPxCh;
        T* gD = static_cast<T*>(p.dX) + (size_t)b * p.C * plane_elems;
        // the s x s block of a cell is walked in windows of SB x SB pixels (one window for s <= 4; at s = 8 a unit is
        // (tile, window), unrolled so that the tap weights stay compile-time constants)
#pragma unroll
        for (int win = 0; win < NWIN; ++win) {
        if (NWIN > 1 && win != my_win) continue;          // (uniform over the CTA)
        const int ky0 = (win / (S / SB)) * SB, kx0 = (win % (S / SB)) * SB;

        // ---- the labels and weights of my window's pixels
        int lab[NP];
        float wq[NP];                      // class weight x pixel weight, 0 where the pixel does not count
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            lab[q] = -1;
            wq[q] = 0.f;
        }
        if (in_map) {
#pragma unroll
            for (int ky = 0; ky < SB; ++ky) {
#pragma unroll
                for (int kx = 0; kx < SB; ++kx) {
                    const size_t off = (size_t)b * label_plane + (size_t)(i * S + ky0 + ky) * Ws + (size_t)(j * S + kx0 + kx);
                    const long long y = p.label[off];
                    const bool cls = y >= 0 && y < p.C;
                    lab[ky * SB + kx] = cls ? (int)y : -1;
                    float w = (cls && y != p.ignore_index) ? 1.f : 0.f;
                    if (w != 0.f && p.class_weight) w = p.class_weight[y];
                    if (w != 0.f && p.pix_weight) w *= p.pix_weight[off];
                    wq[ky * SB + kx] = w;
                }
            }
        }

        // ---------------- sweep 1: per-pixel sum of exponentials (reference = running maximum of the cell's 3 x 3
        // neighbourhood over the channels: an up-sampled value never exceeds it), largest value, value at the label
        float zs[NP], best[NP], vy[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            zs[q] = 0.f;
            best[q] = -3.0e38f;
            vy[q] = -3.0e38f;
        }
        float ref = kUpFloor;
        __syncthreads();
        load_chunk(0, 0);
        for (int ck = 0; ck < n_chunks; ++ck) {
            __syncthreads();
            if (ck + 1 < n_chunks) load_chunk((ck + 1) * kPxCh, (ck + 1) & 1);
            const int nch = min(kPxCh, p.C - ck * kPxCh);
            if (in_map) {
                for (int ch = 0; ch < nch; ++ch) {
                    const int c = ck * kPxCh + ch;
                    float a[3][3], hs[3][SB], dv[2][SB];
                    nbhd(sm.st[ck & 1][ch], a);
                    float m = a[0][0];
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int e = 0; e < 3; ++e) m = fmaxf(m, a[d][e]);
                    up_hrows_win<S, SB>(a, kx0, hs);
                    if (m > ref) {
                        const float f = ref_factor(ref, m, kLog2e);
#pragma unroll
                        for (int q = 0; q < NP; ++q) zs[q] *= f;
                        ref = m;
                    }
                    up_vdiff<SB>(hs, dv);
                    const float r2 = __fmul_rn(ref, kLog2e);
#pragma unroll
                    for (int ky = 0; ky < SB; ++ky) {
#pragma unroll
                        for (int kx = 0; kx < SB; ++kx) {
                            const int q = ky * SB + kx;
                            const float v = up_value_win<S, SB>(hs, dv, ky0 + ky, kx);
                            zs[q] += fast_exp2(fmaf(v, kLog2e, -r2));
                            best[q] = fmaxf(best[q], v);
                            vy[q] = c == lab[q] ? v : vy[q];
                        }
                    }
                }
            }
        }
        // ---- a pixel whose sum vanished against the cell's reference: the CTA sums again, every pixel against its own
        // maximum over the channels (best[], already known)
        bool bad = false;
        if (in_map) {
#pragma unroll
            for (int q = 0; q < NP; ++q) bad = bad || !(zs[q] >= 1e-30f && zs[q] < 3e38f);
        }
        const bool exact = __syncthreads_or(bad) != 0;
        if (exact) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                sm.r2[q][tid] = __fmul_rn(best[q], kLog2e);
                zs[q] = 0.f;
            }
            load_chunk(0, 0);
            for (int ck = 0; ck < n_chunks; ++ck) {
                __syncthreads();
                if (ck + 1 < n_chunks) load_chunk((ck + 1) * kPxCh, (ck + 1) & 1);
                const int nch = min(kPxCh, p.C - ck * kPxCh);
                if (in_map) {
                    for (int ch = 0; ch < nch; ++ch) {
                        float a[3][3], hs[3][SB], dv[2][SB];
                        nbhd(sm.st[ck & 1][ch], a);
                        up_hrows_win<S, SB>(a, kx0, hs);
                        up_vdiff<SB>(hs, dv);
#pragma unroll
                        for (int ky = 0; ky < SB; ++ky) {
#pragma unroll
                            for (int kx = 0; kx < SB; ++kx) {
                                const int q = ky * SB + kx;
                                zs[q] += fast_exp2(fmaf(up_value_win<S, SB>(hs, dv, ky0 + ky, kx), kLog2e, -sm.r2[q][tid]));
                            }
                        }
                    }
                }
            }
        }
        // loss and accuracy of my pixels (owned cells only); then the per-pixel gradient factors
        const float r2 = __fmul_rn(ref, kLog2e);
        if (owned) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                // -log softmax_y = ln Z + ref - v_y = ln2 (log2 Z - (v_y log2e - ref log2e))
                const float nll = kLn2 * (log2f(zs[q]) - fmaf(vy[q], kLog2e, exact ? -sm.r2[q][tid] : -r2));
                if (wq[q] != 0.f) loss_acc = fmaf(wq[q], nll, loss_acc);
                if (lab[q] >= 0 && vy[q] >= best[q]) hit_acc += 1.f;
            }
        }
        // best[] and vy[] are free now: best[q] = coef_q = gscale * w_q, zs[q] = coef_q / Z_q
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            best[q] = p.gscale * wq[q];
            zs[q] = in_map ? __fdividef(best[q], zs[q]) : 0.f;
        }

        // ---------------- sweep 2: per channel, window gradient -> nine contributions -> owned cells
        // (two compiled copies: the cell's reference in a register, or a reference per pixel from shared memory)
        auto sweep2 = [&](auto exact_tag) {
            constexpr bool EXACT = decltype(exact_tag)::value;
            __syncthreads();
            load_chunk(0, 0);
            int cglob = 0;
            for (int ck = 0; ck < n_chunks; ++ck) {
                __syncthreads();
                if (ck + 1 < n_chunks) load_chunk((ck + 1) * kPxCh, (ck + 1) & 1);
                const int nch = min(kPxCh, p.C - ck * kPxCh);
                for (int ch = 0; ch < nch; ++ch, ++cglob) {
                    float* pl = sm.planes[cglob & 1][0];
                    if (in_map) {
                        float a[3][3], hs[3][SB], dv[2][SB];
                        nbhd(sm.st[ck & 1][ch], a);
                        up_hrows_win<S, SB>(a, kx0, hs);
                        up_vdiff<SB>(hs, dv);
                        float m[3][3];
    #pragma unroll
                        for (int d = 0; d < 3; ++d)
    #pragma unroll
                            for (int e = 0; e < 3; ++e) m[d][e] = 0.f;
    #pragma unroll
                        for (int ky = 0; ky < SB; ++ky) {
                            float tr[3] = {0.f, 0.f, 0.f};
    #pragma unroll
                            for (int kx = 0; kx < SB; ++kx) {
                                const int q = ky * SB + kx;
                                const float es = fast_exp2(fmaf(up_value_win<S, SB>(hs, dv, ky0 + ky, kx), kLog2e, EXACT ? -sm.r2[q][tid] : -r2));
                                const float gv = fmaf(es, zs[q], cglob == lab[q] ? -best[q] : 0.f);   // coef (softmax_c - [c == y])
                                const int f = UpW<S>::first(kx0 + kx) + 1;
                                const float w1 = UpW<S>::w1(kx0 + kx);
                                tr[f] = fmaf(1.f - w1, gv, tr[f]);
                                tr[f + 1] = fmaf(w1, gv, tr[f + 1]);
                            }
                            const int f = UpW<S>::first(ky0 + ky) + 1;
                            const float w1 = UpW<S>::w1(ky0 + ky);
    #pragma unroll
                            for (int e = 0; e < 3; ++e) {
                                m[f][e] = fmaf(1.f - w1, tr[e], m[f][e]);
                                m[f + 1][e] = fmaf(w1, tr[e], m[f + 1][e]);
                            }
                        }
                        // taps clamped at the border of the map fall onto the cell itself
                        if (i == 0) {
    #pragma unroll
                            for (int e = 0; e < 3; ++e) { m[1][e] += m[0][e]; m[0][e] = 0.f; }
                        }
                        if (i == p.Hl - 1) {
    #pragma unroll
                            for (int e = 0; e < 3; ++e) { m[1][e] += m[2][e]; m[2][e] = 0.f; }
                        }
                        if (j == 0) {
    #pragma unroll
                            for (int d = 0; d < 3; ++d) { m[d][1] += m[d][0]; m[d][0] = 0.f; }
                        }
                        if (j == p.Wl - 1) {
    #pragma unroll
                            for (int d = 0; d < 3; ++d) { m[d][1] += m[d][2]; m[d][2] = 0.f; }
                        }
    #pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const int ry = ty + d - 1;
                            if (ry >= 0 && ry < kPxTile) {
    #pragma unroll
                                for (int e = 0; e < 3; ++e) pl[(d * 3 + e) * kPxPlane + ry * (kPxTile + 2) + tx + e] = m[d][e];
                            }
                        }
                    }
                    __syncthreads();
                    if (owned) {
                        // plane (d, e) at my position holds what cell (i - d + 1, j - e + 1) sent here
                        const float* q = pl + ty * (kPxTile + 2) + tx + 1;
                        const bool okd[3] = {i + 1 < p.Hl, true, i >= 1};
                        const bool oke[3] = {j + 1 < p.Wl, true, j >= 1};
                        float v = 0.f;
    #pragma unroll
                        for (int d = 0; d < 3; ++d)
    #pragma unroll
                            for (int e = 0; e < 3; ++e) {
                                const float x = q[(d * 3 + e) * kPxPlane];
                                v += (okd[d] && oke[e]) ? x : 0.f;
                            }
                        const size_t off = (size_t)cglob * plane_elems + (size_t)i * p.Wl + j;
                        if (NWIN > 1) p.wpart[((size_t)win * p.B + b) * p.C * plane_elems + off] = v;
                        else up_store<T>(gD + off, v);
                    }
                }
            }
        };
        if (exact) sweep2(std::true_type{});
        else sweep2(std::false_type{});
        }   // windows
    }
    // ---- loss and hit count: CTA partials, the last CTA sums them in a fixed order
    loss_acc = block_sum_n<kPxThreads>(loss_acc, sm.red);
    hit_acc = block_sum_n<kPxThreads>(hit_acc, sm.red);
    __shared__ unsigned ticket_s;
    if (tid == 0) {
        __stcg(&p.part[blockIdx.x], loss_acc);
        __stcg(&p.part[kMaxGrid + blockIdx.x], hit_acc);
        __threadfence();
        ticket_s = atomicAdd(&p.ctrl[0], 1u);
    }
    __syncthreads();
    if (ticket_s == gridDim.x - 1) {
        __threadfence();
        double acc = 0.0, hits = 0.0;
        for (int r = tid; r < (int)gridDim.x; r += kPxThreads) {
            acc += (double)__ldcg(&p.part[r]);
            hits += (double)__ldcg(&p.part[kMaxGrid + r]);
        }
        __shared__ double dred[2][kPxThreads / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_down_sync(0xffffffffu, acc, o);
            hits += __shfl_down_sync(0xffffffffu, hits, o);
        }
        if ((tid & 31) == 0) {
            dred[0][tid >> 5] = acc;
            dred[1][tid >> 5] = hits;
        }
        __syncthreads();
        if (tid == 0) {
            double t = 0.0, h = 0.0;
            for (int w = 0; w < kPxThreads / 32; ++w) {
                t += dred[0][w];
                h += dred[1][w];
            }
            *p.loss = (float)((double)p.lscale * t);
            if (p.acc) *p.acc = (float)((double)p.acc_scale * h);
            atomicExch(&p.ctrl[0], 0u);
        }
    }
}

// dX = sum of the window planes (s = 8), fixed order
template <typename T>
__global__ void __launch_bounds__(256) ce_sum_windows_kernel(const float* __restrict__ wpart, T* __restrict__ dX,
                                                             long long n, int nwin) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        float v = 0.f;
        for (int w = 0; w < nwin; ++w) v += wpart[(size_t)w * n + i];
        up_store<T>(dX + i, v);
    }
}

template <typename T, int S>
static cudaError_t launch_ce_t(const CeParams& p, int sms, cudaStream_t stream) {
    auto k = ce_up_kernel<T, S>;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kPxThreads, 0);
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    constexpr int nwin = S > 4 ? (S / 4) * (S / 4) : 1;
    const long long units = (long long)p.B * ((p.Hl + kPxOwn - 1) / kPxOwn) * ((p.Wl + kPxOwn - 1) / kPxOwn) * nwin;
    long long grid = (long long)sms * occ;
    if (grid > units) grid = units;
    if (grid > kMaxGrid) grid = kMaxGrid;
    k<<<(unsigned)grid, kPxThreads, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || nwin == 1) return e;
    const long long n = (long long)p.B * p.C * p.Hl * p.Wl;
    long long g2 = (n + 255) / 256;
    if (g2 > (long long)sms * 8) g2 = (long long)sms * 8;
    ce_sum_windows_kernel<T><<<(unsigned)g2, 256, 0, stream>>>(p.wpart, static_cast<T*>(p.dX), n, nwin);
    return cudaGetLastError();
}

cudaError_t launch_ce_up(const CeParams& p, bool bf16, int sms, cudaStream_t stream) {
    switch (p.scale) {
        case 1: return bf16 ? launch_ce_t<__nv_bfloat16, 1>(p, sms, stream) : launch_ce_t<float, 1>(p, sms, stream);
        case 2: return bf16 ? launch_ce_t<__nv_bfloat16, 2>(p, sms, stream) : launch_ce_t<float, 2>(p, sms, stream);
        case 4: return bf16 ? launch_ce_t<__nv_bfloat16, 4>(p, sms, stream) : launch_ce_t<float, 4>(p, sms, stream);
        case 8: return bf16 ? launch_ce_t<__nv_bfloat16, 8>(p, sms, stream) : launch_ce_t<float, 8>(p, sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sd
