// Pieces shared by the kernels that up-sample bilinearly on the fly (kl_rows_up.cu: softmax-KL behind the reference's
// resize; ce_up.cu: the student head's cross-entropy behind the same resize): tap weights of an integer scale,
// separable regeneration of a cell's block from its 3 x 3 neighbourhood, the pixel-mode tile geometry.
#pragma once

#include "common.cuh"
#include "params.h"

namespace sd {

constexpr float kUpFloor = -1.0e29f;

template <typename T>
__device__ __forceinline__ float up_load(const T* p);
template <>
__device__ __forceinline__ float up_load<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float up_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void up_store(T* p, float v);
template <>
__device__ __forceinline__ void up_store<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void up_store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// weights of up-sampled index s*i + k on the cells (i-1, i, i+1): PyTorch's source index (k + 1/2)/s - 1/2 + i
template <int S>
struct UpW {
    // tap pair of phase k: (i-1, i) for k < S/2, (i, i+1) otherwise; w1 = weight of the second tap
    static __device__ __forceinline__ constexpr int first(int k) { return k < S / 2 ? -1 : 0; }
    static __device__ __forceinline__ constexpr float w1(int k) {
        return k < S / 2 ? (k + 0.5f) / S + 0.5f : (k + 0.5f) / S - 0.5f;
    }
};

// horizontal pass: h[d][kx] = value of neighbourhood row d at up-sampled column s*j + kx, as x0 + w1 (x1 - x0)
template <int S>
__device__ __forceinline__ void up_hrows(const float (&a)[3][3], float (&h)[3][S]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float dx[2] = {a[d][1] - a[d][0], a[d][2] - a[d][1]};
#pragma unroll
        for (int kx = 0; kx < S; ++kx) {
            const int f = UpW<S>::first(kx) + 1;
            h[d][kx] = fmaf(UpW<S>::w1(kx), dx[f], a[d][f]);
        }
    }
}
// vertical differences of the three interpolated rows: v(ky, kx) = h[f][kx] + w1 dv[f][kx]
template <int S>
__device__ __forceinline__ void up_vdiff(const float (&h)[3][S], float (&dv)[2][S]) {
#pragma unroll
    for (int kx = 0; kx < S; ++kx) {
        dv[0][kx] = h[1][kx] - h[0][kx];
        dv[1][kx] = h[2][kx] - h[1][kx];
    }
}
template <int S>
__device__ __forceinline__ float up_value(const float (&h)[3][S], const float (&dv)[2][S], int ky, int kx) {
    const int f = UpW<S>::first(ky) + 1;
    return fmaf(UpW<S>::w1(ky), dv[f][kx], h[f][kx]);
}

// the same for a window of SB x SB values of the block (columns kx0 .., rows ky0 ..): pixel mode at s = 8 walks the
// block in four 4 x 4 windows so that the per-pixel statistics fit the registers
template <int S, int SB>
__device__ __forceinline__ void up_hrows_win(const float (&a)[3][3], int kx0, float (&h)[3][SB]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float dx[2] = {a[d][1] - a[d][0], a[d][2] - a[d][1]};
#pragma unroll
        for (int kx = 0; kx < SB; ++kx) {
            const int f = UpW<S>::first(kx0 + kx) + 1;
            h[d][kx] = fmaf(UpW<S>::w1(kx0 + kx), dx[f], a[d][f]);
        }
    }
}
template <int S, int SB>
__device__ __forceinline__ float up_value_win(const float (&h)[3][SB], const float (&dv)[2][SB], int ky, int kx) {
    const int f = UpW<S>::first(ky) + 1;     // ky: row within the whole block
    return fmaf(UpW<S>::w1(ky), dv[f][kx], h[f][kx]);
}

template <int NT>
__device__ __forceinline__ float block_sum_n(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) r += red[w];
    return r;
}

// ---------------------------------------------------------------- pixel-mode tiles (one thread per low-res cell)
constexpr int kPxTile = 16;                    // computed cells per tile side = threads per side
constexpr int kPxOwn = kPxTile - 2;            // owned (output) cells per tile side
constexpr int kPxLoad = kPxTile + 2;           // loaded cells per tile side (3 x 3 neighbourhoods)
constexpr int kPxCh = 4;                       // channels per shared-memory stage (static shared memory stays under 48 KB)
constexpr int kPxThreads = kPxTile * kPxTile;
constexpr int kPxPlane = kPxTile * (kPxTile + 2);   // a contribution plane of the tile, padded columns


}  // namespace sd
