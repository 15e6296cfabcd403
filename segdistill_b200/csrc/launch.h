// Host-side launchers implemented next to their kernels; called from cabi.cu only.
#pragma once

#include <cuda_runtime.h>

#include <atomic>

#include "params.h"

namespace sd {

// Function attributes (the opt-in to > 48 KB of dynamic shared memory) and occupancy answers belong to a device
// (context), not to the process: the launchers cache them per device ordinal - one process may drive several GPUs
// (MMDataParallel, `with torch.cuda.device(...)`).  The cached values are idempotent, the flags atomic.
constexpr int kMaxDevices = 64;
inline int device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return (d < 0 || d >= kMaxDevices) ? 0 : d;
}

// kl_rows.cu
cudaError_t launch_kl_rows_tma(const RowsParams& p, bool bf16, int grid, cudaStream_t stream);
cudaError_t launch_kl_rows_generic(const RowsParams& p, bool bf16, cudaStream_t stream);
cudaError_t launch_kl_rows_rm(const RowsParams& p, int grid, cudaStream_t stream);
int kl_rows_tma_chunk_capacity();
cudaError_t launch_kl_rows_pack(const RowsParams& p, bool bf16, int grid, cudaStream_t stream);
// kl_rows_stream.cu
cudaError_t launch_kl_rows_stream(const RowsParams& p, bool bf16, int sms, cudaStream_t stream);
int kl_rows_stream_chunk_capacity();

// kl_rows_group.cu
cudaError_t launch_kl_rows_group(const GroupParams& gp, int max_row_units, bool bf16, int sms, cudaStream_t stream);
int kl_rows_group_chunk_capacity();

// kl_rows_cluster.cu   (probe_only: just answer whether a cluster of g.nc CTAs can be resident)
cudaError_t launch_kl_rows_cluster(const RowsParams& p, const ClusterGeom& g, bool bf16, int sms, cudaStream_t stream,
                                   bool probe_only);

// kl_rows_grid.cu   (probe_only: just answer whether the kernel can be resident on this device)
cudaError_t launch_kl_rows_grid(const RowsParams& p, bool bf16, int sms, cudaStream_t stream, bool probe_only);
cudaError_t launch_kl_rows_grid_group(const GroupParams& gp, int max_row_units, const int* knobs, bool bf16, int sms,
                                      cudaStream_t stream, bool probe_only);

// kl_rows_up.cu
cudaError_t launch_kl_rows_up(const UpParams& p, bool bf16, int sms, cudaStream_t stream);
cudaError_t launch_kl_pixels_up(const UpParams& p, bool bf16, int sms, cudaStream_t stream);   // p.part: [kMaxGrid] CTA partials

// ce_up.cu
cudaError_t launch_ce_up(const CeParams& p, bool bf16, int sms, cudaStream_t stream);

// kl_pixels.cu   (mapS/mapT point at CUtensorMap objects)
cudaError_t launch_kl_pixels_tma(const void* mapS, const void* mapT, const PixParams& p, bool bf16, int cols, int grid,
                                 size_t smem, cudaStream_t stream);
cudaError_t launch_kl_pixels_generic(const PixParams& p, bool bf16, cudaStream_t stream);
// kl_pixels_warp.cu   (no AT term, C <= 256; maps with 128-byte swizzle and boxes of (128 bytes of pixels, C, 1))
cudaError_t launch_kl_pixels_warp(const void* mapS, const void* mapT, const void* mapD, const PixParams& p, bool bf16, int grid,
                                  cudaStream_t stream);
int kl_pixels_warp_tile_pixels(bool bf16);
int pix_warp_stages(int C);    // ring stages that fit next to C channels (0: none)
size_t pix_tma_smem_bytes(int C, int pxt, int nstages, int cols);   // cols: 64 (one CTA per SM) or 32 (bf16: two)
int kl_pixels_tma_max_channels(bool bf16);
int kl_pixels_tile_pixels(bool bf16, int cols);

// corr_gemm.cu   (maps: CUtensorMap[4] built by cabi.cu with the box rows cgd_corr_geometry reports)
size_t cgd_corr_workspace_bytes(int B, int C, int HW, int group);
bool cgd_corr_geometry(int B, int C, int HW, int group, int dtype, int* rows0, int* rows1, int* kbox);
cudaError_t launch_cgd_corr(void* dS, float* loss, int B, int C, int HW, int group, int dtype, float alpha, float grad_scale,
                            void* workspace, const void* maps, cudaStream_t stream);

// mse.cu
cudaError_t launch_mse(const void* S, const void* T, void* dS, float* loss, float* partials, long long n, bool bf16,
                       float gcoef, float scale, int grid, cudaStream_t stream);
cudaError_t launch_scale_grad(void* dS, long long n, bool bf16, const float* g, int grid, cudaStream_t stream,
                              const float* log_values, int log_n, float* log_ring, unsigned* log_cursor, int log_slots);
cudaError_t launch_scale_grad_group(int n_tensors, void* const* dS, const long long* numel, bool bf16, const float* const* g,
                                    int grid, cudaStream_t stream);
cudaError_t launch_scale_grad2(void* dS, long long n, bool bf16, const float* g0, const float* g1, unsigned* flag,
                               int grid, cudaStream_t stream);

cudaError_t launch_log_push(const float* values, int n, float* ring, unsigned* cursor, int slots, cudaStream_t stream);

// ifvd.cu
cudaError_t launch_ifvd_sim(const IfvdParams& p, bool bf16, float loss_scale, cudaStream_t stream);
cudaError_t launch_ifvd_class_map(const long long* target, int* cls, int B, int Ht, int Wt, int h, int w, int C,
                                  cudaStream_t stream);
int ifvd_pix_threads();
int ifvd_max_channels();

}  // namespace sd
