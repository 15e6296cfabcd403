// Host/device parameter blocks and workspace layouts (shared by the kernels and cabi.cu).
#pragma once

#include <stddef.h>
#include <stdint.h>

namespace sd {

// Every workspace, whatever the op or shape, starts with the same COUNTER ARENA: integer words that
// must be zero between launches (each kernel leaves them zero again).  Nothing else is ever stored
// there, so one workspace can serve any sequence of ops and shapes on a stream; the regions behind
// the arena hold floats that are always written before they are read within a launch.
constexpr int kCtrlWords = 64;     // [0] completion ticket, [1] error flag (spin time-out)
constexpr int kRowCntRing = 4096;  // split-row arrival counters, indexed by row % ring (rows in flight
                                   // at once are bounded by the persistent grid size, <= kMaxGrid)
constexpr size_t kArenaBytes = sizeof(unsigned) * (kCtrlWords + kRowCntRing);
constexpr int kMaxGrid = 1024;     // upper bound on persistent-grid size (per-CTA partial slots)
constexpr int kGenericChunk = 4096;  // elements of one channel plane per generic work unit
constexpr int kPartWords = 8;      // floats per published unit partial

// ---------------------------------------------------------------- rows (CD / CGD)
struct RowsParams {
    const void* S;
    const void* T;
    void* dS;
    float* row_kl;        // [R] (user buffer or workspace)
    float* loss;          // [1]
    float* mse_loss;      // [1] or null
    const int32_t* perm;  // [C] or null
    int B, C, HW, g;
    int G;                // groups (rows) per sample = ceil(C/g)
    int G_full;           // complete groups per sample = C/g
    int g_last;           // channels in the ragged last group (0 = none)
    int R;                // B*G
    float c2;             // log2(e)/tau
    float inv_tau;
    float coef;           // grad_scale*alpha/(R*tau)
    float loss_scale;     // alpha/R
    float mse_gcoef;      // grad_scale*2*w/numel   (0 = MSE off)
    float mse_scale;      // w/numel
    // TMA kernel work decomposition: a unit is one chunk of one row
    int chunk_elems;      // logical row elements per chunk (multiple of the vector width)
    int nch_full;         // chunks per complete row
    int nch_last;         // chunks per ragged row
    int units_per_sample;
    long long total_units;
    // generic kernel decomposition: unit = (b, logical channel, plane chunk)
    int KC;               // chunks per channel plane
    // workspace
    unsigned* ctrl;
    float* cta_part;      // [2][kMaxGrid]
    unsigned* row_cnt;    // [kRowCntRing]
    float* unit_part;     // [units][kPartWords]
};

struct RowsWorkspace {
    size_t off_ctrl, off_rowcnt, off_cta, off_rowkl, off_unit, bytes;
};
inline RowsWorkspace rows_workspace_layout(long long B, long long C, long long HW, long long g) {
    RowsWorkspace w;
    const long long G = (C + g - 1) / g;
    const long long R = B * G;
    const long long KC = (HW + kGenericChunk - 1) / kGenericChunk;
    const long long units = B * C * KC;  // >= number of TMA units as well
    size_t o = 0;
    w.off_ctrl = o;   o += sizeof(unsigned) * kCtrlWords;
    w.off_rowcnt = o; o += sizeof(unsigned) * kRowCntRing;   // == kArenaBytes
    w.off_cta = o;    o += sizeof(float) * 2 * kMaxGrid;
    w.off_rowkl = o;  o += sizeof(float) * (size_t)R;
    o = (o + 31) & ~(size_t)31;
    w.off_unit = o;   o += sizeof(float) * kPartWords * (size_t)units;
    w.bytes = (o + 255) & ~(size_t)255;
    return w;
}

// ---------------------------------------------------------------- pixels (PD / AT)
struct PixParams {
    const void* S;
    const void* T;
    void* dS;
    float* row_kl;   // [B*HW]
    float* loss;
    float* at_loss;  // or null
    int B, C, HW;
    int tiles_per_sample;
    long long total_tiles;
    float c2, inv_tau, coef, loss_scale;
    float at_gcoef;  // grad_scale*2*w/(C*B*HW)  (0 = AT term off)
    float at_scale;  // w/(B*HW)
    float inv_C;
    int nstages;
    unsigned stage_bytes;  // bytes of one tensor's tile in shared memory (C*256)
    unsigned* ctrl;
    float* cta_part;  // [2][nparts]
    int nparts;
};

struct PixWorkspace {
    size_t off_ctrl, off_cta, off_rowkl, bytes;
    long long nparts;
};
inline PixWorkspace pix_workspace_layout(long long B, long long C, long long HW) {
    PixWorkspace w;
    (void)C;
    const long long R = B * HW;
    long long nparts = (R + 255) / 256;
    if (nparts < kMaxGrid) nparts = kMaxGrid;
    w.nparts = nparts;
    size_t o = 0;
    w.off_ctrl = o;  o = kArenaBytes;
    w.off_cta = o;   o += sizeof(float) * 2 * (size_t)nparts;
    w.off_rowkl = o; o += sizeof(float) * (size_t)R;
    w.bytes = (o + 255) & ~(size_t)255;
    return w;
}

// ---------------------------------------------------------------- MSE
constexpr int kMseMaxGrid = 148 * 8;

}  // namespace sd
