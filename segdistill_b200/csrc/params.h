// Host/device parameter blocks and workspace layouts (shared by the kernels and cabi.cu).
#pragma once

#include <stddef.h>
#include <stdint.h>

namespace sd {

// Every workspace, whatever the op or shape, starts with the same COUNTER ARENA: integer words that
// must be zero between launches (each kernel leaves them zero again).  Nothing else is ever stored
// there, so one workspace can serve any sequence of ops and shapes on a stream; the regions behind
// the arena hold data that is always written before it is read within a launch, or that is
// validated by a per-launch epoch tag (the unit packets of the split-row exchange).
constexpr int kCtrlWords = 64;  // [0] completion ticket, [1] error flag (spin time-out), [2] launch epoch of the
                                // streaming kernel (tag of its unit packets; counts launches, never reset)
constexpr size_t kArenaBytes = sizeof(unsigned) * kCtrlWords;

constexpr int kMaxGrid = 1024;       // upper bound on persistent-grid size (per-CTA partial slots)
constexpr int kGenericChunk = 4096;  // elements of one channel plane per generic work unit
constexpr int kPartWords = 8;        // floats per published unit partial (generic kernels)
constexpr int kMaxLosses = 2;        // softmax-KL losses fused over one (S, T) pair in one pass
constexpr int kPktWords = 16;        // 8-byte {value, epoch} words per unit packet (6 per loss, padded)
constexpr int kChunkCapMin = 4096;   // smallest unit of the TMA kernels (grid-resident kernel: one 4096-element chunk)
constexpr int kGridChunkElems = 4096;     // grid-resident kernel (kl_rows_grid.cu): elements per chunk and tensor ...
constexpr int kGridUnitMaxChunks = 4;     // ... and chunks per unit at most (two units must fit the 8 TMEM chunk slots)
constexpr int kGridMaxRowUnits = 64;      // ... units of the longest row of any fused loss (two packets per lane of the stats warp)

// one softmax-KL loss over rows of `g` consecutive (gathered) channels x HW
struct RowLoss {
    int g;             // channels per row
    int m;             // rows of l[0] per row of this loss (1 for l[0] itself)
    int G;             // rows per sample = ceil(C/g)
    int R;             // B*G
    float c2;          // log2(e)/tau
    float inv_tau;
    float coef;        // grad_scale*alpha/(R*tau)
    float loss_scale;  // alpha/R
    float* loss;       // [1]
    float* row_kl;     // [R] or null
};

// ---------------------------------------------------------------- rows (CD / CGD)
struct RowsParams {
    const void* S;
    const void* T;
    void* dS;
    const int32_t* perm;  // [C] or null (single-loss launches only)
    int B, C, HW;
    int nl;               // fused losses (1..kMaxLosses); l[0] has the smallest rows and every row of
                          // l[k] is a union of whole rows of l[0]  (l[k].g % l[0].g == 0)
    RowLoss l[kMaxLosses];
    float* mse_loss;      // [1] or null
    float mse_gcoef;      // grad_scale*2*w/numel   (0 = MSE off)
    float mse_scale;      // w/numel
    // TMA kernel work decomposition: a unit is one chunk of one row of l[0]
    int G_full;           // complete l[0] rows per sample = C/g0
    int g_last;           // channels in the ragged last l[0] row (0 = none)
    int chunk_elems;      // logical row elements per chunk (multiple of the vector width)
    int nch_full;         // chunks per complete l[0] row
    int nch_last;         // chunks per ragged l[0] row
    int units_per_sample;
    long long total_units;
    int max_row_units;    // units of the longest row of any fused loss
    int delay;            // streaming kernel: phase 2 trails phase 1 by this many units
    // grid-resident kernel: units [0, units_coarse) are cut as above (chunk_elems ...), the rest - rows from row split_row of
    // l[0] in sample split_b on - into finer units
    long long units_coarse;
    int split_b, split_row;
    int f_chunk_elems, f_nch_full, f_nch_last, f_units_per_sample;
    int grid_knobs[4];    // grid-resident kernel: active gather warps, sleep (ns) of the waits for row statistics / ring slots / packets
    // packed short rows (kl_rows_pack_kernel): a unit is pack_rows whole rows, pack_tpr threads each
    int pack_tpr;         // threads per row (8..256, power of two); row length = 32 * pack_tpr
    int pack_rows;        // rows per unit = 512 / pack_tpr
    long long pack_units; // ceil(R / pack_rows)
    // backward re-runs: device scalars d(total)/d(loss_k) folded into the gradient (null = 1), and a
    // device flag that cancels the launch when it reads 0
    const float* grad_out[kMaxLosses];
    const unsigned* run_if;
    // generic kernel decomposition: unit = (b, logical channel, plane chunk)
    int KC;               // chunks per channel plane
    // workspace
    unsigned* ctrl;
    float* cta_part;            // [kMaxLosses + 1][kMaxGrid]
    unsigned long long* pkt;    // [TMA units][kPktWords]
    float* unit_part;           // [generic units][kPartWords]
    unsigned long long* dbg;    // timing builds (-DSD_GRID_TIMING): 16 counters per CTA, in the row_kl region of the workspace
};

inline long long tma_units_upper(long long B, long long C, long long HW, long long g) {
    if (g > C) g = C;
    const long long full = C / g, rest = C % g;
    const long long per_sample = full * ((g * HW + kChunkCapMin - 1) / kChunkCapMin) +
                                 (rest ? (rest * HW + kChunkCapMin - 1) / kChunkCapMin : 0);
    return B * per_sample;
}

struct RowsWorkspace {
    size_t off_ctrl, off_cta, off_rowkl, off_unit, bytes;
};
// g = the smallest group size of the fused losses
inline RowsWorkspace rows_workspace_layout(long long B, long long C, long long HW, long long g) {
    RowsWorkspace w;
    if (g > C) g = C;
    const long long G = (C + g - 1) / g;
    const long long R = B * G;
    const long long KC = (HW + kGenericChunk - 1) / kGenericChunk;
    const long long gen_units = B * C * KC;
    size_t o = 0;
    w.off_ctrl = o;   o += kArenaBytes;
    w.off_cta = o;    o += sizeof(float) * (kMaxLosses + 1) * kMaxGrid;
    w.off_rowkl = o;  o += sizeof(float) * (size_t)R * kMaxLosses;
    o = (o + 127) & ~(size_t)127;
    w.off_unit = o;
    const size_t gen_bytes = sizeof(float) * kPartWords * (size_t)gen_units;
    const size_t tma_bytes = sizeof(unsigned long long) * kPktWords * (size_t)tma_units_upper(B, C, HW, g);
    o += gen_bytes > tma_bytes ? gen_bytes : tma_bytes;
    w.bytes = (o + 255) & ~(size_t)255;
    return w;
}

// ---------------------------------------------------------------- rows of several (student, teacher) pairs, one launch
constexpr int kMaxSegs = 8;            // pairs per grouped launch
struct GroupSeg {
    const void* S;
    const void* T;
    void* dS;
    float* loss;
    float* row_kl;         // [B*G] or null
    int B, C, HW;
    int g, G, G_full, g_last;          // channels per row, rows per sample, complete rows, channels of the ragged last row
    int chunk_elems, nch_full, nch_last, units_per_sample;
    long long unit0;       // first unit of this pair in the launch's work list
    float c2, coef, loss_scale;
};
struct GroupParams {
    int nseg;
    int delay;
    long long total_units;
    const float* grad_out; // device scalar d(total)/d(losses) folded into the gradient, or null (= 1)
    unsigned* ctrl;
    float* cta_part;       // [kMaxSegs][kMaxGrid]
    unsigned long long* pkt;   // [total units][kPktWords]
    GroupSeg seg[kMaxSegs];
};
struct GroupWorkspace {
    size_t off_ctrl, off_cta, off_pkt, bytes;
};
inline GroupWorkspace group_workspace_layout(long long total_units) {
    GroupWorkspace w;
    size_t o = 0;
    w.off_ctrl = o;  o += kArenaBytes;
    w.off_cta = o;   o += sizeof(float) * kMaxSegs * kMaxGrid;
    o = (o + 127) & ~(size_t)127;
    w.off_pkt = o;   o += sizeof(unsigned long long) * kPktWords * (size_t)total_units;
    w.bytes = (o + 255) & ~(size_t)255;
    return w;
}

// ---------------------------------------------------------------- rows, cluster-resident kernel
constexpr int kClusterMaxSize = 8;     // portable cluster size
constexpr int kClusterMaxChunks = 6;   // 4096-element chunks of a CTA's slice (TMEM: 20 columns per chunk and thread)
constexpr int kClusterChunkVecs = 1024;  // 4-element vectors per chunk and tensor
constexpr int kClusterMaxPieces = 8;   // rows of l[0] that may intersect one slice
// a "super-row" is a row of the loss with the larger group; the cluster keeps it resident, slice by slice
struct ClusterGeom {
    int nc;        // CTAs per cluster
    int g_big;     // channels per super-row
    int G_big;     // super-rows per sample
    int hwv;       // 4-element vectors per channel plane
    int slv;       // vectors per slice (multiple of kClusterChunkVecs)
    int rv0;       // vectors per complete row of l[0]
    int total_sr;  // B * G_big
    int pieces_full;  // pieces (rows of l[0] x slices they intersect) of one complete super-row
};

// ---------------------------------------------------------------- rows on bilinearly up-sampled maps (kl_rows_up.cu)
struct UpParams {
    const void* S;
    const void* T;
    void* dS;                // low resolution, like S
    const int32_t* perm;     // [C] or null
    int B, C, Hl, Wl;        // low-resolution maps
    int scale;               // 2, 4 or 8: the maps are up-sampled to (scale*Hl) x (scale*Wl), bilinear, align_corners=False
    int g, G, R;             // channels per row, rows per sample, B*G
    int SR, NS;              // low-res rows per strip, strips per plane
    long long units;         // B*C*NS
    float c2, inv_c2, inv_tau, coef, loss_scale;
    float inv_coef;          // pixel mode: 1 / coef (the KL's sum p (t - s) term is collected with the gradient weights)
    float inv_Wl;
    float* loss;
    float* row_kl;           // [R]
    float* part;             // [units][8] partial records of kernel 1
    float* wpart;            // pixel mode, scale 8: [4][B*C*Hl*Wl] gradients of the four windows of a block
    unsigned* ctrl;
};
struct UpWorkspace {
    size_t off_ctrl, off_rowkl, off_part, bytes;
};
inline int up_strip_rows(int Hl, int Wl) {
    long long sr = (100 * 1024 / 4 / (long long)Wl - 8) / 11;
    if (sr > 16) sr = 16;
    if (sr > Hl) sr = Hl;
    return (int)sr;            // < 1: the plane is too wide for this kernel
}
inline UpWorkspace up_workspace_layout(long long B, long long C, long long Hl, long long Wl, long long g) {
    UpWorkspace w;
    if (g > C) g = C;
    const long long G = (C + g - 1) / g;
    int sr = up_strip_rows((int)Hl, (int)Wl);
    if (sr < 1) sr = 1;
    const long long NS = (Hl + sr - 1) / sr;
    size_t o = 0;
    w.off_ctrl = o;   o += kArenaBytes;
    w.off_rowkl = o;  o += sizeof(float) * (size_t)(B * G);
    o = (o + 127) & ~(size_t)127;
    w.off_part = o;   o += sizeof(float) * 8 * (size_t)(B * C * NS);
    w.bytes = (o + 255) & ~(size_t)255;
    return w;
}

// ---------------------------------------------------------------- cross-entropy on bilinearly up-sampled logits (ce_up.cu)
struct CeParams {
    const void* X;               // logits (B, C, Hl, Wl)
    const long long* label;      // (B, scale*Hl, scale*Wl) int64
    const float* class_weight;   // [C] or null
    const float* pix_weight;     // (B, scale*Hl, scale*Wl) or null
    void* dX;                    // like X
    float* loss;
    float* acc;                  // top-1 accuracy in percent, or null
    int B, C, Hl, Wl, scale;
    long long ignore_index;
    float gscale;                // grad_scale * loss_weight / denominator
    float lscale;                // loss_weight / denominator
    float acc_scale;             // 100 / number of label pixels
    float* part;                 // [2][kMaxGrid] CTA partials: weighted nll, hit count
    float* wpart;                // scale 8: [4][B*C*Hl*Wl] gradients of the four windows of a block
    unsigned* ctrl;
};

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

// ---------------------------------------------------------------- IFVD similarity term (ifvd.cu)
struct IfvdParams {
    const void* S;      // (B, C, HW)
    const void* T;
    const int* cls;     // (B, HW) class of each pixel in [0, C), or C for "no class"
    void* dS;
    float* loss;
    float* sums;        // [2][B][C+1 classes][C+1] class sums of S, of T: a row of C channel sums + the class count
    float* wsum;        // [B][C+1 classes][C+1]    gradient reaching the class sums of S: C channels + V_k
    float* pix;         // [4][B][HW]       per-pixel backward coefficients
    float* part;        // one loss partial per CTA of the per-pixel kernel
    float* spart;       // [splits][2][B][C+1][C+1] (plain) / [wsplits][B][C+1][C+1] (weighted) class sums per pixel range
    int B, C, HW;
    int splits, wsplits;  // pixel ranges (CTAs) per (sample, 32 channels, tensor) of the plain / weighted class-sum launch
    int vec;            // 16-byte loads: HW a multiple of 16 / sizeof(element), S and T 16-byte aligned
    int accumulate;     // dS += gradient instead of dS = gradient
    float gcoef;        // grad_scale * 2 * weight / (B*HW)
};
struct IfvdWorkspace {
    size_t off_sums, off_wsum, off_pix, off_part, off_spart, bytes;
    long long nparts;
    int splits, wsplits;
};
// Pixel ranges per (sample, 32 channels, tensor) of a class-sum launch.  The CTAs own a whole SM (their bins fill its
// shared memory), so the launch runs in waves of 148; a CTA costs its warps' steps of 32 pixels plus a fixed part
// (zeroing and merging 8 x 151 x 32 bins, about three steps' worth).  Pick the count that minimises
// waves x (steps per warp + 3), with at least two steps per warp; a function of the shape only, so that
// sd_ifvd_sim_workspace_bytes and the launch agree.  `tensors`: 2 for the plain sums (S and T), 1 for the weighted.
inline int ifvd_splits(long long B, long long C, long long HW, int tensors) {
    const long long groups = (C + 1 + 31) / 32;
    const long long base = tensors * B * groups;
    const long long steps = (HW + 31) / 32;
    long long best = 1, best_cost = -1;
    for (long long s = 1; s <= 32; ++s) {
        const long long per_cta = (steps + s - 1) / s;
        if (s > 1 && per_cta < 16) break;
        const long long waves = (base * s + 147) / 148;
        const long long cost = waves * ((per_cta + 7) / 8 + 3);
        if (best_cost < 0 || cost < best_cost) {
            best = s;
            best_cost = cost;
        }
    }
    return (int)best;
}
inline IfvdWorkspace ifvd_workspace_layout(long long B, long long C, long long HW, long long pix_threads) {
    IfvdWorkspace w;
    const size_t K1 = (size_t)C + 1;
    size_t o = kArenaBytes;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.off_sums = o;  o = up(o + sizeof(float) * 2 * (size_t)B * K1 * K1);
    w.off_wsum = o;  o = up(o + sizeof(float) * (size_t)B * K1 * K1);
    w.off_pix = o;   o = up(o + sizeof(float) * 4 * (size_t)B * (size_t)HW);   // 16-byte aligned: read as float4
    w.nparts = B * ((HW + pix_threads - 1) / pix_threads);
    w.off_part = o;  o += sizeof(float) * (size_t)w.nparts;
    w.splits = ifvd_splits(B, C, HW, 2);
    w.wsplits = ifvd_splits(B, C, HW, 1);
    const size_t slabs = (size_t)(2 * w.splits > w.wsplits ? 2 * w.splits : w.wsplits);   // both launches use the buffer in turn
    w.off_spart = o; o += sizeof(float) * slabs * (size_t)B * K1 * K1;
    w.bytes = (o + 255) & ~(size_t)255;
    return w;
}

}  // namespace sd
