// extern "C" entry points of libsegdistill_sm100.so (see include/segdistill.h).
// Argument validation, work decomposition, workspace carving, kernel selection.  No allocation,
// no synchronisation, no exceptions.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <utility>

#include "../../include/segdistill.h"
#include "launch.h"
#include "params.h"

namespace {

std::atomic<uint64_t> g_launches{0};
thread_local const char* t_last_kernel = "";

struct DeviceInfo {
    int sms = 0;
    int cc_major = 0;
    bool ok = false;
};
// one process drives one GPU (torchrun, one rank per device); keep a small per-device cache anyway
DeviceInfo& device_info() {
    static DeviceInfo info[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    DeviceInfo& d = info[dev];
    if (!d.ok) {
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
        d.ok = d.sms > 0;
    }
    return d;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int elem_size(int dtype) { return dtype == SD_BF16 ? 2 : 4; }

PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
    }
    return fn;
}

// (B, C, HW) view of an NCHW map; box = (tile pixels, C, 1); out-of-range pixels read as zero
bool encode_pixel_map(CUtensorMap* m, const void* base, int B, int C, int HW, int dtype, int tile_px, bool swizzle128 = false) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t es = (cuuint64_t)elem_size(dtype);
    cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)HW * es, (cuuint64_t)HW * es * (cuuint64_t)C};
    cuuint32_t box[3] = {(cuuint32_t)tile_px, (cuuint32_t)C, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(m, dtype == SD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// (B*C, HW) view of an NCHW map; box = (128 bytes of HW, rows channels), 128-byte swizzle for the tensor cores
bool encode_rows_map(CUtensorMap* m, const void* base, int B, int C, int HW, int dtype, int box_cols, int box_rows,
                     bool atom32 = false) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t es = (cuuint64_t)elem_size(dtype);
    cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)B * (cuuint64_t)C};
    cuuint64_t gstr[1] = {(cuuint64_t)HW * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(m, dtype == SD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                           const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

extern "C" {

int sd_abi_version(void) { return SD_ABI_VERSION; }

const char* sd_strerror(int rc) {
    switch (rc) {
        case SD_OK: return "ok";
        case SD_ERR_NULL: return "segdistill: a required pointer is NULL";
        case SD_ERR_SHAPE: return "segdistill: invalid shape (non-positive or too large)";
        case SD_ERR_DTYPE: return "segdistill: unsupported dtype (SD_F32 or SD_BF16)";
        case SD_ERR_ALIGN: return "segdistill: base pointers must be 16-byte aligned";
        case SD_ERR_WORKSPACE: return "segdistill: workspace too small";
        case SD_ERR_UNSUPPORTED: return "segdistill: the requested algorithm cannot run this layout";
        case SD_ERR_DEVICE: return "segdistill: kernels are built for sm_100a (B200) only";
        case SD_ERR_VALUE: return "segdistill: invalid scalar argument (tau must be > 0, group >= 1)";
        default: break;
    }
    if (rc > 0) return cudaGetErrorString(static_cast<cudaError_t>(rc));
    return "segdistill: unknown error";
}

int sd_device_check(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return SD_ERR_DEVICE;
    }
    return device_info().cc_major == 10 ? SD_OK : SD_ERR_DEVICE;
}

uint64_t sd_launch_count(void) { return g_launches.load(); }
const char* sd_last_kernel(void) { return t_last_kernel; }

// ============================================================================ rows
size_t sd_kl_rows_workspace_bytes(int B, int C, int HW, int group) {
    if (B <= 0 || C <= 0 || HW <= 0 || group <= 0) return 0;
    return sd::rows_workspace_layout(B, C, HW, group).bytes;
}

namespace {

// phase-2 lag of the streaming kernel, in units (SEGDISTILL_STREAM_DELAY overrides; tuning knob)
int stream_delay() {
    static int d = 0;
    if (d == 0) {
        const char* e = std::getenv("SEGDISTILL_STREAM_DELAY");
        d = e ? std::atoi(e) : 2;
        if (d < 1) d = 1;
        if (d > 64) d = 64;
    }
    return d;
}

// grid-resident kernel: chunks per unit (SEGDISTILL_GRID_UNIT overrides; tuning knob) and whether AUTO prefers it to
// the cluster-resident kernel (SEGDISTILL_PREFER_GRID=0: tuning / A-B knob)
int grid_unit_chunks() {
    static int n = 0;
    if (n == 0) {
        const char* e = std::getenv("SEGDISTILL_GRID_UNIT");
        n = e ? std::atoi(e) : 4;
        if (n < 1) n = 1;
        if (n > sd::kGridUnitMaxChunks) n = sd::kGridUnitMaxChunks;
    }
    return n;
}
// active gather warps and back-off sleeps (ns) of the grid-resident kernel: SEGDISTILL_GRID_KNOBS="gather,fin,tma,stat"
const int* grid_knobs() {
    static int k[4] = {0, 0, 0, 0};
    if (k[0] == 0) {
        int v[4] = {3, 96, 32, 64};
        const char* e = std::getenv("SEGDISTILL_GRID_KNOBS");
        if (e) std::sscanf(e, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]);
        if ((v[0] & 255) < 1 || (v[0] & 255) > 3) v[0] = (v[0] & ~255) | 3;    // (bits 8..: park warps, 8 or 16)
        for (int i = 1; i < 4; ++i) v[i] = v[i] < 0 ? 0 : (v[i] > 100000 ? 100000 : v[i]);
        for (int i = 3; i >= 0; --i) k[i] = v[i];
    }
    return k;
}
// SEGDISTILL_GRID_FINE=0: no fine tail (every unit grid_unit_chunks() chunks), 2: a fine tail whatever the size (A-B knob)
int grid_fine_tail() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("SEGDISTILL_GRID_FINE");
        v = e ? std::atoi(e) : 1;
        if (v < 0) v = 0;
    }
    return v;
}
// SEGDISTILL_GROUP_GRID=0: launches over several pairs keep the two-phase kernel (A-B knob, tests of that kernel)
bool group_prefers_grid() {
    const char* e = std::getenv("SEGDISTILL_GROUP_GRID");
    return e ? (std::atoi(e) != 0) : true;
}
int pix_cols_bf16() {
    static int v = 0;
    if (v == 0) {
        const char* e = std::getenv("SEGDISTILL_PIX_COLS");
        v = e && std::atoi(e) == 64 ? 64 : 32;
    }
    return v;
}
// bf16 rows of exactly the register capacity: kl_rows_rm_kernel (row-maximum references, pipelined loads);
// SEGDISTILL_ROWS_RM=0 keeps kl_rows_tma_kernel
bool rows_rm() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("SEGDISTILL_ROWS_RM");
        v = e ? (std::atoi(e) != 0) : 1;
    }
    return v != 0;
}
// PD (no AT term): kl_pixels_warp_kernel (every warp on its own).  SEGDISTILL_PIX_WARP: 0 = kl_pixels_tma_kernel always,
// 1 = bf16 launches, 2 = fp32 launches too
int pix_warp() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("SEGDISTILL_PIX_WARP");
        v = e ? std::atoi(e) : 1;
    }
    return v;
}
bool prefer_grid() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("SEGDISTILL_PREFER_GRID");
        v = e ? (std::atoi(e) != 0) : 1;
    }
    return v != 0;
}

struct RowsCall {
    const void* S;
    const void* T;
    void* dS;
    const int32_t* perm;
    int B, C, HW, dtype;
    int nl;
    int group[sd::kMaxLosses];
    float tau[sd::kMaxLosses], alpha[sd::kMaxLosses];
    float* loss[sd::kMaxLosses];
    float* row_kl[sd::kMaxLosses];
    const float* grad_out[sd::kMaxLosses];
    const unsigned* run_if;
    float grad_scale;
    float mse_weight;
    float* mse_loss;
    void* workspace;
    size_t workspace_bytes;
    int algo;
    void* stream;
};

int rows_dispatch(RowsCall c) {
    if (!c.S || !c.T || !c.dS || !c.workspace) return SD_ERR_NULL;
    if (c.nl < 1 || c.nl > sd::kMaxLosses) return SD_ERR_VALUE;
    for (int k = 0; k < c.nl; ++k) {
        if (!c.loss[k]) return SD_ERR_NULL;
        if (c.group[k] < 1 || !(c.tau[k] > 0.f)) return SD_ERR_VALUE;
        if (c.group[k] > c.C) c.group[k] = c.C;  // one ragged row per sample either way
    }
    if (c.mse_weight != 0.f && !c.mse_loss) return SD_ERR_NULL;
    if (c.dtype != SD_F32 && c.dtype != SD_BF16) return SD_ERR_DTYPE;
    if (c.B <= 0 || c.C <= 0 || c.HW <= 0) return SD_ERR_SHAPE;
    if (c.nl == 2) {
        if (c.perm || c.mse_weight != 0.f) return SD_ERR_UNSUPPORTED;
        if (c.group[1] < c.group[0]) {  // l[0] must have the smaller rows
            std::swap(c.group[0], c.group[1]);
            std::swap(c.tau[0], c.tau[1]);
            std::swap(c.alpha[0], c.alpha[1]);
            std::swap(c.loss[0], c.loss[1]);
            std::swap(c.row_kl[0], c.row_kl[1]);
            std::swap(c.grad_out[0], c.grad_out[1]);
        }
        // every row of l[1] must be a union of whole rows of l[0]
        if (c.group[1] != c.C && c.group[1] % c.group[0] != 0) return SD_ERR_UNSUPPORTED;
    }
    const int B = c.B, C = c.C, HW = c.HW, g0 = c.group[0];
    const long long numel = (long long)B * C * HW;
    const long long row_len = (long long)g0 * HW;
    if ((long long)c.group[c.nl - 1] * HW >= (1ll << 31) || numel >= (1ll << 40)) return SD_ERR_SHAPE;
    const sd::RowsWorkspace wl = sd::rows_workspace_layout(B, C, HW, g0);
    if (c.workspace_bytes < wl.bytes) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;

    const int es = elem_size(c.dtype);
    const int VE = 16 / es;
    sd::RowsParams p;
    std::memset(&p, 0, sizeof(p));
    p.S = c.S;
    p.T = c.T;
    p.dS = c.dS;
    p.perm = c.perm;
    p.B = B;
    p.C = C;
    p.HW = HW;
    p.nl = c.nl;
    char* ws = static_cast<char*>(c.workspace);
    const int G0 = (C + g0 - 1) / g0;
    for (int k = 0; k < c.nl; ++k) {
        sd::RowLoss& l = p.l[k];
        l.g = c.group[k];
        l.m = k == 0 ? 1 : (c.group[k] == C ? G0 : c.group[k] / g0);
        l.G = (C + l.g - 1) / l.g;
        l.R = B * l.G;
        l.c2 = (float)(1.4426950408889634 / (double)c.tau[k]);
        l.inv_tau = (float)(1.0 / (double)c.tau[k]);
        l.coef = (float)((double)c.grad_scale * (double)c.alpha[k] / ((double)l.R * (double)c.tau[k]));
        l.loss_scale = (float)((double)c.alpha[k] / (double)l.R);
        l.loss = c.loss[k];
        l.row_kl = c.row_kl[k];
        p.grad_out[k] = c.grad_out[k];
    }
    p.run_if = c.run_if;
    p.mse_loss = c.mse_weight != 0.f ? c.mse_loss : nullptr;
    p.mse_gcoef = (float)((double)c.grad_scale * 2.0 * (double)c.mse_weight / (double)numel);
    p.mse_scale = (float)((double)c.mse_weight / (double)numel);
    p.G_full = C / g0;
    p.g_last = C % g0;
    p.KC = (HW + sd::kGenericChunk - 1) / sd::kGenericChunk;
    p.ctrl = reinterpret_cast<unsigned*>(ws + wl.off_ctrl);
    p.cta_part = reinterpret_cast<float*>(ws + wl.off_cta);
    p.pkt = reinterpret_cast<unsigned long long*>(ws + wl.off_unit);
    p.unit_part = reinterpret_cast<float*>(ws + wl.off_unit);
    p.dbg = reinterpret_cast<unsigned long long*>(ws + wl.off_rowkl);

    // TMA paths: rows must start on 16-byte boundaries.  Rows of one loss that fit one CTA's registers
    // (16384 elements) take the register-resident single pass, everything else the streaming kernel.
    const bool layout_ok = ((long long)HW * es) % 16 == 0 && aligned16(c.S) && aligned16(c.T) && aligned16(c.dS);
    const long long longest0 = p.G_full > 0 ? row_len : (long long)p.g_last * HW;
    const bool fits_regs = c.nl == 1 && longest0 <= sd::kl_rows_tma_chunk_capacity();
    const long long gen_units = (long long)B * C * p.KC;
    if (gen_units >= (1ll << 31)) return SD_ERR_SHAPE;

    // cluster-resident single pass: the longest row of the larger-group loss, cut into nc slices, must fit
    // nc CTAs' shared memory, and a slice may intersect only a few rows of the smaller-group loss
    sd::ClusterGeom cg;
    std::memset(&cg, 0, sizeof(cg));
    bool cluster_ok = false;
    // (HW % 128 == 0: a warp's 32 consecutive vectors never straddle two rows of l[0] or the end of a slice)
    if (layout_ok && HW % 128 == 0 && !c.perm && c.mse_weight == 0.f) {
        const int g_big = c.group[c.nl - 1];
        const long long hwv = (long long)HW / 4;   // the cluster kernel works on 4-element vectors
        const long long lv = (long long)g_big * hwv;
        const long long rv0 = (long long)g0 * hwv;
        const long long slice_cap = (long long)sd::kClusterMaxChunks * sd::kClusterChunkVecs;
        // The smallest power-of-two cluster whose slices fit.  Rows that fit ONE CTA are not for this kernel (its summaries
        // travel by st.async into the cluster's shared memory, which a one-CTA "cluster" does not have: compute-sanitizer
        // rejects it) - the grid-resident kernel serves them.
        for (int nc = 1; nc <= sd::kClusterMaxSize && !cluster_ok; nc *= 2) {
            long long slv = (lv + nc - 1) / nc;
            slv = (slv + sd::kClusterChunkVecs - 1) / sd::kClusterChunkVecs * sd::kClusterChunkVecs;
            if (slv > slice_cap) continue;
            if (nc == 1) break;
            if (rv0 < 512) break;
            if (slv / rv0 + 2 > sd::kClusterMaxPieces) continue;  // more, shorter slices
            cg.nc = nc;
            cg.g_big = g_big;
            cg.G_big = (C + g_big - 1) / g_big;
            cg.hwv = (int)hwv;
            cg.slv = (int)slv;
            cg.rv0 = (int)rv0;
            cg.total_sr = B * cg.G_big;
            cg.pieces_full = 0;
            for (int cta = 0; cta < nc; ++cta) {
                const long long cv0 = std::min<long long>(lv, cta * slv), cv1 = std::min<long long>(lv, cv0 + slv);
                if (cv1 > cv0) cg.pieces_full += (int)((cv1 - 1) / rv0 - cv0 / rv0 + 1);
            }
            cluster_ok = sd::launch_kl_rows_cluster(p, cg, c.dtype == SD_BF16, dev.sms, nullptr, true) == cudaSuccess;
        }
    }

    // grid-resident single pass (kl_rows_grid.cu): same layouts as the cluster kernel.  Units of up to grid_uc chunks of
    // 4096 elements for as many whole rounds of the grid as the work list has, units of one chunk (or the smallest
    // size that keeps a row within 64 units) for the rest - so that the last round is shared by every SM.  A row of
    // any fused loss may span at most 64 units and no more than one round of the grid.
    struct GridGeo {
        long long nch_full, nch_last, ups, max_row_units;
    };
    const auto grid_geo = [&](int chunks) {
        GridGeo g;
        const long long cap = (long long)chunks * sd::kGridChunkElems;
        g.nch_full = p.G_full > 0 ? (row_len + cap - 1) / cap : 0;
        g.nch_last = p.g_last ? ((long long)p.g_last * HW + cap - 1) / cap : 0;
        g.ups = p.G_full * g.nch_full + g.nch_last;
        g.max_row_units = std::max(g.nch_full, g.nch_last);
        for (int k = 1; k < c.nl; ++k) {
            const long long m = p.l[k].m;
            long long n = m * g.nch_full;
            if (m > p.G_full) n = (long long)p.G_full * g.nch_full + g.nch_last;
            g.max_row_units = std::max(g.max_row_units, n);
        }
        return g;
    };
    bool grid_ok = false;
    const int grid_uc = grid_unit_chunks();
    int grid_fc = grid_uc;                       // chunks per unit of the fine region (== grid_uc: no fine region)
    GridGeo gc = {0, 0, 0, 0}, gf = {0, 0, 0, 0};
    long long grid_units_coarse = 0, grid_total = 0, grid_max_row_units = 0;
    int grid_split_b = B, grid_split_row = 0;
    if (layout_ok && HW % 128 == 0 && !c.perm && c.mse_weight == 0.f) {
        gc = grid_geo(grid_uc);
        const long long all_coarse = (long long)B * gc.ups;
        const long long sms = std::min<long long>(dev.sms, sd::kMaxGrid);
        for (int fcand = 1; fcand < grid_uc && grid_fine_tail(); fcand *= 2) {
            gf = grid_geo(fcand);
            if (gf.max_row_units <= sd::kGridMaxRowUnits && gf.max_row_units <= sms) {
                grid_fc = fcand;
                break;
            }
        }
        grid_units_coarse = all_coarse;
        grid_total = all_coarse;
        grid_max_row_units = gc.max_row_units;
        // (measured at 16 rounds - 16x150x128x128 - the fine tail costs ~1 us: its units wait for 40 row-mates each; it
        //  pays where the last round is a large part of the work list.  SEGDISTILL_GRID_FINE=2 forces it)
        const bool few_rounds = all_coarse < 8 * sms || grid_fine_tail() > 1;
        if (grid_fc < grid_uc && all_coarse % sms != 0 && few_rounds) {
            // coarse units for floor(all / sms) whole rounds, cut back to a row boundary of the larger-group loss
            const long long budget = all_coarse / sms * sms;
            const int m = c.nl == 2 ? p.l[1].m : 1;
            int sb = (int)(budget / gc.ups);
            const long long rem = budget - (long long)sb * gc.ups;
            int row = gc.nch_full > 0 ? (int)std::min<long long>(rem / gc.nch_full, p.G_full) : 0;
            row = row / m * m;
            grid_split_b = sb;
            grid_split_row = row;
            grid_units_coarse = (long long)sb * gc.ups + (long long)row * gc.nch_full;
            const long long fine_first = (long long)row * gf.nch_full;
            grid_total = grid_units_coarse + (gf.ups - fine_first) + (long long)(B - sb - 1) * gf.ups;
            grid_max_row_units = grid_units_coarse > 0 ? std::max(gc.max_row_units, gf.max_row_units) : gf.max_row_units;
        }
        const long long grid = std::min<long long>(sms, grid_total);
        // (rows of one unit gain nothing from the exchange: the register-resident kernels serve them)
        if (grid_max_row_units <= sd::kGridMaxRowUnits && grid_max_row_units <= grid && row_len >= 2048 &&
            grid_total < (1ll << 31))
            grid_ok = sd::launch_kl_rows_grid(p, c.dtype == SD_BF16, dev.sms, nullptr, true) == cudaSuccess;
    }

    // packed short rows: whole rows of 32 * TPR elements (TPR a power of two), contiguous in memory
    bool pack_ok = false;
    if (layout_ok && fits_regs && !c.perm && C % g0 == 0 && row_len % 32 == 0) {
        const long long tpr = row_len / 32;
        const long long nvec_row = row_len / VE;
        if (tpr >= 8 && tpr <= 256 && (tpr & (tpr - 1)) == 0 && nvec_row <= 1024) {
            pack_ok = true;
            p.pack_tpr = (int)tpr;
            p.pack_rows = 512 / (int)tpr;
            p.pack_units = ((long long)B * (C / g0) + p.pack_rows - 1) / p.pack_rows;
        }
    }

    enum { kGeneric, kRegs, kStream, kCluster, kPack, kGrid } path;
    // (measured on B200: the grid-resident kernel beats the cluster-resident one - 148 SMs instead of 120 -, and that one
    // the streaming kernel, for one loss and for two)
    const auto long_rows = [&]() { return grid_ok && prefer_grid() ? kGrid : (cluster_ok ? kCluster : (grid_ok ? kGrid : kStream)); };
    switch (c.algo) {
        case SD_ALGO_AUTO:
            path = !layout_ok ? kGeneric : (pack_ok ? kPack : fits_regs ? kRegs : long_rows());
            break;
        case SD_ALGO_TMA:
            if (!layout_ok) return SD_ERR_UNSUPPORTED;
            path = pack_ok ? kPack : fits_regs ? kRegs : long_rows();
            break;
        case SD_ALGO_ROWS1:
            if (!layout_ok || !fits_regs) return SD_ERR_UNSUPPORTED;
            path = kRegs;
            break;
        case SD_ALGO_STREAM:
            if (!layout_ok) return SD_ERR_UNSUPPORTED;
            path = kStream;
            break;
        case SD_ALGO_CLUSTER:
            if (!cluster_ok) return SD_ERR_UNSUPPORTED;
            path = kCluster;
            break;
        case SD_ALGO_GRID:
            if (!grid_ok) return SD_ERR_UNSUPPORTED;
            path = kGrid;
            break;
        case SD_ALGO_GENERIC: path = kGeneric; break;
        default: return SD_ERR_VALUE;
    }
    if (path == kGeneric && (c.nl > 1 || c.run_if || c.grad_out[0])) return SD_ERR_UNSUPPORTED;

    const int cap = path == kStream ? sd::kl_rows_stream_chunk_capacity() : sd::kl_rows_tma_chunk_capacity();
    if (path == kGrid) {
        // units of whole chunks; the last unit of a row may be shorter
        p.nch_full = p.G_full > 0 ? (int)gc.nch_full : 1;
        p.chunk_elems = grid_uc * sd::kGridChunkElems;
        p.nch_last = (int)gc.nch_last;
    } else {
        p.nch_full = p.G_full > 0 ? (int)((row_len + cap - 1) / cap) : 1;
        long long ce = p.G_full > 0 ? (row_len + p.nch_full - 1) / p.nch_full : (long long)p.g_last * HW;
        ce = (ce + VE - 1) / VE * VE;
        if (ce > cap) ce = cap;
        p.chunk_elems = (int)ce;
        p.nch_last = p.g_last ? (int)(((long long)p.g_last * HW + ce - 1) / ce) : 0;
    }
    p.units_per_sample = p.G_full * p.nch_full + p.nch_last;
    p.total_units = (long long)B * p.units_per_sample;
    // the longest row of any fused loss, in units
    long long max_row_units = p.nch_full > p.nch_last ? p.nch_full : p.nch_last;
    for (int k = 1; k < c.nl; ++k) {
        const long long m = p.l[k].m;
        long long n = m * p.nch_full;
        if (m > p.G_full) n = (long long)p.G_full * p.nch_full + p.nch_last;
        if (n > max_row_units) max_row_units = n;
    }
    p.units_coarse = p.total_units;
    p.split_b = B;
    if (path == kGrid) {
        p.total_units = grid_total;
        p.units_coarse = grid_units_coarse;
        p.split_b = grid_split_b;
        p.split_row = grid_split_row;
        p.f_chunk_elems = grid_fc * sd::kGridChunkElems;
        p.f_nch_full = p.G_full > 0 ? (int)gf.nch_full : 1;
        p.f_nch_last = (int)gf.nch_last;
        p.f_units_per_sample = (int)gf.ups;
        max_row_units = grid_max_row_units;
    }
    if (max_row_units >= (1ll << 30)) return SD_ERR_SHAPE;
    p.max_row_units = (int)max_row_units;
    p.delay = stream_delay();
    for (int i = 0; i < 4; ++i) p.grid_knobs[i] = grid_knobs()[i];

    cudaStream_t st = static_cast<cudaStream_t>(c.stream);
    cudaError_t e;
    if (path == kPack) {
        int grid = (int)(p.pack_units < dev.sms ? p.pack_units : dev.sms);
        e = sd::launch_kl_rows_pack(p, c.dtype == SD_BF16, grid, st);
        g_launches += 1;
        t_last_kernel = "kl_rows_pack_kernel";
    } else if (path == kRegs) {
        int grid = (int)(p.total_units < dev.sms ? p.total_units : dev.sms);
        const bool mse = p.mse_gcoef != 0.f || p.mse_loss != nullptr;
        const bool whole_rows = p.g_last == 0 && p.nch_full == 1 && row_len == sd::kl_rows_tma_chunk_capacity() &&
                                p.total_units < (1ll << 31);
        if (c.dtype == SD_BF16 && !mse && whole_rows && rows_rm()) {
            e = sd::launch_kl_rows_rm(p, grid, st);
            t_last_kernel = "kl_rows_rm_kernel";
        } else {
            e = sd::launch_kl_rows_tma(p, c.dtype == SD_BF16, grid, st);
            t_last_kernel = "kl_rows_tma_kernel";
        }
        g_launches += 1;
    } else if (path == kCluster) {
        e = sd::launch_kl_rows_cluster(p, cg, c.dtype == SD_BF16, dev.sms, st, false);
        g_launches += 1;
        t_last_kernel = c.nl == 2 ? "kl_rows_cluster_kernel(2 losses)" : "kl_rows_cluster_kernel";
    } else if (path == kGrid) {
        e = sd::launch_kl_rows_grid(p, c.dtype == SD_BF16, dev.sms, st, false);
        g_launches += 1;
        t_last_kernel = c.nl == 2 ? "kl_rows_grid_kernel(2 losses)" : "kl_rows_grid_kernel";
    } else if (path == kStream) {
        e = sd::launch_kl_rows_stream(p, c.dtype == SD_BF16, dev.sms, st);
        g_launches += 1;
        t_last_kernel = c.nl == 2 ? "kl_rows_stream_kernel(2 losses)" : "kl_rows_stream_kernel";
    } else {
        if (!p.l[0].row_kl) p.l[0].row_kl = reinterpret_cast<float*>(ws + wl.off_rowkl);
        e = sd::launch_kl_rows_generic(p, c.dtype == SD_BF16, st);
        g_launches += 3;
        t_last_kernel = "kl_rows_generic";
    }
    return e == cudaSuccess ? SD_OK : (int)e;
}

}  // namespace

int sd_kl_rows_fwd_bwd(const void* S, const void* T, void* dS, float* row_kl, float* loss, const int32_t* chan_perm,
                       int B, int C, int HW, int group, int dtype, float tau, float alpha, float grad_scale,
                       float mse_weight, float* mse_loss, void* workspace, size_t workspace_bytes, int algo,
                       void* stream) {
    RowsCall c;
    std::memset(&c, 0, sizeof(c));
    c.S = S; c.T = T; c.dS = dS; c.perm = chan_perm;
    c.B = B; c.C = C; c.HW = HW; c.dtype = dtype;
    c.nl = 1;
    c.group[0] = group; c.tau[0] = tau; c.alpha[0] = alpha;
    c.loss[0] = loss; c.row_kl[0] = row_kl;
    c.grad_scale = grad_scale;
    c.mse_weight = mse_weight; c.mse_loss = mse_loss;
    c.workspace = workspace; c.workspace_bytes = workspace_bytes;
    c.algo = algo; c.stream = stream;
    return rows_dispatch(c);
}

int sd_kl_rows_multi_fwd_bwd(const void* S, const void* T, void* dS, int n_losses, const int* groups,
                             const float* taus, const float* alphas, float* const* losses, float* const* row_kls,
                             const float* const* grad_outputs, const unsigned* run_if,
                             int B, int C, int HW, int dtype, float grad_scale,
                             void* workspace, size_t workspace_bytes, int algo, void* stream) {
    if (!groups || !taus || !alphas || !losses) return SD_ERR_NULL;
    if (n_losses < 1 || n_losses > sd::kMaxLosses) return SD_ERR_VALUE;
    RowsCall c;
    std::memset(&c, 0, sizeof(c));
    c.S = S; c.T = T; c.dS = dS;
    c.B = B; c.C = C; c.HW = HW; c.dtype = dtype;
    c.nl = n_losses;
    for (int k = 0; k < n_losses; ++k) {
        c.group[k] = groups[k]; c.tau[k] = taus[k]; c.alpha[k] = alphas[k];
        c.loss[k] = losses[k];
        c.row_kl[k] = row_kls ? row_kls[k] : nullptr;
        c.grad_out[k] = grad_outputs ? grad_outputs[k] : nullptr;
    }
    c.run_if = run_if;
    c.grad_scale = grad_scale;
    c.workspace = workspace; c.workspace_bytes = workspace_bytes;
    if (algo == SD_ALGO_GENERIC) return SD_ERR_UNSUPPORTED;
    c.algo = algo == SD_ALGO_AUTO ? SD_ALGO_TMA : algo; c.stream = stream;
    return rows_dispatch(c);
}

// ============================================================================ rows of several pairs, one launch
namespace {

// work decomposition of one pair for the grouped kernel; false: the layout cannot take the TMA path
bool group_seg_geometry(sd::GroupSeg& s, int B, int C, int HW, int group, int es) {
    if (group > C) group = C;
    const int VE = 16 / es;
    const int cap = sd::kl_rows_group_chunk_capacity();
    s.B = B; s.C = C; s.HW = HW;
    s.g = group;
    s.G = (C + group - 1) / group;
    s.G_full = C / group;
    s.g_last = C % group;
    const long long row_len = (long long)group * HW;
    if (row_len >= (1ll << 31) || ((long long)HW * es) % 16 != 0) return false;
    s.nch_full = s.G_full > 0 ? (int)((row_len + cap - 1) / cap) : 1;
    long long ce = s.G_full > 0 ? (row_len + s.nch_full - 1) / s.nch_full : (long long)s.g_last * HW;
    ce = (ce + VE - 1) / VE * VE;
    if (ce > cap) ce = cap;
    s.chunk_elems = (int)ce;
    s.nch_last = s.g_last ? (int)(((long long)s.g_last * HW + ce - 1) / ce) : 0;
    s.units_per_sample = s.G_full * s.nch_full + s.nch_last;
    return true;
}

}  // namespace

size_t sd_kl_rows_group_workspace_bytes(int n_pairs, const int* B, const int* C, const int* HW, const int* groups, int dtype) {
    if (n_pairs < 1 || n_pairs > sd::kMaxSegs || !B || !C || !HW || !groups) return 0;
    long long total = 0;
    for (int k = 0; k < n_pairs; ++k) {
        if (B[k] <= 0 || C[k] <= 0 || HW[k] <= 0 || groups[k] <= 0) return 0;
        sd::GroupSeg s;
        if (!group_seg_geometry(s, B[k], C[k], HW[k], groups[k], elem_size(dtype))) return 0;
        total += (long long)B[k] * s.units_per_sample;
    }
    return sd::group_workspace_layout(total).bytes;
}

int sd_kl_rows_group_fwd_bwd(int n_pairs, const void* const* S, const void* const* T, void* const* dS,
                             float* const* losses, const int* B, const int* C, const int* HW, const int* groups,
                             const float* taus, const float* alphas, int dtype, float grad_scale,
                             const float* grad_output, void* workspace, size_t workspace_bytes, void* stream) {
    if (!S || !T || !dS || !losses || !B || !C || !HW || !groups || !taus || !alphas || !workspace) return SD_ERR_NULL;
    if (n_pairs < 1 || n_pairs > sd::kMaxSegs) return SD_ERR_VALUE;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    sd::GroupParams gp;
    std::memset(&gp, 0, sizeof(gp));
    gp.nseg = n_pairs;
    long long total = 0;
    int max_row_units = 1;
    const int es = elem_size(dtype);
    for (int k = 0; k < n_pairs; ++k) {
        if (!S[k] || !T[k] || !dS[k] || !losses[k]) return SD_ERR_NULL;
        if (B[k] <= 0 || C[k] <= 0 || HW[k] <= 0) return SD_ERR_SHAPE;
        if (groups[k] < 1 || !(taus[k] > 0.f)) return SD_ERR_VALUE;
        if ((long long)B[k] * C[k] * HW[k] >= (1ll << 40)) return SD_ERR_SHAPE;
        sd::GroupSeg& s = gp.seg[k];
        if (!aligned16(S[k]) || !aligned16(T[k]) || !aligned16(dS[k])) return SD_ERR_UNSUPPORTED;
        if (!group_seg_geometry(s, B[k], C[k], HW[k], groups[k], es)) return SD_ERR_UNSUPPORTED;
        s.S = S[k]; s.T = T[k]; s.dS = dS[k];
        s.loss = losses[k];
        s.row_kl = nullptr;
        const double R = (double)B[k] * s.G;
        s.c2 = (float)(1.4426950408889634 / (double)taus[k]);
        s.coef = (float)((double)grad_scale * (double)alphas[k] / (R * (double)taus[k]));
        s.loss_scale = (float)((double)alphas[k] / R);
        s.unit0 = total;
        total += (long long)B[k] * s.units_per_sample;
        max_row_units = std::max(max_row_units, std::max(s.nch_full, s.nch_last));
    }
    if (total >= (1ll << 31)) return SD_ERR_SHAPE;
    const sd::GroupWorkspace wl = sd::group_workspace_layout(total);
    if (workspace_bytes < wl.bytes) return SD_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    gp.total_units = total;
    gp.delay = stream_delay();
    gp.grad_out = grad_output;
    gp.ctrl = reinterpret_cast<unsigned*>(ws + wl.off_ctrl);
    gp.cta_part = reinterpret_cast<float*>(ws + wl.off_cta);
    gp.pkt = reinterpret_cast<unsigned long long*>(ws + wl.off_pkt);

    // The grid-resident kernel (kl_rows_grid.cu: every pair's units parked in tensor memory, one pass over HBM) takes
    // the work list when every map has HW % 128 == 0 and no row spans more than 64 units (of up to 4 chunks of 4096
    // elements) or more than the grid holds; it needs no more packets than the two-phase kernel's work list has.
    if (group_prefers_grid()) {
        sd::GroupParams gg = gp;
        const int uc = grid_unit_chunks();
        const long long cap = (long long)uc * sd::kGridChunkElems;
        long long gtotal = 0, gmax = 1;
        bool ok = true;
        for (int k = 0; k < n_pairs && ok; ++k) {
            sd::GroupSeg& s = gg.seg[k];
            const long long row_len = (long long)s.g * s.HW;
            ok = s.HW % 128 == 0;
            s.chunk_elems = (int)cap;
            s.nch_full = s.G_full > 0 ? (int)((row_len + cap - 1) / cap) : 1;
            s.nch_last = s.g_last ? (int)(((long long)s.g_last * s.HW + cap - 1) / cap) : 0;
            s.units_per_sample = s.G_full * s.nch_full + s.nch_last;
            s.unit0 = gtotal;
            gtotal += (long long)s.B * s.units_per_sample;
            gmax = std::max<long long>(gmax, std::max(s.G_full > 0 ? s.nch_full : 0, s.nch_last));
        }
        const long long grid = std::min<long long>(std::min<long long>(dev.sms, gtotal), sd::kMaxGrid);
        if (ok && gtotal <= total && gmax <= sd::kGridMaxRowUnits && gmax <= grid) {
            gg.total_units = gtotal;
            cudaError_t e = sd::launch_kl_rows_grid_group(gg, (int)gmax, grid_knobs(), dtype == SD_BF16, dev.sms,
                                                          static_cast<cudaStream_t>(stream), false);
            if (e != cudaErrorLaunchOutOfResources) {
                g_launches += 1;
                t_last_kernel = "kl_rows_grid_kernel(group)";
                return e == cudaSuccess ? SD_OK : (int)e;
            }
        }
    }
    cudaError_t e = sd::launch_kl_rows_group(gp, max_row_units, dtype == SD_BF16, dev.sms, static_cast<cudaStream_t>(stream));
    g_launches += 1;
    t_last_kernel = "kl_rows_group_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

// ============================================================================ pixels
size_t sd_kl_pixels_workspace_bytes(int B, int C, int HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return 0;
    return sd::pix_workspace_layout(B, C, HW).bytes;
}

int sd_kl_pixels_fwd_bwd(const void* S, const void* T, void* dS, float* row_kl, float* loss, int B, int C, int HW,
                         int dtype, float tau, float alpha, float grad_scale, float at_weight, float* at_loss,
                         void* workspace, size_t workspace_bytes, int algo, void* stream) {
    if (!S || !T || !dS || !loss || !workspace) return SD_ERR_NULL;
    if (at_weight != 0.f && !at_loss) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (B <= 0 || C <= 0 || HW <= 0) return SD_ERR_SHAPE;
    if (!(tau > 0.f)) return SD_ERR_VALUE;
    const long long R = (long long)B * HW;
    if (R >= (1ll << 31) || (long long)B * C * HW >= (1ll << 40)) return SD_ERR_SHAPE;
    const sd::PixWorkspace wl = sd::pix_workspace_layout(B, C, HW);
    if (workspace_bytes < wl.bytes) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;

    const int es = elem_size(dtype);
    const bool bf16 = dtype == SD_BF16;
    sd::PixParams p;
    std::memset(&p, 0, sizeof(p));
    char* ws = static_cast<char*>(workspace);
    p.S = S;
    p.T = T;
    p.dS = dS;
    p.row_kl = row_kl ? row_kl : reinterpret_cast<float*>(ws + wl.off_rowkl);
    p.loss = loss;
    p.at_loss = at_weight != 0.f ? at_loss : nullptr;
    p.B = B;
    p.C = C;
    p.HW = HW;
    p.c2 = (float)(1.4426950408889634 / (double)tau);
    p.inv_tau = (float)(1.0 / (double)tau);
    p.coef = (float)((double)grad_scale * (double)alpha / ((double)R * (double)tau));
    p.loss_scale = (float)((double)alpha / (double)R);
    p.at_gcoef = (float)((double)grad_scale * 2.0 * (double)at_weight / ((double)C * (double)R));
    p.at_scale = (float)((double)at_weight / (double)R);
    p.inv_C = (float)(1.0 / (double)C);
    p.ctrl = reinterpret_cast<unsigned*>(ws + wl.off_ctrl);
    p.cta_part = reinterpret_cast<float*>(ws + wl.off_cta);
    p.nparts = (int)wl.nparts;

    // bf16: tiles of 32 thread columns (64 pixels), 256 threads, two CTAs per SM (SEGDISTILL_PIX_COLS=64: the one-CTA layout)
    const int cols = bf16 ? pix_cols_bf16() : 64;
    const int tile_px = sd::kl_pixels_tile_pixels(bf16, cols);
    p.tiles_per_sample = (HW + tile_px - 1) / tile_px;
    p.total_tiles = (long long)B * p.tiles_per_sample;
    p.stage_bytes = (unsigned)C * (unsigned)cols * 4u;
    // ring depth: as many stages as fit next to the reduction scratch (at most 4), in the CTA's share of the SM
    const size_t smem_cap = cols == 32 ? 112u * 1024u : 227u * 1024u;
    int nstages = 0;
    for (int n = 4; n >= 1; --n) {
        if (sd::pix_tma_smem_bytes(C, bf16 ? 2 : 1, n, cols) <= smem_cap) {
            nstages = n;
            break;
        }
    }
    p.nstages = nstages;
    const bool layout_ok = ((long long)HW * es) % 16 == 0 && aligned16(S) && aligned16(T) && aligned16(dS);
    const bool tma_ok = layout_ok && nstages >= 1 && C <= sd::kl_pixels_tma_max_channels(bf16) && C <= 256 &&
                        tensor_map_encoder() != nullptr;

    // bf16 PD: the kernel with one warp per pixel column takes up to 256 channels
    const bool at_term = p.at_gcoef != 0.f || p.at_loss != nullptr;
    const int tpx_w = sd::kl_pixels_warp_tile_pixels(bf16);
    const bool warp_can = layout_ok && !at_term && C <= 256 && sd::pix_warp_stages(C) >= 2 && tensor_map_encoder() != nullptr &&
                          (long long)B * ((HW + tpx_w - 1) / tpx_w) < (1ll << 31);
    // (fp32 launches are HBM-bound on kl_pixels_tma_kernel and 3 % slower here - 32-pixel tiles, twice as many: bf16 only)
    const bool warp_ok = warp_can && (algo == SD_ALGO_WARP || (pix_warp() >= (bf16 ? 1 : 2) &&
                                                               (algo == SD_ALGO_AUTO || algo == SD_ALGO_TMA)));

    bool use_tma;
    if (algo == SD_ALGO_TMA) {
        if (!tma_ok && !warp_ok) return SD_ERR_UNSUPPORTED;
        use_tma = true;
    } else if (algo == SD_ALGO_GENERIC) {
        use_tma = false;
    } else if (algo == SD_ALGO_AUTO) {
        use_tma = tma_ok;
    } else if (algo == SD_ALGO_WARP) {
        if (!warp_can) return SD_ERR_UNSUPPORTED;
        use_tma = false;
    } else {
        return SD_ERR_VALUE;
    }

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (warp_ok) {
        alignas(64) CUtensorMap mS, mT, mD;
        const int tpx = tpx_w;
        if (!encode_pixel_map(&mS, S, B, C, HW, dtype, tpx, true) || !encode_pixel_map(&mT, T, B, C, HW, dtype, tpx, true) ||
            !encode_pixel_map(&mD, dS, B, C, HW, dtype, tpx, true))
            return (int)cudaErrorInvalidValue;
        p.tiles_per_sample = (HW + tpx - 1) / tpx;
        p.total_tiles = (long long)B * p.tiles_per_sample;
        p.nstages = sd::pix_warp_stages(C);
        int grid = (int)(p.total_tiles < dev.sms ? p.total_tiles : dev.sms);
        e = sd::launch_kl_pixels_warp(&mS, &mT, &mD, p, bf16, grid, st);
        g_launches += 1;
        t_last_kernel = "kl_pixels_warp_kernel";
    } else if (use_tma) {
        alignas(64) CUtensorMap mS, mT;
        if (!encode_pixel_map(&mS, S, B, C, HW, dtype, tile_px) || !encode_pixel_map(&mT, T, B, C, HW, dtype, tile_px))
            return (int)cudaErrorInvalidValue;
        const long long ctas = (long long)dev.sms * (cols == 32 ? 2 : 1);
        int grid = (int)(p.total_tiles < ctas ? p.total_tiles : ctas);
        e = sd::launch_kl_pixels_tma(&mS, &mT, p, bf16, cols, grid, sd::pix_tma_smem_bytes(C, bf16 ? 2 : 1, nstages, cols), st);
        g_launches += 1;
        t_last_kernel = "kl_pixels_tma_kernel";
    } else {
        e = sd::launch_kl_pixels_generic(p, bf16, st);
        g_launches += 2;
        t_last_kernel = "kl_pixels_generic";
    }
    return e == cudaSuccess ? SD_OK : (int)e;
}

// ============================================================================ rows on up-sampled maps
size_t sd_kl_rows_up_workspace_bytes(int B, int C, int Hl, int Wl, int group) {
    if (B <= 0 || C <= 0 || Hl <= 0 || Wl <= 0 || group <= 0) return 0;
    return sd::up_workspace_layout(B, C, Hl, Wl, group).bytes;
}

int sd_kl_rows_up_fwd_bwd(const void* S, const void* T, void* dS, float* row_kl, float* loss, const int32_t* chan_perm,
                          int B, int C, int Hl, int Wl, int scale, int group, int dtype, float tau, float alpha,
                          float grad_scale, void* workspace, size_t workspace_bytes, void* stream) {
    if (!S || !T || !dS || !loss || !workspace) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (B <= 0 || C <= 0 || Hl <= 0 || Wl <= 0) return SD_ERR_SHAPE;
    if (group < 1 || !(tau > 0.f)) return SD_ERR_VALUE;
    if (scale != 2 && scale != 4 && scale != 8) return SD_ERR_UNSUPPORTED;
    if (group > C) group = C;
    const long long numel = (long long)B * C * Hl * Wl;
    if (numel >= (1ll << 40) || (long long)group * Hl * Wl * scale * scale >= (1ll << 40)) return SD_ERR_SHAPE;
    const int SR = sd::up_strip_rows(Hl, Wl);
    if (SR < 1) return SD_ERR_UNSUPPORTED;
    const sd::UpWorkspace wl = sd::up_workspace_layout(B, C, Hl, Wl, group);
    if (workspace_bytes < wl.bytes) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    sd::UpParams p;
    std::memset(&p, 0, sizeof(p));
    p.S = S; p.T = T; p.dS = dS; p.perm = chan_perm;
    p.B = B; p.C = C; p.Hl = Hl; p.Wl = Wl; p.scale = scale;
    p.g = group;
    p.G = (C + group - 1) / group;
    p.R = B * p.G;
    p.SR = SR;
    p.NS = (Hl + SR - 1) / SR;
    p.units = (long long)B * C * p.NS;
    p.c2 = (float)(1.4426950408889634 / (double)tau);
    p.inv_c2 = (float)((double)tau / 1.4426950408889634);
    p.inv_Wl = 1.0f / (float)Wl;
    p.inv_tau = (float)(1.0 / (double)tau);
    p.coef = (float)((double)grad_scale * (double)alpha / ((double)p.R * (double)tau));
    p.loss_scale = (float)((double)alpha / (double)p.R);
    p.loss = loss;
    char* ws = static_cast<char*>(workspace);
    p.row_kl = row_kl ? row_kl : reinterpret_cast<float*>(ws + wl.off_rowkl);
    p.part = reinterpret_cast<float*>(ws + wl.off_part);
    p.ctrl = reinterpret_cast<unsigned*>(ws + wl.off_ctrl);
    cudaError_t e = sd::launch_kl_rows_up(p, dtype == SD_BF16, dev.sms, static_cast<cudaStream_t>(stream));
    g_launches += 2;
    t_last_kernel = "kl_rows_up_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

size_t sd_kl_pixels_up_workspace_bytes(int B, int C, int Hl, int Wl, int scale) {
    size_t n = sd::kArenaBytes + sizeof(float) * sd::kMaxGrid;
    if (scale == 8 && B > 0 && C > 0 && Hl > 0 && Wl > 0)       // the four window planes of the gradient, fp32
        n += sizeof(float) * 4 * (size_t)B * C * Hl * Wl;
    return (n + 255) & ~(size_t)255;
}

int sd_kl_pixels_up_fwd_bwd(const void* S, const void* T, void* dS, float* loss, int B, int C, int Hl, int Wl, int scale,
                            int dtype, float tau, float alpha, float grad_scale, void* workspace, size_t workspace_bytes,
                            void* stream) {
    if (!S || !T || !dS || !loss || !workspace) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (B <= 0 || C <= 0 || Hl <= 0 || Wl <= 0) return SD_ERR_SHAPE;
    if (!(tau > 0.f) || alpha == 0.f || grad_scale == 0.f) return SD_ERR_VALUE;   // (the kernel divides by their product)
    if (scale != 2 && scale != 4 && scale != 8) return SD_ERR_UNSUPPORTED;
    if ((long long)B * C * Hl * Wl >= (1ll << 40)) return SD_ERR_SHAPE;
    if (workspace_bytes < sd_kl_pixels_up_workspace_bytes(B, C, Hl, Wl, scale)) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    sd::UpParams p;
    std::memset(&p, 0, sizeof(p));
    p.S = S; p.T = T; p.dS = dS;
    p.B = B; p.C = C; p.Hl = Hl; p.Wl = Wl; p.scale = scale;
    const double R = (double)B * Hl * Wl * scale * scale;        // rows = up-sampled pixels
    p.c2 = (float)(1.4426950408889634 / (double)tau);
    p.inv_c2 = (float)((double)tau / 1.4426950408889634);
    p.inv_Wl = 1.0f / (float)Wl;
    p.inv_tau = (float)(1.0 / (double)tau);
    p.coef = (float)((double)grad_scale * (double)alpha / (R * (double)tau));
    p.inv_coef = (float)(R * (double)tau / ((double)grad_scale * (double)alpha));
    p.loss_scale = (float)((double)alpha / R);
    p.loss = loss;
    char* ws = static_cast<char*>(workspace);
    p.ctrl = reinterpret_cast<unsigned*>(ws);
    p.part = reinterpret_cast<float*>(ws + sd::kArenaBytes);
    p.wpart = reinterpret_cast<float*>(ws + sd::kArenaBytes + sizeof(float) * sd::kMaxGrid);
    cudaError_t e = sd::launch_kl_pixels_up(p, dtype == SD_BF16, dev.sms, static_cast<cudaStream_t>(stream));
    g_launches += scale == 8 ? 2 : 1;
    t_last_kernel = "kl_pixels_up_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

// ============================================================================ cross-entropy behind the resize
size_t sd_ce_up_workspace_bytes(int B, int C, int Hl, int Wl, int scale) {
    size_t n = sd::kArenaBytes + sizeof(float) * 2 * sd::kMaxGrid;
    if (scale == 8 && B > 0 && C > 0 && Hl > 0 && Wl > 0)       // the four window planes of the gradient, fp32
        n += sizeof(float) * 4 * (size_t)B * C * Hl * Wl;
    return (n + 255) & ~(size_t)255;
}

int sd_ce_up_fwd_bwd(const void* logits, const int64_t* label, void* dlogits, float* loss, float* acc,
                     const float* class_weight, const float* pixel_weight, int B, int C, int Hl, int Wl, int scale,
                     int dtype, int64_t ignore_index, float loss_weight, double denominator, float grad_scale,
                     void* workspace, size_t workspace_bytes, void* stream) {
    if (!logits || !label || !dlogits || !loss || !workspace) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (B <= 0 || C <= 0 || Hl <= 0 || Wl <= 0) return SD_ERR_SHAPE;
    if (!(denominator > 0.0)) return SD_ERR_VALUE;
    if (scale != 1 && scale != 2 && scale != 4 && scale != 8) return SD_ERR_UNSUPPORTED;
    if ((long long)B * C * Hl * Wl >= (1ll << 40) || (long long)B * Hl * Wl * scale * scale >= (1ll << 40)) return SD_ERR_SHAPE;
    if (workspace_bytes < sd_ce_up_workspace_bytes(B, C, Hl, Wl, scale)) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    sd::CeParams p;
    std::memset(&p, 0, sizeof(p));
    p.X = logits;
    p.label = reinterpret_cast<const long long*>(label);
    p.class_weight = class_weight;
    p.pix_weight = pixel_weight;
    p.dX = dlogits;
    p.loss = loss;
    p.acc = acc;
    p.B = B; p.C = C; p.Hl = Hl; p.Wl = Wl; p.scale = scale;
    p.ignore_index = ignore_index;
    p.gscale = (float)((double)grad_scale * (double)loss_weight / denominator);
    p.lscale = (float)((double)loss_weight / denominator);
    p.acc_scale = (float)(100.0 / ((double)B * Hl * Wl * scale * scale));
    char* ws = static_cast<char*>(workspace);
    p.ctrl = reinterpret_cast<unsigned*>(ws);
    p.part = reinterpret_cast<float*>(ws + sd::kArenaBytes);
    p.wpart = reinterpret_cast<float*>(ws + sd::kArenaBytes + sizeof(float) * 2 * sd::kMaxGrid);
    cudaError_t e = sd::launch_ce_up(p, dtype == SD_BF16, dev.sms, static_cast<cudaStream_t>(stream));
    g_launches += scale == 8 ? 2 : 1;
    t_last_kernel = "ce_up_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

// ============================================================================ IFVD similarity term
size_t sd_ifvd_sim_workspace_bytes(int B, int C, int HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return 0;
    return sd::ifvd_workspace_layout(B, C, HW, sd::ifvd_pix_threads()).bytes;
}

int sd_ifvd_max_channels(void) { return sd::ifvd_max_channels(); }

int sd_ifvd_class_map(const int64_t* target, int32_t* cls, int B, int Ht, int Wt, int h, int w, int C, void* stream) {
    if (!target || !cls) return SD_ERR_NULL;
    if (B <= 0 || Ht <= 0 || Wt <= 0 || h <= 0 || w <= 0 || C <= 0) return SD_ERR_SHAPE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    cudaError_t e = sd::launch_ifvd_class_map(reinterpret_cast<const long long*>(target), cls, B, Ht, Wt, h, w, C,
                                              static_cast<cudaStream_t>(stream));
    g_launches += 1;
    t_last_kernel = "ifvd_class_map_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

int sd_ifvd_sim_fwd_bwd(const void* S, const void* T, const int32_t* cls, void* dS, float* loss, int B, int C, int HW,
                        int dtype, float weight, float grad_scale, int accumulate, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!S || !T || !cls || !dS || !loss || !workspace) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (B <= 0 || C <= 0 || HW <= 0 || B > 32767 || (long long)B * C * HW >= (1ll << 40)) return SD_ERR_SHAPE;
    if (C > sd::ifvd_max_channels()) return SD_ERR_UNSUPPORTED;
    const sd::IfvdWorkspace w = sd::ifvd_workspace_layout(B, C, HW, sd::ifvd_pix_threads());
    if (workspace_bytes < w.bytes) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    char* ws = static_cast<char*>(workspace);
    sd::IfvdParams p;
    std::memset(&p, 0, sizeof(p));
    p.S = S; p.T = T; p.cls = cls; p.dS = dS; p.loss = loss;
    p.sums = reinterpret_cast<float*>(ws + w.off_sums);
    p.wsum = reinterpret_cast<float*>(ws + w.off_wsum);
    p.pix = reinterpret_cast<float*>(ws + w.off_pix);
    p.part = reinterpret_cast<float*>(ws + w.off_part);
    p.spart = reinterpret_cast<float*>(ws + w.off_spart);
    p.B = B; p.C = C; p.HW = HW;
    p.splits = w.splits;
    p.wsplits = w.wsplits;
    p.accumulate = accumulate != 0;
    p.vec = HW % (16 / elem_size(dtype)) == 0 && aligned16(S) && aligned16(T) && aligned16(workspace);
    const double npix = (double)B * (double)HW;
    p.gcoef = (float)((double)grad_scale * 2.0 * (double)weight / npix);
    cudaError_t e = sd::launch_ifvd_sim(p, dtype == SD_BF16, (float)((double)weight / npix),
                                        static_cast<cudaStream_t>(stream));
    g_launches += 5 + (w.splits > 1) + (w.wsplits > 1);
    t_last_kernel = "ifvd_sim (class sums, sim, weighted class sums, grad)";
    return e == cudaSuccess ? SD_OK : (int)e;
}

// ============================================================================ MSE
size_t sd_mse_workspace_bytes(int64_t numel) {
    (void)numel;
    return sd::kArenaBytes + sizeof(float) * sd::kMseMaxGrid;
}

int sd_mse_fwd_bwd(const void* S, const void* T, void* dS, float* loss, int64_t numel, int dtype, float weight,
                   float grad_scale, void* workspace, size_t workspace_bytes, void* stream) {
    if (!S || !T || !dS || !loss || !workspace) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (numel <= 0 || numel >= (1ll << 40)) return SD_ERR_SHAPE;
    if (workspace_bytes < sd_mse_workspace_bytes(numel)) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    const int VE = 16 / elem_size(dtype);
    long long want = (numel / VE + 255) / 256;
    if (want < 1) want = 1;
    int grid = dev.sms * 8;
    if (grid > sd::kMseMaxGrid) grid = sd::kMseMaxGrid;
    if (want < grid) grid = (int)want;
    const float gcoef = (float)((double)grad_scale * 2.0 * (double)weight / (double)numel);
    const float scale = (float)((double)weight / (double)numel);
    cudaError_t e = sd::launch_mse(S, T, dS, loss,
                                   reinterpret_cast<float*>(static_cast<char*>(workspace) + sd::kArenaBytes), numel, dtype == SD_BF16, gcoef,
                                   scale, grid, static_cast<cudaStream_t>(stream));
    g_launches += 2;
    t_last_kernel = "mse_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

int sd_scale_grad_group(int n_tensors, void* const* dS, const int64_t* numel, int dtype, const float* const* grad_outputs,
                        void* stream) {
    if (!dS || !numel || !grad_outputs) return SD_ERR_NULL;
    if (n_tensors < 1 || n_tensors > sd::kMaxSegs) return SD_ERR_VALUE;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    long long nn[sd::kMaxSegs], most = 0;
    for (int k = 0; k < n_tensors; ++k) {
        if (!dS[k] || !grad_outputs[k]) return SD_ERR_NULL;
        if (numel[k] <= 0) return SD_ERR_SHAPE;
        nn[k] = numel[k];
        most = std::max(most, nn[k]);
    }
    const int VE = 16 / elem_size(dtype);
    long long want = (most / VE + 255) / 256;
    if (want < 1) want = 1;
    int grid = dev.sms * 4 / n_tensors;
    if (grid < 1) grid = 1;
    if (want < grid) grid = (int)want;
    cudaError_t e = sd::launch_scale_grad_group(n_tensors, dS, nn, dtype == SD_BF16, grad_outputs, grid,
                                                static_cast<cudaStream_t>(stream));
    g_launches += 1;
    t_last_kernel = "scale_grad_group_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

static int scale_grad_impl(void* dS, int64_t numel, int dtype, const float* grad_output, const float* log_values, int log_n,
                           float* log_ring, unsigned* log_cursor, int log_slots, void* stream) {
    if (!dS || !grad_output) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (numel <= 0) return SD_ERR_SHAPE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    const int VE = 16 / elem_size(dtype);
    long long want = (numel / VE + 255) / 256;
    if (want < 1) want = 1;
    int grid = dev.sms * 2;      // (scale_span: the usual launch has nothing to scale, it only starts and ends)
    if (want < grid) grid = (int)want;
    cudaError_t e = sd::launch_scale_grad(dS, numel, dtype == SD_BF16, grad_output, grid, static_cast<cudaStream_t>(stream),
                                          log_values, log_n, log_ring, log_cursor, log_slots);
    g_launches += 1;
    t_last_kernel = "scale_grad_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

int sd_scale_grad(void* dS, int64_t numel, int dtype, const float* grad_output, void* stream) {
    return scale_grad_impl(dS, numel, dtype, grad_output, nullptr, 0, nullptr, nullptr, 0, stream);
}

int sd_scale_grad_log(void* dS, int64_t numel, int dtype, const float* grad_output, const float* values, int n, float* ring,
                      unsigned* cursor, int slots, void* stream) {
    if (!values || !ring || !cursor) return SD_ERR_NULL;
    if (n <= 0 || slots <= 0) return SD_ERR_SHAPE;
    return scale_grad_impl(dS, numel, dtype, grad_output, values, n, ring, cursor, slots, stream);
}

int sd_log_push(const float* values, int n, float* ring, unsigned* cursor, int slots, void* stream) {
    if (!values || !ring || !cursor) return SD_ERR_NULL;
    if (n <= 0 || slots <= 0) return SD_ERR_SHAPE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    cudaError_t e = sd::launch_log_push(values, n, ring, cursor, slots, static_cast<cudaStream_t>(stream));
    g_launches += 1;
    t_last_kernel = "log_push_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

int sd_scale_grad2(void* dS, int64_t numel, int dtype, const float* grad_output0, const float* grad_output1,
                   unsigned* nonuniform_flag, void* stream) {
    if (!dS || !grad_output0 || !grad_output1 || !nonuniform_flag) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (numel <= 0) return SD_ERR_SHAPE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    const int VE = 16 / elem_size(dtype);
    long long want = (numel / VE + 255) / 256;
    if (want < 1) want = 1;
    int grid = dev.sms * 2;      // (scale_span: the usual launch has nothing to scale, it only starts and ends)
    if (want < grid) grid = (int)want;
    cudaError_t e = sd::launch_scale_grad2(dS, numel, dtype == SD_BF16, grad_output0, grad_output1, nonuniform_flag,
                                           grid, static_cast<cudaStream_t>(stream));
    g_launches += 1;
    t_last_kernel = "scale_grad2_kernel";
    return e == cudaSuccess ? SD_OK : (int)e;
}

}  // extern "C"

// ============================================================================ CGD correlation (extension)
extern "C" {

size_t sd_cgd_corr_workspace_bytes(int B, int C, int HW, int group) {
    if (B <= 0 || C <= 0 || HW <= 0 || group <= 0) return 0;
    return sd::cgd_corr_workspace_bytes(B, C, HW, group);
}

int sd_cgd_corr_fwd_bwd(const void* S, const void* T, void* dS, float* loss, int B, int C, int HW, int group, int dtype,
                        float alpha, float grad_scale, void* workspace, size_t workspace_bytes, void* stream) {
    if (!S || !T || !dS || !loss || !workspace) return SD_ERR_NULL;
    if (dtype != SD_F32 && dtype != SD_BF16) return SD_ERR_DTYPE;
    if (B <= 0 || C <= 0 || HW <= 0) return SD_ERR_SHAPE;
    if (group < 1) return SD_ERR_VALUE;
    if ((long long)B * C >= (1ll << 31) || (long long)B * C * HW >= (1ll << 40)) return SD_ERR_SHAPE;
    if (workspace_bytes < sd::cgd_corr_workspace_bytes(B, C, HW, group)) return SD_ERR_WORKSPACE;
    DeviceInfo& dev = device_info();
    if (dev.cc_major != 10) return SD_ERR_DEVICE;
    const int es = elem_size(dtype);
    // TMA rows and 16-element epilogue vectors
    if (((long long)HW * es) % 16 != 0 || HW % 16 != 0 || !aligned16(S) || !aligned16(T) || !aligned16(dS))
        return SD_ERR_UNSUPPORTED;
    int rows0 = 0, rows1 = 0, kbox = 0;
    if (!sd::cgd_corr_geometry(B, C, HW, group, dtype, &rows0, &rows1, &kbox)) return SD_ERR_UNSUPPORTED;  // group > 256
    // maps 0-3: S, T tiles for the Gram (K-major, 128-byte swizzle); maps 4-5: the S tiles again for the gradient
    // GEMM, where they are read MN-major - 32-bit operands then need the 32-byte-atom flavour of the swizzle
    alignas(64) CUtensorMap maps[6];
    const bool atom32 = dtype == SD_F32;
    if (!encode_rows_map(&maps[0], S, B, C, HW, dtype, kbox, rows0) ||
        !encode_rows_map(&maps[1], T, B, C, HW, dtype, kbox, rows0) ||
        !encode_rows_map(&maps[2], S, B, C, HW, dtype, kbox, rows1 > 0 ? rows1 : rows0) ||
        !encode_rows_map(&maps[3], T, B, C, HW, dtype, kbox, rows1 > 0 ? rows1 : rows0) ||
        !encode_rows_map(&maps[4], S, B, C, HW, dtype, kbox, rows0, atom32) ||
        !encode_rows_map(&maps[5], S, B, C, HW, dtype, kbox, rows1 > 0 ? rows1 : rows0, atom32))
        return (int)cudaErrorInvalidValue;
    cudaError_t e = sd::launch_cgd_corr(dS, loss, B, C, HW, group, dtype, alpha, grad_scale, workspace, maps,
                                        static_cast<cudaStream_t>(stream));
    g_launches += 3;
    t_last_kernel = "cgd_corr (gram, mask, grad: tcgen05)";
    return e == cudaSuccess ? SD_OK : (int)e;
}

}  // extern "C"
