"""ctypes binding of ``libsegdistill_sm100.so`` (C ABI in ``include/segdistill.h``).

PyTorch is plumbing here: it owns device memory and streams; every argument that
crosses the boundary is a raw pointer, a size or a scalar.  There is NO fallback:
if the library is missing, or the device is not sm_100, the calls raise.
"""
from __future__ import annotations

import contextlib
import ctypes
import math
import os
import threading
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libsegdistill_sm100.so')

SD_F32, SD_BF16 = 0, 1
ALGO_AUTO, ALGO_GENERIC, ALGO_TMA, ALGO_STREAM, ALGO_CLUSTER, ALGO_ROWS1, ALGO_GRID, ALGO_WARP = 0, 1, 2, 3, 4, 5, 6, 7
ALGOS = {'auto': ALGO_AUTO, 'generic': ALGO_GENERIC, 'tma': ALGO_TMA, 'stream': ALGO_STREAM,
         'cluster': ALGO_CLUSTER, 'rows1': ALGO_ROWS1, 'grid': ALGO_GRID, 'warp': ALGO_WARP}

# every symbol include/segdistill.h declares (tests check the library exports all of them)
EXPORTS = (
    'sd_abi_version', 'sd_strerror', 'sd_device_check',
    'sd_kl_rows_workspace_bytes', 'sd_kl_rows_fwd_bwd', 'sd_kl_rows_multi_fwd_bwd', 'sd_scale_grad2',
    'sd_kl_pixels_workspace_bytes', 'sd_kl_pixels_fwd_bwd',
    'sd_kl_rows_up_workspace_bytes', 'sd_kl_rows_up_fwd_bwd', 'sd_kl_pixels_up_workspace_bytes', 'sd_kl_pixels_up_fwd_bwd',
    'sd_mse_workspace_bytes', 'sd_mse_fwd_bwd', 'sd_ifvd_sim_workspace_bytes', 'sd_ifvd_sim_fwd_bwd', 'sd_ifvd_max_channels',
    'sd_ifvd_class_map', 'sd_scale_grad', 'sd_scale_grad_log',
    'sd_cgd_corr_workspace_bytes', 'sd_cgd_corr_fwd_bwd', 'sd_log_push', 'sd_ce_up_workspace_bytes', 'sd_ce_up_fwd_bwd',
    'sd_kl_rows_group_workspace_bytes', 'sd_kl_rows_group_fwd_bwd', 'sd_scale_grad_group',
    'sd_launch_count', 'sd_last_kernel',
)

_lib = None
_lock = threading.Lock()


class SegDistillError(RuntimeError):
    pass


class SegDistillUnsupported(SegDistillError):
    """SD_ERR_UNSUPPORTED: the requested kernel cannot run this layout (the caller picks another entry point)."""


def load():
    """Load the CUDA extension; raises (never falls back) when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SegDistillError(
                f'{LIB_PATH} is missing: build it with `python -m segdistill_b200.build` '
                '(nvcc, sm_100a). segdistill_b200 has no CPU or PyTorch fallback.')
        lib = ctypes.CDLL(LIB_PATH)
        c = ctypes
        vp, i32, f32, sz, i64 = c.c_void_p, c.c_int, c.c_float, c.c_size_t, c.c_int64
        lib.sd_abi_version.restype = i32
        lib.sd_strerror.restype = c.c_char_p
        lib.sd_strerror.argtypes = [i32]
        lib.sd_device_check.restype = i32
        lib.sd_kl_rows_workspace_bytes.restype = sz
        lib.sd_kl_rows_workspace_bytes.argtypes = [i32, i32, i32, i32]
        lib.sd_kl_rows_fwd_bwd.restype = i32
        lib.sd_kl_rows_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32,
                                           f32, f32, f32, f32, vp, vp, sz, i32, vp]
        lib.sd_kl_rows_multi_fwd_bwd.restype = i32
        lib.sd_kl_rows_multi_fwd_bwd.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp,
                                                 i32, i32, i32, i32, f32, vp, sz, i32, vp]
        lib.sd_scale_grad2.restype = i32
        lib.sd_scale_grad2.argtypes = [vp, i64, i32, vp, vp, vp, vp]
        lib.sd_kl_pixels_workspace_bytes.restype = sz
        lib.sd_kl_pixels_workspace_bytes.argtypes = [i32, i32, i32]
        lib.sd_kl_pixels_fwd_bwd.restype = i32
        lib.sd_kl_pixels_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32,
                                             f32, f32, f32, f32, vp, vp, sz, i32, vp]
        lib.sd_kl_rows_up_workspace_bytes.restype = sz
        lib.sd_kl_rows_up_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
        lib.sd_kl_rows_up_fwd_bwd.restype = i32
        lib.sd_kl_rows_up_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32,
                                              f32, f32, f32, vp, sz, vp]
        lib.sd_kl_pixels_up_workspace_bytes.restype = sz
        lib.sd_kl_pixels_up_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
        lib.sd_kl_pixels_up_fwd_bwd.restype = i32
        lib.sd_kl_pixels_up_fwd_bwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, f32, vp, sz, vp]
        lib.sd_mse_workspace_bytes.restype = sz
        lib.sd_mse_workspace_bytes.argtypes = [i64]
        lib.sd_mse_fwd_bwd.restype = i32
        lib.sd_mse_fwd_bwd.argtypes = [vp, vp, vp, vp, i64, i32, f32, f32, vp, sz, vp]
        lib.sd_ifvd_sim_workspace_bytes.restype = sz
        lib.sd_ifvd_sim_workspace_bytes.argtypes = [i32, i32, i32]
        lib.sd_ifvd_sim_fwd_bwd.restype = i32
        lib.sd_ifvd_sim_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, i32, vp, sz, vp]
        lib.sd_ifvd_max_channels.restype = i32
        lib.sd_ifvd_class_map.restype = i32
        lib.sd_ifvd_class_map.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
        lib.sd_scale_grad.restype = i32
        lib.sd_scale_grad.argtypes = [vp, i64, i32, vp, vp]
        lib.sd_scale_grad_log.restype = i32
        lib.sd_scale_grad_log.argtypes = [vp, i64, i32, vp, vp, i32, vp, vp, i32, vp]
        lib.sd_scale_grad_group.restype = i32
        lib.sd_scale_grad_group.argtypes = [i32, vp, vp, i32, vp, vp]
        lib.sd_cgd_corr_workspace_bytes.restype = sz
        lib.sd_cgd_corr_workspace_bytes.argtypes = [i32, i32, i32, i32]
        lib.sd_cgd_corr_fwd_bwd.restype = i32
        lib.sd_cgd_corr_fwd_bwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, vp, sz, vp]
        lib.sd_ce_up_workspace_bytes.restype = sz
        lib.sd_ce_up_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
        lib.sd_ce_up_fwd_bwd.restype = i32
        lib.sd_ce_up_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i64, f32, c.c_double, f32,
                                         vp, sz, vp]
        lib.sd_kl_rows_group_workspace_bytes.restype = sz
        lib.sd_kl_rows_group_workspace_bytes.argtypes = [i32, vp, vp, vp, vp, i32]
        lib.sd_kl_rows_group_fwd_bwd.restype = i32
        lib.sd_kl_rows_group_fwd_bwd.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, vp, vp, sz, vp]
        lib.sd_log_push.restype = i32
        lib.sd_log_push.argtypes = [vp, i32, vp, vp, i32, vp]
        lib.sd_launch_count.restype = c.c_uint64
        lib.sd_last_kernel.restype = c.c_char_p
        if lib.sd_abi_version() != 1:
            raise SegDistillError('libsegdistill_sm100.so: ABI version mismatch, rebuild it')
        _lib = lib
    return _lib


def _check(rc: int):
    if rc == -6:
        raise SegDistillUnsupported(f'{load().sd_strerror(rc).decode()} (rc={rc})')
    if rc != 0:
        raise SegDistillError(f'{load().sd_strerror(rc).decode()} (rc={rc})')


def launch_count() -> int:
    return int(load().sd_launch_count())


def last_kernel() -> str:
    return load().sd_last_kernel().decode()


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return SD_F32
    if t.dtype == torch.bfloat16:
        return SD_BF16
    raise SegDistillError(f'unsupported dtype {t.dtype}: float32 or bfloat16 (fp32 accumulation either way)')


def _prep_pair(x_student: torch.Tensor, x_teacher: torch.Tensor):
    if x_student.shape != x_teacher.shape:
        raise SegDistillError(f'student {tuple(x_student.shape)} and teacher {tuple(x_teacher.shape)} '
                              'maps must have the same shape')
    if not x_student.is_cuda or x_teacher.device != x_student.device:
        raise SegDistillError('segdistill_b200 runs on CUDA tensors only (no CPU fallback)')
    if x_student.dtype == torch.float16:      # fp16 features (Fp16OptimizerHook): compute in fp32
        x_student = x_student.float()
    if x_teacher.dtype != x_student.dtype:
        x_teacher = x_teacher.to(x_student.dtype)
    s = x_student.detach().contiguous()
    t = x_teacher.detach().contiguous()
    return s, t, _dtype_code(s)


_NULL_CTX = contextlib.nullcontext()


def _on(device: torch.device):
    """Device guard that costs nothing when `device` already is the current device (the usual case)."""
    if device.index is None or torch.cuda.current_device() == device.index:
        return _NULL_CTX
    return torch.cuda.device(device)


def _dev_index(device: torch.device) -> int:
    return torch.cuda.current_device() if device.index is None else device.index


# ---------------------------------------------------------------- workspaces
_workspaces = {}
_retired_workspaces = []    # outgrown workspaces stay allocated: a captured CUDA graph may hold their address


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch, one per (device, stream); grown on demand, never shared across streams."""
    idx = _dev_index(device)
    key = (idx, _raw_stream(idx))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired_workspaces.append(ws)
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def workspace_error_flag(device=None) -> int:
    """Spin time-out flag of the split-row kernel (0 = never fired); synchronises. Tests only."""
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    idx = _dev_index(device)
    key = (idx, _raw_stream(idx))
    ws = _workspaces.get(key)
    if ws is None:
        return 0
    return int(ws[:8].view(torch.int32)[1].item())


def _raw_stream(index: int) -> int:
    return torch._C._cuda_getCurrentRawStream(index)


def _stream_ptr(device) -> int:
    return _raw_stream(_dev_index(device))


_ws_bytes_cache = {}


def _rows_ws_bytes(lib, B, C, HW, group):
    key = (B, C, HW, group)
    n = _ws_bytes_cache.get(key)
    if n is None:
        n = _ws_bytes_cache[key] = lib.sd_kl_rows_workspace_bytes(B, C, HW, group)
    return n


# ---------------------------------------------------------------- entry points
def kl_rows(x_student, x_teacher, group=1, tau=1.0, alpha=1.0, perm: Optional[torch.Tensor] = None,
            grad_scale=1.0, mse_weight=0.0, algo=ALGO_AUTO, want_row_kl=False, bchw=None):
    """Row-wise softmax-KL (rows = ``group`` channels x HW). Returns (loss, dS, row_kl|None, mse|None).

    ``bchw``: optional (B, C, HW) override to treat the tensor as that 3-D view.
    """
    lib = load()
    s, t, code = _prep_pair(x_student, x_teacher)
    if bchw is None:
        B, C = s.shape[0], s.shape[1]
        HW = math.prod(s.shape[2:])
    else:
        B, C, HW = bchw
    dev = s.device
    with _on(dev):
        ds = torch.empty_like(s)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        G = (C + min(group, C) - 1) // min(group, C)
        row_kl = torch.empty(B * G, dtype=torch.float32, device=dev) if want_row_kl else None
        perm_dev = None
        if perm is not None:
            perm_dev = perm.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
            if perm_dev.numel() != C:
                raise SegDistillError('chan_perm must have C entries')
        ws = _workspace(dev, _rows_ws_bytes(lib, B, C, HW, group))
        rc = lib.sd_kl_rows_fwd_bwd(
            s.data_ptr(), t.data_ptr(), ds.data_ptr(),
            row_kl.data_ptr() if row_kl is not None else None, out.data_ptr(),
            perm_dev.data_ptr() if perm_dev is not None else None,
            B, C, HW, group, code, float(tau), float(alpha), float(grad_scale),
            float(mse_weight), out[1:].data_ptr() if mse_weight != 0 else None,
            ws.data_ptr(), ws.numel(), int(algo), _stream_ptr(dev))
        _check(rc)
    return out[0], ds, row_kl, (out[1] if mse_weight != 0 else None)


_multi_calls = {}


class _MultiCall:
    """Everything of a two-loss launch that does not change from step to step (ctypes arrays, sizes, dtype code)."""
    __slots__ = ('n', 'B', 'C', 'HW', 'code', 'g_arr', 't_arr', 'a_arr', 'l_arr', 'ws_bytes', 'fn')

    def __init__(self, lib, shape, dtype, groups, taus, alphas):
        c = ctypes
        self.n = n = len(groups)
        self.B, self.C = shape[0], shape[1]
        self.HW = math.prod(shape[2:])
        self.code = SD_F32 if dtype == torch.float32 else SD_BF16
        self.g_arr = (c.c_int * n)(*[int(g) for g in groups])
        self.t_arr = (c.c_float * n)(*[float(v) for v in taus])
        self.a_arr = (c.c_float * n)(*[float(v) for v in alphas])
        self.l_arr = (c.c_void_p * n)()
        self.ws_bytes = lib.sd_kl_rows_workspace_bytes(self.B, self.C, self.HW, min(int(g) for g in groups))
        self.fn = lib.sd_kl_rows_multi_fwd_bwd


def kl_rows_multi(x_student, x_teacher, groups, taus, alphas, grad_outputs=None, run_if=None, ds=None,
                  algo=ALGO_AUTO):
    """Two row-wise softmax-KL losses over the same pair in one pass. Returns (losses[n], dS).

    ``grad_outputs``/``run_if``/``ds`` serve the conditional backward re-run (see functional._KLRowsMulti).
    """
    s, t = x_student, x_teacher
    if (s.dtype is not torch.float32 and s.dtype is not torch.bfloat16) or t.dtype is not s.dtype or s.shape != t.shape \
            or not s.is_cuda or t.device != s.device or not s.is_contiguous() or not t.is_contiguous():
        s, t, _ = _prep_pair(x_student, x_teacher)          # conversions, checks with messages
    key = (s.shape, s.dtype, groups, taus, alphas) if type(groups) is tuple else (s.shape, s.dtype, tuple(groups), tuple(taus), tuple(alphas))
    call = _multi_calls.get(key)
    if call is None:
        call = _multi_calls[key] = _MultiCall(load(), tuple(s.shape), s.dtype, groups, taus, alphas)
    dev = s.device
    n = call.n
    with _on(dev):
        if ds is None:
            ds = torch.empty(s.shape, dtype=s.dtype, device=dev)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        base = out.data_ptr()
        for k in range(n):
            call.l_arr[k] = base + 4 * k
        go_arr = None
        if grad_outputs is not None:
            go_arr = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grad_outputs])
        idx = _dev_index(dev)
        stream = _raw_stream(idx)
        ws = _workspaces.get((idx, stream))
        if ws is None or ws.numel() < call.ws_bytes:
            ws = _workspace(dev, call.ws_bytes)
        rc = call.fn(s.data_ptr(), t.data_ptr(), ds.data_ptr(), n, call.g_arr, call.t_arr, call.a_arr, call.l_arr, None,
                     go_arr, run_if.data_ptr() if run_if is not None else None,
                     call.B, call.C, call.HW, call.code, 1.0, ws.data_ptr(), ws.numel(), int(algo), stream)
        if rc:
            _check(rc)
    return out, ds


def multi_supported(shape, groups, elem_size):
    """Host-side mirror of the C ABI's fusability test (nesting + 16-byte rows); the co-residency
    limit is checked by the library, which answers SD_ERR_UNSUPPORTED."""
    C = shape[1]
    HW = 1
    for d in shape[2:]:
        HW *= d
    g = sorted(min(int(v), C) for v in groups)
    if len(g) != 2 or (HW * elem_size) % 16:
        return False
    return g[1] == C or g[1] % g[0] == 0


def scale_grad2_(ds: torch.Tensor, go0: torch.Tensor, go1: torch.Tensor) -> torch.Tensor:
    """dS *= go when go0 == go1; returns the device flag (1 = non-uniform, dS untouched)."""
    lib = load()
    flag = torch.empty(1, dtype=torch.int32, device=ds.device)
    with _on(ds.device):
        rc = lib.sd_scale_grad2(ds.data_ptr(), ds.numel(), _dtype_code(ds), go0.data_ptr(), go1.data_ptr(),
                                flag.data_ptr(), _stream_ptr(ds.device))
        _check(rc)
    return flag


def kl_pixels(x_student, x_teacher, tau=1.0, alpha=1.0, grad_scale=1.0, at_weight=0.0,
              algo=ALGO_AUTO, want_row_kl=False):
    """Per-pixel softmax-KL over channels. Returns (loss, dS, row_kl|None, at_loss|None)."""
    lib = load()
    s, t, code = _prep_pair(x_student, x_teacher)
    B, C = s.shape[0], s.shape[1]
    HW = math.prod(s.shape[2:])
    dev = s.device
    with _on(dev):
        ds = torch.empty_like(s)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        row_kl = torch.empty(B * HW, dtype=torch.float32, device=dev) if want_row_kl else None
        nbytes = lib.sd_kl_pixels_workspace_bytes(B, C, HW)
        ws = _workspace(dev, nbytes)
        rc = lib.sd_kl_pixels_fwd_bwd(
            s.data_ptr(), t.data_ptr(), ds.data_ptr(),
            row_kl.data_ptr() if row_kl is not None else None, out.data_ptr(),
            B, C, HW, code, float(tau), float(alpha), float(grad_scale),
            float(at_weight), out[1:].data_ptr() if at_weight != 0 else None,
            ws.data_ptr(), ws.numel(), int(algo), _stream_ptr(dev))
        _check(rc)
    return out[0], ds, row_kl, (out[1] if at_weight != 0 else None)


UP_SCALES = (2, 4, 8)


def up_supported(hl: int, wl: int) -> bool:
    """Whether a low-resolution plane of this width fits the strip buffers of kl_rows_up (csrc/params.h:up_strip_rows)."""
    return (100 * 1024 // 4 // int(wl) - 8) // 11 >= 1 and hl >= 1


def kl_rows_up(x_student, x_teacher, scale, group=1, tau=1.0, alpha=1.0, perm: Optional[torch.Tensor] = None,
               grad_scale=1.0, want_row_kl=False):
    """Channel-mode softmax-KL on the maps up-sampled ``scale`` x (bilinear, align_corners=False) without
    materialising them.  x_*: low-resolution [B, C, Hl, Wl].  Returns (loss, dS at low resolution, row_kl)."""
    lib = load()
    s, t, code = _prep_pair(x_student, x_teacher)
    if s.dim() != 4:
        raise ValueError('kl_rows_up expects 4-D NCHW maps')
    B, C, Hl, Wl = s.shape
    dev = s.device
    with _on(dev):
        ds = torch.empty_like(s)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        G = -(-C // min(int(group), C))
        row_kl = torch.empty(B * G, dtype=torch.float32, device=dev) if want_row_kl else None
        p = None
        if perm is not None:
            p = perm.to(device=dev, dtype=torch.int32).contiguous()
        ws = _workspace(dev, lib.sd_kl_rows_up_workspace_bytes(B, C, Hl, Wl, int(group)))
        rc = lib.sd_kl_rows_up_fwd_bwd(s.data_ptr(), t.data_ptr(), ds.data_ptr(),
                                       row_kl.data_ptr() if row_kl is not None else None, out.data_ptr(),
                                       p.data_ptr() if p is not None else None, B, C, Hl, Wl, int(scale), int(group),
                                       code, float(tau), float(alpha), float(grad_scale), ws.data_ptr(), ws.numel(),
                                       _stream_ptr(dev))
        _check(rc)
    return out[0], ds, row_kl


UP_PIXEL_SCALES = (2, 4, 8)


def kl_pixels_up(x_student, x_teacher, scale, tau=1.0, alpha=1.0, grad_scale=1.0):
    """Per-pixel softmax-KL over channels on the maps up-sampled ``scale`` x (bilinear, align_corners=False)
    without materialising them.  Returns (loss, dS at low resolution)."""
    lib = load()
    s, t, code = _prep_pair(x_student, x_teacher)
    if s.dim() != 4:
        raise ValueError('kl_pixels_up expects 4-D NCHW maps')
    B, C, Hl, Wl = s.shape
    dev = s.device
    with _on(dev):
        ds = torch.empty_like(s)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        ws = _workspace(dev, lib.sd_kl_pixels_up_workspace_bytes(B, C, Hl, Wl, int(scale)))
        rc = lib.sd_kl_pixels_up_fwd_bwd(s.data_ptr(), t.data_ptr(), ds.data_ptr(), out.data_ptr(), B, C, Hl, Wl,
                                         int(scale), code, float(tau), float(alpha), float(grad_scale),
                                         ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _check(rc)
    return out[0], ds


def mse(x_student, x_teacher, weight=1.0, grad_scale=1.0):
    """``weight * mean((s - t)^2)``. Returns (loss, dS)."""
    lib = load()
    s, t, code = _prep_pair(x_student, x_teacher)
    dev = s.device
    with _on(dev):
        ds = torch.empty_like(s)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        ws = _workspace(dev, lib.sd_mse_workspace_bytes(s.numel()))
        rc = lib.sd_mse_fwd_bwd(s.data_ptr(), t.data_ptr(), ds.data_ptr(), out.data_ptr(), s.numel(), code,
                                float(weight), float(grad_scale), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _check(rc)
    return out[0], ds


def ifvd_sim(x_student, x_teacher, cls, weight=10.0, grad_scale=1.0, ds=None):
    """IFVDLoss similarity term: ``weight * mean((cos(s, centre_s) - cos(t, centre_t))^2)``; ``cls`` is the (B, H*W)
    class index of every pixel, ``C`` meaning "no class". Returns (loss, dS); with ``ds`` given (the gradient of
    another loss on the same student map, same dtype and shape) the gradient is added to it in place."""
    lib = load()
    s, t, code = _prep_pair(x_student, x_teacher)
    B, C = s.shape[0], s.shape[1]
    HW = math.prod(s.shape[2:])
    dev = s.device
    if cls.device != dev or cls.numel() != B * HW:
        raise SegDistillError(f'class map must hold B*H*W = {B * HW} entries on {dev}')
    if ds is not None and (ds.shape != s.shape or ds.dtype != s.dtype or not ds.is_contiguous()):
        raise SegDistillError('ds must be a contiguous tensor of the (converted) student map\'s shape and dtype')
    with _on(dev):
        k = cls.reshape(B, HW).to(torch.int32).contiguous()
        acc = ds is not None
        if ds is None:
            ds = torch.empty_like(s)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        ws = _workspace(dev, lib.sd_ifvd_sim_workspace_bytes(B, C, HW))
        rc = lib.sd_ifvd_sim_fwd_bwd(s.data_ptr(), t.data_ptr(), k.data_ptr(), ds.data_ptr(), out.data_ptr(), B, C, HW,
                                     code, float(weight), float(grad_scale), int(acc), ws.data_ptr(), ws.numel(),
                                     _stream_ptr(dev))
        _check(rc)
    return out[0], ds


def ifvd_max_channels() -> int:
    """Largest C the IFVD class-sum kernels take (their bins must fit one CTA's shared memory)."""
    return int(load().sd_ifvd_max_channels())


def ifvd_class_map(target, n_classes, h, w):
    """(B, h*w) int32 class of every pixel from an integer label map (B, 1, Ht, Wt): nearest-resized to (h, w),
    labels outside [0, n_classes) -> n_classes."""
    lib = load()
    if not target.is_cuda or target.dim() != 4 or target.shape[1] != 1 or target.is_floating_point():
        raise SegDistillError('label map must be an integer CUDA tensor of shape (B, 1, H, W)')
    dev = target.device
    with _on(dev):
        tg = target.to(torch.int64).contiguous()
        B, _, Ht, Wt = tg.shape
        cls = torch.empty(B, h * w, dtype=torch.int32, device=dev)
        _check(lib.sd_ifvd_class_map(tg.data_ptr(), cls.data_ptr(), B, Ht, Wt, int(h), int(w), int(n_classes),
                                     _stream_ptr(dev)))
    return cls


def cgd_corr(x_student, x_teacher, group=10, alpha=1.0, grad_scale=1.0):
    """Per-group Gram-matrix loss (tcgen05). Returns (loss, dS)."""
    lib = load()
    s, t, code = _prep_pair(x_student, x_teacher)
    B, C = s.shape[0], s.shape[1]
    HW = math.prod(s.shape[2:])
    dev = s.device
    with _on(dev):
        ds = torch.empty_like(s)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        ws = _workspace(dev, lib.sd_cgd_corr_workspace_bytes(B, C, HW, group))
        rc = lib.sd_cgd_corr_fwd_bwd(s.data_ptr(), t.data_ptr(), ds.data_ptr(), out.data_ptr(), B, C, HW, group,
                                     code, float(alpha), float(grad_scale), ws.data_ptr(), ws.numel(),
                                     _stream_ptr(dev))
        _check(rc)
    return out[0], ds


# A step's log append that waits for a launch to ride on (dist.DeferredLogs.push(..., in_backward=True)):
# (values, ring, cursor) or None.  The next scaling launch of a backward on the same device takes it.
pending_log = None


def scale_grad_(ds: torch.Tensor, grad_output: torch.Tensor):
    """In place ``ds *= grad_output`` on the device; a no-op launch when grad_output == 1.  Carries a pending log
    append (``pending_log``) when there is one for this device."""
    global pending_log
    lib = _lib or load()
    g = grad_output
    if g.device != ds.device or g.dtype is not torch.float32 or g.numel() != 1 or g.requires_grad:
        g = grad_output.detach().to(device=ds.device, dtype=torch.float32).reshape(1)
    dt = SD_F32 if ds.dtype is torch.float32 else _dtype_code(ds)
    with _on(ds.device):
        log = pending_log
        if log is not None and log[0].device == ds.device:
            pending_log = None
            values, ring, cursor = log
            rc = lib.sd_scale_grad_log(ds.data_ptr(), ds.numel(), dt, g.data_ptr(), values.data_ptr(), values.numel(),
                                       ring.data_ptr(), cursor.data_ptr(), ring.shape[0], _stream_ptr(ds.device))
        else:
            rc = lib.sd_scale_grad(ds.data_ptr(), ds.numel(), dt, g.data_ptr(), _stream_ptr(ds.device))
        if rc:
            _check(rc)
    return ds


def scale_grad_group_(dss, grad_outputs):
    """In place ``dss[k] *= grad_outputs[k]`` for the (same-dtype, same-device) gradients of a grouped launch: ONE launch,
    a no-op per tensor whose factor is 1."""
    lib = _lib or load()
    n = len(dss)
    dev = dss[0].device
    gs = []
    for g in grad_outputs:
        if g.device != dev or g.dtype is not torch.float32 or g.numel() != 1 or g.requires_grad:
            g = g.detach().to(device=dev, dtype=torch.float32).reshape(1)
        gs.append(g)
    ptrs = (ctypes.c_void_p * n)(*[d.data_ptr() for d in dss])
    gptr = (ctypes.c_void_p * n)(*[g.data_ptr() for g in gs])
    numel = (ctypes.c_int64 * n)(*[d.numel() for d in dss])
    with _on(dev):
        rc = lib.sd_scale_grad_group(n, ptrs, numel, _dtype_code(dss[0]), gptr, _stream_ptr(dev))
        if rc:
            _check(rc)
    return dss


def log_push(values: torch.Tensor, ring: torch.Tensor, cursor: torch.Tensor):
    """ring[cursor % slots] = values; cursor += 1, on the device (capturable).  values: n contiguous fp32 on the GPU;
    ring: (slots, n) fp32; cursor: one int32."""
    lib = load()
    n = values.numel()
    if values.dtype != torch.float32 or not values.is_contiguous() or ring.dtype != torch.float32 or ring.shape[1] != n:
        raise SegDistillError('log_push: values must be n contiguous fp32, ring (slots, n) fp32')
    with _on(values.device):
        _check(lib.sd_log_push(values.data_ptr(), n, ring.data_ptr(), cursor.data_ptr(), ring.shape[0],
                               _stream_ptr(values.device)))


CE_SCALES = (1, 2, 4, 8)


def ce_up(logits, label, scale, class_weight=None, pixel_weight=None, ignore_index=255, loss_weight=1.0,
          denominator=None, grad_scale=1.0):
    """Cross-entropy + top-1 accuracy of ``logits`` up-sampled ``scale`` x (bilinear, align_corners=False) against
    ``label`` (B, scale*Hl, scale*Wl), without materialising the up-sampled maps.  Returns (loss, acc, dlogits)."""
    lib = load()
    if not logits.is_cuda or logits.dim() != 4:
        raise SegDistillError('ce_up expects 4-D CUDA logits (no CPU fallback)')
    x = logits.detach()
    if x.dtype == torch.float16:
        x = x.float()
    x = x.contiguous()
    code = _dtype_code(x)
    B, C, Hl, Wl = x.shape
    dev = x.device
    lab = label.reshape(B, Hl * scale, Wl * scale)
    if lab.device != dev or lab.dtype != torch.int64 or not lab.is_contiguous():
        lab = lab.to(device=dev, dtype=torch.int64).contiguous()
    n_pix = B * Hl * Wl * scale * scale
    with _on(dev):
        dx = torch.empty_like(x)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        cw = None
        if class_weight is not None:
            cw = torch.as_tensor(class_weight, dtype=torch.float32, device=dev).contiguous()
            if cw.numel() != C:
                raise SegDistillError('class_weight must have C entries')
        pw = None
        if pixel_weight is not None:
            pw = pixel_weight.to(device=dev, dtype=torch.float32).reshape(B, Hl * scale, Wl * scale).contiguous()
        ws = _workspace(dev, lib.sd_ce_up_workspace_bytes(B, C, Hl, Wl, int(scale)))
        rc = lib.sd_ce_up_fwd_bwd(x.data_ptr(), lab.data_ptr(), dx.data_ptr(), out.data_ptr(), out[1:].data_ptr(),
                                  cw.data_ptr() if cw is not None else None, pw.data_ptr() if pw is not None else None,
                                  B, C, Hl, Wl, int(scale), code, int(ignore_index), float(loss_weight),
                                  float(n_pix if denominator is None else denominator), float(grad_scale),
                                  ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _check(rc)
    return out[0], out[1], dx


MAX_GROUP_PAIRS = 8
_group_calls = {}


class _GroupCall:
    """The step-invariant half of a grouped launch: shapes, group sizes, temperatures, weights as ctypes arrays."""
    __slots__ = ('n', 'code', 'B', 'C', 'HW', 'g', 'tau', 'alpha', 'S', 'T', 'D', 'L', 'ws_bytes', 'fn')

    def __init__(self, lib, shapes, dtype, groups, taus, alphas):
        c = ctypes
        self.n = n = len(shapes)
        self.code = SD_F32 if dtype == torch.float32 else SD_BF16
        self.B = (c.c_int * n)(*[int(s[0]) for s in shapes])
        self.C = (c.c_int * n)(*[int(s[1]) for s in shapes])
        self.HW = (c.c_int * n)(*[int(math.prod(s[2:])) for s in shapes])
        self.g = (c.c_int * n)(*[int(v) for v in groups])
        self.tau = (c.c_float * n)(*[float(v) for v in taus])
        self.alpha = (c.c_float * n)(*[float(v) for v in alphas])
        self.S, self.T, self.D, self.L = ((c.c_void_p * n)() for _ in range(4))
        self.ws_bytes = lib.sd_kl_rows_group_workspace_bytes(n, self.B, self.C, self.HW, self.g, self.code)
        self.fn = lib.sd_kl_rows_group_fwd_bwd


def kl_rows_group(students, teachers, groups, taus, alphas, grad_output=None):
    """Channel-mode softmax-KL of several (student, teacher) pairs in one launch.  Returns (losses[n], [dS_k])."""
    lib = load()
    n = len(students)
    if not 1 <= n <= MAX_GROUP_PAIRS:
        raise SegDistillUnsupported(f'a grouped launch takes 1..{MAX_GROUP_PAIRS} pairs')
    ss, ts = [], []
    for a, b in zip(students, teachers):
        s, t, _ = _prep_pair(a, b)
        ss.append(s)
        ts.append(t)
    dtype, dev = ss[0].dtype, ss[0].device
    if any(s.dtype != dtype or s.device != dev for s in ss):
        raise SegDistillUnsupported('a grouped launch needs one dtype and one device')
    key = (tuple(tuple(s.shape) for s in ss), dtype, tuple(groups), tuple(taus), tuple(alphas))
    call = _group_calls.get(key)
    if call is None:
        call = _group_calls[key] = _GroupCall(lib, key[0], dtype, groups, taus, alphas)
    if call.ws_bytes == 0:
        raise SegDistillUnsupported('grouped launch: a pair\'s rows are not 16-byte aligned')
    with _on(dev):
        out = torch.empty(n, dtype=torch.float32, device=dev)
        dss = [torch.empty_like(s) for s in ss]
        base = out.data_ptr()
        for k in range(n):
            call.S[k], call.T[k], call.D[k], call.L[k] = ss[k].data_ptr(), ts[k].data_ptr(), dss[k].data_ptr(), base + 4 * k
        ws = _workspace(dev, call.ws_bytes)
        rc = call.fn(n, call.S, call.T, call.D, call.L, call.B, call.C, call.HW, call.g, call.tau, call.alpha, call.code,
                     1.0, grad_output.data_ptr() if grad_output is not None else None, ws.data_ptr(), ws.numel(),
                     _stream_ptr(dev))
        _check(rc)
    return out, dss
