"""CPU checks of the C-ABI library: it loads, exports every symbol the header declares, and
rejects bad arguments before touching the device (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from segdistill_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        build.build()
    return _cabi.load()


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'segdistill.h')).read()
    return re.findall(r'SD_API\s+[\w\s\*]+?\b(sd_\w+)\s*\(', text)


def test_every_declared_symbol_is_exported(lib):
    declared = _header_symbols()
    assert len(declared) >= 14
    assert sorted(declared) == sorted(_cabi.EXPORTS)
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None


def test_abi_version_and_error_strings(lib):
    assert lib.sd_abi_version() == 1
    assert lib.sd_strerror(0) == b'ok'
    for rc in range(-8, 0):
        assert b'segdistill' in lib.sd_strerror(rc)


def test_workspace_sizes_are_positive_and_monotone(lib):
    a = lib.sd_kl_rows_workspace_bytes(2, 150, 4096, 1)
    b = lib.sd_kl_rows_workspace_bytes(16, 150, 16384, 10)
    assert 0 < a < b
    assert lib.sd_kl_rows_workspace_bytes(0, 150, 4096, 1) == 0
    assert lib.sd_kl_pixels_workspace_bytes(16, 150, 16384) > 16 * 16384 * 4
    assert lib.sd_mse_workspace_bytes(1 << 20) > 0
    # IFVD: class sums of S, T and the weighted pass (3 * B * 151^2 floats) + 4 floats per pixel, plus the
    # per-pixel-range partial sums; grows with the batch, zero for an empty problem
    small = lib.sd_ifvd_sim_workspace_bytes(2, 150, 16384)
    assert small > 4 * (3 * 2 * 151 * 151 + 4 * 2 * 16384)
    assert lib.sd_ifvd_sim_workspace_bytes(16, 150, 16384) > small
    assert lib.sd_ifvd_sim_workspace_bytes(0, 150, 16384) == 0


def test_argument_errors_are_reported_without_a_device(lib):
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    # NULL pointers
    assert lib.sd_kl_rows_fwd_bwd(None, p, p, None, p, None, 1, 1, 4, 1, 0, 1.0, 1.0, 1.0, 0.0, None,
                                  p, 1 << 20, 0, None) == -1
    # bad dtype
    assert lib.sd_kl_rows_fwd_bwd(p, p, p, None, p, None, 1, 1, 4, 1, 7, 1.0, 1.0, 1.0, 0.0, None,
                                  p, 1 << 20, 0, None) == -3
    # bad shape, bad tau, bad group
    assert lib.sd_kl_rows_fwd_bwd(p, p, p, None, p, None, 0, 1, 4, 1, 0, 1.0, 1.0, 1.0, 0.0, None,
                                  p, 1 << 20, 0, None) == -2
    assert lib.sd_kl_rows_fwd_bwd(p, p, p, None, p, None, 1, 1, 4, 1, 0, 0.0, 1.0, 1.0, 0.0, None,
                                  p, 1 << 20, 0, None) == -8
    assert lib.sd_kl_rows_fwd_bwd(p, p, p, None, p, None, 1, 1, 4, 0, 0, 1.0, 1.0, 1.0, 0.0, None,
                                  p, 1 << 20, 0, None) == -8
    # workspace too small
    assert lib.sd_kl_rows_fwd_bwd(p, p, p, None, p, None, 1, 1, 4, 1, 0, 1.0, 1.0, 1.0, 0.0, None,
                                  p, 16, 0, None) == -5
    # MSE asked for but no output slot
    assert lib.sd_kl_rows_fwd_bwd(p, p, p, None, p, None, 1, 1, 4, 1, 0, 1.0, 1.0, 1.0, 0.5, None,
                                  p, 1 << 20, 0, None) == -1
    assert lib.sd_kl_pixels_fwd_bwd(p, p, None, None, p, 1, 1, 4, 0, 1.0, 1.0, 1.0, 0.0, None,
                                    p, 1 << 20, 0, None) == -1
    assert lib.sd_mse_fwd_bwd(p, p, p, p, 0, 0, 1.0, 1.0, p, 1 << 20, None) == -2
    assert lib.sd_scale_grad(None, 4, 0, p, None) == -1
    # IFVD similarity term: NULL class map, bad dtype, empty shape, workspace too small; class map: NULL, bad size
    assert lib.sd_ifvd_sim_fwd_bwd(p, p, None, p, p, 1, 2, 4, 0, 10.0, 1.0, 0, p, 1 << 20, None) == -1
    assert lib.sd_ifvd_sim_fwd_bwd(p, p, p, p, p, 1, 2, 4, 5, 10.0, 1.0, 0, p, 1 << 20, None) == -3
    assert lib.sd_ifvd_sim_fwd_bwd(p, p, p, p, p, 1, 0, 4, 0, 10.0, 1.0, 0, p, 1 << 20, None) == -2
    assert lib.sd_ifvd_sim_fwd_bwd(p, p, p, p, p, 1, 2, 4, 0, 10.0, 1.0, 0, p, 16, None) == -5
    assert lib.sd_ifvd_class_map(None, p, 1, 4, 4, 2, 2, 3, None) == -1
    assert lib.sd_ifvd_class_map(p, p, 1, 4, 0, 2, 2, 3, None) == -2


def test_no_device_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    assert lib.sd_device_check() == -7
    import segdistill_b200 as sd
    with pytest.raises(_cabi.SegDistillError):
        sd.CDLoss()(torch.randn(1, 2, 4, 4, requires_grad=True), torch.randn(1, 2, 4, 4))
