"""CPU tests of the C-ABI boundary: the library builds (nvcc cross-compiles), loads, exports every symbol the header
declares, validates descriptors, and refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes

import pytest

from fastvideotagging_b200 import _lib, ops


def test_library_exports_every_declared_symbol(lib):
    names = _lib.header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert lib.fvt_version() >= 100


def test_ctypes_signatures_match_the_header(lib):
    """Every function the header declares is bound in _lib.load() with as many ctypes arguments as the C declaration has
    parameters (a drifted binding would pass garbage pointers to a kernel launch instead of failing here)."""
    import re
    with open(_lib.HEADER_PATH) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    decls = re.findall(r"\b(fvt_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(decls) >= 15
    for name, params in decls:
        params = params.strip()
        n_c = 0 if params in ("", "void") else len([q for q in params.split(",") if q.strip()])
        fn = getattr(lib, name)
        assert fn.argtypes is not None, "%s is exported but has no ctypes signature in _lib.load()" % name
        assert len(fn.argtypes) == n_c, (name, len(fn.argtypes), n_c)


def test_conv_descriptor_validation_and_shapes(lib):
    d = ops.conv_desc(2, 8, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert ops.conv_out_shape(d) == (8, 56, 56)
    d = ops.conv_desc(2, 8, 56, 56, 64, 240, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    assert ops.conv_out_shape(d) == (8, 28, 28)                      # floor((x+2p-k)/s)+1
    d = ops.conv_desc(2, 8, 56, 56, 240, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0))
    assert ops.conv_out_shape(d) == (4, 56, 56)
    d = ops.conv_desc(1, 8, 112, 56, 32, 48, (1, 7, 1), (1, 2, 1), (0, 3, 0))
    assert ops.conv_out_shape(d) == (8, 56, 56)
    for cout, bn in ((64, 64), (144, 144), (240, 240), (288, 144), (464, 240), (576, 192), (928, 240), (1152, 192), (512, 256)):
        d = ops.conv_desc(1, 4, 14, 14, 64, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        assert lib.fvt_conv3d_block_n(ctypes.byref(d)) == bn
        rows = (cout + bn - 1) // bn * bn
        assert lib.fvt_conv3d_packed_weight_elems(ctypes.byref(d)) == rows * 9 * 64
    bad = ops.conv_desc(1, 4, 14, 14, 60, 64, (1, 3, 3))
    with pytest.raises(_lib.FvtError, match="cin=60"):
        ops.conv_out_shape(bad)
    bad = ops.conv_desc(1, 4, 14, 14, 64, 64, (1, 3, 3), (1, 9, 1))
    with pytest.raises(_lib.FvtError, match="stride"):
        ops.conv_out_shape(bad)
    bad = ops.conv_desc(1, 1, 2, 2, 64, 64, (1, 7, 7))
    with pytest.raises(_lib.FvtError, match="larger than padded input"):
        ops.conv_out_shape(bad)


def test_philox_host_entry_point_matches_known_answers(lib):
    def run(c, k):
        cc = (ctypes.c_uint32 * 4)(*c)
        kk = (ctypes.c_uint32 * 2)(*k)
        out = (ctypes.c_uint32 * 4)()
        assert lib.fvt_philox4x32_10(cc, kk, out) == 0
        return list(out)
    assert run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    assert lib.fvt_device_check(0) < 0
    from fastvideotagging_b200.model import R2Plus2D, LsepLoss
    net = R2Plus2D(101, 18, final_spatial_kernel=7, final_temporal_kernel=1).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 8, 112, 112))
    with pytest.raises(_lib.FvtError, match="no CPU fallback"):
        LsepLoss()(torch.zeros(2, 4), torch.zeros(2, 4))
