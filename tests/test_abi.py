"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mac_b200.h
declares (no compute calls without a GPU), and the ctypes prototypes cover exactly that set."""
import os
import re

from mac_network_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mac_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(mac_[a-z0-9_]+)\s*\(", src)) - {"mac_b200_count_launch_"}


def test_header_and_binding_agree():
    assert _declared() == set(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.mac_b200_abi_version() == 1
    assert b"workspace" in lib.mac_b200_strerror(-4)


def test_workspace_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.mac_read_workspace_bytes(64, 196, 512, 0) > 2 * 64 * 196 * 512 * 4
    assert lib.mac_write_workspace_bytes(64, 512) > 4096
    assert lib.mac_linear_workspace_bytes(64, 512, 512) > 4096


def test_no_product_import_of_oracle():
    """The product path must never route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "mac_network_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_argument_errors_are_status_codes_not_crashes():
    """Error behaviour of the boundary (include/mac_b200.h): bad sizes / null pointers / misalignment / short workspaces come
    back as negative MAC_ERR_* codes before any CUDA call is made, so this runs without a GPU."""
    import ctypes
    lib = _lib.load()
    INVALID, ALIGN, WORKSPACE = -1, -2, -4
    buf = (ctypes.c_float * 4096)()
    p = ctypes.addressof(buf)
    p = (p + 15) & ~15                                     # 16-byte aligned fake "device" pointer (never dereferenced)
    arr_p = (ctypes.c_void_p * 1)(p)
    one = lambda v: (ctypes.c_int * 1)(v)
    # ops.linear: K not a multiple of 4, null weight, misaligned output
    assert lib.mac_linear_fwd(arr_p, one(6), one(8), 1, p, None, 0.0, 0, p, 8, 4, 8, None, 0, None) == INVALID
    assert lib.mac_linear_fwd(arr_p, one(8), one(8), 1, None, None, 0.0, 0, p, 8, 4, 8, None, 0, None) == INVALID
    assert lib.mac_linear_fwd(arr_p, one(8), one(8), 1, p, None, 0.0, 0, p + 4, 8, 4, 8, None, 0, None) == ALIGN
    # dropout keep outside (0, 1]
    assert lib.mac_dropout_fwd(p, 0.0, 1, 0, 0, p, 16, None) == INVALID
    assert lib.mac_dropout_fwd(p, 1.5, 1, 0, 0, p, 16, None) == INVALID
    # question input unit: embedding width not a multiple of 4; hidden size not a multiple of 8; short workspace
    assert lib.mac_embed_fwd(p, p, 1.0, 1, 0, 0, None, p, 2, 3, 5, 6, None) == INVALID
    assert lib.mac_lstm_fwd(p, p, p, p, p, 1.0, p, None, None, None, None, p, 1 << 20, 2, 3, 12, 2, None) == INVALID
    assert lib.mac_lstm_fwd(p, None, p, None, p, 1.0, p, None, None, None, None, p, 1 << 20, 2, 3, 8, 2, None) == INVALID
    assert lib.mac_lstm_fwd(p, p, p, p, p, 1.0, p, None, None, None, None, p, 16, 2, 3, 8, 2, None) == WORKSPACE
    assert lib.mac_lstm_workspace_bytes(64, 256, 2) >= 3 * 2 * 64 * 256 * 4
    # stem: channel count not a multiple of 4
    assert lib.mac_im2col3x3(p, p, 0, 1.0, 1, 0, 0, 1, 3, 3, 6, None) == INVALID
    assert lib.mac_col2im3x3(p, p, 1.0, 1, 0, 0, 1, 3, 3, 6, None) == INVALID
    assert lib.mac_col2im3x3(p, p + 4, 1.0, 1, 0, 0, 1, 3, 3, 8, None) == ALIGN
    # loss / attention helpers
    assert lib.mac_kb_attend_bwd(None, p, p, p, p, None, None, 2, 3, 8, None) == INVALID
    for code in (INVALID, ALIGN, -3, WORKSPACE, -5):
        assert len(lib.mac_b200_strerror(code)) > 3
