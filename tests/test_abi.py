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
