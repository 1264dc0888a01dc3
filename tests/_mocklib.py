"""Dry-run stand-in for libmac_b200.so used by the CPU tests of the host-side plumbing: every entry point checks its
arguments against the prototype table (`mac_network_b200/_lib.py::PROTOTYPES`, itself checked against include/mac_b200.h)
-- arity, pointer-vs-scalar kinds, integer ranges -- records the call and returns MAC_OK without computing anything.
It lets the host code of the GPU-only paths (argument marshalling, buffer shapes, call order) run on a box without a GPU;
numerics are covered by the `-m gpu` tests."""
import ctypes

from mac_network_b200 import _lib


def _semantic_checks(name, args):
    """The cheap argument rules the C entry points enforce before launching (units.cu / backward.cu / encoder.cu), so that a
    host-side shape that the library would reject with MAC_ERR_INVALID / MAC_ERR_ALIGN fails here too."""
    if name == "mac_linear_fwd":
        k_segs, ldx, nseg, ldy, n_out = args[1], args[2], args[3], args[9], args[11]
        assert 1 <= nseg <= 4 and n_out % 4 == 0 and ldy % 4 == 0, (name, nseg, n_out, ldy)
        assert all(k_segs[i] % 4 == 0 and ldx[i] % 4 == 0 for i in range(nseg)), (name, list(k_segs), list(ldx))
    elif name == "mac_linear_bwd":
        k_segs, nseg, n_out = args[1], args[3], args[13]
        assert 1 <= nseg <= 4 and n_out % 4 == 0 and all(k_segs[i] % 4 == 0 for i in range(nseg)), (name, n_out)
    elif name == "mac_embed_fwd":
        assert args[11] % 4 == 0, (name, "E", args[11])
    elif name in ("mac_lstm_fwd", "mac_lstm_bwd"):
        h, ndir = args[-3], args[-2]
        assert h % 8 == 0 and ndir in (1, 2), (name, h, ndir)
    elif name in ("mac_im2col3x3", "mac_col2im3x3"):
        assert args[-2] % 4 == 0, (name, "C", args[-2])


class MockLib(object):
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name not in _lib.PROTOTYPES:
            raise AttributeError(name)
        restype, argtypes = _lib.PROTOTYPES[name]

        def fn(*args):
            assert len(args) == len(argtypes), "%s: %d arguments, prototype has %d" % (name, len(args), len(argtypes))
            for i, (a, t) in enumerate(zip(args, argtypes)):
                where = "%s argument %d" % (name, i)
                if t is ctypes.c_void_p:
                    assert a is None or isinstance(a, (int, ctypes.c_void_p)) or hasattr(a, "_fields_") \
                        or type(a).__name__ == "CArgObject", (where, type(a))
                elif isinstance(t, type) and issubclass(t, ctypes._Pointer):
                    assert a is None or isinstance(a, ctypes.Array) or type(a).__name__ == "CArgObject", (where, type(a))
                elif t in (ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t, ctypes.c_uint64):
                    assert isinstance(a, int) and not isinstance(a, bool), (where, type(a))
                    if t is ctypes.c_int:
                        assert -2 ** 31 <= a < 2 ** 31, (where, a)
                    if t in (ctypes.c_size_t, ctypes.c_uint64):
                        assert a >= 0, (where, a)
                elif t is ctypes.c_float:
                    assert isinstance(a, (int, float)), (where, type(a))
                else:
                    raise AssertionError("unhandled prototype type %r in %s" % (t, name))
            _semantic_checks(name, args)
            self.calls.append(name)
            if restype is ctypes.c_size_t:
                return 1 << 16
            if restype is ctypes.c_char_p:
                return b"mock"
            if name == "mac_b200_abi_version":
                return 1
            return 0
        return fn


def install(monkeypatch):
    """Route `_lib.load()` to a MockLib and the stream handle to NULL in every host module."""
    import importlib
    mock = MockLib()
    monkeypatch.setattr(_lib, "load", lambda: mock)
    for mod in ("encoder", "stem", "output_unit", "dp", "mac_cell", "autograd", "tape"):
        m = importlib.import_module("mac_network_b200." + mod)
        if hasattr(m, "stream_ptr"):
            monkeypatch.setattr(m, "stream_ptr", lambda: None)
    return mock
