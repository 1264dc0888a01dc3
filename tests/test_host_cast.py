"""Host-side fp32 -> bf16 cast of the inference front end (no GPU): bit-exact against torch's CPU conversion."""
import ctypes

import numpy as np
import torch

from mac_network_b200 import _lib


def test_host_cast_bf16_is_round_to_nearest_even_and_threaded():
    lib = _lib.load()
    rng = np.random.RandomState(3)
    x = rng.standard_normal(1 << 20).astype(np.float32) * np.float32(10.0) ** rng.randint(-20, 20, 1 << 20).astype(np.float32)
    # ties, signed zeros, infinities, denormals, the largest finite value (rounds to inf), a NaN
    special = np.array([0.0, -0.0, np.inf, -np.inf, 1.0 + 2 ** -8, 1.0 + 3 * 2 ** -8, 1e-40, -1e-40, 3.4028235e38, np.nan],
                       np.float32)
    x[:special.size] = special
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy()
    for threads in (1, 3, 8):
        out = np.empty(x.size, np.int16)
        st = lib.mac_host_cast_bf16(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), x.size, threads)
        assert st == 0
        finite = ~np.isnan(x)
        assert np.array_equal(out[finite], ref[finite]), threads
        nan_bits = out[~finite].view(np.uint16)
        assert np.all((nan_bits & 0x7f80) == 0x7f80) and np.all(nan_bits & 0x007f)      # still a NaN
    assert lib.mac_host_cast_bf16(None, None, 4, 1) != 0
