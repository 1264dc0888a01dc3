"""TensorFlow checkpoint ("tensor bundle") interchange without TensorFlow (mac_network_b200/tf_bundle.py, checkpoint.py):
round trips, CRC-32C known answers, the on-disk structure (footer magic, block trailers, sorted keys, header entry), corruption
is detected, and a MAC parameter set travels through a real `weights{epoch}.ckpt` pair under the reference's variable names
(main.py:163-201).  TensorFlow is not installable in this image, so the format itself is restated from its public sources
("unpinned" against a TF-written file; see the module docstring)."""
import os
import struct

import numpy as np
import pytest

from mac_network_b200 import tf_bundle as tb


def test_crc32c_known_answers():
    assert tb.crc32c(b"") == 0
    assert tb.crc32c(b"123456789") == 0xE3069283                      # the CRC-32C check value
    assert tb.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720 B.4: 32 bytes of zeros
    assert tb.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                 # RFC 3720 B.4: 32 bytes of ones
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E                   # RFC 3720 B.4: incrementing
    big = np.random.RandomState(0).bytes(70000)                        # the C helper (>= 4096 bytes) against the table loop
    c = 0xFFFFFFFF
    for b in big:
        c = tb._TABLE_L[(c ^ b) & 0xFF] ^ (c >> 8)
    assert tb.crc32c(big) == (c ^ 0xFFFFFFFF)
    assert tb.mask_crc(0) == 0xA282EAD8


def _tensors():
    r = np.random.RandomState(3)
    return {"macModel/MACnetwork/MACCell/read/linearLayermemKbProj/weights/weight": r.randn(64, 32).astype(np.float32),
            "macModel/MACnetwork/MACCell/read/linearLayermemKbProj/biases/bias": r.randn(32).astype(np.float32),
            "macModel/MACnetwork/initMem": r.randn(32).astype(np.float32),
            "macModel/MACnetwork/initMem/ExponentialMovingAverage": r.randn(32).astype(np.float32),
            "macModel/MACnetwork/MACCell/control/inter2logits/linearLayerlogits/biases/bias": np.float32(0.25),     # 0-d
            "global_step": np.int64(7), "beta1_power": np.float32(0.9), "empty": np.zeros((3, 0, 2), np.float32),
            "macModel/qEmbeddings/emb": r.randn(90, 300).astype(np.float32)}


def test_round_trip_and_structure(tmp_path):
    prefix = str(tmp_path / "weights12.ckpt")
    tens = _tensors()
    names = tb.write_tensor_bundle(prefix, tens)
    assert names == sorted(tens, key=lambda s: s.encode())
    back = tb.read_tensor_bundle(prefix)
    assert set(back) == set(tens)
    for k, v in tens.items():
        v = np.asarray(v)
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack_from("<Q", idx, len(idx) - 8)[0] == tb.MAGIC and len(idx) > 48
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert len(data) == sum(np.asarray(v).nbytes for v in tens.values())          # raw bytes, no padding, key order
    first = np.asarray(tens[names[0]]).tobytes()
    assert data[:len(first)] == first
    sub = tb.read_tensor_bundle(prefix, names={"global_step"})
    assert list(sub) == ["global_step"] and int(sub["global_step"]) == 7


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "w.ckpt")
    tb.write_tensor_bundle(prefix, _tensors())
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[100] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tb.read_tensor_bundle(prefix)
    assert tb.read_tensor_bundle(prefix, verify=False)          # readable when checks are off
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        tb.read_tensor_bundle(prefix)
    with pytest.raises(ValueError):
        open(prefix + ".index", "wb").write(b"\x00" * 64)
        tb.read_tensor_bundle(prefix)


def test_many_entries_exercise_prefix_compression_and_restarts(tmp_path):
    prefix = str(tmp_path / "many.ckpt")
    tens = {"scope/layer%03d/weights/weight" % i: np.full((i % 5 + 1, 3), i, np.float32) for i in range(150)}
    tb.write_tensor_bundle(prefix, tens)
    back = tb.read_tensor_bundle(prefix)
    assert all(np.array_equal(back[k], v) for k, v in tens.items()) and len(back) == 150


def test_mac_parameters_travel_under_reference_names(tmp_path):
    """checkpoint.save_tf_checkpoint / load_tf_checkpoint: the cell's variables (+ EMA shadows) as a real checkpoint pair."""
    from mac_network_b200.checkpoint import load_tf_checkpoint, save_tf_checkpoint, MODEL_SCOPE, EMA_SUFFIX
    from mac_network_b200.config import MACConfig
    from mac_network_b200.params import init_params, perturb_biases
    cfg = MACConfig.args("gqa", netLength=3, memDim=16, ctrlDim=16, attDim=16)
    pv = perturb_biases(init_params(cfg, 3, seed=4), seed=5)
    ema = {k: 0.5 * v for k, v in pv.items()}
    prefix = str(tmp_path / "weights7.ckpt")
    names = save_tf_checkpoint(prefix, pv, ema_values=ema, extra={"global_step": np.int64(7)})
    assert MODEL_SCOPE + "MACnetwork/MACCell/write/linearLayergate/weights/weight" in names
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    back = load_tf_checkpoint(prefix)
    assert set(back) == set(pv) and all(np.array_equal(back[k], pv[k]) for k in pv)
    half = load_tf_checkpoint(prefix, use_ema=True)
    assert all(np.allclose(half[k], 0.5 * pv[k]) for k in pv)
    assert open(str(tmp_path / "checkpoint")).read().startswith('model_checkpoint_path: "weights7.ckpt"')
