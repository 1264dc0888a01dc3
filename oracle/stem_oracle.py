"""ORACLE (test infrastructure only): numpy restatement of the reference's image stem.

  * `model.py:165-204`  stem: `ops.CNNLayer(images[B,H,W,C], [inDim, stemDim.., outDim], dropout=stem dropout)`, then
                        reshape to the knowledge base `[B, H*W, outDim]`
  * `ops.py:380-405`    cnn: dropout on the layer INPUT, conv2d 3x3 stride 1 SAME with an HWIO kernel, + bias, activation
  * `ops.py:423-438`    CNNLayer: activation (RELU -> config.relu) after EVERY layer, including the last
Pinned by `tests/golden/stem_*.npz` (the reference's own `ops.CNNLayer` on the TF1 shim)."""
import numpy as np

from oracle.mac_oracle import elu


def stem_forward(relu, params, images, keep=1.0, uniforms=None, dtype=np.float64):
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    us = iter(uniforms or [])
    x = np.asarray(images, dtype)
    nlayers = len([k for k in p if k.endswith("kernels/kernel")])
    for i in range(nlayers):
        K = p["stem/cnnLayercnn_%d/kernels/kernel" % i]          # [kh, kw, cin, cout]
        b = p["stem/cnnLayercnn_%d/biases/bias" % i]
        if float(keep) != 1.0:
            x = x / dtype(keep) * np.floor(dtype(keep) + np.asarray(next(us), dtype))
        B, H, W, C = x.shape
        kh, kw = K.shape[:2]
        # im2col formulation (independent of the shim's shift-and-add): SAME padding, tap-major then channel columns
        xp = np.zeros((B, H + kh - 1, W + kw - 1, C), dtype=dtype)
        xp[:, (kh - 1) // 2:(kh - 1) // 2 + H, (kw - 1) // 2:(kw - 1) // 2 + W] = x
        cols = np.concatenate([xp[:, i_:i_ + H, j_:j_ + W, :] for i_ in range(kh) for j_ in range(kw)], axis=-1)
        y = cols.reshape(B * H * W, kh * kw * C) @ K.reshape(kh * kw * C, -1) + b
        x = (elu(y) if relu == "ELU" else np.maximum(y, 0)).reshape(B, H, W, -1)
    return x.reshape(x.shape[0], -1, x.shape[-1])
