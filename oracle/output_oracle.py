"""ORACLE (test infrastructure only): numpy restatement of the reference's output unit, classifier and answer loss
 for the shipped flag family (`outQuestion` on; outImage / answerMod / outputBN off).

  * `model.py:512-528`  outputOp:    features = [memory, linear_outQuestion(vecQuestions)]
  * `model.py:547-576`  classifier:  ops.FCLayer(features, [2*memDim] + outClassifierDims + [answerWordsNum], dropout)
  * `ops.py:349-359`    FCLayer:     linear "fc_i" with input dropout, `act` (RELU -> config.relu) between layers
  * `model.py:593-596`  addAnswerLossOp: mean sparse softmax cross entropy
Pinned by `tests/golden/output_*.npz`, produced by running the reference's own `MACnet.outputOp/classifier/addAnswerLossOp`
on the TF1 shim (`oracle/gen_golden.py`)."""
import numpy as np

from oracle.mac_oracle import elu

OUT_PREFIX = ""       # these variables live directly under "macModel/" (model.py:774), beside "MACnetwork/"


def output_forward(cfg_relu, params, memory, vecQuestions, answers, keep=1.0, uniforms=None, dtype=np.float64):
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    us = iter(uniforms or [])

    def dropout(x):
        if float(keep) == 1.0:
            return x
        u = np.asarray(next(us), dtype)
        return x / dtype(keep) * np.floor(dtype(keep) + u)

    def lin(x, scope, drop=False):
        return (dropout(x) if drop else x) @ p[scope + "weights/weight"] + p[scope + "biases/bias"]
    eq = lin(np.asarray(vecQuestions, dtype), "outputUnit/linearLayeroutQuestion/")
    feats = np.concatenate([np.asarray(memory, dtype), eq], axis=-1)                     # ops.concat (ops.py:65-78)
    nfc = len([k for k in p if k.startswith("classifier/linearLayerfc_") and k.endswith("weights/weight")])
    x = feats
    for i in range(nfc):
        x = lin(x, "classifier/linearLayerfc_%d/" % i, drop=True)
        if i < nfc - 1:
            x = elu(x) if cfg_relu == "ELU" else np.maximum(x, 0)
    logits = x
    m = logits.max(-1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(logits - m).sum(-1))
    losses = lse - logits[np.arange(logits.shape[0]), np.asarray(answers)]
    return {"logits": logits, "losses": losses, "loss": losses.mean(), "preds": logits.argmax(-1)}
