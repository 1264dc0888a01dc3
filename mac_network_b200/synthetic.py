"""Seeded synthetic CLEVR-shaped cell inputs.

There is no dataset offline; the cell is fed tensors with the distributions of what the reference's
encoder/stem hand it: the KB is a post-stem ELU activation (`ops.py:403`, `model.py:186-190`), the
contextual words / question vector are bounded LSTM outputs, raw word embeddings are U(-1,1)
(`preprocess.py:585-589`), lengths are U{S/2..S} with max == S (the batch is trimmed to its longest
question, `model.py:681-687`).  numpy legacy RandomState => identical streams on every box.
"""
import numpy as np

# name -> (B, S, N, d, L): BASELINE.json configs (N = H*W KB cells)
SHAPES = {
    "cpu_ref": (32, 20, 196, 512, 4),      # configs[0]/[1]
    "headline": (64, 40, 196, 512, 12),    # configs[2]/[3]
    "gqa": (64, 30, 49, 512, 6),           # configs[4]
}


def make_inputs(B, S, N, d, seed=1234, dtype=np.float32):
    rng = np.random.RandomState(seed)
    kb = rng.standard_normal((B, N, d))
    kb = np.where(kb > 0, kb, np.expm1(np.minimum(kb, 0)))
    cntx = 0.5 * np.tanh(rng.standard_normal((B, S, d)))
    vecq = 0.5 * np.tanh(rng.standard_normal((B, d)))
    words = rng.uniform(-1.0, 1.0, size=(B, S, d))
    lo = max(1, S // 2)
    lengths = rng.randint(lo, S + 1, size=(B,)).astype(np.int32)
    lengths[rng.randint(0, B)] = S
    return {
        "vecQuestions": vecq.astype(dtype),
        "questionWords": words.astype(dtype),
        "questionCntxWords": cntx.astype(dtype),
        "questionLengths": lengths,
        "knowledgeBase": kb.astype(dtype),
    }
