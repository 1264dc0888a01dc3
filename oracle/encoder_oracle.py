"""ORACLE (test infrastructure only): numpy restatement of the reference's question input unit
(SURVEY.md section 8(f) rank 3).  Never imported by the product path.

  * `model.py:208-220`   qEmbeddingsOp: `embeddings = concat([zeros(1, wrdEmbDim), emb])`, `questions = embeddings[qIndices]`
                         (index 0 is the padding row, index i > 0 is row i-1 of the variable `qEmbeddings/emb`)
  * `model.py:279-307`   encoder: `ops.RNNLayer` (bi-LSTM, hDim = encDim/2 per direction), dropout on the question vector
                         (`qDropout`), optional `projCW` / `projQ` linears when `encProj` or `encDim != ctrlDim`
  * `ops.py:859-905`     biRNNLayer: plain dropout on the input sequence (`encInputDropout`), `BasicLSTMCell` fw / bw,
                         `tf.nn.bidirectional_dynamic_rnn(sequence_length=questionLengths)`, outputs concatenated
                         `[fw, bw]`, final state = `[h_fw(last valid step), h_bw(after step 0)]`
  * TensorFlow 1.x (not vendored; published semantics restated): `BasicLSTMCell` -- kernel `[in + h, 4h]`, gate order
    i, j, f, o, `forget_bias = 1.0`, `c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j)`, `h' = tanh(c')*sigmoid(o)`;
    `dynamic_rnn` zeroes outputs and carries the state through for t >= length; the backward direction runs on
    `reverse_sequence(x, lengths)` and its outputs are reversed back.

Pinned by `tests/golden/encoder_*.npz`: the reference's own `MACnet.qEmbeddingsOp` + `MACnet.encoder` run on the TF1 shim
(`oracle/gen_golden.py`), whose LSTM/dynamic_rnn restatement is written independently of this file (time-major loop over
reversed copies there, per-row index arithmetic here)."""
import numpy as np

ENC = "encoder/birnnLayer/bidirectional_rnn/"
ENC_UNI = "encoder/rnnLayer/rnn/"      # ops.fwRNNLayer (`encBi` off, ops.py:797-829): scope "rnnLayer", dynamic_rnn's "rnn"


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def embed(emb, qIndices, dtype=np.float64):
    """model.py:208-220."""
    table = np.concatenate([np.zeros((1, emb.shape[1]), dtype), np.asarray(emb, dtype)], axis=0)
    return table[np.asarray(qIndices).astype(np.int64)]


def lstm_direction(x, lengths, kernel, bias, reverse, forget_bias=1.0):
    """One direction of the bi-LSTM over x [B, S, E]; returns (outputs [B, S, h], final h [B, h], saved per-step data)."""
    B, S, _ = x.shape
    h_dim = kernel.shape[1] // 4
    out = np.zeros((B, S, h_dim), x.dtype)
    c = np.zeros((B, h_dim), x.dtype)
    h = np.zeros((B, h_dim), x.dtype)
    lengths = np.asarray(lengths).astype(np.int64)
    for s in range(S):
        live = s < lengths                                           # rows still inside their question
        t = np.where(reverse, lengths - 1 - s, s)                    # time index this row reads / writes at step s
        t = np.where(live, t, 0)
        xt = x[np.arange(B), t]
        g = np.concatenate([xt, h], axis=1) @ kernel + bias
        i, j, f, o = np.split(g, 4, axis=1)
        c_new = c * _sigmoid(f + forget_bias) + _sigmoid(i) * np.tanh(j)
        h_new = np.tanh(c_new) * _sigmoid(o)
        c = np.where(live[:, None], c_new, c)
        h = np.where(live[:, None], h_new, h)
        rows = np.nonzero(live)[0]
        out[rows, t[rows]] = h_new[rows]
    return out, h


def encoder_forward(params, qIndices, questionLengths, keep_input=1.0, keep_question=1.0, uniforms=None,
                    proj=False, proj_q_act="NON", dtype=np.float64):
    """-> dict(questionWords [B,S,E], questionCntxWords [B,S,encDim or ctrlDim], vecQuestions [B,encDim or ctrlDim])."""
    p = {k: np.asarray(v, dtype) for k, v in params.items()}
    us = iter(uniforms or [])

    def dropout(x, keep):
        if float(keep) == 1.0:
            return x
        return x / dtype(keep) * np.floor(dtype(keep) + np.asarray(next(us), dtype))

    words = embed(p["qEmbeddings/emb"], qIndices, dtype)
    x = dropout(words, keep_input)                                                      # ops.py:877
    if ENC_UNI + "basic_lstm_cell/kernel" in p:                                          # ops.py:797-829, hDim = encDim
        cntx, vecq = lstm_direction(x, questionLengths, p[ENC_UNI + "basic_lstm_cell/kernel"],
                                    p[ENC_UNI + "basic_lstm_cell/bias"], False)
    else:
        fw, h_fw = lstm_direction(x, questionLengths, p[ENC + "fw/basic_lstm_cell/kernel"], p[ENC + "fw/basic_lstm_cell/bias"], False)
        bw, h_bw = lstm_direction(x, questionLengths, p[ENC + "bw/basic_lstm_cell/kernel"], p[ENC + "bw/basic_lstm_cell/bias"], True)
        cntx = np.concatenate([fw, bw], axis=-1)                                        # ops.py:897
        vecq = np.concatenate([h_fw, h_bw], axis=-1)                                    # ops.py:898
    vecq = dropout(vecq, keep_question)                                                 # model.py:297
    if proj:                                                                            # model.py:300-305
        cntx = cntx @ p["encoder/linearLayerprojCW/weights/weight"] + p["encoder/linearLayerprojCW/biases/bias"]
        vecq = vecq @ p["encoder/linearLayerprojQ/weights/weight"] + p["encoder/linearLayerprojQ/biases/bias"]
        if proj_q_act != "NON":          # ops.linear would add the activation AND a nested "projQ_2" layer (ops.py:325-328)
            raise NotImplementedError("encProjQAct != NON (config.py:270 default) is not restated")
    return {"questionWords": words, "questionCntxWords": cntx, "vecQuestions": vecq}
