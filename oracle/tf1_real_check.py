"""ORACLE (test infrastructure only): pin the golden fixtures against a REAL TensorFlow-1.x runtime, when one exists.

    python oracle/tf1_real_check.py [case ...]        # needs TensorFlow 1.x AND /root/reference; neither is in this image

SURVEY.md section 8(c)(iv).  `oracle/gen_golden.py` runs the unmodified reference on a numpy stand-in for TensorFlow, which
pins the reference's graph (op order, scopes, variable names, where dropout is applied) but restates TF's kernels from their
published definitions.  This script closes that last gap wherever a TF-1 runtime is available (e.g. a cp37 container with
`tensorflow==1.15`): it builds the SAME graph with the real library -- the reference's own `MACCell` on placeholders, driven
by the reference's own `parseArgs()` -- loads the fixture's parameter values into the variables the reference creates (matched
by name), runs one `sess.run`, and compares control / memory / attention maps with `tests/golden/<case>.npz`.  Evaluation-mode
cases only (dropout = 1.0 is the identity in TF, so the comparison is deterministic up to fp32 reduction order).

It has NOT been executed in the build container (TensorFlow is not installable offline; the import below fails there with a
clear message).  Nothing under `mac_network_b200/` imports it."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
TOL = 2e-5                      # fp32 TF kernels vs the float64 fixtures, max-abs relative to the tensor's max-abs


def main():
    try:
        import tensorflow as tf
    except ImportError:
        sys.exit("tf1_real_check: TensorFlow is not installed here; run this where a TensorFlow 1.x runtime exists")
    if not tf.__version__.startswith("1."):
        sys.exit("tf1_real_check: needs TensorFlow 1.x (tf.contrib, tf.placeholder); found %s" % tf.__version__)
    if not os.path.isdir(REF):
        sys.exit("tf1_real_check: %s not found" % REF)
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import config as ref_config                       # the reference's module-global config (config.py:92)
    from mac_cell import MACCell                      # the reference's cell, unmodified
    from tests._util import load_golden, rebuild, golden_cases

    cases = sys.argv[1:] or [c for c in golden_cases() if "train" not in c]
    worst = {}
    for case in cases:
        meta, gold = load_golden(case)
        if meta["train"]:
            print("%-22s skipped (training-mode fixtures depend on the recorded uniform draws)" % case)
            continue
        argv = [a if not a.startswith("@configs/") else "@" + os.path.join(REF, a[1:]) for a in meta["argv"]]
        for k in list(vars(ref_config.config).keys()):
            delattr(ref_config.config, k)
        old, sys.argv = sys.argv, ["tf1_real_check"] + argv
        try:
            ref_config.parseArgs()
        finally:
            sys.argv = old
        cfg, inputs, params = rebuild(meta, dtype=np.float32)
        sh = meta["shape"]
        B, L = sh["B"], sh["L"]
        tf.reset_default_graph()
        ph = {k: tf.placeholder(tf.int32 if v.dtype == np.int32 else tf.float32, shape=v.shape, name=k)
              for k, v in inputs.items()}
        with tf.variable_scope("MACnetwork"):                                          # model.py:431
            cell = MACCell(vecQuestions=ph["vecQuestions"], questionWords=ph["questionWords"],
                           questionCntxWords=ph["questionCntxWords"], questionLengths=ph["questionLengths"],
                           knowledgeBase=ph["knowledgeBase"], memoryDropout=1.0, readDropout=1.0, writeDropout=1.0,
                           batchSize=B, train=False, reuse=None)
            state = cell.zero_state(B, tf.float32)                                     # model.py:447
            none = tf.zeros((B, 1), dtype=tf.float32)
            for i in range(L):                                                         # model.py:453-458
                cell.iteration = i
                _, state = cell(none, state)
        fetch = {"control": cell.controls, "memory": cell.memories, "att_question": tf.stack(cell.attentions["question"]),
                 "att_kb": tf.stack(cell.attentions["kb"])}
        assigns, created = [], {}
        for v in tf.global_variables():
            name = v.name.split(":")[0]
            created[name] = v.shape.as_list()
            if name not in params:
                sys.exit("%s: the reference created %s, which the fixture does not carry" % (case, name))
            assigns.append(tf.assign(v, params[name].reshape(created[name])))
        missing = sorted(set(params) - set(created))
        if missing:
            sys.exit("%s: fixture variables the reference did not create: %s" % (case, missing[:4]))
        with tf.Session(config=tf.ConfigProto(device_count={"GPU": 0})) as sess:       # the reference's CPU path
            sess.run(tf.global_variables_initializer())
            sess.run(assigns)
            out = sess.run(fetch, feed_dict={ph[k]: v for k, v in inputs.items()})
        got = {"control": out["control"][:, 1:].transpose(1, 0, 2), "memory": out["memory"][:, 1:].transpose(1, 0, 2),
               "att_question": out["att_question"], "att_kb": out["att_kb"]}
        errs = {k: float(np.max(np.abs(got[k] - gold[k])) / (np.max(np.abs(gold[k])) + 1e-30)) for k in got}
        worst[case] = errs
        ok = all(e < TOL for e in errs.values())
        print("%-22s %s  %s" % (case, "ok " if ok else "MISMATCH", json.dumps({k: float("%.2e" % e) for k, e in errs.items()})))
    bad = {c: e for c, e in worst.items() if any(x >= TOL for x in e.values())}
    if bad:
        sys.exit("tf1_real_check: fixtures disagree with TensorFlow: %s" % sorted(bad))
    print("tf1_real_check: %d fixtures agree with TensorFlow %s within %.0e" % (len(worst), tf.__version__, TOL))


if __name__ == "__main__":
    main()
