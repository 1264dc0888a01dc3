"""ORACLE (test infrastructure only): differentiable fp64 PyTorch restatement of the cell for the shipped flag family,
used to check the hand-written backward kernels against `torch.autograd` (the reference itself uses TF autodiff,
model.py:626-636).  Forward is pinned to the numpy oracle in tests/test_oracle_golden.py.  Dropout consumes the same
uniform draws, in the reference's call order, as `MACOracle` does."""
import torch

PREFIX = "MACnetwork/"


def run(cfg, params_np, inputs_np, L, dropouts=(1.0, 1.0, 1.0), uniforms=None, d_control=None, d_memory=None):
    """Returns (control_L, memory_L, grads) with grads keyed like the product's `mac_backward` output."""
    c = cfg
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    p = {k: t64(v).requires_grad_(True) for k, v in params_np.items()}
    vecQ = t64(inputs_np["vecQuestions"]).requires_grad_(True)
    words = t64(inputs_np["questionCntxWords"]).requires_grad_(True)
    kb = t64(inputs_np["knowledgeBase"]).requires_grad_(True)
    lengths = torch.tensor(inputs_np["questionLengths"]).long()
    us = iter(uniforms or [])
    km, kr, kw = dropouts

    def lin(x, scope, name, bias=0.0):
        sc = PREFIX + scope + "linearLayer" + name + "/"
        W, b = p[sc + "weights/weight"], p[sc + "biases/bias"]
        if W.dim() == 2:
            return x @ W + b + bias
        return (x * W).sum(-1) + b + bias

    def dropout(x, keep):
        if keep == 1.0:
            return x
        u = t64(next(us))
        return x / keep * torch.floor(keep + u)

    B, S, d = words.shape
    control = vecQ if c.initCtrl == "Q" else (p[PREFIX + "initCtrl"].unsqueeze(0).repeat(B, 1) if c.initCtrl == "PRM"
                                              else torch.zeros(B, d, dtype=torch.float64))
    memory = vecQ if c.initMem == "Q" else (p[PREFIX + "initMem"].unsqueeze(0).repeat(B, 1) if c.initMem == "PRM"
                                            else torch.zeros(B, d, dtype=torch.float64))
    controls, memories = control.unsqueeze(1), memory.unsqueeze(1)
    cont_prev = control
    var_mask = None
    if c.memoryVariationalDropout and km < 1.0:
        var_mask = torch.floor(km + t64(next(us)))
    mask = (1 - (torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1)).double()) * (-1e30)
    act_in = {"TANH": torch.tanh, "NON": lambda x: x, "RELU": torch.nn.functional.elu}[c.controlInputAct]
    for i in range(L):
        ci = act_in(lin(vecQ, "MACCell/", "qInput"))
        ci = lin(ci, "MACCell/", ("qInput%d" % i) if c.controlInputUnshared else "qInputU")
        cc = ci
        if c.controlFeedPrev:                                                    # mac_cell.py:141-151
            prev = control if c.controlFeedPrevAtt else cont_prev
            xin = torch.cat([prev, ci], dim=-1) if c.controlFeedInputs else prev
            cc = lin(xin, "MACCell/control/", "contControl")
            if c.controlContAct != "NON":
                cc = torch.tanh(cc) if c.controlContAct == "TANH" else torch.nn.functional.elu(cc)
                cc = lin(cc, "MACCell/control/linearLayercontControl/", "contControl_2")
        cont_prev = cc
        ci_for_selfatt = cc
        logits = lin(cc.unsqueeze(1) * words, "MACCell/control/inter2logits/", "logits")
        qatt = torch.softmax(logits + mask, dim=-1)
        control = (qatt.unsqueeze(-1) * words).sum(-2)
        # read
        if c.memoryVariationalDropout:
            m_in = memory / km * var_mask if km < 1.0 else memory
        else:
            m_in = dropout(memory, km)
        Kd = dropout(kb, kr)
        md = dropout(m_in, kr)
        P = lin(Kd, "MACCell/read/mulmemInter/", "projX")
        y = lin(md, "MACCell/read/mulmemInter/", "projY")
        I0 = torch.cat([P * y.unsqueeze(-2), P], dim=-1)
        H = torch.nn.functional.elu(lin(I0, "MACCell/read/", "memKbProj"))
        I1 = lin(H, "MACCell/read/linearLayermemKbProj/", "memKbProj_2")
        I2 = torch.nn.functional.elu(I1 * control.unsqueeze(-2))
        katt = torch.softmax(lin(dropout(I2, kr), "MACCell/read/inter2att/inter2logits/", "logits"), dim=-1)
        info = (katt.unsqueeze(-1) * kb).sum(-2)
        if c.writeDropout < 1.0:
            info = dropout(info, kw)
        parts = [memory, info]
        if c.writeSelfAtt:
            sc = lin(ci_for_selfatt if c.writeSelfAttMod == "CONT" else control, "MACCell/write/", "ctrlProj")
            satt = torch.softmax(lin(controls * sc.unsqueeze(1), "MACCell/write/inter2attselfAttention/inter2logits/",
                                     "logits"), -1)
            parts.append((satt.unsqueeze(-1) * memories).sum(-2))
        new_mem = lin(torch.cat(parts, dim=-1), "MACCell/write/", "newMemory")
        if c.writeGate:
            z = torch.sigmoid(lin(control, "MACCell/write/", "gate", bias=c.writeGateBias))
            new_mem = new_mem * z + memory * (1 - z)
        memory = new_mem
        controls = torch.cat([controls, control.unsqueeze(1)], dim=1)
        memories = torch.cat([memories, memory.unsqueeze(1)], dim=1)
    loss = 0.0
    if d_control is not None:
        loss = loss + (control * t64(d_control)).sum()
    if d_memory is not None:
        loss = loss + (memory * t64(d_memory)).sum()
    grads = {}
    if d_control is not None or d_memory is not None:
        loss.backward()
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).numpy() for k, v in p.items()}
        grads["knowledgeBase"] = kb.grad.numpy()
        grads["questionCntxWords"] = words.grad.numpy()
        grads["vecQuestions"] = vecQ.grad.numpy()
    return control.detach().numpy(), memory.detach().numpy(), grads
