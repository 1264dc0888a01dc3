"""ORACLE / CPU BASELINE (test infrastructure only): fp32 PyTorch-CPU port of the reference cell at the TF graph's
op granularity (separate matmul / bias / broadcast-mul / concat / ELU / reduce_sum ops, nothing fused), for the
shipped flag family (args, args2, args3, args4 and their union).  It is the "TF1-CPU stand-in" that `bench.py` times
on the GPU box's host cores (TensorFlow itself cannot be installed offline) -- kind "port" in `cpu_baseline`.
`tests/test_oracle_golden.py::test_torch_cpu_port_matches_oracle` pins it to the numpy oracle.

Follows /root/reference: mac_cell.py:133-187 (control), 209-277 (read), 305-375 (write), 420-480 (step),
539-592 (zero_state); ops.py:50-59, 114-150, 243-247, 298-333, 668-725.  Eval mode (dropouts = 1.0, model.py:118-125).
"""
import torch

PREFIX = "MACnetwork/"


class TorchCPUCell(object):
    def __init__(self, cfg, params_np, L):
        self.cfg, self.L = cfg, L
        self.p = {k: torch.from_numpy(v.astype("float32")) for k, v in params_np.items()}

    def lin(self, x, scope, name):
        sc = PREFIX + scope + "linearLayer" + name + "/"
        W, b = self.p[sc + "weights/weight"], self.p[sc + "biases/bias"]
        if W.dim() == 2:
            # ops.multiply (ops.py:50-59): flatten to 2-D, matmul, reshape back; then the bias add (ops.py:319-320)
            y = torch.matmul(x.reshape(-1, W.shape[0]), W).reshape(*x.shape[:-1], W.shape[1])
            return y + b
        return torch.sum(x * W, dim=-1) + b                       # ops.py:316-317

    @torch.no_grad()
    def forward(self, vecQ, cntxWords, lengths, kb):
        c, L = self.cfg, self.L
        B, S, d = cntxWords.shape
        control = vecQ if c.initCtrl == "Q" else self.p[PREFIX + "initCtrl"].unsqueeze(0).repeat(B, 1)
        memory = self.p[PREFIX + "initMem"].unsqueeze(0).repeat(B, 1)
        controls, memories = control.unsqueeze(1), memory.unsqueeze(1)
        mask = (1 - (torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1)).float()) * (-1e30)   # ops.py:243-247
        atts = []
        for i in range(L):
            # control unit
            ci = torch.tanh(self.lin(vecQ, "MACCell/", "qInput"))
            ci = self.lin(ci, "MACCell/", ("qInput%d" % i) if c.controlInputUnshared else "qInputU")
            inter = ci.unsqueeze(1) * cntxWords                                                  # mac_cell.py:155
            logits = self.lin(inter, "MACCell/control/inter2logits/", "logits")
            qatt = torch.softmax(logits + mask, dim=-1)
            control = torch.sum(qatt.unsqueeze(-1) * cntxWords, dim=-2)                          # ops.py:149-150
            # read unit
            P = self.lin(kb, "MACCell/read/mulmemInter/", "projX")
            y = self.lin(memory, "MACCell/read/mulmemInter/", "projY")
            yb = torch.zeros_like(P) + y.unsqueeze(-2)                                           # ops.py:694-697
            I0 = torch.cat([P * yb, P], dim=-1)                                                  # ops.py:700-719
            H = torch.nn.functional.elu(self.lin(I0, "MACCell/read/", "memKbProj"))
            I1 = self.lin(H, "MACCell/read/linearLayermemKbProj/", "memKbProj_2")
            cb = torch.zeros_like(I1) + control.unsqueeze(-2)
            I2 = torch.nn.functional.elu(I1 * cb)                                                # mac_cell.py:248-262
            katt = torch.softmax(self.lin(I2, "MACCell/read/inter2att/inter2logits/", "logits"), dim=-1)
            info = torch.sum(katt.unsqueeze(-1) * kb, dim=-2)
            # write unit
            parts = [memory, info]
            if c.writeSelfAtt:
                sc = self.lin(ci if c.writeSelfAttMod == "CONT" else control, "MACCell/write/", "ctrlProj")
                sint = controls * sc.unsqueeze(1)
                satt = torch.softmax(self.lin(sint, "MACCell/write/inter2attselfAttention/inter2logits/", "logits"), -1)
                parts.append(torch.sum(satt.unsqueeze(-1) * memories, dim=-2))
            new_mem = self.lin(torch.cat(parts, dim=-1), "MACCell/write/", "newMemory")
            if c.writeGate:
                z = torch.sigmoid(self.lin(control, "MACCell/write/", "gate") + c.writeGateBias)
                new_mem = new_mem * z + memory * (1 - z)
            memory = new_mem
            controls = torch.cat([controls, control.unsqueeze(1)], dim=1)
            memories = torch.cat([memories, memory.unsqueeze(1)], dim=1)
            atts.append((qatt, katt))
        return control, memory, atts
