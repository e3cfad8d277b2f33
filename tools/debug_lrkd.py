import torch, numpy as np
from tests.cases import build_case
from deltakd_b200 import functional as Fn
for name in ("lrkd_r32",):
    c = build_case(name, device="cuda")
    a = c.args
    basis = {}
    s_sel = [c.s_feats[0], c.s_feats[1], c.s_feats[-1]]
    t_sel = [c.t_feats[0], c.t_feats[1], c.t_feats[11]]
    kd = Fn.lrkd_layers_loss(s_sel, t_sel, list(c.student.align), a.lrkd_rank, (a.lrkd_alpha, a.lrkd_beta, a.lrkd_gamma), basis_out=basis)
    print("sweeps", basis["sweeps"].tolist())
    for j in range(3):
        V = basis["V"][j].double().cpu()     # [k, Dt]
        am = V.abs().argmax(1)
        vals = V.gather(1, am[:, None])[:, 0]
        bad = (vals <= 0).nonzero()[:, 0].tolist()
        print("layer", j, "bad rows", bad)
        for r in bad[:4]:
            top = V[r].abs().topk(3)
            print("   row", r, "top3 idx", top.indices.tolist(), "vals", V[r][top.indices].tolist())
