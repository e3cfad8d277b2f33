"""GPU: saliency scores (3 methods), score -> mask selection, and the saliency-MGD loss vs oracle / reference goldens.

Bit-exact mask selection is defined GIVEN the score tensor: feeding the reference's recorded scores to
dkd_mask_rank must reproduce the reference's mask and ids_restore exactly; our own scores match the reference's to
fp32 rounding, which at these sizes also yields the identical mask (asserted, with the score margin reported)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import losses as O
from oracle.util import digest, rel_err
from tests.cases import build_case
from deltakd_b200 import heads as H

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4
SCORE_RTOL = 2e-5


def _student(method, device="cuda"):
    from deltakd_b200 import synth
    args = SimpleNamespace(distillation_type="saliency_mgd", saliency_method=method)
    teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
    torch.manual_seed(0)
    H.attach_distillation_heads(student, teacher, args)
    return student.to(device)


@pytest.mark.parametrize("m", [1, 2, 3])
def test_scores_match_reference(golden, m):
    from deltakd_b200 import synth
    from deltakd_b200.misc import saliency_scores, saliency_masking
    from deltakd_b200 import functional as Fn
    student = _student(m)
    _, t_feats = synth.make_features(3, 77, layers=[11])
    t = t_feats[11].cuda()
    score = saliency_scores(student, t, m)
    ref = torch.from_numpy(golden[f"saliency_masking/m{m}/score"])
    assert score.shape == ref.shape and score.dtype == torch.float32
    assert rel_err(score, ref) < SCORE_RTOL
    heads64 = {k: v.detach().double().cpu() for k, v in H.head_tensors(student).items()}
    o = O.saliency_score(m, t_feats[11].double(), heads64)
    assert rel_err(score, o) < SCORE_RTOL
    assert abs(float(score.sum(1).mean()) - float(o.sum(1).mean())) < 1e-5
    # the reference's scores -> the reference's mask, bit for bit
    lk = O.len_keep_of(196, 0.5)
    mask, ids_restore, _ = Fn.mask_rank(ref.cuda(), lk)
    assert np.array_equal(mask.cpu().numpy(), golden[f"saliency_masking/m{m}/mask"])
    assert np.array_equal(ids_restore.cpu().numpy(), golden[f"saliency_masking/m{m}/ids_restore"])
    # our scores: same mask unless two scores straddling the keep boundary are closer than our rounding
    s_feat = torch.randn(3, 196, 192, generator=torch.Generator().manual_seed(1)).cuda()
    x_keep, mask2, ids2 = saliency_masking(student, t, s_feat, 0.5, m)
    srt = np.sort(ref.numpy(), axis=1)
    margin = float(((srt[:, lk] - srt[:, lk - 1]) / srt[:, lk]).min())
    if margin > 10 * SCORE_RTOL:
        assert np.array_equal(mask2.cpu().numpy(), golden[f"saliency_masking/m{m}/mask"]), margin
    assert x_keep.shape == (3, lk, 192) and int(mask2.sum()) == 3 * (196 - lk)
    keep_idx = torch.argsort(ids2, dim=1)[:, :lk]
    assert torch.equal(x_keep, torch.gather(s_feat, 1, keep_idx.unsqueeze(-1).expand(-1, -1, 192)))


@pytest.mark.parametrize("m", [1, 2, 3])
@pytest.mark.parametrize("B,dtype", [(1, torch.float32), (37, torch.float32), (5, torch.bfloat16)])
def test_scores_shapes_dtypes(m, B, dtype):
    from deltakd_b200 import synth
    from deltakd_b200.misc import saliency_scores
    student = _student(m)
    _, t_feats = synth.make_features(B, 5, layers=[11])
    t = t_feats[11].to(dtype).cuda()
    score = saliency_scores(student, t, m)
    heads64 = {k: v.detach().double().cpu() for k, v in H.head_tensors(student).items()}
    o = O.saliency_score(m, t.double().cpu(), heads64)
    assert score.shape == (B, 196)
    assert rel_err(score, o) < (SCORE_RTOL if dtype == torch.float32 else 2e-2)  # bf16: one bf16 pass for q.k
    if m == 3:  # softmax over the patches only: every row sums to 1
        assert float((score.sum(1) - 1).abs().max()) < 1e-5
    assert float(score.min()) > 0


@pytest.mark.parametrize("name", ["salmgd_m1", "salmgd_m2", "salmgd_m3", "salmgd_m1_r07"])
def test_saliency_mgd_matches_reference(golden, name):
    from deltakd_b200 import DistillationLoss, call_base_loss
    c = build_case(name, device="cuda")
    crit = DistillationLoss(call_base_loss(c.args), c.teacher, c.kind, c.alpha, c.tau)
    loss = crit(torch.zeros(c.B, 3, 2, 2, device="cuda"), c.outputs, c.student, c.s_feats, c.labels, c.args)
    loss.backward()
    for tag in ("f32", "f64"):
        ref = float(golden[f"{name}/{tag}/loss"])
        assert abs(loss.item() - ref) <= LOSS_RTOL * abs(ref), (tag, loss.item(), ref)
    heads = H.head_tensors(c.student)
    from tests.gate_search import case_gate_oracle
    ol, grads, ours, n_amb, n_flip = case_gate_oracle(name, c)
    assert abs(loss.item() - ol.item()) <= LOSS_RTOL * abs(ol.item())
    checked = 0
    for k, g in grads.items():
        if float(g.abs().sum()) == 0:
            continue
        assert k in ours, k
        assert rel_err(ours[k], g) < GRAD_RTOL, (k, n_amb, n_flip)
        checked += 1
    assert checked >= 8
    for k, p in heads.items():
        if k.startswith("saliency_attn"):
            assert p.grad is None  # no gradient reaches the scorer
    if n_flip == 0:
        for k, p in heads.items():
            key = f"{name}/f64/g_head/{k}"
            if key in golden.files and p.grad is not None:
                assert rel_err(digest(p.grad), golden[key]) < GRAD_RTOL, key
