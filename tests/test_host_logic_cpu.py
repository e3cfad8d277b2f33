"""CPU: host-side logic around the criterion — selected-layer capture, Mixup's host draws and label
object (rank 4), and the step-epilogue oracle against torch's own GradScaler-free arithmetic (rank 3)."""
import numpy as np
import torch

from deltakd_b200 import synth
from deltakd_b200.features import FrozenTeacher, forward_with_features, needed_layers
from oracle import step as S


def test_needed_layers_match_the_reference_branches():
    a = synth.default_args
    assert needed_layers("soft") == () and needed_layers("hard") == () and needed_layers("none") == ()
    assert needed_layers("mgd") == (11,) and needed_layers("saliency_mgd") == (11,)
    assert needed_layers("wasskd") == (0, 1, 2)
    assert needed_layers("vitkd") == (0, 1, 11) and needed_layers("diffkd") == (0, 1, 11)
    assert set(needed_layers("lrkd")) == {0, 1, 11}
    assert needed_layers("curkd", a(current_epoch=0)) == (0, 1, 2)
    assert needed_layers("curkd", a(current_epoch=99)) == (0, 1, 2)
    assert needed_layers("curkd", a(current_epoch=100)) == (3, 4, 5, 6)
    assert needed_layers("curkd", a(current_epoch=150)) == (3, 4, 5, 6)
    assert needed_layers("curkd", a(current_epoch=151)) == (11,)
    assert needed_layers("aaakd") is None          # unknown types keep the reference behaviour (all blocks)


def test_forward_with_features_hooks_only_selected_layers():
    m = synth.FeatureReplayModel(384)
    feats = [torch.full((2, 3), float(i)) for i in range(12)]
    m.set_outputs(torch.zeros(2, 5), feats)
    out, got = forward_with_features(m, torch.zeros(2, 3))
    assert len(got) == 12 and all(torch.equal(g, f) for g, f in zip(got, feats))
    out, got = forward_with_features(m, torch.zeros(2, 3), layers=(0, -1, 5))
    assert [i for i, g in enumerate(got) if g is not None] == [0, 5, 11]
    assert torch.equal(got[11], feats[11])
    assert all(len(b.mlp._forward_hooks) == 0 for b in m.blocks)        # hooks removed again
    assert forward_with_features(torch.nn.Linear(2, 2), torch.zeros(1, 2)) == (None, None)
    # the frozen-teacher wrapper exposes .blocks / .embed_dim and never trains
    ft = FrozenTeacher(m, dtype=None)
    out, got = forward_with_features(ft, torch.zeros(2, 3), layers=(2,))
    assert got[2] is not None and got[0] is None and ft.embed_dim == 384
    ft.train()
    assert not ft.training and not ft.inner.training


def test_mixup_host_draws_and_dense_labels():
    from deltakd_b200.mixup import MixedLabels, Mixup, rand_bbox
    np.random.seed(0)
    mix = Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, label_smoothing=0.1, num_classes=10)
    seen = set()
    for _ in range(50):
        lam, cm = mix._params_per_batch()
        assert 0.0 <= lam <= 1.0
        seen.add(cm)
    assert seen == {True, False}
    yl, yh, xl, xh = rand_bbox((4, 3, 224, 224), 0.3)
    assert 0 <= yl <= yh <= 224 and 0 <= xl <= xh <= 224
    t = torch.tensor([1, 3, 3, 7])
    lab = MixedLabels(t, torch.tensor([0.25]), 0.1, 10)
    dense = lab.dense()
    assert torch.allclose(dense.sum(1), torch.ones(4))
    assert torch.allclose(dense.double(), S.mixup_target(t, 10, 0.25, 0.1), atol=1e-7)
    assert lab.shape == (4, 10)


def test_step_oracle_equals_plain_adamw_when_nothing_fires():
    """No scaling, no clipping, no EMA: the oracle's epilogue_step is exactly one torch.optim.AdamW step."""
    torch.manual_seed(0)
    w = [torch.nn.Parameter(torch.randn(5, 4, dtype=torch.float64)), torch.nn.Parameter(torch.randn(4, dtype=torch.float64))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in w]
    o1 = torch.optim.AdamW(w, lr=1e-2)
    o2 = torch.optim.AdamW(ref, lr=1e-2)
    for p, r in zip(w, ref):
        g = torch.randn_like(p)
        p.grad, r.grad = g.clone(), g.clone()
    skipped, norm = S.epilogue_step(w, o1, S.ScalerState(scale=1.0, dynamic=False), None, None, None)
    o2.step()
    assert not skipped and norm > 0
    for p, r in zip(w, ref):
        assert torch.equal(p, r)
    acc1, acc5 = S.accuracy(torch.eye(6)[:, :6], torch.arange(6), topk=(1, 5))
    assert acc1.item() == 100.0 and acc5.item() == 100.0
