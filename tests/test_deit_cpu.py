"""CPU: the bench-harness DeiT exposes the surface the loss path touches (models.py:59-75, 96-97, 181-199)."""
import torch

from deltakd_b200 import deit
from deltakd_b200.features import forward_with_features


def test_distilled_student_returns_two_heads_in_training():
    m = deit.create_model("deit_tiny_distilled_patch16_224", num_classes=100)
    m.set_distilled_training(True)
    m.train()
    out = m(torch.randn(2, 3, 224, 224))
    assert isinstance(out, tuple) and out[0].shape == (2, 100) and out[1].shape == (2, 100)
    m.eval()
    assert m(torch.randn(2, 3, 224, 224)).shape == (2, 100)   # eval: mean of the two heads (timm convention)


def test_hooks_capture_pre_residual_mlp_outputs():
    t = deit.create_model("deit_small_distilled_patch16_224").eval()
    s = deit.create_model("deit_tiny_patch16_224")
    x = torch.randn(1, 3, 224, 224)
    with torch.no_grad():
        out, feats = forward_with_features(t, x)
    assert out.shape == (1, 1000) and len(feats) == 12 and feats[0].shape == (1, 198, 384)   # CLS, DIST + 196 patches
    out, feats = forward_with_features(s, x)
    assert feats[11].shape == (1, 197, 192) and feats[11].requires_grad
    assert t.embed_dim == 384 and s.embed_dim == 192
    assert forward_with_features(torch.nn.Linear(3, 3), x) == (None, None)   # models.py:182-183


def test_dkd_row_ops_model_has_same_parameters_and_no_cpu_path():
    import pytest
    a = deit.create_model("deit_tiny_distilled_patch16_224", num_classes=10, row_ops="aten")
    d = deit.create_model("deit_tiny_distilled_patch16_224", num_classes=10, row_ops="dkd")
    assert [(k, tuple(v.shape)) for k, v in a.state_dict().items()] == [(k, tuple(v.shape)) for k, v in d.state_dict().items()]
    d.load_state_dict(a.state_dict())
    with pytest.raises(RuntimeError):          # libdeltakd_sm100's row ops take CUDA tensors only: loud failure, no fallback
        d(torch.randn(1, 3, 224, 224))
    with pytest.raises(ValueError):
        deit.create_model("deit_tiny_patch16_224", row_ops="triton")
