"""CPU: pins the oracle (oracle/losses.py) to what the unmodified reference produced
(tests/golden/reference_v1.npz, written by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import losses as O
from oracle.util import digest, rel_err
from tests.cases import CASES, build_case
from deltakd_b200 import heads as H

# LRKD: the loss depends on LAPACK's arbitrary SVD column signs; fp32 vs fp64 differ
LOSS_TOL = {"f32": 2e-6, "f64": 1e-10}
GRAD_TOL = {"f32": 2e-5, "f64": 1e-8}


def run_oracle(name, dtype):
    c = build_case(name, dtype=dtype)
    heads = H.head_tensors(c.student)
    base_kind = "label_smoothing" if c.int_labels else "soft_target"
    outs = (c.outputs, c.outputs_kd) if c.kind in ("soft", "hard") else c.outputs
    loss = O.distillation_loss(c.kind, outs, c.labels, c.teacher_logits, c.s_feats, c.t_feats, heads,
                               c.args, c.alpha, c.tau, base_kind=base_kind, noise=c.noise)
    loss.backward()
    return c, heads, loss


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_oracle_matches_reference(golden, name, tag):
    dtype = torch.float32 if tag == "f32" else torch.float64
    c, heads, loss = run_oracle(name, dtype)
    lrkd = c.kind == "lrkd"
    ref = float(golden[f"{name}/{tag}/loss"])
    tol_l = 2e-4 if lrkd and tag == "f32" else LOSS_TOL[tag]
    assert abs(loss.item() - ref) <= tol_l * abs(ref), (loss.item(), ref)
    if tag == "f32":
        ind = golden[f"{name}/in_digest"]
        assert np.allclose(digest(c.outputs)[:3], ind[:3], rtol=1e-12), "synthetic-input RNG drift"
    if lrkd and tag == "f32":
        return  # gradient depends on the per-column signs fp32 LAPACK happened to pick; f64 is pinned below
    tol_g = GRAD_TOL[tag]
    assert rel_err(digest(c.outputs.grad), golden[f"{name}/{tag}/g_outputs"]) < tol_g
    if c.kind in ("soft", "hard"):
        assert rel_err(digest(c.outputs_kd.grad), golden[f"{name}/{tag}/g_outputs_kd"]) < tol_g
    if c.needs_feats:
        seen = 0
        for i, f in enumerate(c.s_feats):
            key = f"{name}/{tag}/g_sfeat{i}"
            if key in golden.files:
                assert f.grad is not None, key
                assert rel_err(digest(f.grad), golden[key]) < tol_g, key
                seen += 1
            else:
                assert f.grad is None or float(f.grad.abs().sum()) == 0.0
        assert seen > 0
        for k, p in heads.items():
            key = f"{name}/{tag}/g_head/{k}"
            if key in golden.files:
                assert p.grad is not None, key
                assert rel_err(digest(p.grad), golden[key]) < tol_g, key


@pytest.mark.parametrize("ratio", [0.5, 0.3, 0.75, 0.0])
def test_random_masking_bit_exact(golden, ratio):
    tag = f"random_masking/r{ratio}"
    noise = torch.from_numpy(golden[f"{tag}/noise"])
    x = torch.randn(4, 196, 8, generator=torch.Generator().manual_seed(5))
    x_keep, mask, ids_restore, ids_masked = O.random_masking(x, ratio, noise)
    assert np.array_equal(mask.numpy(), golden[f"{tag}/mask"])
    assert np.array_equal(ids_restore.numpy(), golden[f"{tag}/ids_restore"])
    assert np.array_equal(ids_masked.numpy(), golden[f"{tag}/ids_masked"])
    assert rel_err(digest(x_keep), golden[f"{tag}/x_keep_digest"]) == 0.0
    L = 196
    assert int(mask.sum(1)[0]) == L - int(L * (1 - ratio))


def test_random_masking_ties(golden):
    """Equal noise keys: torch.argsort(stable=False) is formally unspecified (misc.py:17-18).  The
    reference's CPU result (recorded) is *a* valid ranking; ours is defined as lower-index-first,
    which is what torch's CUDA radix sort yields.  Both must be permutations that sort the keys and
    mask exactly L - len_keep tokens; away from tied keys they agree."""
    noise = torch.from_numpy(golden["random_masking/ties/noise"])
    x = torch.randn(3, 196, 8, generator=torch.Generator().manual_seed(5))
    _, mask, ids_restore, _ = O.random_masking(x, 0.5, noise)
    ref_restore = golden["random_masking/ties/ids_restore"]
    ref_mask = golden["random_masking/ties/mask"]
    n = noise.numpy()
    for b in range(3):
        for r in (ids_restore.numpy()[b], ref_restore[b]):
            assert sorted(r.tolist()) == list(range(196))
            inv = np.empty(196, dtype=np.int64)
            inv[r] = np.arange(196)
            assert np.all(np.diff(n[b][inv]) >= 0)
        # ours: stable
        order = np.argsort(ids_restore.numpy()[b])
        same = n[b][order][1:] == n[b][order][:-1]
        assert np.all(order[1:][same] > order[:-1][same])
        assert mask[b].sum() == ref_mask[b].sum() == 98
        # tokens whose key is not the boundary key get the same mask bit in both
        boundary = np.sort(n[b])[97:99]
        free = ~np.isin(n[b], boundary)
        assert np.array_equal(mask.numpy()[b][free], ref_mask[b][free])


@pytest.mark.parametrize("m", [1, 2, 3])
def test_saliency_scores_and_masks(golden, m):
    from deltakd_b200 import synth
    from types import SimpleNamespace
    args = SimpleNamespace(distillation_type="saliency_mgd", saliency_method=m)
    teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
    torch.manual_seed(0)
    H.attach_distillation_heads(student, teacher, args)
    heads = H.head_tensors(student)
    _, t_feats = synth.make_features(3, 77, layers=[11])
    with torch.no_grad():
        score = O.saliency_score(m, t_feats[11], heads)
    ref = golden[f"saliency_masking/m{m}/score"]
    assert rel_err(score, ref) < 2e-6
    # bit-exactness is defined given the same score tensor: feed the reference's scores
    mask, ids_restore, _ = O.mask_from_scores(torch.from_numpy(ref), O.len_keep_of(196, 0.5))
    assert np.array_equal(mask.numpy(), golden[f"saliency_masking/m{m}/mask"])
    assert np.array_equal(ids_restore.numpy(), golden[f"saliency_masking/m{m}/ids_restore"])


def test_lrkd_sign_aligned_f32(golden):
    """fp32 LRKD: singular values match the reference; the projector V_k V_k^T is sign-free."""
    c = build_case("lrkd_r32")
    for j, ti in enumerate((0, 1, 11)):
        A, V, S = O.lrkd_targets(c.t_feats[ti][:, 2:], 32)
        assert rel_err(S, golden[f"lrkd_r32/svd_S{j}"]) < 1e-5
        T = c.t_feats[ti][:, 2:].reshape(-1, 384)
        assert rel_err(T @ V, A) < 1e-4  # U_k S_k == T V_k
