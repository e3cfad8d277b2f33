"""GPU: rank-by-counting mask selection, bit-exact vs the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import losses as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ratio", [0.5, 0.3, 0.75, 0.0])
def test_random_masking_bit_exact_vs_reference(golden, ratio):
    from deltakd_b200 import random_masking
    tag = f"random_masking/r{ratio}"
    noise = torch.from_numpy(golden[f"{tag}/noise"]).cuda()
    x = torch.randn(4, 196, 8, generator=torch.Generator().manual_seed(5)).cuda()
    x_keep, mask, ids_restore, ids_masked = random_masking(x, ratio, noise=noise)
    assert mask.dtype == torch.float32 and ids_restore.dtype == torch.int64
    assert np.array_equal(mask.cpu().numpy(), golden[f"{tag}/mask"])
    assert np.array_equal(ids_restore.cpu().numpy(), golden[f"{tag}/ids_restore"])
    assert np.array_equal(ids_masked.cpu().numpy(), golden[f"{tag}/ids_masked"])
    xk, _, _, _ = O.random_masking(x.cpu(), ratio, noise.cpu())
    assert torch.equal(x_keep.cpu(), xk)


def test_random_masking_draws_like_reference():
    """Without `noise`, the draw is torch.rand(N, L, device=x.device) from the global generator (misc.py:14)."""
    from deltakd_b200 import random_masking
    x = torch.randn(3, 196, 4, device="cuda")
    torch.manual_seed(123)
    _, mask, ids_restore, _ = random_masking(x, 0.5)
    torch.manual_seed(123)
    noise = torch.rand(3, 196, device="cuda")
    m, r, _ = O.mask_from_scores(noise.cpu(), 98)
    assert torch.equal(mask.cpu(), m) and torch.equal(ids_restore.cpu(), r)


def test_ties_rank_lower_index_first(golden):
    from deltakd_b200 import functional as Fn
    noise = torch.from_numpy(golden["random_masking/ties/noise"])
    mask, ids_restore, ids_shuffle = Fn.mask_rank(noise.cuda(), 98)
    m, r, s = O.mask_from_scores(noise, 98)
    assert torch.equal(mask.cpu(), m) and torch.equal(ids_restore.cpu(), r) and torch.equal(ids_shuffle.cpu(), s)


@pytest.mark.parametrize("B,L,keep", [(1, 1, 0), (1, 1, 1), (2, 7, 3), (513, 196, 137), (4096, 196, 98), (3, 1024, 512), (0, 196, 98)])
def test_rank_edge_shapes(B, L, keep):
    from deltakd_b200 import functional as Fn
    score = torch.rand(B, L, generator=torch.Generator().manual_seed(B + L))
    mask, ids_restore, ids_shuffle = Fn.mask_rank(score.cuda(), keep)
    if B == 0:
        assert mask.shape == (0, L)
        return
    m, r, s = O.mask_from_scores(score, keep)
    assert torch.equal(mask.cpu(), m) and torch.equal(ids_restore.cpu(), r) and torch.equal(ids_shuffle.cpu(), s)
    assert torch.all(mask.sum(1) == L - keep)
    # idempotence / permutation property at full size
    assert torch.equal(torch.sort(ids_restore, dim=1).values.cpu(), torch.arange(L).expand(B, L))
