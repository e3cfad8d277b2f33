"""GPU: the LRKD eigensolver on its own (dkd_lrkd_eigensolve) against numpy's LAPACK eigh in fp64 — both algorithms
(cluster-resident / cooperative), spectra the Gram of teacher features can have: flat (Gaussian features, large batch),
wide (few rows), decaying (real features), rank-deficient, diagonal, zero."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N = 384


def _gram(kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "flat":
        x = rng.standard_normal((20000, N)).astype(np.float32).astype(np.float64)
    elif kind == "wide":
        x = rng.standard_normal((588, N)).astype(np.float32).astype(np.float64)
    elif kind == "decay":
        q = np.linalg.qr(rng.standard_normal((N, N)))[0]
        x = (rng.standard_normal((4000, N)) * np.exp(-np.arange(N) / 40.0)[None, :]) @ q
    elif kind == "rank196":
        x = rng.standard_normal((196, N))
    elif kind == "diag":
        return np.diag(np.linspace(1.0, 50.0, N))
    elif kind == "zero":
        return np.zeros((N, N))
    elif kind == "huge":
        x = rng.standard_normal((3000, N)) * 1e9
    else:
        raise KeyError(kind)
    g = x.T @ x
    return 0.5 * (g + g.T)


def _check(g, w, sweeps, top=128):
    lam = np.linalg.norm(w, axis=1)
    order = np.argsort(-lam)
    ev, evec = np.linalg.eigh(g)
    ev, evec = ev[::-1], evec[:, ::-1]
    scale = max(ev[0], 1e-300)
    k = int(min(top, (ev > 1e-10 * scale).sum()))
    assert np.abs(lam[order[:k]] - ev[:k]).max() <= 1e-12 * scale, "eigenvalues"
    v = w[order[:k]] / lam[order[:k], None]
    # residual G v = lambda v and orthonormality of the returned top vectors
    assert np.abs(v @ g - lam[order[:k], None] * v).max() <= 1e-9 * scale, "residual"   # the last sweep saw |cos| <= 1e-6 and rotated those pairs
    assert np.abs(v @ v.T - np.eye(k)).max() <= 1e-10, "orthonormality"
    # against LAPACK where the vector is well defined (relative gap to both neighbours > 1e-6)
    gaps = np.minimum(np.abs(np.diff(ev[:k + 1], prepend=ev[0] + scale)), np.abs(np.diff(ev[:k + 1], append=-scale))[:k + 1])[:k]
    good = gaps > 1e-6 * scale
    dots = np.abs((v * evec[:, :k].T).sum(1))
    if good.any():
        assert np.abs(dots[good] - 1).max() < 1e-9, ("vectors", np.abs(dots[good] - 1).max())
    assert 1 <= int(sweeps) < 30, sweeps   # converged, not capped (cooperative version: cap 14; cluster-resident: 30)


# decaying / rank-deficient spectra: only the cluster-resident version (it stops on the wanted columns; the cooperative
# fallback waits for ALL columns and runs into its cap of 14 sweeps there)
CASES = [("flat", 384, 1), ("flat", 384, 2), ("flat", 64, 1), ("wide", 384, 1), ("wide", 384, 2), ("wide", 32, 1),
         ("decay", 64, 1), ("decay", 128, 1), ("rank196", 128, 1), ("diag", 384, 1), ("diag", 384, 2), ("zero", 384, 1),
         ("zero", 384, 2), ("huge", 384, 1), ("huge", 384, 2)]


@pytest.mark.parametrize("kind,k,algo", CASES)
def test_eigensolve_matches_lapack(kind, k, algo):
    """algo 1 = cluster-resident (k leading pairs wanted), 2 = cooperative (always the full decomposition)."""
    from deltakd_b200 import functional as Fn
    gs = [_gram(kind, s) for s in (1, 2, 3)]
    w, sweeps = Fn.lrkd_eigensolve(torch.tensor(np.stack(gs), device="cuda"), k=k, algo=algo)
    w = w.cpu().numpy()
    for l in range(3):
        if kind == "zero":
            assert np.abs(w[l]).max() == 0.0
            continue
        _check(gs[l], w[l], sweeps[l].item(), top=k)


def test_eigensolve_time_and_default_algo(capsys):
    """The default algorithm is the cluster-resident one on a B200; print the time of both for the record."""
    from deltakd_b200 import functional as Fn
    g = torch.tensor(np.stack([_gram("flat", s) for s in (4, 5, 6)]), device="cuda")
    out = {}
    for algo in (0, 1, 2):
        Fn.lrkd_eigensolve(g, k=64, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w, sw = Fn.lrkd_eigensolve(g, k=64, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        out[algo] = (e0.elapsed_time(e1) * 1e3, sw.tolist(), w)
    with capsys.disabled():
        print("\nLRKD eigensolve, 3 matrices 384x384 fp64, k = 64 (us incl. clone): "
              + ", ".join(f"algo {a}: {t:.0f} us sweeps {s}" for a, (t, s, _) in out.items()))
    assert torch.equal(out[0][2], out[1][2])   # default = cluster-resident, deterministic
