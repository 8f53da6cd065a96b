"""GPU tier: post-correlation array processing: gat_beamform / gat_eigen_weights against float64 numpy
restatements of their formulas.  PARITY UNPINNED against the reference: upstream's `post_corr_filter` is a user
closure, there is no reference implementation or test to pin to."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cuda(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()


@pytest.mark.parametrize("n_ch,L,M", [(1, 1, 1), (7, 3, 5), (64, 11, 16), (3, 3, 32)])
def test_beamform_matches_numpy(engine, n_ch, L, M):
    rng = np.random.default_rng(n_ch * 100 + M)
    acc = rng.normal(size=(n_ch, L, M)) + 1j * rng.normal(size=(n_ch, L, M))
    w = rng.normal(size=(n_ch, M)) + 1j * rng.normal(size=(n_ch, M))
    y_re, y_im = engine.beamform((_cuda(acc.real), _cuda(acc.imag)), (_cuda(w.real), _cuda(w.imag)))
    got = y_re.cpu().numpy() + 1j * y_im.cpu().numpy()
    want = np.einsum("km,klm->kl", w.conj(), acc)
    assert got.shape == (n_ch, L)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max() + 1e-6


def test_eigen_weights_track_the_steering_vector(engine):
    import torch
    rng = np.random.default_rng(3)
    n_ch, L, M, T = 5, 3, 8, 40
    phi = rng.uniform(-1.0, 1.0, n_ch)
    steer = np.exp(1j * np.arange(M)[None, :] * phi[:, None])                       # [n_ch, M]
    cov = (torch.zeros(n_ch, M, M, device="cuda"), torch.zeros(n_ch, M, M, device="cuda"))
    w = (torch.zeros(n_ch, M, device="cuda"), torch.zeros(n_ch, M, device="cuda"))
    R = np.zeros((n_ch, M, M), complex)
    forget = 0.9
    for t in range(T):
        s = np.exp(2j * np.pi * rng.uniform(size=n_ch))[:, None] * 1000.0           # unknown data/carrier phase per period
        p = s * steer + 150.0 * (rng.normal(size=(n_ch, M)) + 1j * rng.normal(size=(n_ch, M)))
        acc = np.zeros((n_ch, L, M), complex)
        acc[:, 1] = p
        acc[:, 0] = acc[:, 2] = 0.5 * p
        engine.eigen_weights((_cuda(acc.real), _cuda(acc.imag)), cov, w, tap=1, forget=forget, iters=3)
        R = forget * R + p.astype(np.complex64)[:, :, None] * p.astype(np.complex64).conj()[:, None, :]
    got_R = cov[0].cpu().numpy() + 1j * cov[1].cpu().numpy()
    assert np.abs(got_R - R).max() <= 1e-4 * np.abs(R).max()
    got_w = w[0].cpu().numpy() + 1j * w[1].cpu().numpy()
    for k in range(n_ch):
        vals, vecs = np.linalg.eigh(R[k])
        v = vecs[:, -1]
        assert abs(np.vdot(v, got_w[k])) > 1 - 1e-4                                  # same direction as the dominant eigenvector
        assert abs(np.linalg.norm(got_w[k]) - 1) < 1e-4 and abs(got_w[k, 0].imag) < 1e-5 and got_w[k, 0].real > 0
        assert abs(np.vdot(steer[k] / np.sqrt(M), got_w[k])) > 0.99                  # ... which is the steering vector


def test_correlate_then_eigen_beamform_on_device(gat, engine):
    """The whole post-correlation stage behind the correlator without leaving the device: an 8-element array,
    per-antenna phase step, noise; the eigen-beamformed prompt collects the full array gain sqrt(M)."""
    import torch
    l1 = gat.GPSL1()
    n, M, fs, L = 20000, 8, 2.0e7, 3
    step = 0.7
    shifts = np.array([-10, 0, 10], np.int32)
    chans = [gat.Channel(l1, 7, 200.0, 1200.0, 0.0)]
    out = (torch.zeros(1, 1, L, M, device="cuda"), torch.zeros(1, 1, L, M, device="cuda"))
    cov = (torch.zeros(1, M, M, device="cuda"), torch.zeros(1, M, M, device="cuda"))
    w = (torch.zeros(1, M, device="cuda"), torch.zeros(1, M, device="cuda"))
    y = None
    for t in range(6):
        engine.gen_signal(30, l1, 7, 1200.0, fs, n, M, start_code_phase=200.0, ant_phase_step=step, noise_sigma=2.0, seed=t)
        engine.correlate_batch([30], [chans], fs, shifts, M, 0, n, out=out)
        engine.eigen_weights(out, cov, w, tap=1, forget=0.9, iters=3)
        y = engine.beamform(out, w)
    prompt = complex(y[0][0, 0, 1].item(), y[1][0, 0, 1].item())
    assert abs(abs(prompt) - n * np.sqrt(M)) < 0.02 * n * np.sqrt(M)
    got_w = (w[0] + 1j * w[1]).cpu().numpy()[0]
    steer = np.exp(1j * step * np.arange(M)) / np.sqrt(M)
    assert abs(np.vdot(steer, got_w)) > 0.995
