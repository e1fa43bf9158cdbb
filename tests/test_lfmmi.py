# SPDX-License-Identifier: MIT
"""LF-MMI loss wrapper (caller pattern examples/test_cuda.jl:118-152):
loss and gradient against the oracle's posteriors, and the gradient against finite differences."""
import numpy as np
import pytest

from test_gpu_parity import dev, gpu_batch, orc_graphs, torch  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def _setup(mm, dtype, B=6, T=30, D=60, seed=7):
    K = mm.LogSemiring[dtype]
    rng = np.random.default_rng(seed)
    den = mm.graphs.denominator(K, n_tokens=300, n_pdf=D, seed=seed)
    nums = [mm.graphs.numerator(K, np.random.default_rng(seed + k), D, n_phones=int(rng.integers(4, 9))) for k in range(B)]
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    lens = rng.integers(T - 6, T + 1, B).astype(np.int32)
    lens[0] = T
    return K, den, nums, V, lens


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lfmmi_loss_and_gradient_vs_oracle(torch, mm, orc, dtype):
    K, den, nums, V, lens = _setup(mm, dtype)
    B, T, D = V.shape
    bnum = gpu_batch(mm, nums, D)
    bden = gpu_batch(mm, [den] * B, D, "shared")
    x = torch.from_numpy(V).cuda().requires_grad_(True)
    loss, zn, zd = mm.lfmmi_loss(x, bnum, bden, seqlengths=lens)
    loss.backward()
    K64 = mm.LogSemiring[np.float64]
    n64 = [(f.astype(K64), p) for f, p in nums]
    d64 = (den[0].astype(K64), den[1])
    V64 = V.astype(np.float64)
    npost, nttl = orc.pdfposteriors(orc_graphs(orc, n64, D), V64, lens)
    dpost, dttl = orc.pdfposteriors(orc_graphs(orc, [d64] * B, D), V64, lens)
    rtol = 1e-4 if dtype == np.float32 else 1e-9
    np.testing.assert_allclose(zn.cpu().numpy(), nttl, rtol=rtol)
    np.testing.assert_allclose(zd.cpu().numpy(), dttl, rtol=rtol)
    np.testing.assert_allclose(float(loss.detach()), -(nttl - dttl).sum(), rtol=10 * rtol)
    want = (dpost - npost).transpose(0, 2, 1)  # (B, T, D)
    got = x.grad.cpu().numpy()
    assert got.shape == (B, T, D)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=2e-6 if dtype == np.float32 else 1e-12)
    for b in range(B):  # frames beyond an utterance's length get no gradient
        assert np.all(got[b, int(lens[b]):] == 0.0)


def test_lfmmi_gradient_finite_differences(torch, mm):
    """d loss / d loglikes by central differences of the Float64 loss itself."""
    K, den, nums, V, lens = _setup(mm, np.float64, B=3, T=26, D=20, seed=11)
    B, T, D = V.shape
    bnum = gpu_batch(mm, nums, D)
    bden = gpu_batch(mm, [den] * B, D, "shared")
    x = torch.from_numpy(V).cuda().requires_grad_(True)
    loss, _, _ = mm.lfmmi_loss(x, bnum, bden, seqlengths=lens)
    assert bool(torch.isfinite(loss))
    loss.backward()
    g = x.grad.cpu().numpy()
    rng = np.random.default_rng(0)
    eps = 1e-5
    for _ in range(12):
        b, d = int(rng.integers(B)), int(rng.integers(D))
        t = int(rng.integers(int(lens[b])))
        Vp, Vm = V.copy(), V.copy()
        Vp[b, t, d] += eps
        Vm[b, t, d] -= eps
        lp = float(mm.lfmmi_loss(torch.from_numpy(Vp).cuda(), bnum, bden, seqlengths=lens)[0])
        lm = float(mm.lfmmi_loss(torch.from_numpy(Vm).cuda(), bnum, bden, seqlengths=lens)[0])
        assert abs((lp - lm) / (2 * eps) - g[b, t, d]) < 1e-6, (b, t, d)


def test_lfmmi_grad_kernel_strided_output(torch, mm):
    """mk_lfmmi_grad writes any (B, T, D) strides; scale is applied."""
    rng = np.random.default_rng(3)
    B, D, N = 5, 37, 9
    num = torch.from_numpy(rng.random((N, D, B)).astype(np.float32)).cuda()
    den = torch.from_numpy(rng.random((N, D, B)).astype(np.float32)).cuda()
    lens = np.array([9, 4, 0, 9, 7], np.int32)
    big = torch.zeros((B, N, 2 * D), device="cuda")
    out = big[:, :, ::2]
    mm.lfmmi_grad(num.permute(2, 1, 0), den.permute(2, 1, 0), lens, scale=0.5, out=out)
    want = 0.5 * (den - num).permute(2, 0, 1).cpu().numpy()
    for b in range(B):
        want[b, int(lens[b]):] = 0
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-6)
    assert float(big[:, :, 1::2].abs().sum()) == 0.0
