"""CPU tests of the oracle: the reference's own known-answer tests, then independent dense formulas.

The reference pins only two numbers for this path (ppca/src/ppca_model.rs:658-671); both literals are
6-7 significant digits, so they are checked to 1e-6 relative.  Everything else is cross-checked against
tests/dense_ref.py (textbook d_obs x d_obs formulas).
"""
import numpy as np
import pytest

import dense_ref
from helpers import init_model, make_data, rel_err
from oracle import oracle as orc

TOY_C = np.array([[1.0, 1.0, 0.0], [1.0, 0.0, 1.0]]).T  # ppca_model.rs:635-656


def test_reference_kat_quadratic_form():
    # ppca_model.rs:658-665
    assert orc.quadratic_form(TOY_C, 0.1, [1.0, 1.0, 1.0]) == pytest.approx(34.219288, rel=1e-6)


def test_reference_kat_covariance_log_det():
    # ppca_model.rs:667-671
    assert orc.covariance_log_det(TOY_C, 0.1) == pytest.approx(-3.49328, rel=1e-6)


def test_reference_llk_sample_matches_dense():
    # ppca_model.rs:673-680 only prints this value; pin it against the dense formula instead
    x = np.array([[1.0, 2.0, 3.0]])
    mu = np.array([0.0, 1.0, 0.0])
    assert orc.llk(x, None, TOY_C, mu, 0.1) == pytest.approx(dense_ref.llk_one(x[0], TOY_C, mu, 0.1), rel=1e-12)


@pytest.mark.parametrize("n,d,k", [(60, 3, 2), (80, 17, 5), (50, 40, 9), (40, 12, 1), (30, 9, 3), (30, 10, 4)])
def test_llks_infer_against_dense(n, d, k):
    X = make_data(n, d, k, 0.3, seed=d, empty_rows=(3,))
    C0, mu0, _ = init_model(d, k)
    for sigma in (1.0, 0.2):
        want = np.array([dense_ref.llk_one(x, C0, mu0, sigma) for x in X])
        assert rel_err(orc.llks(X, C0, mu0, sigma), want) < 1e-10
        Z, COV = orc.infer(X, C0, mu0, sigma)
        for i in range(n):
            z, cov = dense_ref.infer_one(X[i], C0, mu0, sigma)
            assert np.max(np.abs(Z[i] - z)) < 1e-9 * max(1.0, np.max(np.abs(z)))
            assert np.max(np.abs(COV[i] - cov)) < 1e-9
    assert orc.llks(X, C0, mu0, 1.0)[3] == 0.0
    w = np.random.default_rng(0).random(n)
    assert orc.llk(X, w, C0, mu0, 1.0) == pytest.approx(float(w @ orc.llks(X, C0, mu0, 1.0)), rel=1e-12)


@pytest.mark.parametrize("n,d,k", [(100, 3, 2), (120, 15, 4), (90, 30, 6)])
def test_iterate_against_dense(n, d, k):
    X = make_data(n, d, k, 0.25, seed=n, empty_rows=(1,), empty_dims=(d - 1,))
    C, mu, s = init_model(d, k, empty_dims=(d - 1,))
    w = np.random.default_rng(1).random(n) + 0.5
    for _ in range(3):
        got = orc.iterate(X, w, C, mu, s)
        want = dense_ref.iterate(X, w, C, mu, s)
        assert rel_err(got[0], want[0]) < 1e-9 and rel_err(got[1], want[1]) < 1e-9
        assert got[2] == pytest.approx(want[2], rel=1e-10)
        assert np.array_equal(got[0][d - 1], C[d - 1])  # empty dimension keeps the old row (:313-321)
        C, mu, s = got


def test_iterate_priors_against_dense():
    n, d, k = 150, 8, 3
    X = make_data(n, d, k, 0.2, seed=2)
    C, mu, s = init_model(d, k)
    rng = np.random.default_rng(3)
    A = rng.standard_normal((d, d)); cov0 = A @ A.T / d + np.eye(d); m0 = rng.standard_normal(d)
    pr = orc.Prior(mean=m0, mean_covariance=cov0, isotropic_noise_alpha=2.0, isotropic_noise_beta=1.5,
                   transformation_precision=0.7)
    got = orc.iterate(X, None, C, mu, s, pr)
    want = dense_ref.iterate(X, np.ones(n), C, mu, s, tau=0.7, alpha=2.0, beta=1.5, m0=m0, cov0=cov0)
    assert rel_err(got[0], want[0]) < 1e-9 and rel_err(got[1], want[1]) < 1e-9
    assert got[2] == pytest.approx(want[2], rel=1e-10)


def test_smooth_extrapolate_properties():
    X = make_data(70, 11, 3, 0.3, seed=8, empty_rows=(0,))
    C0, mu0, _ = init_model(11, 3)
    sm = orc.smooth(X, C0, mu0, 0.4)
    ex = orc.extrapolate(X, C0, mu0, 0.4)
    fin = np.isfinite(X)
    assert np.array_equal(ex[fin], X[fin])            # ppca_model.rs:246-247: extant values untouched
    assert np.array_equal(ex[~fin], sm[~fin])
    Z, _ = orc.infer(X, C0, mu0, 0.4)
    assert np.allclose(sm, Z @ C0.T + mu0, rtol=0, atol=1e-12)
    assert np.array_equal(sm[0], mu0)                  # empty sample: state 0 -> mean


def test_to_canonical_properties():
    rng = np.random.default_rng(4)
    C = rng.standard_normal((25, 6))
    Cn = orc.to_canonical(C)
    G = Cn.T @ Cn
    assert np.max(np.abs(G - np.diag(np.diag(G)))) < 1e-10          # orthogonal columns = U S
    nrm = np.sqrt(np.diag(G))
    assert np.all(np.diff(nrm) <= 1e-12)                             # singular values descending
    assert np.allclose(nrm, np.linalg.svd(C, compute_uv=False))
    assert np.all(Cn.sum(axis=0) >= 0)                               # sign flip (ppca_model.rs:414-416)
    assert np.allclose(Cn @ Cn.T, C @ C.T, atol=1e-10)               # same output covariance
    X = make_data(50, 25, 6, 0.2, seed=1)
    mu = np.zeros(25)
    assert orc.llk(X, None, Cn, mu, 0.5) == pytest.approx(orc.llk(X, None, C, mu, 0.5), rel=1e-11)  # :395-397


def test_mixture_against_definitions():
    n, d = 80, 6
    X = make_data(n, d, 2, 0.2, seed=6)
    models = [init_model(d, 2, seed=10), init_model(d, 3, seed=11)]
    models = [(C, mu + j, 1.0 + 0.2 * j) for j, (C, mu, s) in enumerate(models)]
    logw = np.log([0.3, 0.7])
    comp = np.stack([orc.llks(X, C, mu, s) for C, mu, s in models], axis=1) + logw
    mx = comp.max(axis=1, keepdims=True)
    lse = (mx + np.log(np.exp(comp - mx).sum(axis=1, keepdims=True)))[:, 0]
    assert np.allclose(orc.mix_llks(X, models, logw), lse, rtol=1e-13)
    lp = orc.mix_infer_cluster(X, models, logw)
    assert np.allclose(lp, comp - lse[:, None], atol=1e-12)
    assert np.allclose(np.exp(lp).sum(axis=1), 1.0)
    # one EM step: each component = weighted single-model iterate with responsibilities (mix.rs:297-330)
    w = np.random.default_rng(0).random(n) + 0.5
    new_models, new_logw = orc.mix_iterate(X, w, models, logw)
    logsum = []
    for j, (C, mu, s) in enumerate(models):
        l = np.log(w) + lp[:, j]
        r = np.exp(l - l.max())
        Cj, muj, sj = orc.iterate(X, r, C, mu, s)
        assert rel_err(Cj, new_models[j][0]) < 1e-11 and abs(sj - new_models[j][2]) < 1e-12 * sj
        logsum.append(np.log(r.sum()) + l.max())
    assert np.allclose(new_logw, orc.log_softmax(np.array(logsum)), atol=1e-13)
    sm = orc.mix_smooth(X, models, logw)
    want = sum(np.exp(lp[:, j:j + 1]) * orc.smooth(X, C, mu, s) for j, (C, mu, s) in enumerate(models))
    assert np.allclose(sm, want, atol=1e-12)


def test_empty_dimensions():
    X = np.array([[1.0, 1.0, np.nan], [1.0, 1.0, np.nan]])  # examples/empty_dimensions.py
    assert orc.empty_dimensions(X) == [2]


def test_reference_noise_floor_is_documented():
    """The reference's posterior formulas cancel (I - T C_o); its own error versus 50-digit arithmetic grows
    like eps (|G|/sigma^2)^2.  The stable oracle variant stays at rounding level.  This is why the GPU parity
    tests compare against both variants (see tests/test_gpu_parity.py::assert_close)."""
    import mpmath as mp
    X = make_data(10, 30, 4, 0.2, seed=1)
    C0, mu0, _ = init_model(30, 4)
    mp.mp.dps = 50
    x = X[0]; m = np.isfinite(x)
    Co = mp.matrix(C0[m].tolist())
    errs = {}
    for sigma in (1.0, 0.01):
        Minv = (Co.T * Co + mp.eye(4) * mp.mpf(sigma) ** 2) ** -1
        truth = np.array([[float(Minv[i, j] * mp.mpf(sigma) ** 2) for j in range(4)] for i in range(4)])
        _, cov_f = orc.infer(X[:1], C0, mu0, sigma)
        with orc.stable():
            _, cov_s = orc.infer(X[:1], C0, mu0, sigma)
        errs[sigma] = (rel_err(cov_f[0], truth), rel_err(cov_s[0], truth))
    assert errs[1.0][0] < 1e-11 and errs[1.0][1] < 1e-13
    assert errs[0.01][0] > 1e-9          # the reference's own noise floor is above the parity tolerance here
    assert errs[0.01][1] < 1e-13         # the cancellation-free form is not


def test_covariance_family_restatement_is_consistent():
    """oracle.smoothed_/extrapolated_covariance(_diagonal) and mix_covariance (ppca_model.rs:471-577, mix.rs:422-501):
    diagonal forms are the diagonals of the full forms, expansion zeroes the observed rows / columns, and the mixture
    form is the law of total covariance."""
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    d, k = 9, 3
    Cm = rng.standard_normal((d, k))
    A = rng.standard_normal((k, k))
    cov = A @ A.T + np.eye(k)
    x = rng.standard_normal(d)
    x[[1, 4, 5]] = np.nan
    full = orc.smoothed_covariance(Cm, 0.3, cov)
    assert np.allclose(np.diag(full), orc.smoothed_covariance_diagonal(Cm, 0.3, cov), rtol=1e-14)
    ex = orc.extrapolated_covariance(Cm, 0.3, cov, x)
    neg = np.isnan(x)
    assert np.array_equal(ex[np.ix_(neg, neg)], np.eye(3) * 0.09 + Cm[neg] @ cov @ Cm[neg].T)
    assert not ex[~neg].any() and not ex[:, ~neg].any()
    assert np.allclose(np.diag(ex), orc.extrapolated_covariance_diagonal(Cm, 0.3, cov, x), rtol=1e-14)
    assert not orc.extrapolated_covariance(Cm, 0.3, cov, np.ones(d)).any()
    # law of total covariance against a Monte-Carlo draw of the two-component Gaussian mixture
    means = [rng.standard_normal(d), rng.standard_normal(d) + 2.0]
    covs = [full, orc.smoothed_covariance(Cm[:, ::-1], 0.5, cov)]
    post = np.array([0.3, 0.7])
    want = orc.mix_covariance(post, means, covs)
    n = 400_000
    comp = rng.random(n) < post[1]
    draws = np.where(comp[:, None], rng.multivariate_normal(means[1], covs[1], n), rng.multivariate_normal(means[0], covs[0], n))
    assert np.max(np.abs(np.cov(draws.T) - want)) < 0.03 * np.max(np.abs(want))
