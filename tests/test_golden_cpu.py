"""The committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py) against the oracle."""
import glob
import os

import numpy as np
import pytest

from helpers import rel_err
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["toy_d3_k2", "ragged_d37_k5", "c2shape_d200_k16"])
def test_oracle_reproduces_golden_single(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    X, w, C, mu, s = g["X"], g["w"], g["C0"], g["mu0"], float(g["s0"])
    assert rel_err(orc.llks(X, C, mu, s), g["llks0"]) < 1e-12
    Z, COV = orc.infer(X, C, mu, s)
    assert rel_err(Z, g["Z0"]) < 1e-12 and rel_err(COV, g["COV0"]) < 1e-12
    assert rel_err(orc.extrapolate(X, C, mu, s), g["extrapolate0"]) < 1e-12
    for it in range(g["C_traj"].shape[0]):
        assert orc.llk(X, w, C, mu, s) == pytest.approx(float(g["llk_traj"][it]), rel=1e-11)
        C, mu, s = orc.iterate(X, w, C, mu, s)
        assert rel_err(C, g["C_traj"][it]) < 1e-10 and rel_err(mu, g["mu_traj"][it]) < 1e-10
        assert s == pytest.approx(float(g["s_traj"][it]), rel=1e-10)
        # teacher-forced: the next step starts from the committed model, so the (thread-count dependent) OpenMP
        # reduction order of the oracle cannot compound along the trajectory
        C, mu, s = g["C_traj"][it], g["mu_traj"][it], float(g["s_traj"][it])
    assert np.all(np.diff(g["llk_traj"]) >= -1e-9 * np.abs(g["llk_traj"][1:]))   # ppca_model.rs:263-265


def test_oracle_reproduces_golden_mixture():
    g = np.load(os.path.join(GOLDEN, "mix_d12_k232.npz"))
    X, w, logw = g["X"], g["w"], g["logw0"]
    models = [(g[f"C0_{j}"], g[f"mu0_{j}"], float(g[f"s0_{j}"])) for j in range(len(g["ks"]))]
    assert rel_err(orc.mix_llks(X, models, logw), g["llks0"]) < 1e-12
    assert np.max(np.abs(orc.mix_infer_cluster(X, models, logw) - g["logpost0"])) < 1e-11
    for it in range(3):
        models, logw = orc.mix_iterate(X, w, models, logw)
        assert np.max(np.abs(logw - g[f"logw_{it}"])) < 1e-10
        for j, (C, mu, s) in enumerate(models):
            assert rel_err(C, g[f"C_{it}_{j}"]) < 1e-9 and s == pytest.approx(float(g[f"s_{it}_{j}"]), rel=1e-9)


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLDEN, "*.npz"))) >= 4
