"""PPCA mixture: model selection over the number of components, then inference (cf. examples/ppca_mixture.py)."""
import _alias  # noqa: F401
import numpy as np
from ppca_rs import PPCAMix, PPCAMixTrainer
from ppca_rs.ppca_rs import PPCAModel

truth = PPCAMix(
    [
        PPCAModel(transform=np.array([[1, 0, 0], [0, 0, 1]], dtype="float64").T, isotropic_noise=0.1,
                  mean=np.array([[1, 1, 1]], dtype="float64").T),
        PPCAModel(transform=np.array([[1, 1, 0], [1, 0, 1]], dtype="float64").T, isotropic_noise=0.1,
                  mean=np.array([[0, 1, 0]], dtype="float64").T),
    ],
    log_weights=np.log([0.33333, 0.66667]),
)
sample = truth.sample(500, 0.1)
for n_models in (1, 2, 3):
    model = PPCAMixTrainer(sample).train(n_models=n_models, state_size=2, n_iters=30, quiet=True)
    print(f"{n_models} component(s): llk = {model.llk(sample):.2f}, weights = {np.round(model.weights, 3)}")
print("smoothed:", model.smooth(sample).numpy()[:2])
print("extrapolated:", model.extrapolate(sample).numpy()[:2])
print("cluster posteriors:", np.exp(model.infer_cluster(sample)[:3]))
