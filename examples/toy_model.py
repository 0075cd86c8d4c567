"""Small synthetic PPCA with masked values (BASELINE configs[0]; cf. the reference's examples/toy_model.py)."""
import _alias  # noqa: F401
import numpy as np
from ppca_rs import PPCAModel

truth = PPCAModel(
    transform=np.array([[1, 1], [0, 1], [0, 1]], dtype="float64"),
    isotropic_noise=0.1,
    mean=np.array([[0], [1], [0]], dtype="float64"),
)
sample = truth.sample(100, mask_prob=0.2)
model = PPCAModel.init(2, sample)
for it in range(100):
    if it % 20 == 0:
        print(f"iteration {it + 1}: llk = {model.llk(sample):.4f}")
    model = model.iterate(sample)
model = model.to_canonical()
print(model, model.singular_values)
inferred = model.infer(sample)
print("posterior std of the first samples:\n", inferred.smoothed_covariances_diagonal(model).numpy()[:3] ** 0.5)
print("extrapolated:\n", model.extrapolate(sample).numpy()[:3])
