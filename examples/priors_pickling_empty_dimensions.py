"""Priors (MAP estimation), pickling and Dataset.empty_dimensions (cf. examples/priors.py, pickling.py, empty_dimensions.py)."""
import _alias  # noqa: F401
import pickle

import numpy as np
from ppca_rs import Dataset, PPCAModel, Prior

truth = PPCAModel(transform=np.array([[1, 1, 0], [1, 0, 1]], dtype="float64").T, isotropic_noise=0.1,
                  mean=np.array([[0, 1, 0]], dtype="float64").T)
sample = truth.sample(100, mask_prob=0.2)
prior = (Prior().with_isotropic_noise_prior(100.0, 100.0)
         .with_mean_prior(np.array([1.0, 0.0, 1.0]), 0.0001 * np.eye(3)))
model = PPCAModel.init(2, sample)
for _ in range(100):
    model = model.iterate_with_prior(sample, prior)
model = model.to_canonical()
print(model, model.isotropic_noise)

copy = pickle.loads(pickle.dumps(model))
assert np.array_equal(copy.transform, model.transform) and copy.isotropic_noise == model.isotropic_noise
print(copy)

dataset = Dataset(np.array([[1.0, 1.0, np.nan], [1.0, 1.0, np.nan]]), weights=np.array([1.0, 2.0]))
print("empty dimensions:", dataset.empty_dimensions())
