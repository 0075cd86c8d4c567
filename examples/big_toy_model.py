"""examples/big_toy_model.py scale (BASELINE configs[1] shape): d = 200, k = 16, 20 % missing."""
import _alias  # noqa: F401
import numpy as np
from ppca_rs import PPCAModel, PPCATrainer

rng = np.random.default_rng(0)
truth = PPCAModel(transform=(rng.random((200, 16)) < 0.1).astype("float64"), isotropic_noise=0.1, mean=np.zeros((200, 1)))
sample = truth.sample(100_000, 0.2)
model = PPCATrainer(sample).train(state_size=16, n_iters=24)
print(model)
