"""What the B200 engine adds around the reference's calls (no reference equivalent):

* `Dataset.from_device`: a torch / CuPy / DLPack array that already lives on the GPU becomes a Dataset without touching the
  host (the reference copies numpy -> Rust element by element, src/python_bindings.rs:41-54);
* `GeneratedDataset`: a synthetic job larger than GPU memory, regenerated chunk by chunk inside every EM step;
* posterior samples, drawn on the device, from the same inferred object the reference returns.
"""
import _alias  # noqa: F401  (installs ppca_rs_b200 under the reference's module names)
import numpy as np
import torch
from ppca_rs import Dataset, PPCATrainer
from ppca_rs.ppca_rs import PPCAModel

import ppca_rs_b200 as engine

rng = np.random.default_rng(0)
truth = PPCAModel(0.1, rng.standard_normal((12, 3)), np.zeros(12))
sample = truth.sample(20_000, 0.2)

# 1. the same data as a CUDA tensor: ingested in place, same model as the host-built dataset
on_gpu = sample.to_torch()
ds = Dataset.from_device(on_gpu, torch.ones(len(sample), dtype=torch.float64, device="cuda"))
model = PPCATrainer(ds).train(state_size=3, n_iters=25, quiet=True)
print("trained from a device tensor: sigma = %.4f, llk per sample = %.4f" % (model.isotropic_noise, model.llk(ds) / len(ds)))

# 2. posterior samples of the missing entries, on the device
draws = model.infer(ds).posterior_sampler().sample()
print("posterior draw:", draws.numpy().shape, "finite:", bool(np.isfinite(draws.numpy()).all()))

# 3. a job that is never stored: rows [1000, 51000) of the synthetic dataset, regenerated inside each EM step
job = engine.GeneratedDataset(50_000, 64, 8, sigma_true=0.1, mask_prob=0.3, seed=7, row_begin=1000)
m = PPCAModel.init(8, job.materialize().chunks(50).__next__())    # initialise from the first 1000 rows
for _ in range(10):
    m = m.iterate(job)
print("out-of-core EM over regenerated chunks: sigma = %.4f (truth 0.1)" % m.isotropic_noise)
