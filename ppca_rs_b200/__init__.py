"""ppca_rs_b200 — B200-native EM engine for probabilistic PCA and PPCA mixtures with missing values.

Drop-in for the training / inference path of the `ppca_rs` Python package (python/ppca_rs/__init__.py):
the same `Dataset`, `PPCAModel`, `PPCAMix`, `Prior`, `PPCATrainer`, `PPCAMixTrainer` names and semantics,
with the data-parallel work running as hand-written sm_100a CUDA behind the C ABI in include/ppca_b200.h.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional

import numpy as np

from ._native import Context, NativeError, device_count, get_context, set_context
from .adapter import DataFrameAdapter, DataFrameAdapterDescription
from .model import (
    Dataset,
    DatasetChunks,
    GeneratedDataset,
    HostDataset,
    InferredMasked,
    InferredMaskedMix,
    PosteriorSampler,
    PosteriorSamplerMix,
    PPCAMix,
    PPCAModel,
    Prior,
)

__version__ = "0.1.0"

__all__ = [
    "Dataset", "DatasetChunks", "GeneratedDataset", "HostDataset", "InferredMasked", "InferredMaskedMix", "PosteriorSampler", "PosteriorSamplerMix",
    "PPCAMix", "PPCAModel", "Prior", "PPCATrainer", "PPCAMixTrainer", "TrainMetrics", "DataFrameAdapter",
    "DataFrameAdapterDescription", "Context", "NativeError",
    "device_count", "get_context", "set_context",
]


@dataclass(frozen=True)
class TrainMetrics:
    """python/ppca_rs/__init__.py:14-18."""

    llk: float
    aic: float
    bic: float


def _metrics(llk: float, n_parameters: int, n: int) -> TrainMetrics:
    return TrainMetrics(  # python/ppca_rs/__init__.py:52-57
        llk=llk / n,
        aic=2.0 * (n_parameters - llk) / n,
        bic=(llk - n_parameters * np.log(n)) / n,
    )


@dataclass
class PPCATrainer:
    """A trainer for a PPCA Model over masked data (python/ppca_rs/__init__.py:21-67)."""

    dataset: Dataset

    def train(
        self,
        *,
        start: Optional[PPCAModel] = None,
        prior: Optional[Prior] = None,
        state_size: int,
        n_iters: int = 10,
        metric: Literal["aic", "bic", "llk"] = "aic",
        quiet: bool = False,
    ) -> PPCAModel:
        model = start or PPCAModel.init(state_size, self.dataset)
        for idx in range(n_iters):
            # The reference calls model.llk(dataset) and then model.iterate(dataset) (two passes); the
            # engine's E-step returns the log-likelihood of the input model as a by-product of iterate.
            new_model, llk = model._iterate(self.dataset, prior)
            if not quiet:
                metrics = _metrics(llk, model.n_parameters, len(self.dataset))
                print(f"Masked PPCA iteration {idx + 1}: {metric}={getattr(metrics, metric)}")
            model = new_model
        return model.to_canonical()


@dataclass
class PPCAMixTrainer:
    """A trainer for a PPCA Mixture Model over masked data (python/ppca_rs/__init__.py:70-118)."""

    dataset: Dataset

    def train(
        self,
        *,
        start: Optional[PPCAMix] = None,
        prior: Optional[Prior] = None,
        n_models: int,
        state_size: int,
        n_iters: int = 10,
        metric: Literal["aic", "bic", "llk"] = "aic",
        quiet: bool = False,
    ) -> PPCAMix:
        model = start or PPCAMix.init(n_models, state_size, self.dataset)
        for idx in range(n_iters):
            new_model, llk = model._iterate(self.dataset, prior)
            if not quiet:
                metrics = _metrics(llk, model.n_parameters, len(self.dataset))
                print(f"Masked PPCA mix iteration {idx + 1}: {metric}={getattr(metrics, metric)}")
            model = new_model
        return model.to_canonical()
