"""DataFrame <-> Dataset adapters (mirror of python/ppca_rs/__init__.py:121-433 in the reference).

A long-format table (one row per (sample key, dimension, metric value)) becomes the dense `n_samples x n_dimensions`
matrix a `Dataset` wants, missing combinations staying NaN (= masked), and model outputs (`smooth`, `extrapolate`,
covariance diagonals ...) go back to the long format with their keys and dimension columns.

Same class names, fields, method names and results as the reference; the implementation is vectorised (key
factorisation + one scatter) instead of a Python loop over groups, because on a GPU-sized dataset the loop would
dominate.  pandas is imported lazily (duck-typed dependency, as in the reference); the polars path needs polars,
which is not installed in this image and is therefore untested here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from .model import Dataset

DIM = "__dim_idx"
SAMPLE = "__sample_idx"


def _as_list(x) -> List[str]:
    return [x] if isinstance(x, str) else list(x)


@dataclass
class DataFrameAdapter:
    """Utility class to facilitate the transformation of DataFrames into Datasets
    (python/ppca_rs/__init__.py:121-146)."""

    keys: List[str]            # columns that uniquely define a sample
    dimensions: List[str]      # columns that define a dimension of the output space
    metric: str                # column that populates the output space
    dimension_idx: Any         # DataFrame: dimensions + "__dim_idx"
    sample_idx: Any            # DataFrame: keys + "__sample_idx"
    dataset: Any               # the mapped Dataset
    origin: str                # "pandas" | "polars"

    # ---------------------------------------------------------------------------------------------
    @classmethod
    def from_pandas(cls, df, *, keys: List[str], dimensions: Optional[List[str]] = None, dimension_idx=None,
                    metric: str, dataset_factory: Callable[[np.ndarray], Any] = Dataset) -> "DataFrameAdapter":
        """python/ppca_rs/__init__.py:148-210.  `dataset_factory` (extension) builds the dataset from the dense
        matrix — `Dataset` (device resident) by default, `HostDataset` for out-of-core training."""
        import pandas as pd

        keys = _as_list(keys)
        if dimension_idx is None:
            if dimensions is None:
                raise ValueError("either `dimensions` or `dimension_idx` must be given")
            dimensions = _as_list(dimensions)
            # reproducible dimension numbering: distinct dimension tuples in sorted order (:163-171)
            dimension_idx = df[dimensions].drop_duplicates().sort_values(dimensions).reset_index(drop=True)
            dimension_idx.insert(0, DIM, np.arange(len(dimension_idx), dtype=np.int64))
        elif dimensions is None:
            dimensions = [col for col in dimension_idx.columns if col != DIM]
        else:
            dimensions = _as_list(dimensions)

        # rows whose dimensions are unknown to the index are dropped (inner join, :178)
        joined = df[[*keys, *dimensions, metric]].merge(dimension_idx[[DIM, *dimensions]], on=dimensions)
        # samples are numbered in sorted key order, like groupby(keys) iterates (:178,192-196)
        uniq = joined[keys].drop_duplicates().sort_values(keys).reset_index(drop=True)
        uniq[SAMPLE] = np.arange(len(uniq), dtype=np.int64)
        rows = joined.merge(uniq, on=keys)[SAMPLE].to_numpy()
        cols = joined[DIM].to_numpy().astype(np.int64)
        dense = np.full((len(uniq), len(dimension_idx)), np.nan)
        if len(joined):
            dense[rows, cols] = joined[metric].to_numpy(dtype=np.float64)
        sample_idx = uniq[[*keys, SAMPLE]]
        return cls(keys, dimensions, metric, dimension_idx, sample_idx, dataset_factory(dense), origin="pandas")

    @classmethod
    def from_polars(cls, df, *, keys: List[str], dimensions: Optional[List[str]] = None, dimension_idx=None,
                    metric: str, dataset_factory: Callable[[np.ndarray], Any] = Dataset) -> "DataFrameAdapter":
        """python/ppca_rs/__init__.py:212-267 (needs polars)."""
        import polars as pl

        keys = _as_list(keys)

        def with_rows(frame, name):
            return frame.with_row_index(name) if hasattr(frame, "with_row_index") else frame.with_row_count(name)

        if dimension_idx is None:
            if dimensions is None:
                raise ValueError("either `dimensions` or `dimension_idx` must be given")
            dimensions = _as_list(dimensions)
            dimension_idx = with_rows(df.select(dimensions).unique(maintain_order=False).sort(dimensions), DIM)
        elif dimensions is None:
            dimensions = [col for col in dimension_idx.columns if col != DIM]
        else:
            dimensions = _as_list(dimensions)
        joined = df.select([*keys, *dimensions, metric]).join(dimension_idx, on=dimensions)
        uniq = with_rows(joined.select(keys).unique(maintain_order=False).sort(keys), SAMPLE)
        located = joined.join(uniq, on=keys)
        dense = np.full((len(uniq), len(dimension_idx)), np.nan)
        if len(located):
            dense[located[SAMPLE].to_numpy().astype(np.int64), located[DIM].to_numpy().astype(np.int64)] = \
                located[metric].to_numpy().astype(np.float64)
        sample_idx = uniq.select([*keys, SAMPLE])
        return cls(keys, dimensions, metric, dimension_idx, sample_idx, dataset_factory(dense), origin="polars")

    # ---------------------------------------------------------------------------------------------
    def description(self) -> "DataFrameAdapterDescription":
        """A data-free, serialisable description of this adapter (:269-295)."""
        if self.origin == "pandas":
            ordered = self.dimension_idx.sort_values(DIM)
            table = [[row[col] for col in self.dimensions] for _, row in ordered.iterrows()]
        elif self.origin == "polars":
            ordered = self.dimension_idx.sort(DIM)
            table = [[ordered[col][i] for col in self.dimensions] for i in range(len(ordered))]
        else:
            raise Exception(f"Unknown origin {self.origin}")
        table = [[v.item() if hasattr(v, "item") else v for v in row] for row in table]
        return DataFrameAdapterDescription(keys=self.keys, dimensions=self.dimensions, metric=self.metric,
                                           dimension_idx=table)

    def convert_dataset(self, dataset, *, column_name: str):
        return self.convert_datasets({column_name: dataset})

    def convert_datasets(self, datasets: Dict[str, Any]):
        """Datasets shaped like `self.dataset` back to the long format: one row per (sample, dimension) with the key
        and dimension columns and one column per dataset (:300-352)."""
        n, d = len(self.sample_idx), len(self.dimension_idx)
        data = {}
        for name, ds in datasets.items():
            flat = np.asarray(ds.numpy(), dtype=np.float64).reshape(-1)
            if flat.shape[0] != n * d:
                raise ValueError(f"dataset {name!r} has {flat.shape[0]} entries, adapter expects {n} x {d}")
            data[name] = flat
        s_of = np.repeat(np.arange(n, dtype=np.int64), d)
        d_of = np.tile(np.arange(d, dtype=np.int64), n)
        if self.origin == "pandas":
            import pandas as pd

            samples = self.sample_idx.sort_values(SAMPLE)
            dims = self.dimension_idx.sort_values(DIM)
            out = {}
            for col in self.keys:
                out[col] = samples[col].to_numpy()[s_of]
            for col in self.dimensions:
                out[col] = dims[col].to_numpy()[d_of]
            out.update(data)
            return pd.DataFrame(out)
        elif self.origin == "polars":
            import polars as pl

            frame = pl.DataFrame({**data, SAMPLE: s_of.astype(np.uint32), DIM: d_of.astype(np.uint32)})
            return (frame.join(self.dimension_idx.with_columns(pl.col(DIM).cast(pl.UInt32)), on=DIM)
                    .join(self.sample_idx.with_columns(pl.col(SAMPLE).cast(pl.UInt32)), on=SAMPLE)
                    .select([*self.keys, *self.dimensions, *data.keys()]))
        raise Exception(f"Unknown origin {self.origin}")


@dataclass
class DataFrameAdapterDescription:
    """How to adapt a DataFrame to a Dataset, free of actual data: suitable for serialising next to a trained model
    (:355-433)."""

    keys: List[str]
    dimensions: List[str]
    metric: str
    dimension_idx: List[List]   # dimension_idx[i] = values of `dimensions` for output index i

    def _columns(self) -> dict:
        return {DIM: np.arange(len(self.dimension_idx), dtype=np.int64),
                **{dim: [item[i] for item in self.dimension_idx] for i, dim in enumerate(self.dimensions)}}

    @property
    def dimension_idx_pandas(self) -> Any:
        import pandas as pd

        return pd.DataFrame(self._columns())

    @property
    def dimension_idx_polars(self) -> Any:
        import polars as pl

        cols = self._columns()
        cols[DIM] = cols[DIM].astype(np.uint32)
        return pl.DataFrame(cols)

    @classmethod
    def from_json(cls, value: dict) -> "DataFrameAdapterDescription":
        return cls(**value)

    def to_json(self) -> dict:
        return {"keys": self.keys, "dimensions": self.dimensions, "metric": self.metric,
                "dimension_idx": self.dimension_idx}

    def adapt_pandas(self, df, **kw) -> DataFrameAdapter:
        return DataFrameAdapter.from_pandas(df, keys=self.keys, dimension_idx=self.dimension_idx_pandas,
                                            metric=self.metric, **kw)

    def adapt_polars(self, df, **kw) -> DataFrameAdapter:
        return DataFrameAdapter.from_polars(df, keys=self.keys, dimension_idx=self.dimension_idx_polars,
                                            metric=self.metric, **kw)
