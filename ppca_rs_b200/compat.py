"""Opt-in aliasing so unmodified `ppca_rs` user code runs on the B200 engine:

    import ppca_rs_b200.compat as compat; compat.install_as_ppca_rs()
    from ppca_rs import Dataset, PPCATrainer          # python/ppca_rs/__init__.py names
    from ppca_rs.ppca_rs import PPCAModel             # the pyO3 extension module's names (src/python_bindings.rs:15-26)
"""
from __future__ import annotations

import sys
import types

import ppca_rs_b200 as _pk

# classes the reference registers in its extension module `ppca_rs.ppca_rs`
_NATIVE = ("Dataset", "Prior", "PPCAModel", "PPCAMix", "InferredMasked", "InferredMaskedMix")


def install_as_ppca_rs(force: bool = False) -> None:
    """Registers `ppca_rs` and `ppca_rs.ppca_rs` in sys.modules as views of this package.  Refuses to shadow an already
    imported real `ppca_rs` unless `force`."""
    existing = sys.modules.get("ppca_rs")
    if existing is not None and getattr(existing, "__ppca_b200_alias__", False) is False and not force:
        raise ImportError("a real `ppca_rs` is already imported; pass force=True to shadow it")
    top = types.ModuleType("ppca_rs", "alias of ppca_rs_b200 (B200 engine)")
    top.__ppca_b200_alias__ = True
    for name in _pk.__all__:
        setattr(top, name, getattr(_pk, name))
    native = types.ModuleType("ppca_rs.ppca_rs", "alias of the ppca_rs_b200 classes the pyO3 module exports")
    native.__ppca_b200_alias__ = True
    for name in _NATIVE:
        setattr(native, name, getattr(_pk, name))
    top.ppca_rs = native
    top.__path__ = []  # lets `import ppca_rs.ppca_rs` resolve through sys.modules
    sys.modules["ppca_rs"] = top
    sys.modules["ppca_rs.ppca_rs"] = native
