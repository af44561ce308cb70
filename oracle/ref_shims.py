"""
ORACLE (test infrastructure) - import the reference's OWN loop code under stubs.

Only usable where /root/reference exists (the authoring container).  Used by
tests/golden/make_golden.py to produce the committed golden vectors and by
tests marked `needs_reference` to re-check the restatement live.  Nothing in
the `-m gpu` tests, smoke() or bench.py goes through here.

Nothing arithmetic lives in any stubbed module: matplotlib,
pytorch_lightning, biotite, huggingface_hub and (for modelling.py only) the
transformers BERT classes are import-time dependencies of files whose
sampling / schedule / noise functions are pure torch + numpy.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FOLDINGDIFF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "foldingdiff", "sampling.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so `import a.b` resolves
    sys.modules[name] = m
    return m


def install() -> None:
    """Idempotently install the stubs and put the reference on sys.path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch

    if "matplotlib" not in sys.modules:
        def _noop(*a, **k):
            return None
        plt = _stub("matplotlib.pyplot", subplots=_noop, figure=_noop, cm=types.SimpleNamespace())
        colors = _stub("matplotlib.colors", LogNorm=object)
        _stub("matplotlib", pyplot=plt, colors=colors)
    if "pytorch_lightning" not in sys.modules:
        class LightningModule(torch.nn.Module):
            pass
        util = _stub("pytorch_lightning.utilities", rank_zero_info=lambda *a, **k: None,
                     rank_zero_only=lambda f: f)
        _stub("pytorch_lightning", LightningModule=LightningModule, utilities=util)
    if "biotite" not in sys.modules:
        pdb = _stub("biotite.structure.io.pdb", PDBFile=object)
        io = _stub("biotite.structure.io", pdb=pdb)
        struc = _stub("biotite.structure", io=io)
        seq = _stub("biotite.sequence", ProteinSequence=object)
        _stub("biotite", structure=struc, sequence=seq)
    try:
        import huggingface_hub  # noqa: F401
    except Exception:  # pragma: no cover
        _stub("huggingface_hub", snapshot_download=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("offline")))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """-> namespace with the reference's sampling, datasets, beta_schedules, utils modules."""
    install()
    from foldingdiff import beta_schedules, utils  # type: ignore
    from foldingdiff import datasets  # type: ignore
    from foldingdiff import sampling  # type: ignore
    return types.SimpleNamespace(sampling=sampling, datasets=datasets,
                                 beta_schedules=beta_schedules, utils=utils)
