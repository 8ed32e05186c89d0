"""Import harness for the UNMODIFIED reference (steventan0110/nFST) -- TEST INFRASTRUCTURE ONLY.

This module is the only place in the repo that touches ``/root/reference``.  It is used
by ``tests/golden/make_golden.py`` (run in the build container, where the reference is
mounted) to generate the committed golden vectors under ``tests/golden/``; nothing in
the product path, in ``-m gpu`` tests, in ``smoke()`` or in ``bench.py`` imports it --
``/root/reference`` does not exist on the GPU box.

The reference is pure Python but imports three packages that are absent from this image
(``mfst``, ``pynini`` -- both OpenFst wrappers -- and ``bidict``).  None of them is used
by the lattice dynamic programme (reference ``src/modules/scorers.py:692-856``), so they
are stubbed in ``sys.modules`` before ``src.*`` is imported (recipe recorded in
SURVEY.md, Appendix A).  No reference source is copied: the reference's own functions
are called as they are.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NFST_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "modules"))


def _install_stubs() -> None:
    if "bidict" not in sys.modules:
        m = types.ModuleType("bidict")

        class bidict(dict):  # noqa: N801 - mirrors the package's class name
            """dict with an ``.inverse`` view (used at preprocess_util.py:24-28,61)."""

            @property
            def inverse(self):
                return {v: k for k, v in self.items()}

        m.bidict = bidict
        sys.modules["bidict"] = m
    if "mfst" not in sys.modules:
        m = types.ModuleType("mfst")

        class _Any:
            def __init__(self, *a, **k):
                pass

            def create_from_string(self, *a, **k):
                return self

            def __getattr__(self, name):
                # permissive: any attribute is a callable returning self
                def _f(*a, **k):
                    return self

                return _f

        class FST(_Any):
            pass

        class BooleanSemiringWeight(_Any):
            pass

        class AbstractSemiringWeight(_Any):
            pass

        m.FST = FST
        m.BooleanSemiringWeight = BooleanSemiringWeight
        m.AbstractSemiringWeight = AbstractSemiringWeight
        sys.modules["mfst"] = m
    if "pynini" not in sys.modules:
        m = types.ModuleType("pynini")
        m.Fst = object
        m.Weight = object
        sys.modules["pynini"] = m


def load():
    """Import the reference and return a namespace with the classes on the path.

    Vocabulary ids follow the shipped pipeline (``fsm/tr.py:242-247`` +
    ``preprocess_util.py:42-53``): bos=1, eos=2, pad=3, then the three mark names that
    ``FSAGRUScorer.__init__`` looks up (``scorers.py:974-976``), then the label alphabet.
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.util.preprocess_util import Vocab  # type: ignore

        for w in ("<bos>", "<eos>", "<pad>", "input-mark", "output-mark", "insertion-mark"):
            Vocab.add_word(w)
        from src.modules.scorers import FSAGRUScorer, WFSTScorer  # type: ignore
        from src.modules.samplers import Sampler  # type: ignore
        from src.modules.estimatros import Estimators  # type: ignore

    ns = types.SimpleNamespace(
        Vocab=Vocab,
        FSAGRUScorer=FSAGRUScorer,
        WFSTScorer=WFSTScorer,
        Sampler=Sampler,
        Estimators=Estimators,
        bos=Vocab.lookup("<bos>"),
        eos=Vocab.lookup("<eos>"),
        pad=Vocab.lookup("<pad>"),
    )
    _loaded = ns
    return ns


def make_scorer(hid_dim: int, vocab_size: int, *, seed: int, zero_wh: bool, double: bool = True):
    """Build the reference's beta-capable proposal scorer (scorers.py:920-993)."""
    import torch
    import warnings

    ns = load()
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ns.FSAGRUScorer(
            hid_dim,
            vocab_size,
            pad=ns.pad,
            bos=ns.bos,
            eos=ns.eos,
            use_beta=True,
            max_length=64,
            dropout=0.0,
        ).eval()
    if zero_wh:
        with torch.no_grad():
            m.Wh.zero_()
    if double:
        m = m.double()
    return m


def arc_theta(m):
    """Per-label arc score when Wh == 0: theta[l] = W . tanh(Wx e_l + b) (scorers.py:732-738)."""
    import torch

    with torch.no_grad():
        e = m.embeddings.weight  # [V, H]
        return (m.W @ torch.tanh(m.Wx @ e.T + m.beta_bias[:, None]))[0]


class default_dtype:
    """Context manager: run the reference with torch's default dtype switched.

    The reference allocates its work tensors with ``torch.zeros(...)`` (default dtype,
    ``scorers.py:699-701,780-789``), so a float64 run needs the default switched as well
    as the module cast.
    """

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        import torch

        self.prev = torch.get_default_dtype()
        torch.set_default_dtype(self.dtype)

    def __exit__(self, *exc):
        import torch

        torch.set_default_dtype(self.prev)
        return False
