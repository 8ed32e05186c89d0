"""Host-side mirror of the reference's beta interface (``FSAGRUScorer``, scorers.py:858-918).

The reference module is driven as ``set_masks(emission, transition)`` -> ``set_k(k)`` ->
``compute_beta()`` (called from ``Sampler.stateful_sample``, samplers.py:196-198) and
returns real-space ``beta[B*k, S]``.  ``LatticeBetaScorer`` offers the same three calls on
top of the CUDA path; ``patch_compute_beta`` swaps ``compute_beta`` on an existing reference
module for the Wh = 0 regime, where the arc weight is a function of the arc label alone:

    theta[l] = W . tanh(Wx e_l + b)          (scorers.py:732-738 with Wh = 0)

so that beta is the log-semiring backward pass with arc score theta[label].
"""
from __future__ import annotations

import types
from typing import Optional

import torch

from .ops import beta_dense, lattice_backward
from .pack import PackedLattices, pack_dense


def label_scores(embeddings: torch.Tensor, Wx: torch.Tensor, W: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """theta[V] = W . tanh(Wx e_l + b): the reference's message weight when Wh = 0.  A [V,H]
    x [H,H] product -- the neural arc scorer, outside the lattice path."""
    return (torch.tanh(embeddings @ Wx.T + bias) @ W.reshape(-1)).to(torch.float32)


class LatticeBetaScorer:
    """set_masks / set_k / compute_beta with the reference's argument meaning and errors."""

    def __init__(self, theta: Optional[torch.Tensor] = None):
        self.theta = theta
        self.emission = self.transition = None
        self.k = 1
        self._packed: Optional[PackedLattices] = None

    def set_masks(self, emission: torch.Tensor, transition: torch.Tensor):
        assert len(emission.shape) == 3  # scorers.py:878
        assert len(transition.shape) == 3  # scorers.py:879
        self.emission, self.transition = emission.contiguous(), transition.contiguous()
        self._packed = None  # packed lazily, once per batch, reused by every compute_beta call

    def set_k(self, k: int):
        self.k = int(k)  # the k-fold expansion of scorers.py:887-918 is never materialised here

    @property
    def packed(self) -> PackedLattices:
        if self._packed is None:
            if self.transition is None:
                raise AssertionError("set_masks() must be called first")  # cf. scorers.py:1038
            self._packed = pack_dense(self.emission, self.transition, weighted=self.emission.is_floating_point())
        return self._packed

    def compute_beta(self, theta: Optional[torch.Tensor] = None) -> torch.Tensor:
        """real-space beta[B*k, S] float32 (scorers.py:854: ``beta.repeat_interleave(k)``)."""
        theta = self.theta if theta is None else theta
        if theta is None:
            raise ValueError("compute_beta needs theta[V]")
        r = lattice_backward(self.packed, None, theta, want_beta=True)
        return beta_dense(self.packed, r["beta"], self.k)


def patch_compute_beta(module) -> None:
    """Replace ``module.compute_beta`` (a reference ``FSAGRUScorer`` built with
    ``use_beta=True``) by the CUDA path.  Raises if the module's Wh is not zero: the
    beta-hat recurrence (Wh != 0) is a level-stepped computation that this path does not
    cover."""
    if float(module.Wh.detach().abs().max()) != 0.0:
        raise NotImplementedError("compute_beta drop-in covers the Wh = 0 regime only")
    impl = LatticeBetaScorer()

    def compute_beta(self):
        if impl.transition is not self.transition:
            impl.set_masks(self.emission, self.transition)
        impl.set_k(self.k)
        theta = label_scores(self.embeddings.weight.detach(), self.Wx.detach(), self.W.detach(), self.beta_bias.detach())
        return impl.compute_beta(theta)

    module.compute_beta = types.MethodType(compute_beta, module)
