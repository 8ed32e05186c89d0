"""The operator surface ``JointProb`` offers its callers, computed exactly.

``JointProb.forward(num_emission, num_transition, denom_emission, denom_transition, gs, ps, return_samples)``
(``src/modules/lightning.py:442-480``) returns ``(num_prob[B], denom_prob[B] = 0)`` -- ``num_prob`` is the k-sample
importance-weighted estimate of the log-marginal of the mark lattice (``log_marginalize`` ``:408-440`` ->
``Estimators.iwae``, ``estimatros.py:32-44``) -- and, with ``return_samples``, the best of the k samples
(``:474-479``); ``decode_from_npz`` (``:647-658``, called by ``decode/decoder.py:77-79``) wraps it for one file.

For an arc-factored model (``WFSTScorer``, ``scorers.py:1663-1687``: path score = sum of ``theta[mark]``, ``pad``
scores 0) both quantities have exact counterparts on the lattice, and this adapter returns them with the
reference's shapes: the log-marginal is ``logZ - theta[bos]`` (the sampler consumes ``bos`` as its first input,
``scorers.py:230-231``, so the reference's samples and their scores never contain it), and "the best sample" is the
Viterbi path (labels after ``bos``, ``eos`` included, cf. ``samplers.py:304-307``).  ``num_prob`` carries autograd:
its gradient w.r.t. ``theta`` is the expected label count (arc posteriors summed by label).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .data import load_fsa_from_npz
from .ops import lattice_log_partition, lattice_viterbi_padded
from .pack import PackedLattices, pack_dense


class ExactJointProb(torch.nn.Module):
    """Exact stand-in for the numerator path of ``JointProb`` with an arc-factored model.

    ``theta``: ``[V]`` label scores (a Parameter, a tensor, or a callable returning one -- e.g. the
    ``wfst_dist_over_next`` of the reference's ``WFSTScorer``, ``scorers.py:1681-1683``)."""

    def __init__(self, theta, bos: int = 1, eos: int = 2, pad: int = 3, device=None):
        super().__init__()
        if isinstance(theta, torch.nn.Parameter):
            self.theta = theta
        elif callable(theta):
            self._theta_fn = theta
        else:
            self.register_buffer("theta", torch.as_tensor(theta, dtype=torch.float32))
        self.__bos__, self.__eos__, self.__pad__ = int(bos), int(eos), int(pad)
        self._device = device

    def _theta(self) -> torch.Tensor:
        fn = getattr(self, "_theta_fn", None)
        return fn() if fn is not None else self.theta

    def _packed(self, emission, transition) -> PackedLattices:
        if isinstance(emission, PackedLattices):
            return emission
        dev = self._device or (transition.device if transition.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        return pack_dense(emission.to(dev), transition.to(dev), weighted=emission.is_floating_point())

    def log_marginalize(self, emission, transition, proposal=None, x=None, y=None, k=None):
        """(log_marginalized[B], None, None, None): the reference's tuple (``lightning.py:440``) with the estimate
        replaced by the exact value; there are no samples, so the other three are None."""
        packed = self._packed(emission, transition)
        theta = self._theta().to(packed.device)
        logz = lattice_log_partition(packed, theta=theta)
        return (logz - theta[self.__bos__]).to(torch.float32), None, None, None

    def forward(self, numerator_emission, numerator_transition, denom_emission=None, denom_transition=None, gs=None, ps=None,
                return_samples: bool = False):
        """``(num_prob[B], denom_prob[B])`` and, with ``return_samples``, the best path's labels: ``[T]`` for a
        batch of one (what ``decode_from_npz`` gets), else ``[B, T]`` padded with ``pad``.  ``numerator_emission``
        may already be a ``PackedLattices`` (cached examples)."""
        packed = self._packed(numerator_emission, numerator_transition)
        num_prob, _, _, _ = self.log_marginalize(packed, None)
        denom_prob = torch.zeros_like(num_prob)  # lightning.py:473
        if not return_samples:
            return num_prob, denom_prob
        # the best path's labels after bos, one padded row per lattice, written on the device; ONE host read (the
        # longest path) fixes T
        _, labels, lengths = lattice_viterbi_padded(packed, theta=self._theta().detach().to(packed.device),
                                                    pad_label=self.__pad__, skip_label=self.__bos__)
        T = int(lengths.max()) if packed.n_lattices else 0
        best = labels[:, :T]
        if packed.n_lattices == 1:
            return num_prob, denom_prob, best[0]
        return num_prob, denom_prob, best.contiguous()

    def decode_from_npz(self, npz_path: str, vocab_size: Optional[int] = None, pad: Optional[int] = None):
        """``(prob, mark)`` of one example file, as ``JointProb.decode_from_npz`` (``lightning.py:647-658``)."""
        single_batch = tuple(torch.from_numpy(np.ascontiguousarray(a)).unsqueeze(0)
                             for a in load_fsa_from_npz(npz_path, None, vocab_size, pad))
        ne, nt = single_batch[0], single_batch[1].to(torch.int64)
        out = self.forward(ne, nt, *single_batch[2:], return_samples=True)
        prob = (out[0] - out[1]).flatten()[0].item()
        return prob, out[2]
