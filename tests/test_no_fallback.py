"""There is no CPU or PyTorch path behind the operator surface: every op refuses CPU inputs with a RuntimeError
that says so, and a missing ``libnfst_b200.so`` is an error at load time -- never a silent fallback (the reference
interface these stand in for is listed in ``nfst_b200/__init__.py`` and INTEGRATION.md)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import joint, ops, sampler, scorer
from oracle import lattice_oracle as lo
from tests.lattice_gen import PAD, random_mark_lattice

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = 24


def _cpu_batch():
    rng = np.random.default_rng(0)
    tabs = [random_mark_lattice(rng, 6, V)[1] for _ in range(3)]
    tr = torch.from_numpy(lo.collate_pad(tabs, PAD))
    return tr, nb.pack_dense(None, tr)


def test_every_op_refuses_cpu_inputs():
    tr, p = _cpu_batch()
    theta = torch.zeros(V)
    beta = torch.ones(p.n_states)
    state = p.start_state.to(torch.int32)
    calls = {
        "lattice_forward": lambda: ops.lattice_forward(p, theta=theta),
        "lattice_pull": lambda: ops.lattice_pull(p, theta=theta),
        "lattice_backward": lambda: ops.lattice_backward(p, theta=theta, want_beta=True),
        "lattice_forward_backward": lambda: ops.lattice_forward_backward(p, theta=theta),
        "lattice_log_partition": lambda: ops.lattice_log_partition(p, theta=theta),
        "lattice_viterbi": lambda: ops.lattice_viterbi(p, theta=theta),
        "lattice_viterbi_padded": lambda: ops.lattice_viterbi_padded(p, theta=theta, pad_label=PAD),
        "lattice_beta_hat": lambda: ops.lattice_beta_hat(p, torch.zeros(V, 4), torch.zeros(4, 4), torch.zeros(4)),
        "compute_beta": lambda: ops.compute_beta(tr != 0, tr, theta),
        "LatticeBetaScorer.compute_beta": lambda: _scorer(tr).compute_beta(theta),
        "LatticeWalker": lambda: sampler.LatticeWalker(p, 2, beta, PAD),
        "walk_step": lambda: sampler.walk_step(p, 1, state, torch.zeros(3, V), beta, PAD),
        "sample_paths": lambda: sampler.sample_paths(p, 2, theta=theta),
        "stripping_pad": lambda: sampler.stripping_pad(torch.full((2, 3), PAD, dtype=torch.int64), PAD),
        "ExactJointProb.forward": lambda: joint.ExactJointProb(theta)(p, None),
    }
    for name, call in calls.items():
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA|NVIDIA"):
            call()
            pytest.fail(f"{name} accepted CPU inputs")


def _scorer(tr):
    s = scorer.LatticeBetaScorer()
    s.set_masks(tr != 0, tr)
    s.set_k(2)
    return s


def test_reference_style_argument_errors():
    s = scorer.LatticeBetaScorer()
    with pytest.raises(AssertionError):  # scorers.py:878-879: both tables are [B, S, V]
        s.set_masks(torch.zeros(4, V, dtype=torch.bool), torch.zeros(4, V, dtype=torch.int64))
    with pytest.raises(AssertionError):  # compute_beta before set_masks (cf. scorers.py:1038)
        s.packed
    tr, _ = _cpu_batch()
    s.set_masks(tr != 0, tr)
    with pytest.raises(ValueError, match="theta"):
        s.compute_beta()


def test_missing_library_is_an_error_not_a_fallback(tmp_path):
    code = ("import nfst_b200._lib as l\n"
            "try:\n    l.load()\nexcept RuntimeError as e:\n    assert 'no CPU or PyTorch fallback' in str(e), e\n    print('raised')\n")
    env = dict(os.environ, NFST_LIB=str(tmp_path / "absent.so"), PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert out.returncode == 0 and "raised" in out.stdout, out.stderr
