"""nfst_b200 -- B200-native (sm_100a) lattice dynamic programming for neuralized FSTs.

One hot path of steventan0110/nFST, rebuilt from scratch: log-semiring forward/backward
(log-partition, arc posteriors as the autograd gradient of arc scores) and tropical
Viterbi over batched, topologically levelled, CSR-packed lattices.  Hand-written CUDA
behind a C ABI (``include/nfst_b200.h``); no Triton, no CPU fallback.
"""
from . import tiles  # noqa: F401
from .pack import PackedLattices, pack_arcs, pack_dense, dense_arcs  # noqa: F401
from .ops import (  # noqa: F401
    CapturedForwardBackward,
    LatticeLogPartition,
    beta_dense,
    compute_beta,
    lattice_backward,
    lattice_beta_hat,
    lattice_forward,
    lattice_forward_backward,
    lattice_log_partition,
    lattice_viterbi,
    lattice_viterbi_padded,
)

from .sampler import LatticeWalker, sample_paths, stripping_pad, walk_step  # noqa: F401
from .joint import ExactJointProb  # noqa: F401
from . import data  # noqa: F401
from . import construct  # noqa: F401
from .construct import edit_lattices  # noqa: F401

__version__ = "0.1.0"
