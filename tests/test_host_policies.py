"""Host-side policies that decide HOW a batch runs (no GPU needed): the state dtype "auto" picks, the launch groups of
deep and shallow tile-stream lattices, and the one-launch / two-launch choice for small-lattice groups."""
import torch

from nfst_b200 import ops, synth
from nfst_b200 import tiles as T
from nfst_b200.pack import concat_packed


def test_auto_state_dtype_follows_depth_and_layout():
    # small-lattice / CSR kernels form posteriors from float32 log-values rounded at every level: float64 from 32 levels
    assert ops.F64_DEPTH_PLAIN == 32 and ops.F64_DEPTH == T.F64_LEVELS == 96
    p, _ = synth.transliteration_batch(4, seed=0).pack()  # ~40-60 levels, small-lattice kernel
    assert p.max_levels > 32 and ops.resolve_state_dtype(p) == torch.float64
    p, _ = synth.random_dag_batch(2, 1500, levels=24, seed=1).pack()  # 24 levels, narrow: small-lattice kernel
    assert not p.has_columns and ops.resolve_state_dtype(p) == torch.float32
    p, _ = synth.random_dag_batch(2, 20_000, levels=64, seed=1).pack()  # column-major (conditionals + flow): float32 up to 96
    assert p.has_tiles and ops.resolve_state_dtype(p) == torch.float32
    p, _ = synth.random_dag_batch(2, 40_000, levels=130, seed=1).pack()
    assert p.has_tiles and ops.resolve_state_dtype(p) == torch.float64
    assert ops.resolve_state_dtype(p, torch.float32) == torch.float32  # the caller's choice wins


def test_deep_and_shallow_tile_lattices_do_not_share_a_launch():
    shallow, _ = synth.random_dag_batch(2, 20_000, levels=64, seed=2).pack()
    deep, _ = synth.random_dag_batch(2, 40_000, levels=130, seed=3).pack()
    assert shallow.has_tiles and deep.has_tiles
    assert shallow.groups[0].block_threads == deep.groups[0].block_threads  # same warp count: one launch before the split
    both = concat_packed([shallow, deep])
    tile_groups = [g for g in both.groups if g.tiles]
    assert len(tile_groups) == 2
    levels = both.n_levels.tolist()
    for g in tile_groups:
        depth = {levels[b] > T.F64_LEVELS for b in g.ids.tolist()}
        assert len(depth) == 1, "a launch group holds deep lattices (float64 DP ring) or shallow ones (float32 ring), not both"
        assert g.n_levels == max(levels[b] for b in g.ids.tolist())


def test_small_groups_run_in_one_launch_only_while_resident_at_once():
    few, _ = synth.transliteration_batch(32, seed=0).pack()
    g = few.groups[0]
    assert g.small_max_arcs > 0 and ops._small_fused_fits(g, few.vocab, torch.float64)
    many = concat_packed([synth.transliteration_batch(512, seed=4 + o).pack()[0] for o in range(0, 4096, 512)])
    g = many.groups[0]
    assert g.n == 4096 and not ops._small_fused_fits(g, many.vocab, torch.float32)


def test_algorithmic_bytes_are_the_survey_figures():
    """fwd+bwd = 20 B per arc + 20 B per state; Viterbi = 8 A + 12 S + 8 |path| -- the numerators of
    every roofline fraction bench.py prints."""
    import numpy as np
    import torch

    import nfst_b200 as nb
    from oracle import lattice_oracle as lo
    from tests.lattice_gen import PAD, random_mark_lattice

    rng = np.random.default_rng(0)
    tabs = [random_mark_lattice(rng, 7, 24)[1] for _ in range(3)]
    p = nb.pack_dense(None, torch.from_numpy(lo.collate_pad(tabs, PAD)))
    A, S = p.n_arcs, p.n_states
    assert A > 0 and p.algorithmic_bytes_fwd_bwd() == 20 * A + 20 * S
    assert p.algorithmic_bytes_viterbi() == 8 * A + 12 * S
    assert p.algorithmic_bytes_viterbi(17) == 8 * A + 12 * S + 8 * 17
