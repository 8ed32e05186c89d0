"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It calls the reference's own ``FSAGRUScorer.compute_beta_per_sample`` (scorers.py:692-751),
``compute_beta_parallel`` (:753-856) and ``Estimators.iwae`` (estimatros.py:32-44, with
``WFSTScorer`` scorers.py:1663-1687 as the model and ``Sampler`` samplers.py:137-335 as the
proposal) through ``oracle/ref_harness.py`` and stores inputs + outputs as small ``.npz``
fixtures.  The fixtures are committed; the GPU box never needs the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from tests.lattice_gen import PAD, chain_with_skips, random_mark_lattice  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
V, H = 24, 8


def params_of(m):
    return dict(
        emb=m.embeddings.weight.detach().double().numpy(),
        Wx=m.Wx.detach().double().numpy(),
        Wh=m.Wh.detach().double().numpy(),
        W=m.W.detach().double().numpy(),
        bias=m.beta_bias.detach().double().numpy(),
    )


def gen_per_sample():
    """float64 runs of compute_beta_per_sample: Wh=0 (log-semiring bridge) and Wh!=0."""
    rng = np.random.default_rng(20260101)
    cases = []
    for i, n_inner in enumerate([1, 2, 3, 5, 8, 12, 17, 25, 40]):
        cases.append(random_mark_lattice(rng, n_inner, V, parallel_arcs=False)[1])
    for n_inner in [6, 15, 30]:
        cases.append(random_mark_lattice(rng, n_inner, V, parallel_arcs=True)[1])
    cases.append(chain_with_skips(60, V)[1])
    out = {"n_cases": np.int64(len(cases)), "vocab": np.int64(V), "hid": np.int64(H)}
    m0 = rh.make_scorer(H, V, seed=11, zero_wh=True, double=True)
    m1 = rh.make_scorer(H, V, seed=12, zero_wh=False, double=True)
    out["theta0"] = rh.arc_theta(m0).numpy()
    for k, v in params_of(m1).items():
        out["p1_" + k] = v
    with rh.default_dtype(torch.float64), torch.no_grad():
        for i, tr in enumerate(cases):
            t = torch.from_numpy(tr).int()
            out[f"tr_{i}"] = tr
            out[f"beta0_{i}"] = m0.compute_beta_per_sample(t).numpy()  # real space, float64
            out[f"beta1_{i}"] = m1.compute_beta_per_sample(t).numpy()
    np.savez_compressed(os.path.join(OUT, "beta_per_sample.npz"), **out)
    print("beta_per_sample.npz:", len(cases), "cases")


def gen_parallel():
    """float32 runs of the batched compute_beta_parallel on collate-padded batches (Q5),
    lattices without parallel arcs (Q4); k=2 exercises repeat_interleave (:854)."""
    from oracle.lattice_oracle import collate_pad

    rng = np.random.default_rng(20260202)
    out = {"vocab": np.int64(V), "hid": np.int64(H), "k": np.int64(2)}
    m0 = rh.make_scorer(H, V, seed=21, zero_wh=True, double=False)
    out["theta0"] = rh.arc_theta(m0).double().numpy()
    batches = [[4, 9, 6], [7, 7, 2, 11]]
    out["n_batches"] = np.int64(len(batches))
    for bi, sizes in enumerate(batches):
        tabs = [random_mark_lattice(rng, n, V, parallel_arcs=False)[1] for n in sizes]
        tr = collate_pad(tabs, PAD)
        em = collate_pad([t != 0 for t in tabs], PAD)  # bool array padded with pad id -> True
        with torch.no_grad():
            m0.set_masks(emission=torch.from_numpy(em), transition=torch.from_numpy(tr))
            m0.set_k(2)
            beta = m0.compute_beta()  # [B*k, S] float32 real space
        out[f"tr_{bi}"] = tr
        out[f"n_states_{bi}"] = np.array([t.shape[0] for t in tabs], dtype=np.int64)
        out[f"beta_{bi}"] = beta.numpy()
    np.savez_compressed(os.path.join(OUT, "beta_parallel.npz"), **out)
    print("beta_parallel.npz:", len(batches), "batches")


def gen_iwae():
    """The reference's importance-sampling estimate of the log-marginal with an
    arc-factored model (WFSTScorer) -- a statistical pin for the exact logZ.  The sampler
    consumes bos as its initial input (scorers.py:230-231), so the estimate targets
    logZ - theta[bos]."""
    ns = rh.load()
    rng = np.random.default_rng(20260303)
    out = {"vocab": np.int64(V)}
    m = rh.make_scorer(H, V, seed=31, zero_wh=True, double=False)
    theta = rh.arc_theta(m).double().numpy()
    out["theta"] = theta
    th_t = torch.from_numpy(theta).float()
    model = ns.WFSTScorer(ns.pad, ns.bos, ns.eos, scorer=lambda t: th_t[t])
    sampler = ns.Sampler(m)
    sizes = [5, 10]
    out["n_cases"] = np.int64(len(sizes))
    k, reps = 512, 12
    for i, n in enumerate(sizes):
        em, tr = random_mark_lattice(rng, n, V, parallel_arcs=False)
        sampler.set_masks(transition=torch.from_numpy(tr)[None], emission=torch.from_numpy(em)[None])
        ests = []
        for r in range(reps):
            torch.manual_seed(1000 * i + r)
            with torch.no_grad():
                est, _, _, _ = ns.Estimators.iwae(sampler, model, 1, k, 0, query_args=None)
            ests.append(float(est[0]))
        out[f"tr_{i}"] = tr
        out[f"iwae_{i}"] = np.asarray(ests)
    out["k"] = np.int64(k)
    np.savez_compressed(os.path.join(OUT, "iwae.npz"), **out)
    print("iwae.npz:", len(sizes), "cases")


class _Fixed(torch.nn.Module):
    """stands in for the scorer's beta_scorer MLP: returns a recorded tensor"""

    def __init__(self, value):
        super().__init__()
        self.value = value

    def forward(self, h):
        return self.value


def gen_walk():
    """A few time steps of the reference's own sampling loop (samplers.py:243-297) through
    FSAGRUScorer.left_to_right_score (scorers.py:340-366): masked logits, sampled symbols, log_prob, zs and
    the FSA states before / after.  The network's beta_scorer output is replaced by a recorded random tensor
    (an attribute of the module instance; no reference source is touched), everything else is the reference."""
    from torch.distributions import Categorical

    from oracle.lattice_oracle import collate_pad

    ns = rh.load()
    rng = np.random.default_rng(20260404)
    k, T, steps = 3, 0.8, 7
    m = rh.make_scorer(H, V, seed=41, zero_wh=True, double=False)
    tabs = [random_mark_lattice(rng, n, V, parallel_arcs=False)[1] for n in (4, 7, 5)]
    tr = collate_pad(tabs, PAD)
    em = collate_pad([t != 0 for t in tabs], PAD)
    out = {"vocab": np.int64(V), "k": np.int64(k), "temperature": np.float64(T), "steps": np.int64(steps),
           "pad": np.int64(ns.pad), "bos": np.int64(ns.bos), "tr": tr, "em": em,
           "n_states": np.array([t.shape[0] for t in tabs], dtype=np.int64)}
    torch.manual_seed(4242)
    with torch.no_grad():
        m.set_masks(emission=torch.from_numpy(em), transition=torch.from_numpy(tr))
        m.set_k(k)
        beta = m.compute_beta()  # [B*k, S] real space
        out["beta"] = beta.numpy()
        N = beta.shape[0]
        hx, inp, metadata = m.get_init_states(N, device=torch.device("cpu"))
        for t in range(steps):
            out[f"state_old_{t}"] = metadata["state"].numpy().copy()
            out[f"inp_{t}"] = inp.numpy().copy()
            base = super(ns.FSAGRUScorer, m).mask_out_invalid(inp, metadata)  # vocabulary mask, scorers.py:314-338
            prefix = torch.randn(N, V)
            m.beta_scorer = _Fixed(prefix)
            new_h, masked, upd = m.left_to_right_score(left_h=hx, left_inp=inp, metadata=metadata, beta=beta)
            masked = masked / T
            dist = Categorical(logits=masked)
            sym = dist.sample()
            out[f"state_new_{t}"] = upd["state"].numpy().copy()
            out[f"prefix_{t}"] = prefix.numpy()
            out[f"base_{t}"] = base.numpy().copy()
            out[f"masked_{t}"] = masked.numpy().copy()
            out[f"sym_{t}"] = sym.numpy().copy()
            out[f"logp_{t}"] = dist.log_prob(sym).numpy().copy()
            out[f"zs_{t}"] = torch.logsumexp(masked, dim=1).numpy().copy()
            out[f"next_{t}"] = m.update_fsa_state(sym, upd["state"]).numpy().copy()
            metadata = m.metadata_callback(new_h, sym, upd)
            inp, hx = sym, new_h
    np.savez_compressed(os.path.join(OUT, "walk_step.npz"), **out)
    print("walk_step.npz:", steps, "steps,", N, "rows")


def gen_stateful():
    """ONE FULL CALL of the reference's ``Sampler.stateful_sample`` (samplers.py:182-335) on a real
    ``FSAGRUScorer(use_beta=True)``: ``compute_beta()`` at its top (:196-198), then the whole loop with the module's own
    GRU, ``beta_scorer`` and masks.  Recorded without touching the reference's source: the instance's ``beta_scorer`` is
    wrapped by a module that calls the real one and keeps its output, and the parent class's ``mask_out_invalid`` is
    wrapped the same way for the duration of the call (the vocabulary mask of every step).  Stored: tables, k, the
    module's parameters that define theta, beta, per step {beta_scorer output, vocabulary mask}, the sampled
    sequences and the summed log-probabilities the call returns."""
    from oracle.lattice_oracle import collate_pad

    ns = rh.load()
    rng = np.random.default_rng(20260505)
    k, T = 4, 0.9
    m = rh.make_scorer(H, V, seed=51, zero_wh=True, double=False)
    tabs = [random_mark_lattice(rng, n, V, parallel_arcs=False)[1] for n in (5, 8, 6)]
    tr = collate_pad(tabs, PAD)
    em = collate_pad([t != 0 for t in tabs], PAD)
    prefixes, bases = [], []
    real_scorer = m.beta_scorer

    class Recorder(torch.nn.Module):
        def forward(self, h):
            out = real_scorer(h)
            prefixes.append(out.detach().clone())
            return out

    parent = ns.FSAGRUScorer.__mro__[1]
    orig_mask = parent.mask_out_invalid

    def recording_mask(self, inp, metadata):
        r = orig_mask(self, inp, metadata)
        bases.append(r.detach().clone())
        return r

    sampler = ns.Sampler(m)
    with torch.no_grad():
        m.set_masks(emission=torch.from_numpy(em), transition=torch.from_numpy(tr))
        m.set_k(k)
        N = tr.shape[0] * k
        beta = m.compute_beta().clone()  # what stateful_sample computes at its top (same inputs, deterministic)
        m.beta_scorer = Recorder()
        parent.mask_out_invalid = recording_mask
        try:
            torch.manual_seed(5151)
            summed, seqs, _ = sampler.stateful_sample(N, temperature=T)
        finally:
            if "mask_out_invalid" in parent.__dict__ and parent.__dict__["mask_out_invalid"] is recording_mask:
                parent.mask_out_invalid = orig_mask
            m.beta_scorer = real_scorer
    steps = len(prefixes)
    assert steps == len(bases) == seqs.shape[1] + 1
    out = {"vocab": np.int64(V), "k": np.int64(k), "temperature": np.float64(T), "steps": np.int64(steps), "pad": np.int64(ns.pad),
           "bos": np.int64(ns.bos), "tr": tr, "em": em, "theta": rh.arc_theta(m).numpy(), "beta": beta.numpy(),
           "summed_log_probs": summed.numpy(), "sequences": seqs.numpy()}
    for name, v in params_of(m).items():
        out["p_" + name] = v.astype(np.float32)
    for t in range(steps):
        out[f"prefix_{t}"] = prefixes[t].numpy()
        out[f"base_{t}"] = bases[t].numpy()
    np.savez_compressed(os.path.join(OUT, "stateful_sample.npz"), **out)
    print("stateful_sample.npz:", steps, "steps,", N, "rows; mean log q", float(summed.mean()))


def gen_state_mask():
    """The reference's table-format function ``FSAGRUScorer.get_state_mask_pynini`` (scorers.py:995-1035), unmodified,
    on OpenFst-shaped stand-ins (tests/lattice_gen.FakeFst; ``pynini.Weight`` replaced by FakeWeight in the stub
    module -- pynini itself is absent).  Unweighted only: the weighted branch names ``np.float`` (:1008), which current
    numpy no longer has."""
    from tests.lattice_gen import FakeFst, FakeWeight, random_fst_arrays

    ns = rh.load()
    sys.modules["pynini"].Weight = FakeWeight
    rng = np.random.default_rng(20260404)
    out = {"vocab": np.int64(V), "n_cases": np.int64(8)}
    for i, n in enumerate([3, 4, 5, 7, 10, 16, 25, 40]):
        arrs = random_fst_arrays(rng, n, V)
        em, tr = ns.FSAGRUScorer.get_state_mask_pynini(FakeFst(*arrs), V, PAD, to_numpy=True)
        for k, a in zip(("n", "src", "ilabel", "nextstate", "weight", "is_final"), arrs):
            out[f"{k}_{i}"] = np.asarray(a)
        out[f"emission_{i}"] = em
        out[f"transition_{i}"] = tr
    np.savez_compressed(os.path.join(OUT, "state_mask.npz"), **out)
    print("state_mask.npz: 8 machines")


EDIT_PAIRS = [([7, 8, 9], [10, 11]), ([7], [12, 12, 13]), ([8, 9, 7, 7], [10, 13, 11, 12]), ([], [10]), ([9, 8], [])]


def edit_tables(x, y, sub):
    """dense tables of the edit lattice of (x, y) the way the PRODUCT writes them: host construction (the test oracle of
    nfst_edit_lattice_arcs) -> pack_arcs -> data.packed_to_dense."""
    import nfst_b200 as nb
    from nfst_b200 import data as nd
    from oracle import edit_lattice_oracle as elo

    arcs, n = elo.edit_lattice(x, y, bos=1, eos=2, input_mark=4, output_mark=5, sub_mark=6 if sub else None)
    src, lab, dst = (torch.tensor([a[i] for a in arcs]) for i in (0, 1, 2))
    p = nb.pack_arcs(torch.zeros(len(arcs), dtype=torch.int64), src, dst, lab, torch.tensor([n]), V)
    return nd.packed_to_dense(p, PAD)


def gen_edit_tables():
    """The unmodified reference's compute_beta_per_sample (float64, Wh = 0 and Wh != 0) on edit lattices that THIS repo
    constructs and writes as dense tables (data.packed_to_dense): the reference accepts them, and its beta pins
    construction + table format + DP end to end."""
    m0 = rh.make_scorer(H, V, seed=31, zero_wh=True)
    m1 = rh.make_scorer(H, V, seed=32, zero_wh=False)
    out = {"vocab": np.int64(V), "hid": np.int64(H), "n_cases": np.int64(2 * len(EDIT_PAIRS))}
    for k, v in params_of(m0).items():
        out[f"p0_{k}"] = v
    for k, v in params_of(m1).items():
        out[f"p1_{k}"] = v
    i = 0
    with rh.default_dtype(torch.float64), torch.no_grad():
        for sub in (False, True):
            for x, y in EDIT_PAIRS:
                em, tr = edit_tables(x, y, sub)
                out[f"tr_{i}"] = tr[0].numpy()
                out[f"beta0_{i}"] = m0.compute_beta_per_sample(tr[0].int()).numpy()
                out[f"beta1_{i}"] = m1.compute_beta_per_sample(tr[0].int()).numpy()
                i += 1
    np.savez_compressed(os.path.join(OUT, "edit_tables.npz"), **out)
    print("edit_tables.npz:", i, "lattices")


if __name__ == "__main__":
    if not rh.available():
        raise SystemExit("reference not mounted; golden vectors can only be regenerated in the build container")
    gen_per_sample()
    gen_parallel()
    gen_iwae()
    gen_walk()
    gen_stateful()
    gen_state_mask()
    gen_edit_tables()
