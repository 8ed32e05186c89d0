"""CPU restatement of ONE time step of the reference's lattice-constrained sampling loop -- TEST
INFRASTRUCTURE ONLY (imported by tests/ alone; the product path is nfst_b200/csrc/nfst_walk.cu).

Follows the reference line by line on the DENSE tables, like the reference does:
  * scorers.py:586-592   transition row of the (stale, quirk Q9) state, beta_logits = gather(beta, 1, row),
                         final = beta_logits + prefix
  * scorers.py:182-187   pad_masking: the score of the pad column is multiplied by 0
  * scorers.py:340-357   + base vocabulary mask + emission mask; activation = identity
  * scorers.py:1037-1054 emission mask of the ADVANCED state: 0 / -inf (bool tables) or the float row itself
  * samplers.py:251-283  / temperature; Categorical(logits): log_softmax, log_prob, logsumexp ("zs")
  * scorers.py:683-690   next state = transition_k[row, state][symbol]
Pinned against outputs of the unmodified reference in tests/golden/walk_step.npz
(tests/test_walk_golden.py).
"""
from __future__ import annotations

import numpy as np


def walk_step_dense(emission, transition, k, beta, state_old, state_new, prefix, base_mask, pad, temperature, symbols):
    """emission/transition [B, S, V]; beta [B*k, S] real space; state_old/state_new/symbols [N = B*k];
    prefix/base_mask [N, V].  Returns (masked_logits [N, V], log_prob [N], logsumexp [N], next_state [N])."""
    N, V = prefix.shape
    masked = np.empty((N, V), dtype=np.float64)
    logp = np.empty(N)
    logz = np.empty(N)
    nxt = np.empty(N, dtype=np.int64)
    for n in range(N):
        b = n // k
        row = transition[b, state_old[n]]  # scorers.py:586-589 (the state before the previous symbol)
        final = beta[n, row].astype(np.float64) + prefix[n].astype(np.float64)  # :590-592
        final[pad] *= 0.0  # pad_masking, :182-187
        em = emission[b, state_new[n]]
        mask = np.where(em, 0.0, -np.inf) if em.dtype == np.bool_ else em.astype(np.float64)  # :1042-1053
        with np.errstate(invalid="ignore"):
            v = (final + base_mask[n].astype(np.float64) + mask) / temperature  # :357, samplers.py:251
        masked[n] = v
        m = v.max()
        logz[n] = m + np.log(np.exp(v - m).sum()) if np.isfinite(m) else -np.inf
        logp[n] = v[symbols[n]] - logz[n]
        nxt[n] = transition[b, state_new[n]][symbols[n]]  # update_fsa_state, :683-690
    return masked, logp, logz, nxt
