/*
 * lattice_oracle.c -- CPU ORACLE (plain C) for the nFST lattice DP.  TEST INFRASTRUCTURE,
 * NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load this; the nfst_b200 package never does.
 *
 * Restates the reference's per-sample recurrence (steventan0110/nFST,
 * src/modules/scorers.py:692-751) with Wh = 0 in log space, float64:
 *   - graph: arcs are given explicitly (the dense-table edge rule of :704-716 is applied
 *     by the caller); a state with no outgoing arc is an end state with beta = 1
 *     (:715-720; the batched form :795-805 admits several);
 *   - order: Kahn from the end states -- a state fires when all its outgoing arcs have
 *     delivered their message (:741-749);
 *   - message over c --j--> n : exp(theta_j) * beta[n]  (:736-738)  ==> in log space
 *     beta[c] = logsumexp_a (w_a + beta[dst_a]).
 * The forward pass, posteriors and the tropical (Viterbi) pass have no counterpart in
 * the reference ("parity unpinned", see oracle/lattice_oracle.py); they mirror the
 * numpy oracle, which the tests hold them to.
 *
 * Pinned by tests/test_oracle_golden.py::test_c_oracle_matches_reference_golden against
 * the reference's own outputs in tests/golden/beta_per_sample.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double lse2(double m, double s) { return (m == -INFINITY) ? -INFINITY : m + log(s); }

/* Build CSR (ptr, order) of arcs grouped by key[a] in [0, n); stable. */
static void group_by(int n, int na, const int32_t* key, int32_t* ptr, int32_t* order) {
  memset(ptr, 0, sizeof(int32_t) * (size_t)(n + 1));
  for (int a = 0; a < na; ++a) ptr[key[a] + 1]++;
  for (int i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  memcpy(fill, ptr, sizeof(int32_t) * (size_t)n);
  for (int a = 0; a < na; ++a) order[fill[key[a]]++] = a;
  free(fill);
}

/*
 * One lattice.  Kahn order computed from `tails` towards `heads`:
 *   backward pass: tails = dst, heads = src  (fires from the sinks)
 *   forward  pass: tails = src, heads = dst  (fires from the sources)
 * value[h] = logsumexp over arcs a with heads[a]==h of (w[a] + value[tails[a]]).
 * States with no arc in the `heads` role get init[h] (0 for end states / start, -inf
 * otherwise).  Returns the number of states that fired (== n for acyclic input).
 */
static int kahn_lse(int n, int na, const int32_t* heads, const int32_t* tails, const double* w, const double* init,
                    double* value) {
  int32_t* hptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
  int32_t* horder = (int32_t*)malloc(sizeof(int32_t) * (size_t)(na > 0 ? na : 1));
  int32_t* tptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
  int32_t* torder = (int32_t*)malloc(sizeof(int32_t) * (size_t)(na > 0 ? na : 1));
  int32_t* pending = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  group_by(n, na, heads, hptr, horder);
  group_by(n, na, tails, tptr, torder);
  int qh = 0, qt = 0;
  for (int s = 0; s < n; ++s) {
    pending[s] = hptr[s + 1] - hptr[s];
    value[s] = -INFINITY;
    if (pending[s] == 0) {
      value[s] = init[s];
      queue[qt++] = s;
    }
  }
  while (qh < qt) {
    const int t = queue[qh++];
    /* every arc whose tail is t delivers its message to its head */
    for (int i = tptr[t]; i < tptr[t + 1]; ++i) {
      const int h = heads[torder[i]];
      if (--pending[h] == 0) {
        double m = -INFINITY;
        for (int j = hptr[h]; j < hptr[h + 1]; ++j) {
          const int a = horder[j];
          const double v = w[a] + value[tails[a]];
          if (v > m) m = v;
        }
        double s = 0.0;
        if (m > -INFINITY)
          for (int j = hptr[h]; j < hptr[h + 1]; ++j) {
            const int a = horder[j];
            s += exp(w[a] + value[tails[a]] - m);
          }
        value[h] = lse2(m, s);
        queue[qt++] = h;
      }
    }
  }
  free(hptr); free(horder); free(tptr); free(torder); free(pending); free(queue);
  return qt;
}

/* beta, alpha, logZ = beta[start], posteriors.  Any of alpha/post may be NULL.
 * Returns 0, or -1 if the lattice is cyclic. */
int oracle_forward_backward(int n_states, int n_arcs, const int32_t* src, const int32_t* dst, const double* w,
                            int start, double* alpha, double* beta, double* post, double* logz) {
  double* init = (double*)malloc(sizeof(double) * (size_t)(n_states > 0 ? n_states : 1));
  for (int s = 0; s < n_states; ++s) init[s] = 0.0; /* beta = 1 at every end state */
  int fired = kahn_lse(n_states, n_arcs, src, dst, w, init, beta);
  int rc = fired == n_states ? 0 : -1;
  if (logz) *logz = beta[start];
  if (alpha) {
    for (int s = 0; s < n_states; ++s) init[s] = (s == start) ? 0.0 : -INFINITY;
    kahn_lse(n_states, n_arcs, dst, src, w, init, alpha);
    if (post) {
      const double z = beta[start];
      for (int a = 0; a < n_arcs; ++a) {
        const double e = alpha[src[a]] + w[a] + beta[dst[a]] - z;
        post[a] = (e != e) ? 0.0 : exp(e);
      }
    }
  }
  free(init);
  return rc;
}

/* Tropical pass, float32, backward from the sinks; ties -> first arc in input order
 * (arcs must be in scan order: state ascending, label ascending).  bp[s] = arc index or
 * -1.  Returns the best score delta[start]; writes the path (arc indices) and its
 * length. */
float oracle_viterbi_f32(int n_states, int n_arcs, const int32_t* src, const int32_t* dst, const float* w, int start,
                         float* delta, int32_t* bp, int32_t* path, int32_t* path_len) {
  int32_t* optr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_states + 1));
  int32_t* oorder = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_arcs > 0 ? n_arcs : 1));
  int32_t* iptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_states + 1));
  int32_t* iorder = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_arcs > 0 ? n_arcs : 1));
  int32_t* pending = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_states > 0 ? n_states : 1));
  int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_states > 0 ? n_states : 1));
  group_by(n_states, n_arcs, src, optr, oorder);
  group_by(n_states, n_arcs, dst, iptr, iorder);
  int qh = 0, qt = 0;
  for (int s = 0; s < n_states; ++s) {
    pending[s] = optr[s + 1] - optr[s];
    delta[s] = -INFINITY;
    bp[s] = -1;
    if (pending[s] == 0) {
      delta[s] = 0.0f;
      queue[qt++] = s;
    }
  }
  while (qh < qt) {
    const int n = queue[qh++];
    for (int i = iptr[n]; i < iptr[n + 1]; ++i) {
      const int c = src[iorder[i]];
      if (--pending[c] == 0) {
        float best = -INFINITY;
        int arg = -1;
        for (int j = optr[c]; j < optr[c + 1]; ++j) {
          const int a = oorder[j];
          const volatile float cand = w[a] + delta[dst[a]]; /* one fp32 add, no contraction */
          if (arg < 0 || cand > best) {
            best = cand;
            arg = a;
          }
        }
        delta[c] = best;
        bp[c] = arg;
        queue[qt++] = c;
      }
    }
  }
  int k = 0, s = start;
  while (bp[s] >= 0) {
    if (path) path[k] = bp[s];
    ++k;
    s = dst[bp[s]];
  }
  if (path_len) *path_len = k;
  free(optr); free(oorder); free(iptr); free(iorder); free(pending); free(queue);
  return delta[start];
}

/*
 * Batch driver (OpenMP over lattices): arcs grouped by lattice, local state ids, float32
 * scores as the GPU path receives them, float64 arithmetic.  Outputs may be NULL except
 * logz.  Returns the number of cyclic lattices.
 */
int oracle_batch_forward_backward(int n_lattices, const int64_t* state_off, const int64_t* arc_off,
                                  const int32_t* src, const int32_t* dst, const float* w, int start, int want_post,
                                  double* alpha, double* beta, double* post, double* logz, int n_threads) {
  int bad = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
  for (int b = 0; b < n_lattices; ++b) {
    const int64_t a0 = arc_off[b], s0 = state_off[b];
    const int na = (int)(arc_off[b + 1] - a0), ns = (int)(state_off[b + 1] - s0);
    double* wd = (double*)malloc(sizeof(double) * (size_t)(na > 0 ? na : 1));
    for (int a = 0; a < na; ++a) wd[a] = (double)w[a0 + a];
    double* be = beta ? beta + s0 : (double*)malloc(sizeof(double) * (size_t)ns);
    double* al = NULL;
    if (want_post) al = alpha ? alpha + s0 : (double*)malloc(sizeof(double) * (size_t)ns);
    double* po = (want_post && post) ? post + a0 : NULL;
    bad += oracle_forward_backward(ns, na, src + a0, dst + a0, wd, start, al, be, po, &logz[b]) ? 1 : 0;
    if (!beta) free(be);
    if (want_post && !alpha) free(al);
    free(wd);
  }
  return bad;
}

int oracle_batch_viterbi_f32(int n_lattices, const int64_t* state_off, const int64_t* arc_off, const int32_t* src,
                             const int32_t* dst, const float* w, int start, float* score, int32_t* path,
                             int32_t* path_len, int n_threads) {
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < n_lattices; ++b) {
    const int64_t a0 = arc_off[b], s0 = state_off[b];
    const int na = (int)(arc_off[b + 1] - a0), ns = (int)(state_off[b + 1] - s0);
    float* delta = (float*)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
    int32_t* bp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ns > 0 ? ns : 1));
    /* path slots: lattice b owns path[s0 .. s0+ns) (a path has fewer arcs than states) */
    score[b] = oracle_viterbi_f32(ns, na, src + a0, dst + a0, w + a0, start, delta, bp, path ? path + s0 : NULL,
                                  &path_len[b]);
    free(delta);
    free(bp);
  }
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
