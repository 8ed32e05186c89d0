// nfst_walk.cu -- one step of the lattice-constrained sampling / scoring loop (sm_100a).
//
// The reference's proposal walks the lattice one symbol per time step (Sampler.stateful_sample,
// src/modules/samplers.py:243-297).  Per step and per row it
//   * gathers the dense row transition_k[row, state, :]            (scorers.py:586-589),
//   * looks the successor states' beta up, beta_logits = gather(beta, 1, transition) and adds the
//     network's prefix score                                        (scorers.py:590-592),
//   * zeroes the pad column (pad_masking, scorers.py:182-187), adds the vocabulary mask and the
//     emission mask of the current state -- 0 / -inf for boolean tables, the arc's log-weight for
//     weighted ones                                                 (scorers.py:340-357, :1037-1054),
//   * divides by the temperature, builds Categorical(logits) and samples or scores the next symbol,
//     log_prob and logsumexp ("zs")                                 (samplers.py:251-283),
//   * advances the state, transition_k[row, state, symbol]          (scorers.py:683-690)
// -- about a dozen eager kernels over [B*k, V] (and two gathers of V-wide rows from the k-times
// expanded [B*k, S, V] tables).  Here the step is ONE launch over the packed CSR arcs: a row only
// visits the out-arcs of its current state (1-3 for mark lattices), the k-fold table expansion
// (scorers.py:887-918) never happens (a row maps to its lattice by row / k).
//
// One thread per row; arcs of a state are visited in label order, so sampling by inverse CDF is
// deterministic for a given uniform number.  A state without arcs is the reference's absorbing sink
// (its only arc is the pad self-loop, which the edge rule drops): it emits pad with probability 1.
#include "nfst_b200.h"

#include <cuda_runtime.h>

int nfst_fail_msg(int code, const char* fmt, ...);  // nfst_kernels.cu

namespace {

constexpr float kNegInf = -__builtin_huge_valf();

__global__ void walk_step_kernel(const nfst_packed_lattices_t L, int n_rows, int rows_per_lattice,
                                 const int32_t* __restrict__ state, const int32_t* __restrict__ look_state,
                                 const float* __restrict__ prefix,
                                 const float* __restrict__ base_mask, const float* __restrict__ beta_real,
                                 const float* __restrict__ arc_static, float inv_temperature, int pad_id,
                                 const int32_t* __restrict__ given_sym, const float* __restrict__ uniform,
                                 int32_t* __restrict__ sym_out, float* __restrict__ logp_out,
                                 int32_t* __restrict__ next_state, float* __restrict__ logz_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int V = L.vocab;
  const int s = state[r];
  const int a0 = L.out_ptr[s], a1 = L.out_ptr[s + 1];
  const float* pre = prefix + static_cast<size_t>(r) * V;
  const float* bm = base_mask ? base_mask + static_cast<size_t>(r) * V : nullptr;
  // masked logit of arc a (label l, destination d), the reference's order of operations:
  //   ((beta[d] + prefix[l]) * (l != pad)) + base_mask[l] + emission_mask, then / temperature
  // beta look-ahead.  Aligned (look_state == NULL): beta of the arc's own destination.  Reference-faithful
  // (quirk Q9, scorers.py:584 vs :679): the reference gathers beta through the transition row of the state
  // BEFORE the previous symbol was consumed -- the label's successor from THAT state, or dense state 0 (the
  // start state) where that state has no such arc (transition = 0).
  const int ls = look_state ? look_state[r] : -1;
  const int b0 = look_state ? L.out_ptr[ls] : 0, b1 = look_state ? L.out_ptr[ls + 1] : 0;
  const float beta_none = look_state ? beta_real[L.start_state[r / rows_per_lattice]] : 0.0f;
  auto look = [&](int a, int l) -> float {
    if (!look_state) return beta_real[L.dst_out[a]];
    for (int b = b0; b < b1; ++b)
      if (L.label_out[b] == l) return beta_real[L.dst_out[b]];
    return beta_none;
  };
  auto logit = [&](int a) -> float {
    const int l = L.label_out[a];
    float v = (l == pad_id) ? 0.0f : look(a, l) + pre[l];
    if (bm) v += bm[l];
    if (arc_static) v += arc_static[a];
    return v * inv_temperature;
  };
  if (a1 == a0) {  // absorbing sink: only pad -> itself
    const float v = (bm ? bm[pad_id] : 0.0f) * inv_temperature;
    sym_out[r] = pad_id;
    next_state[r] = s;
    logp_out[r] = (given_sym && given_sym[r] != pad_id) ? kNegInf : 0.0f;
    if (logz_out) logz_out[r] = v;
    return;
  }
  float m = kNegInf;
  for (int a = a0; a < a1; ++a) m = fmaxf(m, logit(a));
  float sum = 0.0f;
  for (int a = a0; a < a1; ++a) sum += (m > kNegInf) ? expf(logit(a) - m) : 0.0f;
  const float lz = (m > kNegInf) ? m + logf(sum) : kNegInf;
  int pick = -1;
  if (given_sym) {  // evaluate_only: score the given symbol (samplers.py:258-259)
    const int want = given_sym[r];
    for (int a = a0; a < a1; ++a)
      if (L.label_out[a] == want) pick = a;
  } else {  // inverse CDF over the arcs in label order
    const float target = uniform[r] * sum;
    float cum = 0.0f;
    for (int a = a0; a < a1; ++a) {
      const float e = expf(logit(a) - m);
      cum += e;
      if (e > 0.0f) pick = a;  // the last arc with mass catches round-off at the top
      if (cum > target && e > 0.0f) break;
    }
  }
  if (pick < 0) {  // the given symbol is not allowed here: probability 0, the state does not move
    sym_out[r] = given_sym ? given_sym[r] : pad_id;
    next_state[r] = s;
    logp_out[r] = kNegInf;
  } else {
    sym_out[r] = L.label_out[pick];
    next_state[r] = L.dst_out[pick];
    logp_out[r] = logit(pick) - lz;
  }
  if (logz_out) logz_out[r] = lz;
}

}  // namespace

extern "C" int nfst_walk_step_f32(const nfst_packed_lattices_t* lat, int32_t n_rows, int32_t rows_per_lattice,
                                  const int32_t* state, const int32_t* look_state, const float* prefix,
                                  const float* base_mask,
                                  const float* beta_real, const float* arc_static, float temperature, int32_t pad_id,
                                  const int32_t* given_sym, const float* uniform, int32_t* sym_out, float* logp_out,
                                  int32_t* next_state, float* logz_out, void* cuda_stream) {
  if (!lat || !state || !prefix || !beta_real || !sym_out || !logp_out || !next_state)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_walk_step_f32: null argument");
  if ((given_sym == nullptr) == (uniform == nullptr))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_walk_step_f32: exactly one of given_sym (score) and uniform (sample) is required");
  if (!(temperature > 0.0f)) return nfst_fail_msg(NFST_ERR_BAD_ARG, "temperature must be positive");
  if (pad_id < 0 || pad_id >= lat->vocab) return nfst_fail_msg(NFST_ERR_BAD_ARG, "pad_id=%d outside the vocabulary", pad_id);
  if (n_rows < 0 || rows_per_lattice <= 0) return nfst_fail_msg(NFST_ERR_BAD_ARG, "bad row counts");
  if (n_rows == 0) return 0;
  const int threads = 128;
  walk_step_kernel<<<(n_rows + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      *lat, n_rows, rows_per_lattice, state, look_state, prefix, base_mask, beta_real, arc_static, 1.0f / temperature,
      pad_id,
      given_sym, uniform, sym_out, logp_out, next_state, logz_out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nfst_fail_msg(NFST_ERR_CUDA, "walk_step_kernel: %s", cudaGetErrorString(e));
  return 0;
}

// =====================================================================================
// Whole sampling loop for an ARC-FACTORED proposal (second half).
//
// When the proposal's score of a path is the sum of its arc scores (the reference's WFSTScorer,
// scorers.py:1663-1687), the look-ahead beta is exact and the loop of Sampler.stateful_sample
// (samplers.py:243-297) needs no network between steps: from state s the next arc is drawn with
// probability exp(w_a + beta[dst_a] - beta[s]), and the walk ends at a state without arcs.  The
// paths are then EXACT samples of the posterior over paths, log q(z) = score(z) - logZ, so every
// importance weight log p~(z) - log q(z) of Estimators.iwae (estimatros.py:10-44) equals logZ: the
// k-sample estimate has zero variance.  One thread per row (lattice, sample) walks the whole path in one
// launch; inverse CDF over the arcs of a state in label order, one uniform number per step.
// =====================================================================================
namespace {

template <typename ST>
__global__ void sample_paths_kernel(const nfst_packed_lattices_t L, int n_rows, int rows_per_lattice, int max_len,
                                    const float* __restrict__ arc_scores, const float* __restrict__ theta,
                                    const ST* __restrict__ beta, const float* __restrict__ uniform, int pad_id,
                                    int32_t* __restrict__ labels, int32_t* __restrict__ arcs,
                                    int32_t* __restrict__ length, float* __restrict__ log_q) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int s = L.start_state[r / rows_per_lattice];
  double lq = 0.0;
  int t = 0;
  for (; t < max_len; ++t) {
    const int a0 = L.out_ptr[s], a1 = L.out_ptr[s + 1];
    if (a1 == a0) break;  // the sink (its pad loop is not an arc)
    const double bs = static_cast<double>(beta[s]);
    const float u = uniform[static_cast<size_t>(r) * max_len + t];
    auto logc = [&](int a) -> double {  // log P(arc | state)
      float w = arc_scores ? arc_scores[a] : 0.0f;
      if (theta) w += theta[L.label_out[a]];
      return static_cast<double>(w) + static_cast<double>(beta[L.dst_out[a]]) - bs;
    };
    int pick = -1;
    double cum = 0.0, lp = 0.0;
    for (int a = a0; a < a1; ++a) {
      const double lc = logc(a);
      const double c = exp(lc);
      if (c > 0.0) {  // the last arc with mass catches round-off at the top of the CDF
        pick = a;
        lp = lc;
      }
      cum += c;
      if (cum > static_cast<double>(u) && c > 0.0) break;
    }
    if (pick < 0) break;  // every arc has probability 0 (scores -inf): the walk cannot continue
    lq += lp;
    labels[static_cast<size_t>(r) * max_len + t] = L.label_out[pick];
    if (arcs) arcs[static_cast<size_t>(r) * max_len + t] = pick;
    s = L.dst_out[pick];
  }
  length[r] = t;
  log_q[r] = static_cast<float>(lq);
  for (int i = t; i < max_len; ++i) {
    labels[static_cast<size_t>(r) * max_len + i] = pad_id;
    if (arcs) arcs[static_cast<size_t>(r) * max_len + i] = -1;
  }
}

}  // namespace

extern "C" int nfst_sample_paths_f32(const nfst_packed_lattices_t* lat, int32_t n_rows, int32_t rows_per_lattice,
                                     int32_t max_len, const nfst_scores_t* scores, const void* beta, int beta_f64,
                                     const float* uniform, int32_t pad_id, int32_t* labels, int32_t* arcs,
                                     int32_t* length, float* log_q, void* cuda_stream) {
  if (!lat || !scores || !beta || !uniform || !labels || !length || !log_q)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_sample_paths_f32: null argument");
  if (!scores->arc_scores && !scores->theta) return nfst_fail_msg(NFST_ERR_BAD_ARG, "arc_scores or theta is required");
  if (n_rows < 0 || rows_per_lattice <= 0 || max_len <= 0) return nfst_fail_msg(NFST_ERR_BAD_ARG, "bad row counts");
  if (n_rows == 0) return 0;
  const int threads = 128;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (beta_f64)
    sample_paths_kernel<double><<<(n_rows + threads - 1) / threads, threads, 0, st>>>(
        *lat, n_rows, rows_per_lattice, max_len, scores->arc_scores, scores->theta, static_cast<const double*>(beta),
        uniform, pad_id, labels, arcs, length, log_q);
  else
    sample_paths_kernel<float><<<(n_rows + threads - 1) / threads, threads, 0, st>>>(
        *lat, n_rows, rows_per_lattice, max_len, scores->arc_scores, scores->theta, static_cast<const float*>(beta),
        uniform, pad_id, labels, arcs, length, log_q);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return nfst_fail_msg(NFST_ERR_CUDA, "sample_paths_kernel: %s", cudaGetErrorString(e));
  return 0;
}
