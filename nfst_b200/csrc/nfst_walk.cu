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

// canonical id of the arc at CSR position i (= out_ptr[s] + k: the k-th arc of state s in label order).  CSR
// lattices: the identity.  Column-major lattices (sliced-column / tile-stream groups) store a state's arcs apart;
// nfst_packed_lattices_t.out_arc lists them.
__device__ __forceinline__ int arc_at(const nfst_packed_lattices_t& L, int i) { return L.out_arc ? __ldg(L.out_arc + i) : i; }

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
    for (int i = b0; i < b1; ++i) {
      const int b = arc_at(L, i);
      if (L.label_out[b] == l) return beta_real[L.dst_out[b]];
    }
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
  for (int i = a0; i < a1; ++i) m = fmaxf(m, logit(arc_at(L, i)));
  float sum = 0.0f;
  for (int i = a0; i < a1; ++i) sum += (m > kNegInf) ? expf(logit(arc_at(L, i)) - m) : 0.0f;
  const float lz = (m > kNegInf) ? m + logf(sum) : kNegInf;
  int pick = -1;
  if (given_sym) {  // evaluate_only: score the given symbol (samplers.py:258-259)
    const int want = given_sym[r];
    for (int i = a0; i < a1; ++i) {
      const int a = arc_at(L, i);
      if (L.label_out[a] == want) pick = a;
    }
  } else {  // inverse CDF over the arcs in label order
    const float target = uniform[r] * sum;
    float cum = 0.0f;
    for (int i = a0; i < a1; ++i) {
      const int a = arc_at(L, i);
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
    for (int i = a0; i < a1; ++i) {
      const int a = arc_at(L, i);
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

// ---- Sampler.stripping_pad (src/modules/samplers.py:162-180) ----
// Pass 1: which columns hold `pad` in EVERY row (the reference stops at the first such column).
__global__ void strip_pad_columns_kernel(const int64_t* __restrict__ seq, int n_rows, int T, int64_t pad_id,
                                         int32_t* __restrict__ col_not_pad) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int64_t* row = seq + static_cast<size_t>(warp) * T;
  for (int i = lane; i < T; i += 32)
    if (row[i] != pad_id) col_not_pad[i] = 1;  // benign race: every writer stores 1
}
// Pass 2: one warp per row left-compacts the symbols != 0 of columns 0..last (last = the first all-pad column, or
// T - 1), exactly as the reference's loop does: a symbol is written at the row's cursor, the cursor advances unless
// the symbol is 0 -- so a 0 is overwritten by whatever follows, and survives only behind the last non-zero symbol.
__global__ void strip_pad_compact_kernel(const int64_t* __restrict__ seq, int n_rows, int T, int64_t pad_id,
                                         const int32_t* __restrict__ col_not_pad, int64_t* __restrict__ out,
                                         int32_t* __restrict__ width_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  int last = T - 1;  // every lane scans the flags: T is a few hundred
  for (int i0 = 0; i0 < T; i0 += 32) {
    const int i = i0 + lane;
    const unsigned all_pad = __ballot_sync(0xffffffffu, i < T && col_not_pad[i] == 0);
    if (all_pad) {
      last = i0 + __ffs(all_pad) - 1;
      break;
    }
  }
  if (warp == 0 && lane == 0) *width_out = last + 1;
  const int64_t* row = seq + static_cast<size_t>(warp) * T;
  int64_t* dst = out + static_cast<size_t>(warp) * T;
  int cursor = 0;
  bool trailing_zero = false;
  for (int i0 = 0; i0 <= last; i0 += 32) {
    const int i = i0 + lane;
    const bool in = i <= last;
    const int64_t v = in ? row[i] : 0;
    const unsigned nz = __ballot_sync(0xffffffffu, in && v != 0);
    const unsigned zr = __ballot_sync(0xffffffffu, in && v == 0);
    if (in && v != 0) dst[cursor + __popc(nz & ((1u << lane) - 1u))] = v;
    cursor += __popc(nz);
    // a zero behind the last non-zero symbol seen so far stays at the cursor
    if (zr) trailing_zero = nz == 0u ? true : (31 - __clz(zr)) > (31 - __clz(nz));
    else if (nz) trailing_zero = false;
  }
  for (int i = cursor + lane; i <= last; i += 32) dst[i] = (i == cursor && trailing_zero) ? 0 : pad_id;
}

}  // namespace

extern "C" int nfst_strip_pad(const int64_t* sequences, int32_t n_rows, int32_t seq_len, int64_t pad_id, int64_t* out,
                              int32_t* col_flags, int32_t* width_out, void* cuda_stream) {
  if (!sequences || !out || !col_flags || !width_out) return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_strip_pad: null argument");
  if (n_rows <= 0 || seq_len <= 0) return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_strip_pad: empty input");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  cudaError_t e = cudaMemsetAsync(col_flags, 0, sizeof(int32_t) * seq_len, st);
  if (e != cudaSuccess) return nfst_fail_msg(NFST_ERR_CUDA, "nfst_strip_pad: %s", cudaGetErrorString(e));
  const int threads = 128, blocks = (n_rows * 32 + threads - 1) / threads;
  strip_pad_columns_kernel<<<blocks, threads, 0, st>>>(sequences, n_rows, seq_len, pad_id, col_flags);
  strip_pad_compact_kernel<<<blocks, threads, 0, st>>>(sequences, n_rows, seq_len, pad_id, col_flags, out, width_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return nfst_fail_msg(NFST_ERR_CUDA, "strip_pad kernels: %s", cudaGetErrorString(e));
  return 0;
}

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
