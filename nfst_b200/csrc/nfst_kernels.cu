// nfst_kernels.cu -- sm_100a kernels and C ABI of the lattice dynamic-programming library.
//
// Hot path of steventan0110/nFST restated for B200: log-semiring forward/backward and
// tropical Viterbi over batched, topologically levelled, CSR-packed lattices.  The
// recurrence is the reference's FSAGRUScorer.compute_beta_per_sample
// (src/modules/scorers.py:692-751) / compute_beta_parallel (:753-856) with Wh = 0, in
// log space:  beta[c] = logsumexp_{c -j-> n} (theta_j + beta[n]).
//
// Execution model (v1): one thread block per lattice, level-synchronous.  Inside a level
// 2^g lanes cooperate on one state (g chosen per lattice at pack time from the mean
// degree): the lanes stride over the state's contiguous CSR segment with coalesced
// loads, keep a lane-local online (max, sum) pair, and combine with xor-shuffles.  The
// per-state DP vectors live in shared memory when the lattice fits, otherwise in global
// memory (L1/L2-resident; the same block wrote them, so block-scope barriers order
// them).  The fused backward emits beta, arc posteriors (scaled by the incoming
// gradient), the per-label gradient and the Viterbi delta/backpointer in one pass over
// the outgoing arcs.
#include "nfst_b200.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define NFST_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return fail(NFST_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

constexpr float kNegInf = -__builtin_huge_valf();

// ---- online logsumexp pair (m, s): value = m + log(s) --------------------------------
__device__ __forceinline__ void lse_add(float& m, float& s, float v) {
  if (v > m) {
    s = s * __expf(m - v) + 1.0f;  // m == -inf: 0 * 0 + 1
    m = v;
  } else if (v > kNegInf) {
    s += __expf(v - m);
  }
}
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  if (m2 > m) {
    s = s * __expf(m - m2) + s2;
    m = m2;
  } else if (m2 > kNegInf) {
    s += s2 * __expf(m2 - m);
  }
}
__device__ __forceinline__ float lse_value(float m, float s) { return (m == kNegInf) ? kNegInf : m + logf(s); }

// ---- shared memory carve-up (identical on host and device) ---------------------------
struct SmemPlan {
  int lvl, theta, dtheta, st0, st1, words;
};
__host__ __device__ inline SmemPlan smem_plan(int level_cap, int state_cap, int vocab, int n_state_arrays,
                                              bool with_theta, bool with_dtheta) {
  SmemPlan p;
  int w = 0;
  p.lvl = w;
  w += level_cap > 0 ? level_cap + 1 : 0;
  p.theta = w;
  w += with_theta ? vocab : 0;
  p.dtheta = w;
  w += with_dtheta ? vocab : 0;
  p.st0 = w;
  w += n_state_arrays >= 1 ? state_cap : 0;
  p.st1 = w;
  w += n_state_arrays >= 2 ? state_cap : 0;
  p.words = w;
  return p;
}

// block-wide combine of (m, s) pairs; result valid in thread 0
__device__ float block_lse(float m, float s) {
  __shared__ float red_m[32], red_s[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    red_m[wid] = m;
    red_s[wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
    m = lane < nw ? red_m[lane] : kNegInf;
    s = lane < nw ? red_s[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      float s2 = __shfl_xor_sync(0xffffffffu, s, o);
      lse_merge(m, s, m2, s2);
    }
  }
  return lse_value(m, s);
}

// =====================================================================================
// forward: alpha
// =====================================================================================
template <bool SMEM_STATE>
__global__ void __launch_bounds__(1024) nfst_fwd_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids,
                                                        int level_cap, int state_cap,
                                                        const float* __restrict__ arc_scores,
                                                        const float* __restrict__ theta, int theta_smem,
                                                        float* alpha, float* __restrict__ logz) {
  extern __shared__ __align__(16) float smem[];
  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int s0 = L.state_off[b];
  const int lvl0 = L.level_off[b];
  const int nlev = L.level_off[b + 1] - lvl0 - 1;
  const SmemPlan plan = smem_plan(level_cap, state_cap, L.vocab, 1, theta_smem != 0, false);

  const int32_t* lp = L.level_ptr + lvl0;
  if (level_cap > 0 && nlev <= level_cap) {
    int32_t* slp = reinterpret_cast<int32_t*>(smem + plan.lvl);
    for (int i = threadIdx.x; i <= nlev; i += blockDim.x) slp[i] = lp[i];
    lp = slp;
  }
  const float* th = theta;
  if (theta && theta_smem) {
    float* sth = smem + plan.theta;
    for (int i = threadIdx.x; i < L.vocab; i += blockDim.x) sth[i] = theta[i];
    th = sth;
  }
  float* sA = smem + plan.st0;
  __syncthreads();

  const int lg0 = L.lanes_in_log2[b];
  const int start = L.start_state[b];

  // level 0 holds the start state only (unreachable states are trimmed at pack time)
  for (int s = lp[0] + threadIdx.x; s < lp[1]; s += blockDim.x) {
    const float v = (s == start) ? 0.0f : kNegInf;
    if (SMEM_STATE) sA[s - s0] = v;
    alpha[s] = v;
  }
  __syncthreads();

  for (int l = 1; l < nlev; ++l) {
    const int sb = lp[l], se = lp[l + 1];
    // lanes per state: the lattice's default, widened while the level leaves lanes idle
    int lg = lg0;
    while (lg < 5 && ((se - sb) << (lg + 1)) <= static_cast<int>(blockDim.x)) ++lg;
    const int G = 1 << lg;
    const int lane_g = threadIdx.x & (G - 1);
    const int grp = threadIdx.x >> lg;
    const int ngrp = blockDim.x >> lg;
    for (int base = sb; base < se; base += ngrp) {
      const int s = base + grp;
      const bool valid = s < se;
      int a0 = 0, a1 = 0;
      if (valid) {
        a0 = L.in_ptr[s];
        a1 = L.in_ptr[s + 1];
      }
      float m = kNegInf, sum = 0.0f;
      for (int a = a0 + lane_g; a < a1; a += G) {
        const int src = L.src_in[a];
        float w = 0.0f;
        if (arc_scores) w = arc_scores[L.in2out[a]];
        if (th) w += th[L.label_in[a]];
        const float av = SMEM_STATE ? sA[src - s0] : alpha[src];
        lse_add(m, sum, w + av);
      }
      for (int o = G >> 1; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
        lse_merge(m, sum, m2, s2);
      }
      if (valid && lane_g == 0) {
        const float v = lse_value(m, sum);
        if (SMEM_STATE) sA[s - s0] = v;
        alpha[s] = v;
      }
    }
    __syncthreads();
  }

  // logZ = logsumexp over the sinks of alpha  (every zero-out-degree state has beta = 1,
  // scorers.py:795-805)
  float m = kNegInf, sum = 0.0f;
  for (int i = L.sink_off[b] + threadIdx.x; i < L.sink_off[b + 1]; i += blockDim.x) {
    const int s = L.sinks[i];
    lse_add(m, sum, SMEM_STATE ? sA[s - s0] : alpha[s]);
  }
  const float z = block_lse(m, sum);
  if (threadIdx.x == 0) logz[b] = z;
}

// =====================================================================================
// fused backward: beta (+ posteriors, dtheta) and/or Viterbi delta + backpointer
// =====================================================================================
template <bool SMEM_STATE, bool LOGS, bool TROP>
__global__ void __launch_bounds__(1024)
    nfst_bwd_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int level_cap, int state_cap,
                    const float* __restrict__ arc_scores, const float* __restrict__ theta, int theta_smem,
                    int dtheta_smem, const float* __restrict__ alpha, const float* __restrict__ logz,
                    const float* __restrict__ grad_logz, float* beta, float* __restrict__ logz_bwd,
                    float* __restrict__ post, float* __restrict__ dtheta, float* delta, int32_t* __restrict__ backptr,
                    float* __restrict__ vit_score) {
  extern __shared__ __align__(16) float smem[];
  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int s0 = L.state_off[b];
  const int lvl0 = L.level_off[b];
  const int nlev = L.level_off[b + 1] - lvl0 - 1;
  constexpr int kArrays = (LOGS ? 1 : 0) + (TROP ? 1 : 0);
  const SmemPlan plan = smem_plan(level_cap, state_cap, L.vocab, kArrays, theta_smem != 0, dtheta_smem != 0);

  const int32_t* lp = L.level_ptr + lvl0;
  if (level_cap > 0 && nlev <= level_cap) {
    int32_t* slp = reinterpret_cast<int32_t*>(smem + plan.lvl);
    for (int i = threadIdx.x; i <= nlev; i += blockDim.x) slp[i] = lp[i];
    lp = slp;
  }
  const float* th = theta;
  if (theta && theta_smem) {
    float* sth = smem + plan.theta;
    for (int i = threadIdx.x; i < L.vocab; i += blockDim.x) sth[i] = theta[i];
    th = sth;
  }
  float* hist = nullptr;
  if (LOGS && dtheta) {
    if (dtheta_smem) {
      hist = smem + plan.dtheta;
      for (int i = threadIdx.x; i < L.vocab; i += blockDim.x) hist[i] = 0.0f;
    } else {
      hist = dtheta;
    }
  }
  float* sB = smem + plan.st0;                   // beta (LOGS) or delta (!LOGS)
  float* sD = LOGS ? smem + plan.st1 : sB;       // delta
  __syncthreads();

  const int lg0 = L.lanes_out_log2[b];
  const bool want_post = LOGS && (post != nullptr || hist != nullptr);
  const bool need_label = (th != nullptr) || (hist != nullptr);
  float lz = 0.0f, gscale = 1.0f;
  if (want_post) {
    lz = logz[b];
    if (grad_logz) gscale = grad_logz[b];
  }

  for (int l = nlev - 1; l >= 0; --l) {
    const int sb = lp[l], se = lp[l + 1];
    int lg = lg0;
    while (lg < 5 && ((se - sb) << (lg + 1)) <= static_cast<int>(blockDim.x)) ++lg;
    const int G = 1 << lg;
    const int lane_g = threadIdx.x & (G - 1);
    const int grp = threadIdx.x >> lg;
    const int ngrp = blockDim.x >> lg;
    for (int base = sb; base < se; base += ngrp) {
      const int s = base + grp;
      const bool valid = s < se;
      int a0 = 0, a1 = 0;
      float am = 0.0f;
      if (valid) {
        a0 = L.out_ptr[s];
        a1 = L.out_ptr[s + 1];
        if (want_post) am = alpha[s] - lz;
      }
      float m = kNegInf, sum = 0.0f;
      float bt = kNegInf;
      int ba = 0x7fffffff;
      for (int a = a0 + lane_g; a < a1; a += G) {
        const int d = L.dst_out[a];
        int lab = 0;
        if (need_label) lab = L.label_out[a];
        float w = 0.0f;
        if (arc_scores) w = arc_scores[a];
        if (th) w += th[lab];
        if (LOGS) {
          const float u = w + (SMEM_STATE ? sB[d - s0] : beta[d]);
          lse_add(m, sum, u);
          if (want_post) {
            const float p = __expf(am + u) * gscale;
            if (post) post[a] = p;
            if (hist) atomicAdd(&hist[lab], p);
          }
        }
        if (TROP) {
          const float t = __fadd_rn(w, SMEM_STATE ? sD[d - s0] : delta[d]);
          if (t > bt || (t == bt && a < ba)) {
            bt = t;
            ba = a;
          }
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) {
        if (LOGS) {
          const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
          const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
          lse_merge(m, sum, m2, s2);
        }
        if (TROP) {
          const float t2 = __shfl_xor_sync(0xffffffffu, bt, o);
          const int a2 = __shfl_xor_sync(0xffffffffu, ba, o);
          if (t2 > bt || (t2 == bt && a2 < ba)) {
            bt = t2;
            ba = a2;
          }
        }
      }
      if (valid && lane_g == 0) {
        const bool sink = (a0 == a1);
        if (LOGS) {
          const float v = sink ? 0.0f : lse_value(m, sum);
          if (SMEM_STATE) sB[s - s0] = v;
          if (!SMEM_STATE || beta) beta[s] = v;
        }
        if (TROP) {
          const float v = sink ? 0.0f : bt;
          if (SMEM_STATE) sD[s - s0] = v;
          if (!SMEM_STATE || delta) delta[s] = v;
          backptr[s] = sink ? -1 : ba;
        }
      }
    }
    __syncthreads();
  }

  if (threadIdx.x == 0) {
    const int start = L.start_state[b];
    if (LOGS && logz_bwd) logz_bwd[b] = SMEM_STATE ? sB[start - s0] : beta[start];
    if (TROP && vit_score) vit_score[b] = SMEM_STATE ? sD[start - s0] : delta[start];
  }
  if (LOGS && dtheta && dtheta_smem) {
    for (int i = threadIdx.x; i < L.vocab; i += blockDim.x) {
      const float v = hist[i];
      if (v != 0.0f) atomicAdd(&dtheta[i], v);
    }
  }
}

// =====================================================================================
// small kernels
// =====================================================================================
__global__ void nfst_backtrace_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ backptr,
                                      const int32_t* __restrict__ path_off, int32_t* __restrict__ path_arcs,
                                      int32_t* __restrict__ path_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= L.n_lattices) return;
  const int off = path_off[b];
  const int cap = path_off[b + 1] - off;
  int s = L.start_state[b];
  int k = 0;
  while (k < cap) {
    const int a = backptr[s];
    if (a < 0) break;
    path_arcs[off + k] = a;
    ++k;
    s = L.dst_out[a];
  }
  path_len[b] = k;
}

__global__ void nfst_beta_to_dense_kernel(const nfst_packed_lattices_t L, const float* __restrict__ beta,
                                          const int32_t* __restrict__ orig_state, int k, int dense_states,
                                          float* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L.n_states) return;
  // lattice of state s: last b with state_off[b] <= s
  int lo = 0, hi = L.n_lattices;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (L.state_off[mid] <= s) lo = mid; else hi = mid;
  }
  const float v = expf(beta[s]);
  const size_t row0 = static_cast<size_t>(lo) * k;
  const int col = orig_state[s];
  for (int j = 0; j < k; ++j) out[(row0 + j) * dense_states + col] = v;
}

// dense-table edge rule (scorers.py:704-716): one warp per table row
__global__ void nfst_dense_count_kernel(const int64_t* __restrict__ tr, int64_t n_rows, int S, int V,
                                        int32_t* __restrict__ counts) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t self = row % S;
  const int64_t* p = tr + row * V;
  int c = 0;
  for (int j = lane; j < V; j += 32) {
    const int64_t t = p[j];
    c += (t != 0 && t != self) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[row] = c;
}

__global__ void nfst_dense_extract_kernel(const int64_t* __restrict__ tr, int64_t n_rows, int S, int V,
                                          const int64_t* __restrict__ row_start, int32_t* __restrict__ arc_row,
                                          int32_t* __restrict__ arc_label, int32_t* __restrict__ arc_dst) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t self = row % S;
  const int64_t* p = tr + row * V;
  int64_t pos = row_start[row];
  for (int j0 = 0; j0 < V; j0 += 32) {
    const int j = j0 + lane;
    int64_t t = 0;
    if (j < V) t = p[j];
    const bool keep = (j < V) && t != 0 && t != self;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int64_t at = pos + __popc(mask & ((1u << lane) - 1u));
      arc_row[at] = static_cast<int32_t>(row);
      arc_label[at] = j;
      arc_dst[at] = static_cast<int32_t>(t);
    }
    pos += __popc(mask);
  }
}

// ---- host helpers ------------------------------------------------------------------
int check_launch(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch) {
  if (!lat || !launch) return fail(NFST_ERR_BAD_ARG, "null lattice or launch descriptor");
  if (launch->n_ids < 0 || launch->n_ids > lat->n_lattices)
    return fail(NFST_ERR_BAD_ARG, "n_ids=%d out of range (B=%d)", launch->n_ids, lat->n_lattices);
  if (launch->block_threads < 32 || launch->block_threads > 1024 || (launch->block_threads & 31))
    return fail(NFST_ERR_BAD_ARG, "block_threads=%d must be a multiple of 32 in [32,1024]", launch->block_threads);
  if (launch->state_smem_cap < 0 || launch->level_smem_cap < 0)
    return fail(NFST_ERR_BAD_ARG, "negative shared-memory capacity");
  return NFST_OK;
}

template <typename K>
int prepare_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    int dev = 0;
    NFST_CUDA_OK(cudaGetDevice(&dev));
    int optin = 0;
    NFST_CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (bytes > static_cast<size_t>(optin))
      return fail(NFST_ERR_TOO_LARGE, "launch needs %zu B of shared memory, device allows %d", bytes, optin);
    NFST_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  }
  return NFST_OK;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int nfst_abi_version(void) { return NFST_ABI_VERSION; }

const char* nfst_last_error_string(void) { return g_last_error.c_str(); }

int nfst_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* max_smem_optin) {
  cudaDeviceProp prop;
  NFST_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (max_smem_optin) *max_smem_optin = prop.sharedMemPerBlockOptin;
  if (prop.major != 10)
    return fail(NFST_ERR_UNSUPPORTED_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  return NFST_OK;
}

size_t nfst_launch_smem_bytes(const nfst_launch_t* launch, int32_t vocab, int n_state_arrays, int with_theta,
                              int with_dtheta) {
  if (!launch) return 0;
  const bool small_v = vocab <= NFST_THETA_SMEM_MAX;
  const SmemPlan p = smem_plan(launch->level_smem_cap, launch->state_smem_cap, vocab, n_state_arrays,
                               with_theta && small_v, with_dtheta && small_v);
  return static_cast<size_t>(p.words) * sizeof(float);
}

int nfst_fwd_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                 float* alpha, float* logz, void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!scores || !alpha || !logz) return fail(NFST_ERR_BAD_ARG, "nfst_fwd_f32: null scores/alpha/logz");
  if (launch->n_ids == 0) return NFST_OK;
  const int theta_smem = scores->theta && lat->vocab <= NFST_THETA_SMEM_MAX;
  const size_t bytes = nfst_launch_smem_bytes(launch, lat->vocab, 1, scores->theta != nullptr, 0);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (launch->state_smem_cap > 0) {
    if (int rc = prepare_smem(nfst_fwd_kernel<true>, bytes)) return rc;
    nfst_fwd_kernel<true><<<launch->n_ids, launch->block_threads, bytes, st>>>(
        *lat, launch->lattice_ids, launch->level_smem_cap, launch->state_smem_cap, scores->arc_scores, scores->theta,
        theta_smem, alpha, logz);
  } else {
    if (int rc = prepare_smem(nfst_fwd_kernel<false>, bytes)) return rc;
    nfst_fwd_kernel<false><<<launch->n_ids, launch->block_threads, bytes, st>>>(
        *lat, launch->lattice_ids, launch->level_smem_cap, 0, scores->arc_scores, scores->theta, theta_smem, alpha,
        logz);
  }
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_bwd_fused_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                       const float* alpha, const float* logz, const float* grad_logz, float* beta, float* logz_bwd,
                       float* post, float* dtheta, float* delta, int32_t* backptr, float* vit_score,
                       void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!scores) return fail(NFST_ERR_BAD_ARG, "nfst_bwd_fused_f32: null scores");
  const bool smem_state = launch->state_smem_cap > 0;
  const bool want_post = post || dtheta;
  // the log pass runs when any of its outputs is requested; in global-state mode it
  // needs the beta buffer as its working vector
  const bool logs = beta || logz_bwd || want_post;
  const bool trop = delta || backptr || vit_score;
  if (!logs && !trop) return fail(NFST_ERR_BAD_ARG, "nfst_bwd_fused_f32: no output requested");
  if (logs && !smem_state && !beta)
    return fail(NFST_ERR_BAD_ARG, "log pass with state vectors in global memory needs a beta[S] buffer");
  if (trop && !backptr) return fail(NFST_ERR_BAD_ARG, "tropical pass needs backptr[S]");
  if (trop && !smem_state && !delta)
    return fail(NFST_ERR_BAD_ARG, "tropical pass with state vectors in global memory needs a delta[S] buffer");
  if (want_post && (!alpha || !logz))
    return fail(NFST_ERR_BAD_ARG, "posteriors / dtheta need alpha[S] and logz[B] from nfst_fwd_f32");
  if (launch->n_ids == 0) return NFST_OK;

  const bool small_v = lat->vocab <= NFST_THETA_SMEM_MAX;
  const int theta_smem = scores->theta && small_v;
  const int dtheta_smem = dtheta && small_v;
  const int n_arrays = (logs ? 1 : 0) + (trop ? 1 : 0);
  const size_t bytes =
      nfst_launch_smem_bytes(launch, lat->vocab, n_arrays, scores->theta != nullptr, logs && dtheta != nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int scap = smem_state ? launch->state_smem_cap : 0;

#define NFST_LAUNCH_BWD(SM, LG, TR)                                                                              \
  do {                                                                                                           \
    if (int rc = prepare_smem(nfst_bwd_kernel<SM, LG, TR>, bytes)) return rc;                                    \
    nfst_bwd_kernel<SM, LG, TR><<<launch->n_ids, launch->block_threads, bytes, st>>>(                            \
        *lat, launch->lattice_ids, launch->level_smem_cap, scap, scores->arc_scores, scores->theta, theta_smem,  \
        dtheta_smem, alpha, logz, grad_logz, beta, logz_bwd, post, dtheta, delta, backptr, vit_score);           \
  } while (0)

  if (smem_state) {
    if (logs && trop) NFST_LAUNCH_BWD(true, true, true);
    else if (logs) NFST_LAUNCH_BWD(true, true, false);
    else NFST_LAUNCH_BWD(true, false, true);
  } else {
    if (logs && trop) NFST_LAUNCH_BWD(false, true, true);
    else if (logs) NFST_LAUNCH_BWD(false, true, false);
    else NFST_LAUNCH_BWD(false, false, true);
  }
#undef NFST_LAUNCH_BWD
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_viterbi_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                     float* delta, int32_t* backptr, float* vit_score, void* cuda_stream) {
  return nfst_bwd_fused_f32(lat, launch, scores, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, delta,
                            backptr, vit_score, cuda_stream);
}

int nfst_backtrace(const nfst_packed_lattices_t* lat, const int32_t* backptr, const int32_t* path_off,
                   int32_t* path_arcs, int32_t* path_len, void* cuda_stream) {
  if (!lat || !backptr || !path_off || !path_arcs || !path_len)
    return fail(NFST_ERR_BAD_ARG, "nfst_backtrace: null argument");
  if (lat->n_lattices == 0) return NFST_OK;
  const int threads = 128;
  const int blocks = (lat->n_lattices + threads - 1) / threads;
  nfst_backtrace_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(*lat, backptr, path_off,
                                                                                        path_arcs, path_len);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_beta_to_dense_f32(const nfst_packed_lattices_t* lat, const float* beta, const int32_t* orig_state,
                           int32_t k, int32_t dense_states, float* out, void* cuda_stream) {
  if (!lat || !beta || !orig_state || !out || k < 1 || dense_states < 1)
    return fail(NFST_ERR_BAD_ARG, "nfst_beta_to_dense_f32: bad argument");
  if (lat->n_states == 0) return NFST_OK;
  const int threads = 256;
  const int blocks = (lat->n_states + threads - 1) / threads;
  nfst_beta_to_dense_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(*lat, beta, orig_state, k,
                                                                                            dense_states, out);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_dense_count_arcs(const int64_t* transition, int64_t n_rows, int32_t states_per_lattice, int32_t vocab,
                          int32_t* row_counts, void* cuda_stream) {
  if (!transition || !row_counts || n_rows < 0 || states_per_lattice < 1 || vocab < 1)
    return fail(NFST_ERR_BAD_ARG, "nfst_dense_count_arcs: bad argument");
  if (n_rows == 0) return NFST_OK;
  const int threads = 256;  // 8 rows per block
  const int64_t blocks = (n_rows * 32 + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return fail(NFST_ERR_TOO_LARGE, "too many table rows");
  nfst_dense_count_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      transition, n_rows, states_per_lattice, vocab, row_counts);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_dense_extract_arcs(const int64_t* transition, int64_t n_rows, int32_t states_per_lattice, int32_t vocab,
                            const int64_t* row_start, int32_t* arc_row, int32_t* arc_label, int32_t* arc_dst,
                            void* cuda_stream) {
  if (!transition || !row_start || !arc_row || !arc_label || !arc_dst || n_rows < 0 || states_per_lattice < 1 ||
      vocab < 1)
    return fail(NFST_ERR_BAD_ARG, "nfst_dense_extract_arcs: bad argument");
  if (n_rows == 0) return NFST_OK;
  const int threads = 256;
  const int64_t blocks = (n_rows * 32 + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return fail(NFST_ERR_TOO_LARGE, "too many table rows");
  nfst_dense_extract_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      transition, n_rows, states_per_lattice, vocab, row_start, arc_row, arc_label, arc_dst);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

}  // extern "C"
