// nfst_kernels.cu -- sm_100a kernels and C ABI of the lattice dynamic-programming library.
//
// Hot path of steventan0110/nFST restated for B200: log-semiring forward/backward and
// tropical Viterbi over batched, topologically levelled, CSR-packed lattices.  The
// recurrence is the reference's FSAGRUScorer.compute_beta_per_sample
// (src/modules/scorers.py:692-751) / compute_beta_parallel (:753-856) with Wh = 0, in
// log space:  beta[c] = logsumexp_{c -j-> n} (theta_j + beta[n]).
//
// Execution model (v2).  One thread block walks one lattice chunk by chunk (a chunk = a
// run of consecutive states of one level + their contiguous CSR arc range, built at pack
// time; at most 8*blockDim arcs are staged at once):
//   phase 1 (arc-parallel)   every thread owns 4 consecutive arcs: 128-bit coalesced loads
//            of the arc arrays -- issued one to two chunks AHEAD into registers, so HBM
//            streaming is decoupled from the per-level barriers -- then gathers of the
//            neighbour state's DP value from a shared-memory window of the most recent
//            `window_states` values (global memory beyond it), and a store of
//            w + value into a shared-memory tile;
//   phase 2 (state-parallel) 2^g lanes per state reduce the state's segment of the tile
//            (max pass, sum-of-exp pass, xor-shuffle combine) and write the new DP value
//            to the window and to global memory; the fused backward also emits the arc
//            posteriors (scaled by the incoming gradient), the per-label gradient and the
//            Viterbi delta / backpointer from the same tile.
// State vectors are float32 or float64 (template ST); the tropical pass is always fp32
// and uses only one add and exact comparisons, so it is bit-reproducible.
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

#define NFST_CUDA_OK(expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return fail(NFST_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

}  // namespace

// error reporting for the library's other translation units (nfst_sell.cu)
int nfst_fail_msg(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

namespace {

constexpr float kNegInf = -__builtin_huge_valf();
#ifndef NFST_MIN_BLOCKS
#define NFST_MIN_BLOCKS 3  // resident blocks per SM the register allocation aims for
#endif

__device__ __forceinline__ int4 chunk_at(const nfst_chunk_t* c, int i, int end) {
  return i < end ? __ldg(reinterpret_cast<const int4*>(c + i)) : make_int4(0, 0, 0, 0);
}

// ---- online logsumexp (m, s): value = m + log(s); m in the state type, s in fp32 --------
template <typename ST>
__device__ __forceinline__ void lse_add(ST& m, float& s, ST v) {
  if (v > m) {
    s = s * __expf(static_cast<float>(m - v)) + 1.0f;  // m == -inf: 0 * 0 + 1
    m = v;
  } else if (v > static_cast<ST>(kNegInf)) {
    s += __expf(static_cast<float>(v - m));
  }
}
template <typename ST>
__device__ __forceinline__ void lse_merge(ST& m, float& s, ST m2, float s2) {
  if (m2 > m) {
    s = s * __expf(static_cast<float>(m - m2)) + s2;
    m = m2;
  } else if (m2 > static_cast<ST>(kNegInf)) {
    s += s2 * __expf(static_cast<float>(m2 - m));
  }
}
template <typename ST>
__device__ __forceinline__ ST lse_value(ST m, float s) {
  return (m == static_cast<ST>(kNegInf)) ? static_cast<ST>(kNegInf) : m + static_cast<ST>(logf(s));
}

// block-wide combine; result valid in every thread
template <typename ST>
__device__ ST block_lse(ST m, float s) {
  __shared__ double red_m[32];
  __shared__ float red_s[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const ST m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();  // protect red_* against a previous call
  if (lane == 0) {
    red_m[wid] = static_cast<double>(m);
    red_s[wid] = s;
  }
  __syncthreads();
  m = lane < nw ? static_cast<ST>(red_m[lane]) : static_cast<ST>(kNegInf);
  s = lane < nw ? red_s[lane] : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const ST m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  return lse_value(m, s);
}

// ---- shared memory carve-up (identical on host and device), offsets in bytes ----------
// [window ST[W]] [delta window f32[W]] [2 stages x {nbr int[cap+8], aux int[cap+8], lab int[cap+8],
// ptr int[cap+8]}] [theta f32[V]] [dtheta f32[V]]
struct SmemPlan {
  size_t win, dwin, stage0, stage_bytes, nbr, aux, wsc, lab, ptr, theta, dtheta, bytes;
};
__host__ __device__ inline SmemPlan smem_plan(int W, int cap, int st_bytes, int vocab, bool logs, bool trop,
                                              bool with_scores, bool with_labels, bool with_theta, bool with_dtheta,
                                              int n_stages = 2, bool gathered_scores = false) {
  SmemPlan p;
  size_t o = 0;
  const size_t arr = static_cast<size_t>(cap + 8) * 4;
  p.win = o;
  o += logs ? static_cast<size_t>(W) * st_bytes : 0;
  p.dwin = o;
  o += trop ? static_cast<size_t>(W) * 4 : 0;
  o = (o + 15) & ~static_cast<size_t>(15);
  p.stage0 = o;
  size_t q = 0;
  p.nbr = q; q += arr;                      // neighbour state of every arc (src_in / dst_out)
  p.aux = q; q += with_scores ? arr : 0;    // forward: in2out index; backward: the score itself
  p.wsc = p.aux; (void)gathered_scores;      // forward: scores are gathered IN PLACE over their indices
  p.lab = q; q += with_labels ? arr : 0;
  p.ptr = q; q += arr;                      // CSR row pointers of the chunk's states
  p.stage_bytes = q;
  o += static_cast<size_t>(n_stages) * q;
  p.theta = o;
  o += with_theta ? static_cast<size_t>(vocab) * 4 : 0;
  p.dtheta = o;
  o += with_dtheta ? static_cast<size_t>(vocab) * 4 : 0;
  p.bytes = (o + 15) & ~static_cast<size_t>(15);
  return p;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// 2^x and log2(x) on the SFU, flush-to-zero (MUFU.EX2 / MUFU.LG2, no range fix-up code)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Online logsumexp with a FINITE floor for the running max (kFloor instead of -inf), so the
// update is branch-free and NaN-free: v = -inf gives d = -inf, e = 0; the first finite v
// gives d ~ +1e30, e = 0, (m, s) = (v, 1).  Exactly one exp per element.
constexpr float kFloor = -1.0e30f;
template <typename ST>
__device__ __forceinline__ void lse_push(ST& m, float& s, ST v, ST) {
  const float d = static_cast<float>(v - m);
  const float e = ex2_approx(-fabsf(d) * kLog2e);
  const bool up = d > 0.0f;
  s = up ? fmaf(s, e, 1.0f) : s + e;
  m = up ? v : m;
}
// Four values at once: one max tree, then five INDEPENDENT exps (4 terms + the rescale of the
// running sum) instead of a 4-long dependent chain through (m, s).
template <typename ST>
__device__ __forceinline__ void lse_push4(ST& m, float& s, ST v0, ST v1, ST v2, ST v3) {
  const ST mx = max(max(max(v0, v1), max(v2, v3)), m);
  const float r = ex2_approx(static_cast<float>(m - mx) * kLog2e);
  const float e0 = ex2_approx(static_cast<float>(v0 - mx) * kLog2e);
  const float e1 = ex2_approx(static_cast<float>(v1 - mx) * kLog2e);
  const float e2 = ex2_approx(static_cast<float>(v2 - mx) * kLog2e);
  const float e3 = ex2_approx(static_cast<float>(v3 - mx) * kLog2e);
  s = fmaf(s, r, (e0 + e1) + (e2 + e3));
  m = mx;
}
template <typename ST>
__device__ __forceinline__ void lse_join(ST& m, float& s, ST m2, float s2, ST) {
  const float d = static_cast<float>(m2 - m);
  const float e = ex2_approx(-fabsf(d) * kLog2e);
  const bool up = d > 0.0f;
  s = up ? fmaf(s, e, s2) : fmaf(s2, e, s);
  m = up ? m2 : m;
}
template <typename ST>
__device__ __forceinline__ ST lse_finish(ST m, float s, ST neg_inf) {
  return (m > static_cast<ST>(kFloor)) ? m + static_cast<ST>(lg2_approx(s) * kLn2) : neg_inf;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// All dynamic shared memory is addressed as 32-bit words off one typed extern array (keeps
// the accesses in the shared address space for the compiler: plain LDS/STS with immediate
// offsets, no generic-pointer conversions in the loops).
extern __shared__ __align__(16) float smem_f[];

// Stage the arc arrays and row pointers of chunk k into the stage at word offset `st`
// (16-byte cp.async copies, all threads).  Copies start at the 4-aligned position below the
// chunk and may run a few elements past it; the arrays are zero-padded, so whatever is
// over-read is a valid index.
template <bool AUX, bool LAB>
__device__ __forceinline__ void stage_chunk(const int4& k, int cap, int st, const SmemPlan& plan,
                                            const int32_t* __restrict__ nbr, const void* __restrict__ aux,
                                            const int32_t* __restrict__ lab, const int32_t* __restrict__ ptr) {
  if (k.w > k.z && k.y - k.x <= cap) {
    const int base4 = k.x & ~3;
    const int n4 = (k.y - base4 + 3) >> 2;
    float* s_nbr = smem_f + st + plan.nbr / 4;
    float* s_aux = smem_f + st + plan.aux / 4;
    float* s_lab = smem_f + st + plan.lab / 4;
    float* s_ptr = smem_f + st + plan.ptr / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      cp_async16(s_nbr + 4 * i, nbr + base4 + 4 * i);
      if (AUX) cp_async16(s_aux + 4 * i, static_cast<const int32_t*>(aux) + base4 + 4 * i);
      if (LAB) cp_async16(s_lab + 4 * i, lab + base4 + 4 * i);
    }
    const int pb = k.z & ~3;
    const int p4 = (k.w + 1 - pb + 3) >> 2;
    for (int i = threadIdx.x; i < p4; i += blockDim.x) cp_async16(s_ptr + 4 * i, ptr + pb + 4 * i);
  }
}

// Gather the scores of a staged chunk through its staged index array, straight into shared
// memory (4-byte cp.async, no registers): scores[idx[pos]] -> wsc[pos] for every staged slot.
__device__ __forceinline__ void gather_scores(const int4& k, int cap, int st, const SmemPlan& plan,
                                              const float* __restrict__ scores) {
  if (k.w > k.z && k.y - k.x <= cap) {
    const int n_slots = ((k.y - (k.x & ~3) + 3) >> 2) << 2;
    const int* s_idx = reinterpret_cast<const int*>(smem_f + st + plan.aux / 4);
    float* s_w = smem_f + st + plan.wsc / 4;
    // in place: each slot's index is read (LDS) before the copy that overwrites it is issued
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) cp_async4(s_w + i, scores + s_idx[i]);
  }
}

// rare path: a neighbour older than the shared-memory window (kept out of line so that the
// hot loops carry no 64-bit address arithmetic)
template <typename T>
__device__ __noinline__ T load_behind_window(const T* p, int i) { return p[i]; }

// ---- optional per-phase cycle counters (build with -DNFST_TIMING; see tools/phase_timing.py)
#ifdef NFST_TIMING
__device__ unsigned long long g_dbg[32];
#define NFST_T(var) const long long var = clock64()
#define NFST_ACC(cond, slot, cycles) \
  do { if (cond) atomicAdd(&g_dbg[slot], static_cast<unsigned long long>(cycles)); } while (0)
#else
#define NFST_T(var) do {} while (0)
#define NFST_ACC(cond, slot, cycles) do {} while (0)
#endif

// =====================================================================================
// forward: alpha
// =====================================================================================
// SC: per-arc scores given; TH: theta[label] given (at least one of them).
template <typename ST, bool SC, bool TH>
__global__ void __launch_bounds__(256, NFST_MIN_BLOCKS)
    nfst_fwd_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int W, int cap,
                    const float* __restrict__ arc_scores, const float* __restrict__ theta, int theta_smem, ST* alpha,
                    ST* __restrict__ logz) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const SmemPlan plan = smem_plan(W, cap, sizeof(ST), L.vocab, true, false, SC, TH, theta_smem != 0, false, 3, true);
  ST* const win = reinterpret_cast<ST*>(smem_f + plan.win / 4);
  const ST neg_inf = static_cast<ST>(kNegInf);
  const int stage0 = static_cast<int>(plan.stage0 / 4), stage_words = static_cast<int>(plan.stage_bytes / 4);
  const int o_nbr = static_cast<int>(plan.nbr / 4), o_wsc = static_cast<int>(plan.wsc / 4),
            o_lab = static_cast<int>(plan.lab / 4), o_ptr = static_cast<int>(plan.ptr / 4);

  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int base_s = L.state_off[b];
  const int start = L.start_state[b];
  const int wmask = W - 1;
  const bool whole = (L.state_off[b + 1] - base_s) <= W;  // the whole lattice fits the window
  const float* th = theta;
  if (TH && theta_smem) {
    float* sth = smem_f + plan.theta / 4;
    for (int i = tid; i < L.vocab; i += NT) sth[i] = theta[i];
    th = sth;
  }

  const nfst_chunk_t* chunks = L.fwd_chunks;
  int c = L.fwd_chunk_off[b];
  const int c_end = L.fwd_chunk_off[b + 1];

  // Three-stage software pipeline, all through cp.async (no staging registers):
  //   chunk c   : arc arrays + gathered scores resident  -> reduced now
  //   chunk c+1 : arc arrays resident                     -> its scores are gathered now (4 B copies)
  //   chunk c+2 :                                            its arc arrays stream in now (16 B copies)
  int4 k0 = chunk_at(chunks, c, c_end), k1 = chunk_at(chunks, c + 1, c_end), k2 = chunk_at(chunks, c + 2, c_end);
  int s_cur = stage0, s_nxt = stage0 + stage_words, s_nn = stage0 + 2 * stage_words;
  stage_chunk<SC, TH>(k0, cap, s_cur, plan, L.src_in, L.in2out, L.label_in, L.in_ptr);
  stage_chunk<SC, TH>(k1, cap, s_nxt, plan, L.src_in, L.in2out, L.label_in, L.in_ptr);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  if (SC) gather_scores(k0, cap, s_cur, plan, arc_scores);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

#pragma unroll 1
  for (; c < c_end; ++c) {
    const int4 k3 = chunk_at(chunks, c + 3, c_end);  // descriptors run ahead of their use
    stage_chunk<SC, TH>(k2, cap, s_nn, plan, L.src_in, L.in2out, L.label_in, L.in_ptr);
    if (SC) gather_scores(k1, cap, s_nxt, plan, arc_scores);
    cp_async_commit();
    const int* s_src = reinterpret_cast<const int*>(smem_f + s_cur + o_nbr);
    const float* s_w = smem_f + s_cur + o_wsc;
    const int* s_lab = reinterpret_cast<const int*>(smem_f + s_cur + o_lab);

    const int a0 = k0.x, a1 = k0.y, s0 = k0.z, s1 = k0.w;
    const int n = a1 - a0, ns = s1 - s0;
    const int base4 = a0 & ~3;
    // row pointers relative to the stage arrays
    const int* s_ptr = reinterpret_cast<const int*>(smem_f + s_cur + o_ptr) + (s0 & 3);
    // readable window: states [s1 - W, s0).  The slots of [s0 - W, s1 - W) are being
    // overwritten by this chunk's own results, so those states are re-read from global.
    const int lo = whole ? static_cast<int>(0x80000000) : s1 - W;
    auto value_of = [&](int src) -> ST {
      if (src >= lo) return win[(src - base_s) & wmask];
      return load_behind_window(alpha, src);
    };
    auto arc_value = [&](int i) -> ST {
      float w = SC ? s_w[i] : 0.0f;
      if (TH) w += th[s_lab[i]];
      return value_of(s_src[i]) + static_cast<ST>(w);
    };
    if (n <= cap) {
      if (ns * 2 > NT) {
        // ---- wide chunk: one thread per state, arcs straight from the staged arrays
#pragma unroll 1
        for (int j = tid; j < ns; j += NT) {
          int i = s_ptr[j] - base4;
          const int b1 = s_ptr[j + 1] - base4;
          ST m = static_cast<ST>(kFloor);
          float sum = 0.0f;
#pragma unroll 1
          for (; i + 3 < b1; i += 4) {  // 4 arcs per trip: independent loads first
            const ST v0 = arc_value(i), v1 = arc_value(i + 1), v2 = arc_value(i + 2), v3 = arc_value(i + 3);
            lse_push4(m, sum, v0, v1, v2, v3);
          }
#pragma unroll 1
          for (; i < b1; ++i) lse_push(m, sum, arc_value(i), neg_inf);
          const int s = s0 + j;
          const ST v = (s == start) ? static_cast<ST>(0) : lse_finish(m, sum, neg_inf);
          if (s >= s1 - W) win[(s - base_s) & wmask] = v;  // only the newest W states own a slot
          alpha[s] = v;
        }
      } else {
        // ---- narrow chunk: 2^lg lanes per state so that the block stays busy
        int lg = 1;
        while (lg < 5 && (ns << (lg + 1)) <= NT) ++lg;
        const int G = 1 << lg, lane_g = tid & (G - 1), ngrp = NT >> lg;
#pragma unroll 1
        for (int jb = 0; jb < ns; jb += ngrp) {
          const int j = jb + (tid >> lg);
          const bool valid = j < ns;
          const int b0 = valid ? s_ptr[j] - base4 : 0, b1 = valid ? s_ptr[j + 1] - base4 : 0;
          ST m = static_cast<ST>(kFloor);
          float sum = 0.0f;
#pragma unroll 1
          for (int i = b0 + lane_g; i < b1; i += G) lse_push(m, sum, arc_value(i), neg_inf);
          for (int o = G >> 1; o > 0; o >>= 1) {
            const ST m2 = __shfl_xor_sync(0xffffffffu, m, o);
            const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
            lse_join(m, sum, m2, s2, neg_inf);
          }
          if (valid && lane_g == 0) {
            const int s = s0 + j;
            const ST v = (s == start) ? static_cast<ST>(0) : lse_finish(m, sum, neg_inf);
            if (s >= s1 - W) win[(s - base_s) & wmask] = v;
            alpha[s] = v;
          }
        }
      }
    } else {
      // ---- oversize chunk (a state with more than `cap` incoming arcs): block-wide, from global
      for (int j = 0; j < ns; ++j) {
        const int s = s0 + j;
        const int b0 = L.in_ptr[s], b1 = L.in_ptr[s + 1];
        ST m = neg_inf;
        float sum = 0.0f;
        for (int a = b0 + tid; a < b1; a += NT) {
          float w = 0.0f;
          if (SC) w = arc_scores[L.in2out[a]];
          if (TH) w += th[L.label_in[a]];
          lse_add(m, sum, static_cast<ST>(w) + value_of(L.src_in[a]));
        }
        const ST v0 = block_lse(m, sum);
        if (tid == 0) {
          const ST v = (s == start) ? static_cast<ST>(0) : v0;
          win[(s - base_s) & wmask] = v;
          alpha[s] = v;
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();  // alpha of chunk c visible; scores of c+1 and arrays of c+2 landed; stage of c free
    k0 = k1; k1 = k2; k2 = k3;
    const int t = s_cur; s_cur = s_nxt; s_nxt = s_nn; s_nn = t;
  }

  // logZ = logsumexp over the sinks of alpha (every zero-out-degree state has beta = 1,
  // scorers.py:795-805); alpha was written by this block and the loop ended on a barrier
  ST m = neg_inf;
  float sum = 0.0f;
  for (int i = L.sink_off[b] + tid; i < L.sink_off[b + 1]; i += NT) lse_add(m, sum, alpha[L.sinks[i]]);
  const ST z = block_lse(m, sum);
  if (tid == 0) logz[b] = z;
}

// =====================================================================================
// fused backward: beta (+ posteriors, dtheta) and/or Viterbi delta + backpointer
// =====================================================================================
// LOGS / TROP: semirings; SC / TH: score sources; POST: posteriors (post and/or dtheta).
// dtheta histogram of one block (= one lattice's arcs): in shared memory the posteriors are accumulated WITHOUT the
// lattice's gradient scale, in fixed point with native integer atomics (a float add there is a compare-and-swap loop,
// and thousands of float adds per label lose ~1e-5: measured 2.4e-5 on 200-level bigram cipher lattices); a label's
// expected count is at most the lattice's level count, so the unit is 2^31 / (levels rounded up to a power of two).
// Without the shared-memory table (V > NFST_THETA_SMEM_MAX) the scaled posterior goes to dtheta with a float atomic.
struct Hist {
  unsigned* fix;  // shared memory, fixed point (or null)
  float* glob;    // dtheta itself (or null)
  float scale, inv;
};
__device__ __forceinline__ Hist hist_make(unsigned* fix, float* glob, int levels) {
  Hist h;
  h.fix = fix;
  h.glob = glob;
  const int lv = levels < 1 ? 1 : levels;
  const float p2 = static_cast<float>(1u << (32 - __clz(lv - 1 > 0 ? lv - 1 : 0)));  // >= levels, a power of two (1 for lv = 1)
  h.scale = 2147483648.0f / p2;
  h.inv = p2 / 2147483648.0f;
  return h;
}
// (callers guard the call with `if (h.fix || h.glob)`: the label expression may read an array that only exists with a histogram)
__device__ __forceinline__ void hist_add(const Hist& h, int lab, float pu, float gscale) {
  if (h.fix) atomicAdd(&h.fix[lab], __float2uint_rn(pu * h.scale));
  else if (h.glob) atomicAdd(&h.glob[lab], pu * gscale);
}
__device__ __forceinline__ void hist_flush(const Hist& h, float* dtheta, int vocab, float gscale, int tid, int nt) {
  const float sc = h.inv * gscale;
  for (int i = tid; i < vocab; i += nt) {
    const unsigned v = h.fix[i];
    if (v) atomicAdd(&dtheta[i], static_cast<float>(v) * sc);
  }
}

template <typename ST, bool LOGS, bool TROP, bool SC, bool TH, bool POST>
__global__ void __launch_bounds__(256, NFST_MIN_BLOCKS)
    nfst_bwd_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int W, int cap,
                    const float* __restrict__ arc_scores, const float* __restrict__ theta, int theta_smem,
                    int dtheta_smem, const ST* __restrict__ alpha, const ST* __restrict__ logz,
                    const float* __restrict__ grad_logz, ST* beta, ST* __restrict__ logz_bwd, float* __restrict__ post,
                    float* __restrict__ dtheta, float* delta, int32_t* __restrict__ backptr,
                    float* __restrict__ vit_score) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const bool want_hist = POST && dtheta != nullptr;
  const bool need_label = TH || want_hist;
  const SmemPlan plan =
      smem_plan(W, cap, sizeof(ST), L.vocab, LOGS, TROP, SC, need_label, theta_smem != 0, dtheta_smem != 0);
  ST* const win = reinterpret_cast<ST*>(smem_f + plan.win / 4);
  float* const dwin = smem_f + plan.dwin / 4;
  const ST neg_inf = static_cast<ST>(kNegInf);
  const int stage0 = static_cast<int>(plan.stage0 / 4), stage_words = static_cast<int>(plan.stage_bytes / 4);
  const int o_nbr = static_cast<int>(plan.nbr / 4), o_aux = static_cast<int>(plan.aux / 4),
            o_lab = static_cast<int>(plan.lab / 4), o_ptr = static_cast<int>(plan.ptr / 4);

  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int base_s = L.state_off[b];
  const int wmask = W - 1;
  const bool whole = (L.state_off[b + 1] - base_s) <= W;
  const float* th = theta;
  if (TH && theta_smem) {
    float* sth = smem_f + plan.theta / 4;
    for (int i = tid; i < L.vocab; i += NT) sth[i] = theta[i];
    th = sth;
  }
  unsigned* hist_fix = nullptr;
  if (want_hist && dtheta_smem) {
    hist_fix = reinterpret_cast<unsigned*>(smem_f + plan.dtheta / 4);
    for (int i = tid; i < L.vocab; i += NT) hist_fix[i] = 0u;
  }
  const Hist H = hist_make(hist_fix, want_hist && !dtheta_smem ? dtheta : nullptr, L.level_off[b + 1] - L.level_off[b] - 1);
  ST lz = 0;
  float gscale = 1.0f;
  if (POST) {
    lz = logz[b];
    if (grad_logz) gscale = grad_logz[b];
  }

  const nfst_chunk_t* chunks = L.bwd_chunks;
  int c = L.bwd_chunk_off[b];
  const int c_end = L.bwd_chunk_off[b + 1];
  const int32_t* lab_arr = need_label ? L.label_out : nullptr;

  int4 k0 = chunk_at(chunks, c, c_end), k1 = chunk_at(chunks, c + 1, c_end);
  int s_cur = stage0, s_nxt = stage0 + stage_words;
  auto stage_in = [&](const int4& k, int st) {
    if (need_label) stage_chunk<SC, true>(k, cap, st, plan, L.dst_out, arc_scores, lab_arr, L.out_ptr);
    else stage_chunk<SC, false>(k, cap, st, plan, L.dst_out, arc_scores, lab_arr, L.out_ptr);
  };
  stage_in(k0, s_cur);
  cp_async_commit();
  int sid_cur = (LOGS && k0.z + tid < k0.w) ? __ldg(L.bwd_order + k0.z + tid) : 0;
  cp_async_wait_all();
  __syncthreads();

#pragma unroll 1
  for (; c < c_end; ++c) {
    const int4 k2 = chunk_at(chunks, c + 2, c_end);
    stage_in(k1, s_nxt);
    cp_async_commit();
    const int* s_dst = reinterpret_cast<const int*>(smem_f + s_cur + o_nbr);
    const float* s_w = smem_f + s_cur + o_aux;
    const int* s_lab = reinterpret_cast<const int*>(smem_f + s_cur + o_lab);

    const int a0 = k0.x, a1 = k0.y, s0 = k0.z, s1 = k0.w;
    const int n = a1 - a0, ns = s1 - s0;
    const int base4 = a0 & ~3;
    const int* s_ptr = reinterpret_cast<const int*>(smem_f + s_cur + o_ptr) + (s0 & 3);
    // readable window: states [s1, s0 + W) (see the forward kernel)
    const int hi = whole ? 0x7fffffff : s0 + W;
    auto beta_of = [&](int d) -> ST {
      if (d < hi) return win[(d - base_s) & wmask];
      return load_behind_window(beta, d);
    };
    auto delta_of = [&](int d) -> float {
      if (d < hi) return dwin[(d - base_s) & wmask];
      return load_behind_window(delta, d);
    };
    // one arc of the staged chunk: log-semiring push (+ posterior) and/or tropical candidate
    auto visit = [&](int i, ST am, ST& m, float& sum, float& bt, int& bi) {
      const int d = s_dst[i];
      float w = SC ? s_w[i] : 0.0f;
      int lab = 0;
      if (TH || POST) { if (need_label) lab = s_lab[i]; }
      if (TH) w += th[lab];
      if (LOGS) {
        const ST u = static_cast<ST>(w) + beta_of(d);
        lse_push(m, sum, u, neg_inf);
        if (POST) {
          const float pu = ex2_approx(static_cast<float>(am + u) * kLog2e);
          const float p = pu * gscale;
          if (post) post[base4 + i] = p;
          if (H.fix || H.glob) hist_add(H, lab, pu, gscale);
        }
      }
      if (TROP) {
        const float t = __fadd_rn(w, delta_of(d));
        // strict: within a lane arcs come in label order; the first candidate is taken whatever its score, so a
        // state whose arcs all score -inf (or NaN) still gets a valid backpointer
        if (t > bt || bi == 0x7fffffff) { bt = t; bi = i; }
      }
    };
    // four consecutive arcs of one state: independent loads and exps (see lse_push4)
    auto visit4 = [&](int i, ST am, ST& m, float& sum, float& bt, int& bi) {
      ST u[4];
      int lab[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = s_dst[i + k];
        float w = SC ? s_w[i + k] : 0.0f;
        if (TH || POST) { if (need_label) lab[k] = s_lab[i + k]; }
        if (TH) w += th[lab[k]];
        if (LOGS) u[k] = static_cast<ST>(w) + beta_of(d);
        if (TROP) {
          const float t = __fadd_rn(w, delta_of(d));
          if (t > bt || bi == 0x7fffffff) { bt = t; bi = i + k; }
        }
      }
      if (LOGS) {
        lse_push4(m, sum, u[0], u[1], u[2], u[3]);
        if (POST) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float pu = ex2_approx(static_cast<float>(am + u[k]) * kLog2e);
          const float p = pu * gscale;
            if (post) post[base4 + i + k] = p;
            if (H.fix || H.glob) hist_add(H, lab[k], pu, gscale);
          }
        }
      }
    };
    auto finish = [&](int s, bool sink, ST m, float sum, float bt, int bi) {
      // sinks: beta = 1 (scorers.py:720), delta = 0
      if (LOGS) {
        const ST v = sink ? static_cast<ST>(0) : lse_finish(m, sum, neg_inf);
        if (s < s0 + W) win[(s - base_s) & wmask] = v;
        beta[s] = v;
      }
      if (TROP) {
        const float v = sink ? 0.0f : bt;
        if (s < s0 + W) dwin[(s - base_s) & wmask] = v;
        delta[s] = v;
        backptr[s] = sink ? -1 : base4 + bi;
      }
    };
    // first state of this thread in the NEXT chunk's processing order (consumed next iteration)
    const int sid_next = (LOGS && k1.z + tid < k1.w) ? __ldg(L.bwd_order + k1.z + tid) : 0;
    if (n <= cap) {
      if (ns * 2 > NT) {
#pragma unroll 1
        for (int j = tid; j < ns; j += NT) {
          // states are visited in out-degree order, so a warp's 32 states have equal trip counts
          // (log semiring only: the tropical-only pass has too little work per arc to pay for it)
          const int s = !LOGS ? s0 + j : ((j == tid) ? sid_cur : __ldg(L.bwd_order + s0 + j));
          const int b0 = s_ptr[s - s0] - base4, b1 = s_ptr[s - s0 + 1] - base4;
          ST am = 0;
          if (POST) am = alpha[s] - lz;  // issued before the arc loop, consumed inside it
          ST m = static_cast<ST>(kFloor);
          float sum = 0.0f, bt = kNegInf;
          int bi = 0x7fffffff;
          int i = b0;
#pragma unroll 1
          for (; i + 3 < b1; i += 4) visit4(i, am, m, sum, bt, bi);
#pragma unroll 1
          for (; i < b1; ++i) visit(i, am, m, sum, bt, bi);
          finish(s, b0 == b1, m, sum, bt, bi);
        }
      } else {
        int lg = 1;
        while (lg < 5 && (ns << (lg + 1)) <= NT) ++lg;
        const int G = 1 << lg, lane_g = tid & (G - 1), ngrp = NT >> lg;
#pragma unroll 1
        for (int jb = 0; jb < ns; jb += ngrp) {
          const int j = jb + (tid >> lg);
          const bool valid = j < ns;
          const int b0 = valid ? s_ptr[j] - base4 : 0, b1 = valid ? s_ptr[j + 1] - base4 : 0;
          const int s = s0 + j;
          ST am = 0;
          if (POST && valid) am = alpha[s] - lz;
          ST m = static_cast<ST>(kFloor);
          float sum = 0.0f, bt = kNegInf;
          int bi = 0x7fffffff;
#pragma unroll 1
          for (int i = b0 + lane_g; i < b1; i += G) visit(i, am, m, sum, bt, bi);
          for (int o = G >> 1; o > 0; o >>= 1) {
            if (LOGS) {
              const ST m2 = __shfl_xor_sync(0xffffffffu, m, o);
              const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
              lse_join(m, sum, m2, s2, neg_inf);
            }
            if (TROP) {
              const float t2 = __shfl_xor_sync(0xffffffffu, bt, o);
              const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
              if (t2 > bt || (t2 == bt && i2 < bi)) { bt = t2; bi = i2; }
            }
          }
          if (valid && lane_g == 0) finish(s, b0 == b1, m, sum, bt, bi);
        }
      }
    } else {
      // ---- oversize chunk: one state at a time, block-wide, from global
      for (int j = 0; j < ns; ++j) {
        const int s = s0 + j;
        const int b0 = L.out_ptr[s], b1 = L.out_ptr[s + 1];
        ST am = 0;
        if (POST) am = alpha[s] - lz;
        ST m = neg_inf;
        float sum = 0.0f;
        float bt = kNegInf;
        int bi = 0x7fffffff;
        for (int a = b0 + tid; a < b1; a += NT) {
          const int d = L.dst_out[a];
          int lab = 0;
          if (need_label) lab = L.label_out[a];
          float w = SC ? arc_scores[a] : 0.0f;
          if (TH) w += th[lab];
          if (LOGS) {
            const ST u = static_cast<ST>(w) + beta_of(d);
            lse_add(m, sum, u);
            if (POST) {
              const float pu = ex2_approx(static_cast<float>(am + u) * kLog2e);
          const float p = pu * gscale;
              if (post) post[a] = p;
              if (H.fix || H.glob) hist_add(H, lab, pu, gscale);
            }
          }
          if (TROP) {
            const float t = __fadd_rn(w, delta_of(d));
            if (t > bt || (t == bt && a < bi)) { bt = t; bi = a; }
          }
        }
        if (LOGS) {
          const ST v0 = block_lse(m, sum);
          if (tid == 0) {
            const ST v = (b0 == b1) ? static_cast<ST>(0) : v0;
            win[(s - base_s) & wmask] = v;
            beta[s] = v;
          }
        }
        if (TROP) {
          __shared__ float red_t[32];
          __shared__ int red_i[32];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float t2 = __shfl_xor_sync(0xffffffffu, bt, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
            if (t2 > bt || (t2 == bt && i2 < bi)) { bt = t2; bi = i2; }
          }
          __syncthreads();
          if ((tid & 31) == 0) { red_t[tid >> 5] = bt; red_i[tid >> 5] = bi; }
          __syncthreads();
          if (tid == 0) {
            for (int w2 = 1; w2 < (NT + 31) / 32; ++w2)
              if (red_t[w2] > bt || (red_t[w2] == bt && red_i[w2] < bi)) { bt = red_t[w2]; bi = red_i[w2]; }
            const bool sink = (b0 == b1);
            const float v = sink ? 0.0f : bt;
            dwin[(s - base_s) & wmask] = v;
            delta[s] = v;
            backptr[s] = sink ? -1 : bi;
          }
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
    k0 = k1; k1 = k2; sid_cur = sid_next;
    const int t = s_cur; s_cur = s_nxt; s_nxt = t;
  }

  if (tid == 0) {
    const int start = L.start_state[b];
    if (LOGS && logz_bwd) logz_bwd[b] = beta[start];
    if (TROP && vit_score) vit_score[b] = delta[start];
  }
  if (want_hist && dtheta_smem) hist_flush(H, dtheta, L.vocab, gscale, tid, NT);
}

// =====================================================================================
// small lattices (the size nFST's own lattices have: hundreds to a few thousand arcs): the
// WHOLE lattice -- CSR arrays, scores in both arc orders, DP vectors -- is loaded into shared
// memory once, then every level costs shared-memory latency + one barrier.  Forward, backward,
// posteriors and the Viterbi recursion can run in ONE launch (DO_FWD && DO_BWD), with alpha
// never leaving the SM; or as two launches around an autograd boundary.
// =====================================================================================
// optional best-path read-out fused into the tropical pass (small-lattice kernel): all null = off
struct PathOut {
  const int32_t* path_off;  // [B+1] capacity offsets into path_arcs
  int32_t* path_arcs;       // canonical arc ids, start -> sink
  int32_t* path_len;        // [B]
};

struct SmallPlan {  // offsets in 32-bit words
  int a, b, d, lp, lg, inp, outp, src, win, dst, wout, lab, bp, hist, words;
};
// Lattice-local state and arc indices are 16-bit (a small lattice has < 65536 of either).
typedef unsigned short small_idx_t;
__host__ __device__ inline SmallPlan small_plan(int S, int A, int Lv, int st_words, bool fwd, bool bwd, bool trop,
                                                bool labels, int hist_words) {
  SmallPlan p;
  int o = 0;
  auto take = [&](int n) { const int at = o; o += (n + 3) & ~3; return at; };
  auto take16 = [&](int n) { return take((n + 1) >> 1); };
  p.a = take(S * st_words);
  p.b = take(bwd ? S * st_words : 0);
  p.d = take(trop ? S : 0);
  p.bp = take(trop ? S : 0);
  p.lp = take(Lv + 1);
  p.lg = take(Lv + 1);
  p.inp = take16(fwd ? S + 1 : 0);
  p.outp = take16(bwd ? S + 1 : 0);
  p.src = take16(fwd ? A : 0);
  p.win = take(fwd ? A : 0);
  p.dst = take16(bwd ? A : 0);
  p.wout = take(bwd ? A : 0);
  p.lab = take(labels ? A : 0);
  p.hist = take(hist_words);
  p.words = o;
  return p;
}

template <typename ST, bool SC, bool TH, bool DO_FWD, bool DO_BWD, bool LOGS, bool TROP, bool POST>
__global__ void __launch_bounds__(256, 4)
    nfst_small_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int capS, int capA, int capL,
                      const float* __restrict__ arc_scores, const float* __restrict__ theta, int dtheta_smem,
                      ST* __restrict__ alpha_g, ST* __restrict__ logz_g, const float* __restrict__ grad_logz,
                      ST* __restrict__ beta_g, ST* __restrict__ logz_bwd, float* __restrict__ post,
                      float* __restrict__ dtheta, float* __restrict__ delta_g, int32_t* __restrict__ backptr,
                      float* __restrict__ vit_score, const PathOut paths) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const ST neg_inf = static_cast<ST>(kNegInf);
  const bool want_hist = POST && dtheta != nullptr;
  const SmallPlan P = small_plan(capS, capA, capL, sizeof(ST) / 4, DO_FWD, DO_BWD, DO_BWD && TROP, want_hist,
                                 (want_hist && dtheta_smem) ? L.vocab : 0);
  ST* const sA = reinterpret_cast<ST*>(smem_f + P.a);
  ST* const sB = reinterpret_cast<ST*>(smem_f + P.b);
  float* const sD = smem_f + P.d;
  int* const sBP = reinterpret_cast<int*>(smem_f + P.bp);
  int* const s_lp = reinterpret_cast<int*>(smem_f + P.lp);
  small_idx_t* const s_inp = reinterpret_cast<small_idx_t*>(smem_f + P.inp);
  small_idx_t* const s_outp = reinterpret_cast<small_idx_t*>(smem_f + P.outp);
  small_idx_t* const s_src = reinterpret_cast<small_idx_t*>(smem_f + P.src);
  float* const s_win = smem_f + P.win;
  small_idx_t* const s_dst = reinterpret_cast<small_idx_t*>(smem_f + P.dst);
  float* const s_wout = smem_f + P.wout;
  int* const s_lab = reinterpret_cast<int*>(smem_f + P.lab);

  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int s_lo = L.state_off[b], Sb = L.state_off[b + 1] - s_lo;
  const int a_lo = L.out_ptr[s_lo], Ab = L.out_ptr[s_lo + Sb] - a_lo;  // arcs are grouped by lattice in both orders
  const int lvl0 = L.level_off[b], nlev = L.level_off[b + 1] - lvl0 - 1;
  const int start = L.start_state[b] - s_lo;
  unsigned* hist_fix = nullptr;
  if (want_hist && dtheta_smem) {
    hist_fix = reinterpret_cast<unsigned*>(smem_f + P.hist);
    for (int i = tid; i < L.vocab; i += NT) hist_fix[i] = 0u;
  }
  const Hist H = hist_make(hist_fix, want_hist && !dtheta_smem ? dtheta : nullptr, L.level_off[b + 1] - L.level_off[b] - 1);

  NFST_T(t_start);
  // ---- load the lattice (everything made lattice-local); all global loads are independent
  for (int i = tid; i <= nlev; i += NT) s_lp[i] = L.level_ptr[lvl0 + i] - s_lo;
  for (int i = tid; i <= Sb; i += NT) {
    if (DO_FWD) s_inp[i] = static_cast<small_idx_t>(L.in_ptr[s_lo + i] - a_lo);
    if (DO_BWD) s_outp[i] = static_cast<small_idx_t>(L.out_ptr[s_lo + i] - a_lo);
  }
  for (int i = tid; i < Ab; i += NT) {
    if (DO_BWD) {
      float w = SC ? arc_scores[a_lo + i] : 0.0f;
      if (TH || want_hist) {
        const int lab = L.label_out[a_lo + i];
        if (TH) w += __ldg(theta + lab);
        if (want_hist) s_lab[i] = lab;
      }
      s_wout[i] = w;
      s_dst[i] = static_cast<small_idx_t>(L.dst_out[a_lo + i] - s_lo);
    }
    if (DO_FWD) {
      float w = SC ? arc_scores[L.in2out[a_lo + i]] : 0.0f;  // scores in in-order
      if (TH) w += __ldg(theta + L.label_in[a_lo + i]);
      s_win[i] = w;
      s_src[i] = static_cast<small_idx_t>(L.src_in[a_lo + i] - s_lo);
    }
  }
  if (!DO_FWD && DO_BWD && POST)
    for (int i = tid; i < Sb; i += NT) sA[i] = alpha_g[s_lo + i];
  if (DO_FWD && tid == 0) sA[start] = static_cast<ST>(0);
  __syncthreads();
  // Level schedule, once: 2^lg lanes per state for every level and direction -- about two arcs
  // per lane, never more lanes than fill the block; lg = 0 (one thread per state) is the lean
  // path below.  Kept in the spare high bits of the level pointers (bits 24-26 fwd, 28-30 bwd).
  constexpr int kLvMask = 0xffffff;
  int* const s_lg = reinterpret_cast<int*>(smem_f + P.lg);
  for (int l = tid; l < nlev; l += NT) {
    const int sb = s_lp[l], se = s_lp[l + 1], ns = se - sb;
    int enc = 0;
    if (DO_FWD) {
      const int na = s_inp[se] - s_inp[sb];
      int lg = 0;
      while (lg < 5 && (ns << (lg + 1)) <= NT && (ns << (lg + 2)) <= na) ++lg;
      enc |= lg << 24;
    }
    if (DO_BWD) {
      const int na = s_outp[se] - s_outp[sb];
      int lg = 0;
      while (lg < 5 && (ns << (lg + 1)) <= NT && (ns << (lg + 2)) <= na) ++lg;
      enc |= lg << 28;
    }
    s_lg[l] = enc;
  }
  __syncthreads();
  for (int l = tid; l < nlev; l += NT) s_lp[l] |= s_lg[l];
  __syncthreads();

  NFST_T(t_loaded);
  NFST_ACC(tid == 0, 16, t_loaded - t_start);
  NFST_ACC(tid == 0, 20, 1);
  NFST_ACC(tid == 0, 21, nlev);
  // One warp per lattice needs no block barrier.
  auto level_barrier = [&]() {
    if (NT == 32) __syncwarp(); else __syncthreads();
  };
  // Software pipeline over levels (registers): none of a level's index data depends on DP
  // values, so the level pointers are fetched three levels ahead, this thread's CSR range two
  // levels ahead and its first two arcs one level ahead.  What is left on the level-to-level
  // critical path is: DP value of the neighbour (LDS) -> add -> [max, exp, log] -> store -> barrier.
  struct Range { int b0, b1; };
  struct Arc2 { int n0, n1; float w0, w1; };
  auto load_range = [&](int enc, int enc_next, const small_idx_t* s_ptr) {
    const int s = (enc & kLvMask) + tid;
    Range r{0, 0};
    if (s < (enc_next & kLvMask)) { r.b0 = s_ptr[s]; r.b1 = s_ptr[s + 1]; }
    return r;
  };
  auto load_arcs = [&](const Range& r, const small_idx_t* s_nbr, const float* s_w) {
    Arc2 q{0, 0, 0.0f, 0.0f};
    if (r.b0 < r.b1) { q.n0 = s_nbr[r.b0]; q.w0 = s_w[r.b0]; }
    if (r.b0 + 1 < r.b1) { q.n1 = s_nbr[r.b0 + 1]; q.w1 = s_w[r.b0 + 1]; }
    return q;
  };

  ST lz = 0;
  if (DO_FWD) {
    // ---- forward over levels 1..L-1 (level 0 is the start state)
    auto lv = [&](int l) { return s_lp[min(l, max(nlev, 0))]; };
    // one state from its arcs: logsumexp in two passes (max, then exp-sum); one arc = plain add
    auto fwd_state = [&](int s, const Range& r, const Arc2& q) {
      const int deg = r.b1 - r.b0;
      const ST u0 = sA[q.n0] + static_cast<ST>(q.w0);
      ST v = u0;
      if (deg > 1) {
        const ST u1 = sA[q.n1] + static_cast<ST>(q.w1);
        ST m = max(max(u0, u1), static_cast<ST>(kFloor));
        for (int i = r.b0 + 2; i < r.b1; ++i) m = max(m, sA[s_src[i]] + static_cast<ST>(s_win[i]));
        float sum = ex2_approx(static_cast<float>(u0 - m) * kLog2e) + ex2_approx(static_cast<float>(u1 - m) * kLog2e);
        for (int i = r.b0 + 2; i < r.b1; ++i)
          sum += ex2_approx(static_cast<float>(sA[s_src[i]] + static_cast<ST>(s_win[i]) - m) * kLog2e);
        v = sum > 0.0f ? m + static_cast<ST>(lg2_approx(sum) * kLn2) : neg_inf;
      }
      sA[s] = v;
    };
    int e0 = lv(1), e1 = lv(2), e2 = lv(3), e3 = lv(4);
    Range r0 = load_range(e0, e1, s_inp), r1 = load_range(e1, e2, s_inp);
    Arc2 q0 = load_arcs(r0, s_src, s_win);
    for (int l = 1; l < nlev; ++l) {
      const int e4 = lv(l + 4);
      const Range r2 = load_range(e2, e3, s_inp);
      const Arc2 q1 = load_arcs(r1, s_src, s_win);
      const int sb = e0 & kLvMask, se = e1 & kLvMask, lg = (e0 >> 24) & 7;
      if (lg == 0) {
        int s = sb + tid;
        if (s < se) fwd_state(s, r0, q0);
        for (s += NT; s < se; s += NT) {
          const Range r{s_inp[s], s_inp[s + 1]};
          fwd_state(s, r, load_arcs(r, s_src, s_win));
        }
      } else {
        // lanes-per-state path: G lanes stride over a state's arcs, two shuffle trees (max, sum)
        const int G = 1 << lg, lane_g = tid & (G - 1), ngrp = NT >> lg, ns = se - sb;
        for (int jb = 0; jb < ns; jb += ngrp) {
          const int j = jb + (tid >> lg);
          const bool valid = j < ns;
          const int b0 = valid ? s_inp[sb + j] : 0, b1 = valid ? s_inp[sb + j + 1] : 0;
          const int i0 = b0 + lane_g, i1 = i0 + G;
          ST u0 = neg_inf, u1 = neg_inf;
          if (i0 < b1) u0 = sA[s_src[i0]] + static_cast<ST>(s_win[i0]);
          if (i1 < b1) u1 = sA[s_src[i1]] + static_cast<ST>(s_win[i1]);
          ST m = max(max(u0, u1), static_cast<ST>(kFloor));
          for (int i = i1 + G; i < b1; i += G) m = max(m, sA[s_src[i]] + static_cast<ST>(s_win[i]));
          for (int o = G >> 1; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
          float sum = ex2_approx(static_cast<float>(u0 - m) * kLog2e) + ex2_approx(static_cast<float>(u1 - m) * kLog2e);
          for (int i = i1 + G; i < b1; i += G)
            sum += ex2_approx(static_cast<float>(sA[s_src[i]] + static_cast<ST>(s_win[i]) - m) * kLog2e);
          for (int o = G >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (valid && lane_g == 0)
            sA[sb + j] = sum > 0.0f ? m + static_cast<ST>(lg2_approx(sum) * kLn2) : neg_inf;
        }
      }
      level_barrier();
      e0 = e1; e1 = e2; e2 = e3; e3 = e4;
      r0 = r1; r1 = r2; q0 = q1;
    }
    NFST_T(t_fwd);
    NFST_ACC(tid == 0, 17, t_fwd - t_loaded);
    // logZ = logsumexp over the sinks (scorers.py:795-805: every arc-less state has beta = 1)
    ST m = neg_inf;
    float sum = 0.0f;
    for (int i = L.sink_off[b] + tid; i < L.sink_off[b + 1]; i += NT) lse_add(m, sum, sA[L.sinks[i] - s_lo]);
    lz = block_lse(m, sum);
    if (tid == 0) logz_g[b] = lz;
    if (alpha_g)
      for (int i = tid; i < Sb; i += NT) alpha_g[s_lo + i] = sA[i];
  } else if (POST) {
    lz = logz_g[b];
  }

  if (DO_BWD) {
    float gscale = 1.0f;
    if (POST && grad_logz) gscale = grad_logz[b];
    // ---- backward over levels L-1..0: beta, posteriors, and/or the tropical recursion
    NFST_T(t_bwd0);
    auto lv = [&](int l) { return s_lp[max(l, 0)]; };       // level l's encoded pointer
    auto lv_end = [&](int l) { return s_lp[min(max(l, 0) + 1, max(nlev, 0))]; };  // its end = next level's pointer
    // pass 1 of one arc: posterior + tropical candidate; returns the log-semiring term
    auto bwd_state = [&](int s, const Range& r, const Arc2& q, ST am) {
      const int deg = r.b1 - r.b0;
      if (deg == 0) {  // sinks: beta = 1 (scorers.py:720), delta = 0
        if (LOGS) sB[s] = static_cast<ST>(0);
        if (TROP) { sD[s] = 0.0f; sBP[s] = -1; }
        return;
      }
      float bt = kNegInf;
      int bi = 0x7fffffff;
      auto visit = [&](int i, int d, float w) -> ST {
        ST u = neg_inf;
        if (LOGS) {
          u = static_cast<ST>(w) + sB[d];
          if (POST) {
            const float pzu = ex2_approx(static_cast<float>(am + u) * kLog2e);
          const float pz = pzu * gscale;
            if (post) post[a_lo + i] = pz;
            if (H.fix || H.glob) hist_add(H, s_lab[i], pzu, gscale);
          }
        }
        if (TROP) {
          const float t = __fadd_rn(w, sD[d]);
          if (t > bt || bi == 0x7fffffff) { bt = t; bi = i; }
        }
        return u;
      };
      const ST u0 = visit(r.b0, q.n0, q.w0);
      ST v = u0;
      if (deg > 1) {
        const ST u1 = visit(r.b0 + 1, q.n1, q.w1);
        ST m = max(max(u0, u1), static_cast<ST>(kFloor));
        for (int i = r.b0 + 2; i < r.b1; ++i) {
          const ST u = visit(i, s_dst[i], s_wout[i]);
          if (LOGS) m = max(m, u);
        }
        if (LOGS) {
          float sum = ex2_approx(static_cast<float>(u0 - m) * kLog2e) + ex2_approx(static_cast<float>(u1 - m) * kLog2e);
          for (int i = r.b0 + 2; i < r.b1; ++i)
            sum += ex2_approx(static_cast<float>(static_cast<ST>(s_wout[i]) + sB[s_dst[i]] - m) * kLog2e);
          v = sum > 0.0f ? m + static_cast<ST>(lg2_approx(sum) * kLn2) : neg_inf;
        }
      }
      if (LOGS) sB[s] = v;
      if (TROP) { sD[s] = bt; sBP[s] = a_lo + bi; }
    };
    int e0 = lv(nlev - 1), f0 = lv_end(nlev - 1), e1 = lv(nlev - 2), e2 = lv(nlev - 3), e3 = lv(nlev - 4);
    Range r0 = load_range(e0, f0, s_outp), r1 = load_range(e1, e0, s_outp);
    Arc2 q0 = load_arcs(r0, s_dst, s_wout);
    ST am0 = 0;
    if (POST && (e0 & kLvMask) + tid < (f0 & kLvMask)) am0 = sA[(e0 & kLvMask) + tid] - lz;
    for (int l = nlev - 1; l >= 0; --l) {
      const int e4 = lv(l - 4);
      const Range r2 = load_range(e2, e1, s_outp);
      const Arc2 q1 = load_arcs(r1, s_dst, s_wout);
      ST am1 = 0;
      if (POST && (e1 & kLvMask) + tid < (e0 & kLvMask)) am1 = sA[(e1 & kLvMask) + tid] - lz;
      const int sb = e0 & kLvMask, se = f0 & kLvMask, lg = (e0 >> 28) & 7;
      if (lg == 0) {
        int s = sb + tid;
        if (s < se) bwd_state(s, r0, q0, am0);
        for (s += NT; s < se; s += NT) {
          const Range r{s_outp[s], s_outp[s + 1]};
          ST am = 0;
          if (POST) am = sA[s] - lz;
          bwd_state(s, r, load_arcs(r, s_dst, s_wout), am);
        }
      } else {
        const int G = 1 << lg, lane_g = tid & (G - 1), ngrp = NT >> lg, ns = se - sb;
        for (int jb = 0; jb < ns; jb += ngrp) {
          const int j = jb + (tid >> lg);
          const bool valid = j < ns;
          const int s = sb + j;
          const int b0 = valid ? s_outp[s] : 0, b1 = valid ? s_outp[s + 1] : 0;
          ST am = 0;
          if (POST && valid) am = sA[s] - lz;
          ST m = static_cast<ST>(kFloor), u0 = neg_inf, u1 = neg_inf;
          float sum = 0.0f, bt = kNegInf;
          int bi = 0x7fffffff;
          auto visit = [&](int i) -> ST {
            const int d = s_dst[i];
            const float w = s_wout[i];
            ST u = neg_inf;
            if (LOGS) {
              u = static_cast<ST>(w) + sB[d];
              if (POST) {
                const float pzu = ex2_approx(static_cast<float>(am + u) * kLog2e);
          const float pz = pzu * gscale;
                if (post) post[a_lo + i] = pz;
                if (H.fix || H.glob) hist_add(H, s_lab[i], pzu, gscale);
              }
            }
            if (TROP) {
              const float t = __fadd_rn(w, sD[d]);
              if (t > bt || bi == 0x7fffffff) { bt = t; bi = i; }
            }
            return u;
          };
          const int i0 = b0 + lane_g, i1 = i0 + G;
          if (i0 < b1) u0 = visit(i0);
          if (i1 < b1) u1 = visit(i1);
          if (LOGS) m = max(max(u0, u1), m);
          for (int i = i1 + G; i < b1; i += G) {
            const ST u = visit(i);
            if (LOGS) m = max(m, u);
          }
          for (int o = G >> 1; o > 0; o >>= 1) {
            if (LOGS) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (TROP) {
              const float t2 = __shfl_xor_sync(0xffffffffu, bt, o);
              const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
              if (t2 > bt || (t2 == bt && i2 < bi)) { bt = t2; bi = i2; }
            }
          }
          if (LOGS) {
            sum = ex2_approx(static_cast<float>(u0 - m) * kLog2e) + ex2_approx(static_cast<float>(u1 - m) * kLog2e);
            for (int i = i1 + G; i < b1; i += G)
              sum += ex2_approx(static_cast<float>(static_cast<ST>(s_wout[i]) + sB[s_dst[i]] - m) * kLog2e);
            for (int o = G >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          }
          if (valid && lane_g == 0) {
            const bool sink = (b0 == b1);
            if (LOGS) sB[s] = sink ? static_cast<ST>(0) : (sum > 0.0f ? m + static_cast<ST>(lg2_approx(sum) * kLn2) : neg_inf);
            if (TROP) {
              sD[s] = sink ? 0.0f : bt;
              sBP[s] = sink ? -1 : a_lo + bi;
            }
          }
        }
      }
      level_barrier();
      f0 = e0; e0 = e1; e1 = e2; e2 = e3; e3 = e4;
      r0 = r1; r1 = r2; q0 = q1; am0 = am1;
    }
    NFST_T(t_bwd1);
    NFST_ACC(tid == 0, 18, t_bwd1 - t_bwd0);
    for (int i = tid; i < Sb; i += NT) {
      if (LOGS && beta_g) beta_g[s_lo + i] = sB[i];
      if (TROP) {
        if (delta_g) delta_g[s_lo + i] = sD[i];
        backptr[s_lo + i] = sBP[i];
      }
    }
    if (tid == 0) {
      if (LOGS && logz_bwd) logz_bwd[b] = sB[start];
      if (TROP && vit_score) vit_score[b] = sD[start];
      if (TROP && paths.path_arcs) {  // follow the backpointers while they are still in shared memory
        const int off = paths.path_off[b], room = paths.path_off[b + 1] - off;
        int s = start, k = 0;
        while (k < room) {
          const int a = sBP[s];
          if (a < 0) break;
          paths.path_arcs[off + k] = a;
          ++k;
          s = s_dst[a - a_lo];
        }
        paths.path_len[b] = k;
      }
    }
    if (want_hist && dtheta_smem) {
      __syncthreads();
      hist_flush(H, dtheta, L.vocab, gscale, tid, NT);
    }
  }
}

// =====================================================================================
// level-major kernels: one launch per topological level, one block per chunk of that level,
// over ALL lattices of the launch group.  For lattices whose levels are wide (thousands of arcs)
// this exposes every chunk of a level at once: no in-kernel barrier, no per-lattice serial
// chain; the neighbour DP values were written by earlier launches and are read through L2/L1.
// =====================================================================================
template <typename ST, bool SC, bool TH>
__global__ void __launch_bounds__(256, 4)
    nfst_fwd_level_kernel(const nfst_packed_lattices_t L, const nfst_chunk_t* __restrict__ chunks, int cap,
                          const float* __restrict__ arc_scores, const float* __restrict__ th,
                          ST* __restrict__ alpha) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const ST neg_inf = static_cast<ST>(kNegInf);
  const int4 k = __ldg(reinterpret_cast<const int4*>(chunks) + blockIdx.x);
  const int s0 = k.z, ns = k.w - k.z, n = k.y - k.x;
  const int32_t* __restrict__ in_ptr = L.in_ptr;
  const int32_t* __restrict__ src_in = L.src_in;
  const int32_t* __restrict__ in2out = L.in2out;
  const int32_t* __restrict__ label_in = L.label_in;
  auto arc_value = [&](int a) -> ST {
    float w = SC ? __ldg(arc_scores + __ldg(in2out + a)) : 0.0f;
    if (TH) w += __ldg(th + __ldg(label_in + a));
    return alpha[__ldg(src_in + a)] + static_cast<ST>(w);  // alpha of earlier levels: earlier launches
  };
  if (n <= cap) {
    if (ns * 2 > NT) {
#pragma unroll 1
      for (int j = tid; j < ns; j += NT) {
        int a = __ldg(in_ptr + s0 + j);
        const int b1 = __ldg(in_ptr + s0 + j + 1);
        ST m = static_cast<ST>(kFloor);
        float sum = 0.0f;
#pragma unroll 1
        for (; a + 3 < b1; a += 4) {
          const ST v0 = arc_value(a), v1 = arc_value(a + 1), v2 = arc_value(a + 2), v3 = arc_value(a + 3);
          lse_push4(m, sum, v0, v1, v2, v3);
        }
#pragma unroll 1
        for (; a < b1; ++a) lse_push(m, sum, arc_value(a), neg_inf);
        alpha[s0 + j] = lse_finish(m, sum, neg_inf);
      }
    } else {
      int lg = 1;
      while (lg < 5 && (ns << (lg + 1)) <= NT) ++lg;
      const int G = 1 << lg, lane_g = tid & (G - 1), ngrp = NT >> lg;
#pragma unroll 1
      for (int jb = 0; jb < ns; jb += ngrp) {
        const int j = jb + (tid >> lg);
        const bool valid = j < ns;
        const int b0 = valid ? __ldg(in_ptr + s0 + j) : 0, b1 = valid ? __ldg(in_ptr + s0 + j + 1) : 0;
        ST m = static_cast<ST>(kFloor);
        float sum = 0.0f;
#pragma unroll 1
        for (int a = b0 + lane_g; a < b1; a += G) lse_push(m, sum, arc_value(a), neg_inf);
        for (int o = G >> 1; o > 0; o >>= 1) {
          const ST m2 = __shfl_xor_sync(0xffffffffu, m, o);
          const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
          lse_join(m, sum, m2, s2, neg_inf);
        }
        if (valid && lane_g == 0) alpha[s0 + j] = lse_finish(m, sum, neg_inf);
      }
    }
  } else {
    for (int j = 0; j < ns; ++j) {  // heavy chunk: block-wide
      const int s = s0 + j;
      const int b0 = in_ptr[s], b1 = in_ptr[s + 1];
      ST m = neg_inf;
      float sum = 0.0f;
      for (int a = b0 + tid; a < b1; a += NT) lse_add(m, sum, arc_value(a));
      const ST v0 = block_lse(m, sum);
      if (tid == 0) alpha[s] = v0;
    }
  }
}

// alpha[start] = 0 before the level launches; logZ = logsumexp_{sinks} alpha after them
template <typename ST>
__global__ void nfst_fwd_init_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int n,
                                     ST* __restrict__ alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) alpha[L.start_state[ids ? ids[i] : i]] = static_cast<ST>(0);
}
template <typename ST>
__global__ void __launch_bounds__(128)
    nfst_logz_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, const ST* __restrict__ alpha,
                     ST* __restrict__ logz) {
  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  ST m = static_cast<ST>(kNegInf);
  float sum = 0.0f;
  for (int i = L.sink_off[b] + threadIdx.x; i < L.sink_off[b + 1]; i += blockDim.x) lse_add(m, sum, alpha[L.sinks[i]]);
  const ST z = block_lse(m, sum);
  if (threadIdx.x == 0) logz[b] = z;
}

template <typename ST, bool LOGS, bool TROP, bool SC, bool TH, bool POST>
__global__ void __launch_bounds__(256, 4)
    nfst_bwd_level_kernel(const nfst_packed_lattices_t L, const nfst_chunk_t* __restrict__ chunks,
                          const int32_t* __restrict__ chunk_lat, int cap, const float* __restrict__ arc_scores,
                          const float* __restrict__ th, int dtheta_smem, const ST* __restrict__ alpha,
                          const ST* __restrict__ logz, const float* __restrict__ grad_logz, ST* __restrict__ beta,
                          float* __restrict__ post, float* __restrict__ dtheta, float* __restrict__ delta,
                          int32_t* __restrict__ backptr) {
  const int NT = blockDim.x, tid = threadIdx.x;
  const ST neg_inf = static_cast<ST>(kNegInf);
  const bool want_hist = POST && dtheta != nullptr;
  const bool need_label = TH || want_hist;
  const int4 k = __ldg(reinterpret_cast<const int4*>(chunks) + blockIdx.x);
  const int s0 = k.z, ns = k.w - k.z, n = k.y - k.x;
  const int32_t* __restrict__ out_ptr = L.out_ptr;
  const int32_t* __restrict__ dst_out = L.dst_out;
  const int32_t* __restrict__ label_out = L.label_out;
  unsigned* hist_fix = nullptr;
  if (want_hist && dtheta_smem) {
    hist_fix = reinterpret_cast<unsigned*>(smem_f);
    for (int i = tid; i < L.vocab; i += NT) hist_fix[i] = 0u;
    __syncthreads();
  }
  ST lz = 0;
  float gscale = 1.0f;
  int hist_levels = 1;
  if (POST) {
    const int b = __ldg(chunk_lat + blockIdx.x);
    lz = logz[b];
    if (grad_logz) gscale = grad_logz[b];
    hist_levels = L.level_off[b + 1] - L.level_off[b] - 1;
  }
  const Hist H = hist_make(hist_fix, want_hist && !dtheta_smem ? dtheta : nullptr, hist_levels);
  auto visit = [&](int a, ST am, ST& m, float& sum, float& bt, int& bi) {
    const int d = __ldg(dst_out + a);
    float w = SC ? __ldg(arc_scores + a) : 0.0f;
    int lab = 0;
    if (TH || POST) { if (need_label) lab = __ldg(label_out + a); }
    if (TH) w += __ldg(th + lab);
    if (LOGS) {
      const ST u = static_cast<ST>(w) + beta[d];
      lse_push(m, sum, u, neg_inf);
      if (POST) {
        const float pu = ex2_approx(static_cast<float>(am + u) * kLog2e);
          const float p = pu * gscale;
        if (post) post[a] = p;
        if (H.fix || H.glob) hist_add(H, lab, pu, gscale);
      }
    }
    if (TROP) {
      const float t = __fadd_rn(w, delta[d]);
      if (t > bt || bi == 0x7fffffff) { bt = t; bi = a; }
    }
  };
  auto visit4 = [&](int a, ST am, ST& m, float& sum, float& bt, int& bi) {
    ST u[4];
    int lab[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int d = __ldg(dst_out + a + q);
      float w = SC ? __ldg(arc_scores + a + q) : 0.0f;
      if (TH || POST) { if (need_label) lab[q] = __ldg(label_out + a + q); }
      if (TH) w += __ldg(th + lab[q]);
      if (LOGS) u[q] = static_cast<ST>(w) + beta[d];
      if (TROP) {
        const float t = __fadd_rn(w, delta[d]);
        if (t > bt || bi == 0x7fffffff) { bt = t; bi = a + q; }
      }
    }
    if (LOGS) {
      lse_push4(m, sum, u[0], u[1], u[2], u[3]);
      if (POST) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float pu = ex2_approx(static_cast<float>(am + u[q]) * kLog2e);
          const float p = pu * gscale;
          if (post) post[a + q] = p;
          if (H.fix || H.glob) hist_add(H, lab[q], pu, gscale);
        }
      }
    }
  };
  auto finish = [&](int s, bool sink, ST m, float sum, float bt, int bi) {
    if (LOGS) beta[s] = sink ? static_cast<ST>(0) : lse_finish(m, sum, neg_inf);
    if (TROP) {
      delta[s] = sink ? 0.0f : bt;
      backptr[s] = sink ? -1 : bi;
    }
  };
  if (n <= cap) {
    if (ns * 2 > NT) {
#pragma unroll 1
      for (int j = tid; j < ns; j += NT) {
        const int s = s0 + j;
        const int b0 = __ldg(out_ptr + s), b1 = __ldg(out_ptr + s + 1);
        ST am = 0;
        if (POST) am = alpha[s] - lz;
        ST m = static_cast<ST>(kFloor);
        float sum = 0.0f, bt = kNegInf;
        int bi = 0x7fffffff;
        int a = b0;
#pragma unroll 1
        for (; a + 3 < b1; a += 4) visit4(a, am, m, sum, bt, bi);
#pragma unroll 1
        for (; a < b1; ++a) visit(a, am, m, sum, bt, bi);
        finish(s, b0 == b1, m, sum, bt, bi);
      }
    } else {
      int lg = 1;
      while (lg < 5 && (ns << (lg + 1)) <= NT) ++lg;
      const int G = 1 << lg, lane_g = tid & (G - 1), ngrp = NT >> lg;
#pragma unroll 1
      for (int jb = 0; jb < ns; jb += ngrp) {
        const int j = jb + (tid >> lg);
        const bool valid = j < ns;
        const int b0 = valid ? __ldg(out_ptr + s0 + j) : 0, b1 = valid ? __ldg(out_ptr + s0 + j + 1) : 0;
        const int s = s0 + j;
        ST am = 0;
        if (POST && valid) am = alpha[s] - lz;
        ST m = static_cast<ST>(kFloor);
        float sum = 0.0f, bt = kNegInf;
        int bi = 0x7fffffff;
#pragma unroll 1
        for (int a = b0 + lane_g; a < b1; a += G) visit(a, am, m, sum, bt, bi);
        for (int o = G >> 1; o > 0; o >>= 1) {
          if (LOGS) {
            const ST m2 = __shfl_xor_sync(0xffffffffu, m, o);
            const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
            lse_join(m, sum, m2, s2, neg_inf);
          }
          if (TROP) {
            const float t2 = __shfl_xor_sync(0xffffffffu, bt, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
            if (t2 > bt || (t2 == bt && i2 < bi)) { bt = t2; bi = i2; }
          }
        }
        if (valid && lane_g == 0) finish(s, b0 == b1, m, sum, bt, bi);
      }
    }
  } else {
    for (int j = 0; j < ns; ++j) {  // heavy chunk: block-wide
      const int s = s0 + j;
      const int b0 = out_ptr[s], b1 = out_ptr[s + 1];
      ST am = 0;
      if (POST) am = alpha[s] - lz;
      ST m = neg_inf;
      float sum = 0.0f, bt = kNegInf;
      int bi = 0x7fffffff;
      for (int a = b0 + tid; a < b1; a += NT) {
        const int d = dst_out[a];
        int lab = 0;
        if (need_label) lab = label_out[a];
        float w = SC ? arc_scores[a] : 0.0f;
        if (TH) w += th[lab];
        if (LOGS) {
          const ST u = static_cast<ST>(w) + beta[d];
          lse_add(m, sum, u);
          if (POST) {
            const float pu = ex2_approx(static_cast<float>(am + u) * kLog2e);
          const float p = pu * gscale;
            if (post) post[a] = p;
            if (H.fix || H.glob) hist_add(H, lab, pu, gscale);
          }
        }
        if (TROP) {
          const float t = __fadd_rn(w, delta[d]);
          if (t > bt || (t == bt && a < bi)) { bt = t; bi = a; }
        }
      }
      if (LOGS) {
        const ST v0 = block_lse(m, sum);
        if (tid == 0) beta[s] = (b0 == b1) ? static_cast<ST>(0) : v0;
      }
      if (TROP) {
        __shared__ float red_t[32];
        __shared__ int red_i[32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float t2 = __shfl_xor_sync(0xffffffffu, bt, o);
          const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
          if (t2 > bt || (t2 == bt && i2 < bi)) { bt = t2; bi = i2; }
        }
        __syncthreads();
        if ((tid & 31) == 0) { red_t[tid >> 5] = bt; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
          for (int w2 = 1; w2 < (NT + 31) / 32; ++w2)
            if (red_t[w2] > bt || (red_t[w2] == bt && red_i[w2] < bi)) { bt = red_t[w2]; bi = red_i[w2]; }
          const bool sink = (b0 == b1);
          delta[s] = sink ? 0.0f : bt;
          backptr[s] = sink ? -1 : bi;
        }
      }
    }
  }
  if (want_hist && dtheta_smem) {
    __syncthreads();
    hist_flush(H, dtheta, L.vocab, gscale, tid, NT);
  }
}

// logz_bwd[b] = beta[start], vit_score[b] = delta[start]
template <typename ST>
__global__ void nfst_pick_start_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int n,
                                       const ST* __restrict__ beta, ST* __restrict__ logz_bwd,
                                       const float* __restrict__ delta, float* __restrict__ vit_score) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = ids ? ids[i] : i;
  const int s = L.start_state[b];
  if (beta && logz_bwd) logz_bwd[b] = beta[s];
  if (delta && vit_score) vit_score[b] = delta[s];
}

// =====================================================================================
// small kernels
// =====================================================================================
__global__ void nfst_backtrace_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int n,
                                      const int32_t* __restrict__ backptr, const int32_t* __restrict__ path_off,
                                      int32_t* __restrict__ path_arcs, int32_t* __restrict__ path_len) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int b = ids ? ids[t] : t;
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

// ragged result: path slots (capacity offsets) -> dense runs (out_off), plus the labels; one warp per lattice
__global__ void nfst_compact_paths_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ path_off,
                                          const int32_t* __restrict__ path_len, const int32_t* __restrict__ path_buf,
                                          const int64_t* __restrict__ out_off, int32_t* __restrict__ out_arcs,
                                          int32_t* __restrict__ out_labels) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= L.n_lattices) return;
  const int lane = threadIdx.x & 31;
  const int src0 = path_off[b], len = path_len[b];
  const int64_t dst0 = out_off[b];
  for (int k = lane; k < len; k += 32) {
    const int a = path_buf[src0 + k];
    out_arcs[dst0 + k] = a;
    if (out_labels) out_labels[dst0 + k] = L.label_out[a];
  }
}

// padded result: row b of out_labels[B, T] = the labels of lattice b's path (without a leading skip_label: the
// reference's samples never hold bos, scorers.py:230-231), then pad_label; out_len[b] = labels written; one warp
// per lattice, no host read anywhere (the best-sample read-out of lightning.py:474-479 as a [B, T] tensor)
__global__ void nfst_pad_paths_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ path_off,
                                      const int32_t* __restrict__ path_len, const int32_t* __restrict__ path_buf,
                                      int skip_label, int64_t pad_label, int T, int64_t* __restrict__ out_labels,
                                      int32_t* __restrict__ out_len) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= L.n_lattices) return;
  const int lane = threadIdx.x & 31;
  const int src0 = path_off[b];
  int len = path_len[b];
  int skip = 0;
  if (len > 0 && skip_label >= 0 && L.label_out[path_buf[src0]] == skip_label) skip = 1;
  len = min(len - skip, T);
  int64_t* const row = out_labels + static_cast<int64_t>(b) * T;
  for (int k = lane; k < T; k += 32) row[k] = k < len ? static_cast<int64_t>(L.label_out[path_buf[src0 + skip + k]]) : pad_label;
  if (lane == 0 && out_len) out_len[b] = len;
}

// =====================================================================================
// beta-hat recurrence (the reference's full compute_beta, Wh != 0; scorers.py:692-751):
//   message over arc c --j--> n :  m_hat = tanh(Wx e_j + Wh beta_hat[n] + b)     (:732-735)
//                                  m     = exp(W . m_hat) * beta[n]              (:736-738)
//   aggregation at c            :  beta[c] = sum m ;  beta_hat[c] = sum (m / beta[c]) m_hat   (:741-747)
// The arc weight depends on the destination's beta_hat, so the pass is stepped level by level
// (one launch per topological level, deepest first); within a launch one thread block per
// state, one thread per hidden unit.  Wx e_j + b is a per-LABEL table and Wh beta_hat[n] a
// per-STATE vector (computed once, when n is finished), so an arc costs H tanh + one dot
// product.  beta is kept in log space (online logsumexp over the arcs).
// =====================================================================================
__global__ void nfst_beta_hat_level_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ states, int H,
                                           const float* __restrict__ label_proj,  // [V, H] = Wx e_l + b
                                           const float* __restrict__ wh_t,        // [H, H], wh_t[k*H + i] = Wh[i][k]
                                           const float* __restrict__ w,           // [H]
                                           float* log_beta, float* beta_hat, float* h_proj) {
  extern __shared__ float sh[];  // [blockDim.x] beta_hat of this state (for the projection)
  __shared__ float red[32];
  const int s = states[blockIdx.x];
  const int i = threadIdx.x;
  const bool act = i < H;
  const int nw = (blockDim.x + 31) >> 5;
  const int b0 = L.out_ptr[s], b1 = L.out_ptr[s + 1];
  const size_t row = static_cast<size_t>(s) * H;
  if (b0 == b1) {  // sink: beta = 1 (scorers.py:720), beta_hat = 0
    if (i == 0) log_beta[s] = 0.0f;
    if (act) { beta_hat[row + i] = 0.0f; h_proj[row + i] = 0.0f; }
    return;
  }
  const float wi = act ? w[i] : 0.0f;
  float mx = kNegInf, sum = 0.0f, acc = 0.0f;
  for (int ai = b0; ai < b1; ++ai) {
    const int a = L.out_arc ? L.out_arc[ai] : ai;  // column-major lattices list a state's arcs through out_arc
    const int n = L.dst_out[a], lab = L.label_out[a];
    const float mh = act ? tanhf(label_proj[static_cast<size_t>(lab) * H + i] + h_proj[static_cast<size_t>(n) * H + i]) : 0.0f;
    float part = wi * mh;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncthreads();  // red[] free
    if ((i & 31) == 0) red[i >> 5] = part;
    __syncthreads();
    float score = 0.0f;
    for (int q = 0; q < nw; ++q) score += red[q];  // same order in every thread
    const float lm = score + log_beta[n];
    if (lm > mx) {
      const float r = expf(mx - lm);  // mx = -inf: 0
      sum = sum * r + 1.0f;
      acc = acc * r + mh;
      mx = lm;
    } else if (lm > kNegInf) {
      const float e = expf(lm - mx);
      sum += e;
      acc += e * mh;
    }
  }
  const float bh = (sum > 0.0f) ? acc / sum : 0.0f;
  if (i == 0) log_beta[s] = (sum > 0.0f) ? mx + logf(sum) : kNegInf;
  if (act) beta_hat[row + i] = bh;
  sh[i] = act ? bh : 0.0f;
  __syncthreads();
  if (act) {  // h = Wh beta_hat: column i of wh_t, coalesced across the block
    float h = 0.0f;
    for (int k = 0; k < H; ++k) h = fmaf(wh_t[static_cast<size_t>(k) * H + i], sh[k], h);
    h_proj[row + i] = h;
  }
}

template <typename ST>
__global__ void nfst_beta_to_dense_kernel(const nfst_packed_lattices_t L, const ST* __restrict__ beta,
                                          const int32_t* __restrict__ orig_state, int k, int dense_states,
                                          float* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L.n_states) return;
  int lo = 0, hi = L.n_lattices;  // lattice of state s: last b with state_off[b] <= s
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (L.state_off[mid] <= s) lo = mid; else hi = mid;
  }
  const float v = static_cast<float>(exp(static_cast<double>(beta[s])));
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
                                          const int64_t* __restrict__ row_start, int64_t capacity, int32_t* __restrict__ arc_row,
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
    const int64_t at = pos + __popc(mask & ((1u << lane) - 1u));
    if (keep && at < capacity) {  // arcs beyond the caller's capacity are dropped (the caller sees the total in row_start)
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
  if (launch->sell)
    return fail(NFST_ERR_BAD_ARG, "sliced-column launch group: its arcs are not CSR, use nfst_sell_pull_f32 / nfst_sell_flow_f32");
  if (launch->tiles)
    return fail(NFST_ERR_BAD_ARG, "tile-stream launch group: its arcs are not CSR, use nfst_tile_pull_f32 / nfst_tile_flow_f32");
  if (launch->n_ids < 0 || launch->n_ids > lat->n_lattices)
    return fail(NFST_ERR_BAD_ARG, "n_ids=%d out of range (B=%d)", launch->n_ids, lat->n_lattices);
  const int bt = launch->block_threads;
  if (bt != 32 && bt != 64 && bt != 128 && bt != 256)
    return fail(NFST_ERR_BAD_ARG, "block_threads=%d must be 32, 64, 128 or 256", bt);
  const int w = launch->window_states;
  if (w < 32 || (w & (w - 1))) return fail(NFST_ERR_BAD_ARG, "window_states=%d must be a power of two >= 32", w);
  if (launch->chunk_cap < 32 || (launch->chunk_cap & 7))
    return fail(NFST_ERR_BAD_ARG, "chunk_cap=%d must be a multiple of 8, >= 32", launch->chunk_cap);
  if (launch->small_max_arcs > 65535 || launch->small_max_states > 65535 || launch->small_max_levels >= (1 << 24))
    return fail(NFST_ERR_TOO_LARGE, "small-lattice group: at most 65535 states and arcs per lattice (got %d, %d)",
                launch->small_max_states, launch->small_max_arcs);
  return NFST_OK;
}

template <typename K>
int prepare_smem(K kernel, size_t bytes) {
  if (bytes > 40 * 1024) {  // static shared memory counts against the default 48 KB too
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

template <typename ST, bool SC, bool TH>
int launch_fwd2(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                void* alpha, void* logz, cudaStream_t st) {
  const int theta_smem = TH && lat->vocab <= NFST_THETA_SMEM_MAX;
  const size_t bytes = nfst_launch_smem_bytes(launch, lat->vocab, 0, 1, 0, 0, SC, TH, 0);
  if (int rc = prepare_smem(nfst_fwd_kernel<ST, SC, TH>, bytes)) return rc;
  nfst_fwd_kernel<ST, SC, TH><<<launch->n_ids, launch->block_threads, bytes, st>>>(
      *lat, launch->lattice_ids, launch->window_states, launch->chunk_cap, scores->arc_scores, scores->theta,
      theta_smem, static_cast<ST*>(alpha), static_cast<ST*>(logz));
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

template <typename ST, bool SC, bool TH, bool DO_FWD, bool DO_BWD, bool LOGS, bool TROP, bool POST>
int launch_small(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                 void* alpha, void* logz, const float* grad_logz, void* beta, void* logz_bwd, float* post, float* dtheta,
                 float* delta, int32_t* backptr, float* vit_score, cudaStream_t st, PathOut paths = PathOut{nullptr, nullptr, nullptr}) {
  const bool want_hist = POST && dtheta != nullptr;
  const int dtheta_smem = want_hist && lat->vocab <= NFST_THETA_SMEM_MAX;
  const SmallPlan P = small_plan(launch->small_max_states, launch->small_max_arcs, launch->small_max_levels,
                                 sizeof(ST) / 4, DO_FWD, DO_BWD, DO_BWD && TROP, want_hist, dtheta_smem ? lat->vocab : 0);
  const size_t bytes = static_cast<size_t>(P.words) * 4;
  auto kernel = nfst_small_kernel<ST, SC, TH, DO_FWD, DO_BWD, LOGS, TROP, POST>;
  if (int rc = prepare_smem(kernel, bytes)) return rc;
  kernel<<<launch->n_ids, launch->block_threads, bytes, st>>>(
      *lat, launch->lattice_ids, launch->small_max_states, launch->small_max_arcs, launch->small_max_levels,
      scores->arc_scores, scores->theta, dtheta_smem, static_cast<ST*>(alpha), static_cast<ST*>(logz), grad_logz,
      static_cast<ST*>(beta), static_cast<ST*>(logz_bwd), post, dtheta, delta, backptr, vit_score, paths);
  if (cudaError_t e = cudaGetLastError())
    return fail(NFST_ERR_CUDA, "small-lattice launch (grid %d, block %d, %zu B shared): %s", launch->n_ids,
                launch->block_threads, bytes, cudaGetErrorString(e));
  return NFST_OK;
}

// score-mode dispatch for the small path
template <typename ST, bool DO_FWD, bool DO_BWD, bool LOGS, bool TROP, bool POST>
int launch_small_sc(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                    void* alpha, void* logz, const float* grad_logz, void* beta, void* logz_bwd, float* post,
                    float* dtheta, float* delta, int32_t* backptr, float* vit_score, cudaStream_t st,
                    PathOut paths = PathOut{nullptr, nullptr, nullptr}) {
  const bool sc = scores->arc_scores != nullptr, th = scores->theta != nullptr;
#define NFST_SM(SC, TH)                                                                                          \
  return launch_small<ST, SC, TH, DO_FWD, DO_BWD, LOGS, TROP, POST>(lat, launch, scores, alpha, logz, grad_logz, beta, \
                                                                    logz_bwd, post, dtheta, delta, backptr, vit_score, st, paths)
  if (sc && th) NFST_SM(true, true);
  if (sc) NFST_SM(true, false);
  NFST_SM(false, true);
#undef NFST_SM
}

template <typename ST, bool SC, bool TH>
int launch_fwd_levels(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                      void* alpha, void* logz, cudaStream_t st) {
  ST* a = static_cast<ST*>(alpha);
  const int n = launch->n_ids;
  nfst_fwd_init_kernel<ST><<<(n + 127) / 128, 128, 0, st>>>(*lat, launch->lattice_ids, n, a);
  for (int l = 1; l < launch->n_levels; ++l) {  // level 0 holds the start states only
    const int c0 = launch->fwd_level_off[l], c1 = launch->fwd_level_off[l + 1];
    if (c1 > c0)
      nfst_fwd_level_kernel<ST, SC, TH><<<c1 - c0, launch->block_threads, 0, st>>>(
          *lat, launch->fwd_level_chunks + c0, launch->chunk_cap, scores->arc_scores, scores->theta, a);
  }
  nfst_logz_kernel<ST><<<n, 128, 0, st>>>(*lat, launch->lattice_ids, a, static_cast<ST*>(logz));
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

template <typename ST>
int launch_fwd(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores, void* alpha,
               void* logz, cudaStream_t st) {
  const bool sc = scores->arc_scores != nullptr, th = scores->theta != nullptr;
  if (launch->small_max_arcs > 0)
    return launch_small_sc<ST, true, false, true, false, false>(lat, launch, scores, alpha, logz, nullptr, nullptr,
                                                                nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st);
  if (launch->fwd_level_chunks) {
    if (sc && th) return launch_fwd_levels<ST, true, true>(lat, launch, scores, alpha, logz, st);
    if (sc) return launch_fwd_levels<ST, true, false>(lat, launch, scores, alpha, logz, st);
    return launch_fwd_levels<ST, false, true>(lat, launch, scores, alpha, logz, st);
  }
  if (sc && th) return launch_fwd2<ST, true, true>(lat, launch, scores, alpha, logz, st);
  if (sc) return launch_fwd2<ST, true, false>(lat, launch, scores, alpha, logz, st);
  return launch_fwd2<ST, false, true>(lat, launch, scores, alpha, logz, st);
}

template <typename ST, bool LOGS, bool TROP, bool SC, bool TH, bool POST>
int launch_bwd3(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                const void* alpha, const void* logz, const float* grad_logz, void* beta, void* logz_bwd, float* post,
                float* dtheta, float* delta, int32_t* backptr, float* vit_score, cudaStream_t st) {
  const bool small_v = lat->vocab <= NFST_THETA_SMEM_MAX;
  const int theta_smem = TH && small_v;
  const int dtheta_smem = POST && dtheta && small_v;
  const size_t bytes = nfst_launch_smem_bytes(launch, lat->vocab, 1, LOGS, TROP, POST, SC, TH, POST && dtheta != nullptr);
  if (int rc = prepare_smem(nfst_bwd_kernel<ST, LOGS, TROP, SC, TH, POST>, bytes)) return rc;
  nfst_bwd_kernel<ST, LOGS, TROP, SC, TH, POST><<<launch->n_ids, launch->block_threads, bytes, st>>>(
      *lat, launch->lattice_ids, launch->window_states, launch->chunk_cap, scores->arc_scores, scores->theta,
      theta_smem, dtheta_smem,
      static_cast<const ST*>(alpha), static_cast<const ST*>(logz), grad_logz, static_cast<ST*>(beta),
      static_cast<ST*>(logz_bwd), post, dtheta, delta, backptr, vit_score);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

template <typename ST, bool LOGS, bool TROP, bool SC, bool TH, bool POST>
int launch_bwd_levels(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                      const void* alpha, const void* logz, const float* grad_logz, void* beta, void* logz_bwd,
                      float* post, float* dtheta, float* delta, int32_t* backptr, float* vit_score, cudaStream_t st) {
  const int dtheta_smem = POST && dtheta && lat->vocab <= NFST_THETA_SMEM_MAX;
  const size_t bytes = dtheta_smem ? static_cast<size_t>(lat->vocab) * 4 : 0;
  for (int l = launch->n_levels - 1; l >= 0; --l) {
    const int c0 = launch->bwd_level_off[l], c1 = launch->bwd_level_off[l + 1];
    if (c1 > c0)
      nfst_bwd_level_kernel<ST, LOGS, TROP, SC, TH, POST><<<c1 - c0, launch->block_threads, bytes, st>>>(
          *lat, launch->bwd_level_chunks + c0, launch->bwd_level_lat + c0, launch->chunk_cap, scores->arc_scores,
          scores->theta, dtheta_smem, static_cast<const ST*>(alpha), static_cast<const ST*>(logz), grad_logz,
          static_cast<ST*>(beta), post, dtheta, delta, backptr);
  }
  const int n = launch->n_ids;
  nfst_pick_start_kernel<ST><<<(n + 127) / 128, 128, 0, st>>>(*lat, launch->lattice_ids, n,
                                                               LOGS ? static_cast<const ST*>(beta) : nullptr,
                                                               static_cast<ST*>(logz_bwd), TROP ? delta : nullptr,
                                                               vit_score);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

template <typename ST, bool LOGS, bool TROP>
int launch_bwd(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
               const void* alpha, const void* logz, const float* grad_logz, void* beta, void* logz_bwd, float* post,
               float* dtheta, float* delta, int32_t* backptr, float* vit_score, cudaStream_t st) {
  const bool sc = scores->arc_scores != nullptr, th = scores->theta != nullptr;
  const bool ps = LOGS && (post || dtheta);
  if (launch->small_max_arcs > 0) {
    void* a = const_cast<void*>(alpha);
    void* z = const_cast<void*>(logz);
    if (ps) return launch_small_sc<ST, false, true, LOGS, TROP, LOGS>(lat, launch, scores, a, z, grad_logz, beta, logz_bwd,
                                                                      post, dtheta, delta, backptr, vit_score, st);
    return launch_small_sc<ST, false, true, LOGS, TROP, false>(lat, launch, scores, a, z, grad_logz, beta, logz_bwd, post,
                                                               dtheta, delta, backptr, vit_score, st);
  }
  if (launch->bwd_level_chunks) {
#define NFST_GOL(SC, TH, PS)                                                                                     \
  return launch_bwd_levels<ST, LOGS, TROP, SC, TH, (LOGS && PS)>(lat, launch, scores, alpha, logz, grad_logz, beta, \
                                                                 logz_bwd, post, dtheta, delta, backptr, vit_score, st)
    if (sc && th) { if (ps) NFST_GOL(true, true, true); NFST_GOL(true, true, false); }
    if (sc) { if (ps) NFST_GOL(true, false, true); NFST_GOL(true, false, false); }
    if (ps) NFST_GOL(false, true, true);
    NFST_GOL(false, true, false);
#undef NFST_GOL
  }
#define NFST_GO(SC, TH, PS)                                                                                      \
  return launch_bwd3<ST, LOGS, TROP, SC, TH, (LOGS && PS)>(lat, launch, scores, alpha, logz, grad_logz, beta, logz_bwd, \
                                                           post, dtheta, delta, backptr, vit_score, st)
  if (sc && th) { if (ps) NFST_GO(true, true, true); NFST_GO(true, true, false); }
  if (sc) { if (ps) NFST_GO(true, false, true); NFST_GO(true, false, false); }
  if (ps) NFST_GO(false, true, true);
  NFST_GO(false, true, false);
#undef NFST_GO
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int nfst_abi_version(void) { return NFST_ABI_VERSION; }

#ifdef NFST_TIMING
int nfst_debug_read(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 32);
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_dbg, z, sizeof(z));
  }
  return 0;
}
#endif

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

size_t nfst_launch_smem_bytes(const nfst_launch_t* launch, int32_t vocab, int pass, int with_log, int with_trop,
                              int with_post, int with_scores, int with_theta, int with_dtheta) {
  if (!launch) return 0;
  const bool small_v = vocab <= NFST_THETA_SMEM_MAX;
  (void)with_post;
  const bool bwd = pass != 0;
  const SmemPlan p = smem_plan(launch->window_states, launch->chunk_cap,
                               launch->state_f64 ? 8 : 4, vocab, !bwd || with_log != 0, bwd && with_trop != 0,
                               with_scores != 0, with_theta != 0 || (bwd && with_dtheta != 0), with_theta && small_v,
                               bwd && with_dtheta && small_v, bwd ? 2 : 3, !bwd);
  return p.bytes;
}

int nfst_fwd_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                 void* alpha, void* logz, void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!scores || !alpha || !logz) return fail(NFST_ERR_BAD_ARG, "nfst_fwd_f32: null scores/alpha/logz");
  if (!scores->arc_scores && !scores->theta) return fail(NFST_ERR_BAD_ARG, "need arc_scores and/or theta");
  if (launch->n_ids == 0) return NFST_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  return launch->state_f64 ? launch_fwd<double>(lat, launch, scores, alpha, logz, st)
                           : launch_fwd<float>(lat, launch, scores, alpha, logz, st);
}

int nfst_bwd_fused_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                       const void* alpha, const void* logz, const float* grad_logz, void* beta, void* logz_bwd,
                       float* post, float* dtheta, float* delta, int32_t* backptr, float* vit_score,
                       void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!scores) return fail(NFST_ERR_BAD_ARG, "nfst_bwd_fused_f32: null scores");
  if (!scores->arc_scores && !scores->theta) return fail(NFST_ERR_BAD_ARG, "need arc_scores and/or theta");
  const bool want_post = post || dtheta;
  const bool logs = beta || logz_bwd || want_post;
  const bool trop = delta || backptr || vit_score;
  if (!logs && !trop) return fail(NFST_ERR_BAD_ARG, "nfst_bwd_fused_f32: no output requested");
  const bool small = launch->small_max_arcs > 0;  // small lattices keep their working vectors in shared memory
  if (logs && !beta && !small)
    return fail(NFST_ERR_BAD_ARG, "the log-semiring pass needs a beta[S] buffer (its working vector)");
  if (trop && (!backptr || (!delta && !small)))
    return fail(NFST_ERR_BAD_ARG, "the tropical pass needs delta[S] and backptr[S]");
  if (want_post && (!alpha || !logz))
    return fail(NFST_ERR_BAD_ARG, "posteriors / dtheta need alpha[S] and logz[B] from nfst_fwd_f32");
  if (launch->n_ids == 0) return NFST_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
#define NFST_BWD(ST, LG, TR) \
  launch_bwd<ST, LG, TR>(lat, launch, scores, alpha, logz, grad_logz, beta, logz_bwd, post, dtheta, delta, backptr, vit_score, st)
  if (launch->state_f64) {
    if (logs && trop) return NFST_BWD(double, true, true);
    if (logs) return NFST_BWD(double, true, false);
    return NFST_BWD(float, false, true);
  }
  if (logs && trop) return NFST_BWD(float, true, true);
  if (logs) return NFST_BWD(float, true, false);
  return NFST_BWD(float, false, true);
#undef NFST_BWD
}

int nfst_fwd_bwd_small_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                           const float* grad_logz, void* alpha, void* logz, void* beta, void* logz_bwd, float* post,
                           float* dtheta, void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!scores || (!scores->arc_scores && !scores->theta)) return fail(NFST_ERR_BAD_ARG, "need arc_scores and/or theta");
  if (launch->small_max_arcs <= 0) return fail(NFST_ERR_BAD_ARG, "nfst_fwd_bwd_small_f32 needs a small-lattice launch group");
  if (!logz) return fail(NFST_ERR_BAD_ARG, "nfst_fwd_bwd_small_f32: logz is required");
  if (launch->n_ids == 0) return NFST_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const bool ps = post || dtheta;
  if (launch->state_f64) {
    if (ps) return launch_small_sc<double, true, true, true, false, true>(lat, launch, scores, alpha, logz, grad_logz, beta, logz_bwd, post, dtheta, nullptr, nullptr, nullptr, st);
    return launch_small_sc<double, true, true, true, false, false>(lat, launch, scores, alpha, logz, grad_logz, beta, logz_bwd, post, dtheta, nullptr, nullptr, nullptr, st);
  }
  if (ps) return launch_small_sc<float, true, true, true, false, true>(lat, launch, scores, alpha, logz, grad_logz, beta, logz_bwd, post, dtheta, nullptr, nullptr, nullptr, st);
  return launch_small_sc<float, true, true, true, false, false>(lat, launch, scores, alpha, logz, grad_logz, beta, logz_bwd, post, dtheta, nullptr, nullptr, nullptr, st);
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
  nfst_backtrace_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      *lat, nullptr, lat->n_lattices, backptr, path_off, path_arcs, path_len);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_viterbi_paths_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                           float* delta, int32_t* backptr, float* vit_score, const int32_t* path_off,
                           int32_t* path_arcs, int32_t* path_len, void* cuda_stream) {
  if (!launch || !(launch->sell || launch->tiles)) {
    if (int rc = check_launch(lat, launch)) return rc;
  }
  if (!scores || (!scores->arc_scores && !scores->theta)) return fail(NFST_ERR_BAD_ARG, "need arc_scores and/or theta");
  if (!backptr || !path_off || !path_arcs || !path_len)
    return fail(NFST_ERR_BAD_ARG, "nfst_viterbi_paths_f32: backptr, path_off, path_arcs and path_len are required");
  if (launch->n_ids == 0) return NFST_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (launch->small_max_arcs > 0)  // backpointers are followed inside the kernel, from shared memory
    return launch_small_sc<float, false, true, false, true, false>(lat, launch, scores, nullptr, nullptr, nullptr, nullptr,
                                                                   nullptr, nullptr, nullptr, delta, backptr, vit_score, st,
                                                                   PathOut{path_off, path_arcs, path_len});
  if (launch->tiles) {  // tile-stream group: tropical pull pass (nfst_tiles.cu), then the same read-out
    if (int rc = nfst_tile_pull_f32(lat, launch, scores, nullptr, nullptr, nullptr, delta, backptr, vit_score, cuda_stream))
      return rc;
  } else if (launch->sell) {  // sliced-column group: tropical pull pass (nfst_sell.cu), then the same read-out
    if (int rc = nfst_sell_pull_f32(lat, launch, scores, nullptr, nullptr, nullptr, delta, backptr, vit_score, cuda_stream))
      return rc;
  } else if (int rc = nfst_bwd_fused_f32(lat, launch, scores, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                         delta, backptr, vit_score, cuda_stream)) {
    return rc;
  }
  const int threads = 128;
  nfst_backtrace_kernel<<<(launch->n_ids + threads - 1) / threads, threads, 0, st>>>(
      *lat, launch->lattice_ids, launch->n_ids, backptr, path_off, path_arcs, path_len);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_compact_paths(const nfst_packed_lattices_t* lat, const int32_t* path_off, const int32_t* path_len,
                       const int32_t* path_buf, const int64_t* out_off, int32_t* out_arcs, int32_t* out_labels,
                       void* cuda_stream) {
  if (!lat || !path_off || !path_len || !path_buf || !out_off || !out_arcs)
    return fail(NFST_ERR_BAD_ARG, "nfst_compact_paths: null argument");
  if (lat->n_lattices == 0) return NFST_OK;
  const int threads = 256;
  const int blocks = (lat->n_lattices * 32 + threads - 1) / threads;
  nfst_compact_paths_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      *lat, path_off, path_len, path_buf, out_off, out_arcs, out_labels);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_pad_paths(const nfst_packed_lattices_t* lat, const int32_t* path_off, const int32_t* path_len, const int32_t* path_buf,
                   int32_t skip_label, int64_t pad_label, int32_t row_len, int64_t* out_labels, int32_t* out_len,
                   void* cuda_stream) {
  if (!lat || !path_off || !path_len || !path_buf || !out_labels || row_len < 0)
    return fail(NFST_ERR_BAD_ARG, "nfst_pad_paths: bad argument");
  if (lat->n_lattices == 0) return NFST_OK;
  const int threads = 256;
  const int blocks = (lat->n_lattices * 32 + threads - 1) / threads;
  nfst_pad_paths_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      *lat, path_off, path_len, path_buf, skip_label, pad_label, row_len, out_labels, out_len);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_beta_to_dense(const nfst_packed_lattices_t* lat, const void* beta, int beta_f64, const int32_t* orig_state,
                       int32_t k, int32_t dense_states, float* out, void* cuda_stream) {
  if (!lat || !beta || !orig_state || !out || k < 1 || dense_states < 1)
    return fail(NFST_ERR_BAD_ARG, "nfst_beta_to_dense: bad argument");
  if (lat->n_states == 0) return NFST_OK;
  const int threads = 256;
  const int blocks = (lat->n_states + threads - 1) / threads;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (beta_f64)
    nfst_beta_to_dense_kernel<double><<<blocks, threads, 0, st>>>(*lat, static_cast<const double*>(beta), orig_state, k,
                                                                  dense_states, out);
  else
    nfst_beta_to_dense_kernel<float><<<blocks, threads, 0, st>>>(*lat, static_cast<const float*>(beta), orig_state, k,
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
                            const int64_t* row_start, int64_t arc_capacity, int32_t* arc_row, int32_t* arc_label,
                            int32_t* arc_dst, void* cuda_stream) {
  if (!transition || !row_start || !arc_row || !arc_label || !arc_dst || n_rows < 0 || states_per_lattice < 1 ||
      vocab < 1 || arc_capacity < 0)
    return fail(NFST_ERR_BAD_ARG, "nfst_dense_extract_arcs: bad argument");
  if (n_rows == 0) return NFST_OK;
  const int threads = 256;
  const int64_t blocks = (n_rows * 32 + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return fail(NFST_ERR_TOO_LARGE, "too many table rows");
  nfst_dense_extract_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      transition, n_rows, states_per_lattice, vocab, row_start, arc_capacity, arc_row, arc_label, arc_dst);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

int nfst_beta_hat_level_f32(const nfst_packed_lattices_t* lat, const int32_t* states, int32_t n_states, int32_t hidden,
                            const float* label_proj, const float* wh_t, const float* w, float* log_beta,
                            float* beta_hat, float* h_proj, void* cuda_stream) {
  if (!lat || !states || !label_proj || !wh_t || !w || !log_beta || !beta_hat || !h_proj)
    return fail(NFST_ERR_BAD_ARG, "nfst_beta_hat_level_f32: null argument");
  if (hidden < 1 || hidden > 1024) return fail(NFST_ERR_BAD_ARG, "hidden=%d must be in 1..1024", hidden);
  if (n_states < 0) return fail(NFST_ERR_BAD_ARG, "n_states=%d", n_states);
  if (n_states == 0) return NFST_OK;
  const int threads = (hidden + 31) & ~31;
  nfst_beta_hat_level_kernel<<<n_states, threads, threads * sizeof(float), static_cast<cudaStream_t>(cuda_stream)>>>(
      *lat, states, hidden, label_proj, wh_t, w, log_beta, beta_hat, h_proj);
  NFST_CUDA_OK(cudaGetLastError());
  return NFST_OK;
}

}  // extern "C"
