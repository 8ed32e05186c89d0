// nfst_sell.cu -- "sliced column" execution model of the lattice DP (sm_100a).
//
// Same recurrence as nfst_kernels.cu (the reference's FSAGRUScorer.compute_beta_per_sample,
// src/modules/scorers.py:692-751, with Wh = 0, in log space), for lattices whose levels are at
// least a warp wide.  The packer stores the arcs of such a lattice slice by slice: a SLICE is
// 32 consecutive states of one topological level (sorted by out-degree, descending), and the
// slice's arcs are laid out column-major without padding -- column k holds the k-th arc of every
// state of the slice that has more than k arcs, in lane order.  Lane i of a warp owns state i of
// the slice, so the k-th arcs of 32 states are 128 consecutive bytes: no row pointers (one 16-byte
// descriptor per slice and one degree byte per state).
//
// A thread block owns one lattice and walks its levels with one barrier per level; the slices of a
// level are dealt to the warps in serpentine rounds (see Pos).  Every warp owns a small stage in shared
// memory and keeps it filled with cp.async: the descriptor and degree bytes of the slice two ahead,
// the arc columns (dst, score or conditional, label) of the next slice -- issued the moment the
// current slice's columns have been read out into registers, so the copy travels while the current
// slice is computed and no warp waits a memory latency per slice.  Nothing about the slices ahead is
// carried in registers (ptxas spills loop-carried descriptors at 64 registers, and a local-memory
// reload on every slice's critical path doubled the kernel time).  The per-slice code exists in
// several widths (slices whose states have at most 2 / 4 / 6 arcs, and the rest), chosen by a
// warp-uniform branch; inside a width the columns are straight-line and predicated per lane: no warp
// votes, no per-column branches.  128-thread blocks are compiled for 7 resident blocks per SM
// (72 registers: 1024 lattices run in ONE wave), 256-thread blocks for 4 (64 registers).  The only
// data that crosses levels, the per-state DP value, lives in a shared-memory ring.
//
//   pull pass (deepest level first): beta[s] = logsumexp_k (w_k + beta[dst_k]) -- per-arc terms as float32
//       offsets from a reference arc of the state (float64 with a float64 state) -- and the arc's
//       conditional probability cond[a] = exp(w + beta[dst] - beta[s])
//       written in real space next to it; or the tropical recursion delta / backpointer
//       (one fp32 add per arc, first maximum in label order wins -- bit-exact rule).
//   flow pass (start level first): gamma[start] = g_b; post[a] = gamma[src] * cond[a];
//       gamma[dst] += post[a] (shared-memory atomics).  post is d logZ / d w_a scaled by the
//       incoming gradient -- no second logsumexp, no alpha pass, both passes stream the same
//       arrays in the same order.  alpha[s] = log gamma[s] + logZ - beta[s] on request.
#include "nfst_b200.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <utility>

int nfst_fail_msg(int code, const char* fmt, ...);  // nfst_kernels.cu (sets the thread's last error)

namespace {

#define SELL_CUDA_OK(expr)                                                                                  \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) return nfst_fail_msg(NFST_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#ifndef SELL_FLOW_MIN_BLOCKS
#define SELL_FLOW_MIN_BLOCKS 4
#endif
// resident blocks the pull kernels are compiled for (register budget): 128-thread / 256-thread instantiations
#ifndef SELL_PULL_MIN_BLOCKS_128
#define SELL_PULL_MIN_BLOCKS_128 7
#endif
#ifndef SELL_PULL_MIN_BLOCKS_256
#define SELL_PULL_MIN_BLOCKS_256 4
#endif
constexpr int KU = 8;  // arc columns held in registers per slice; deeper columns take the tail loops
constexpr int TW = 4;  // columns per tail window (loads issued together)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kFloor = -1.0e30f;
constexpr float kNegInf = -__builtin_huge_valf();

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {  // 1 ulp; __frcp_rn costs a Newton step and a range branch
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

extern __shared__ __align__(16) unsigned char sell_smem[];

// position of a warp in a lattice: level l, round j of that level.  The slices of a level are dealt to the warps
// in rounds of nw, in SERPENTINE order (round 0 to warps 0..nw-1, round 1 to warps nw-1..0, ...): degrees
// descend along a level, so the slices of a round get lighter from first to last; the warp that takes the
// heaviest slice of the level (the one with the many-arc states, whose arcs beyond the register window cost
// extra memory round trips) takes the lightest of the next round, and the leftover round -- the lightest slices
// -- never lands on that warp when the level has a multiple of nw slices plus one.
struct Pos {
  int l, j;
  bool ok;
};
__device__ __forceinline__ int slice_of(const Pos& p, int warp, int nw) {
  return p.j * nw + ((p.j & 1) ? nw - 1 - warp : warp);
}
// degree byte + descriptor of a slice (loaded two slices ahead; nothing here waits for a load)
struct StA {
  int s;     // this lane's state (global packed id); 0x7fffffff for idle lanes
  int degb;  // its out-degree byte (0 for idle lanes; 255 = saturated, see slice_degree)
  int4 d;    // nfst_packed_lattices_t.sell_desc: arc range, column starts, largest degree
};
// start of column K (0..7) of a slice, relative to its first arc -- no warp vote on the hot path
template <int K>
__device__ __forceinline__ int col_start(const int4& d) {
  if (K == 0) return 0;
  const unsigned w = static_cast<unsigned>(K <= 4 ? d.z : d.w);
  return (w >> (8 * ((K - 1) & 3))) & 0xff;
}
__device__ __forceinline__ int slice_dmax8(const int4& d) { return static_cast<unsigned>(d.w) >> 24; }

// lsl[l] = index of the first slice of level l (lsl[n_levels] closes the last level)
template <bool DESC>
__device__ __forceinline__ void advance(Pos& p, const int* lsl, int n_levels, int warp, int nw) {
  p.j += 1;
  while (DESC ? p.l >= 0 : p.l < n_levels) {
    if (slice_of(p, warp, nw) < lsl[p.l + 1] - lsl[p.l]) {
      p.ok = true;
      return;
    }
    p.l += DESC ? -1 : 1;
    p.j = 0;
  }
  p.ok = false;
}

// ---- slice meta data through shared memory ----
// The descriptor and the degree bytes of a slice travel like its arc columns: cp.async into a per-warp buffer two
// slices ahead (two buffers, by the parity of the warp's slice counter), read when the slice is reached.
// Nothing about the slices ahead is carried in registers: at 64 registers ptxas spilled exactly those
// loop-carried values, and a local-memory reload sits on every slice's critical path.
constexpr int META = 16;  // ints: descriptor (4) + the word-aligned superset of 32 degree bytes (<= 9 words)
__device__ __forceinline__ void meta_issue(const Pos& p, const int* lvl, const int* lsl, int lane, int warp, int nw,
                                           int n_states, const uint8_t* __restrict__ deg8,
                                           const int4* __restrict__ desc, unsigned buf);
__device__ __forceinline__ StA meta_read(const Pos& p, const int* lvl, int lane, int warp, int nw, const int32_t* buf) {
  StA a;
  const int first = lvl[p.l] + 32 * slice_of(p, warp, nw);
  a.d = *reinterpret_cast<const int4*>(buf);
  a.s = 0x7fffffff;
  a.degb = 0;
  if (first + lane < lvl[p.l + 1]) {
    a.s = first + lane;
    a.degb = reinterpret_cast<const uint8_t*>(buf + 4)[(first & 3) + lane];
  }
  return a;
}
// true out-degree (the byte saturates at 255)
__device__ __forceinline__ int slice_degree(const StA& a, const int32_t* __restrict__ out_ptr) {
  int deg = a.degb;
  if (deg == 255) deg = __ldg(out_ptr + a.s + 1) - __ldg(out_ptr + a.s);
  return deg;
}
// Visit columns 0..KU-1 of a slice: f(k, on, arc) with k a compile-time constant for the caller's register
// arrays.  Straight-line and predicated per lane (on = the lane's state has a k-th arc): no warp vote, no
// branch on the slice's largest degree -- the branches cost more in reconvergence and register moves than the
// idle columns do.
template <int K, int N, typename F>
__device__ __forceinline__ void for_columns(const int4& d, int deg, int lane, F&& f) {
  if constexpr (K < N) {
    f(std::integral_constant<int, K>{}, deg > K, d.x + col_start<K>(d) + lane);
    for_columns<K + 1, N>(d, deg, lane, f);
  }
}
// The per-slice code exists in several widths: for slices whose states have at most 2 / 4 / 6 arcs (degrees descend
// along a level: most slices of a level are narrow) and for the others.  One warp-uniform branch tree per phase
// picks the instantiation; inside it the columns stay straight-line, so their loads and arithmetic interleave.
// Four widths for the flow and the tropical pass; two (<= KL, KU) for the log-semiring pull pass, whose per-column
// code is the longest: with four its loop outgrows the instruction cache (measured 0.563 vs 0.541 ms).
constexpr int KL = 4;
template <int WIDTHS, typename F>
__device__ __forceinline__ void by_width(int dmax, F&& f) {
  if constexpr (WIDTHS == 4) {
    if (dmax <= 4) {
      if (dmax <= 2) f(std::integral_constant<int, 2>{});
      else f(std::integral_constant<int, 4>{});
    } else {
      if (dmax <= 6) f(std::integral_constant<int, 6>{});
      else f(std::integral_constant<int, KU>{});
    }
  } else {
    if (dmax <= KL) f(std::integral_constant<int, KL>{});
    else f(std::integral_constant<int, KU>{});
  }
}
// ---- per-warp staging of a slice's arc columns (cp.async, 16 bytes per lane) ----
// The columns 0..KU-1 of a slice lie inside its first 32*KU arcs.  Each warp owns one stage of SA elements per
// streamed array; the copy of the NEXT slice's arcs is issued as soon as the current slice's columns have been
// read out into registers, so it travels while the current slice is computed: one slice of every warp is
// always in flight and no warp waits a full memory latency per slice.  Copies are 16-byte chunks of the
// aligned superset of the range (the arrays are 16-byte aligned; the last chunk of the last slice of the batch
// is copied by element).
constexpr int SA = 32 * KU + 8;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async_16(unsigned smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(unsigned smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <bool A1, bool A2>
__device__ __forceinline__ void stage_issue(const int4& d, int lane, int n_arcs, unsigned stage, const int32_t* g0,
                                            const void* g1, const void* g2) {
  // stage: 32-bit shared-memory address of the warp's stage (arrays SA * 4 bytes apart)
  const int base = d.x & ~3;
  const int n16 = (min(d.y, d.x + 32 * KU) - base + 3) >> 2;  // <= 65, the same for the whole warp
  const int32_t* h1 = static_cast<const int32_t*>(g1);
  const int32_t* h2 = static_cast<const int32_t*>(g2);
  constexpr unsigned AB = SA * 4;
  if (base + 4 * n16 <= n_arcs) {
    // every slice but the very last of the batch: whole 16-byte chunks; a slice of average size needs one
    // round, the second and third run under warp-uniform branches
    const int e0 = base + 4 * lane;
    const unsigned st = stage + 16 * lane;
    if (lane < n16) {
      cp_async_16(st, g0 + e0);
      if (A1) cp_async_16(st + AB, h1 + e0);
      if (A2) cp_async_16(st + 2 * AB, h2 + e0);
    }
    if (n16 > 32) {
      if (lane + 32 < n16) {
        cp_async_16(st + 512, g0 + e0 + 128);
        if (A1) cp_async_16(st + AB + 512, h1 + e0 + 128);
        if (A2) cp_async_16(st + 2 * AB + 512, h2 + e0 + 128);
      }
      if (lane + 64 < n16) {
        cp_async_16(st + 1024, g0 + e0 + 256);
        if (A1) cp_async_16(st + AB + 1024, h1 + e0 + 256);
        if (A2) cp_async_16(st + 2 * AB + 1024, h2 + e0 + 256);
      }
    }
  } else {
    for (int j = lane; j < 4 * n16; j += 32) {
      const int e = base + j;
      if (e < n_arcs) {
        cp_async_4(stage + 4 * j, g0 + e);
        if (A1) cp_async_4(stage + AB + 4 * j, h1 + e);
        if (A2) cp_async_4(stage + 2 * AB + 4 * j, h2 + e);
      }
    }
  }
}

__device__ __forceinline__ void meta_issue(const Pos& p, const int* lvl, const int* lsl, int lane, int warp, int nw,
                                           int n_states, const uint8_t* __restrict__ deg8,
                                           const int4* __restrict__ desc, unsigned buf) {
  // one 4-byte copy per lane, no divergent branch: lanes 0..3 the descriptor, lanes 4..12 the degree words
  const int jj = slice_of(p, warp, nw);
  const int first = lvl[p.l] + 32 * jj;
  const int w = (first >> 2) + lane - 4;  // out_deg8 is 4-byte aligned and padded to a whole word
  const int32_t* const dsrc = reinterpret_cast<const int32_t*>(desc + lsl[p.l] + jj) + lane;
  const int32_t* const gsrc = reinterpret_cast<const int32_t*>(deg8) + w;
  const int32_t* src;  // a select, not a branch: both addresses are a handful of instructions
  asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %3, 4;\n\tselp.b64 %0, %1, %2, p;\n\t}" : "=l"(src) : "l"(dsrc), "l"(gsrc), "r"(lane));
  if (lane < 4 || (lane < 13 && 4 * w < n_states)) cp_async_4(buf + 4 * lane, src);
}

// rare path: an arc longer than the shared-memory ring (kept out of line: no 64-bit address arithmetic in
// the hot loops).  The value was written by this block before an earlier level barrier.
#ifdef SELL_FAR_INLINE
#define SELL_FAR_ATTR __forceinline__
#else
#define SELL_FAR_ATTR __noinline__
#endif
template <typename T>
__device__ SELL_FAR_ATTR T far_load(const T* p, int i) {
  return *reinterpret_cast<const volatile T*>(p + i);
}

// ---- columns k >= KU of a slice (states with more arcs than the register window holds) ----
// tail_windows: while more than one state is still active, hands out further windows of TW columns --
// win(a, on): a[j] = this lane's arc in column k + j, on[j] = the lane has one -- so that a window's loads are
// issued together (one memory latency per window, not per column).  Returns with (col, k) at the first
// column that has at most one active state.  Degrees descend inside a slice, so that state is lane 0's and
// the rest of ITS arcs are contiguous from `col` (single-entry columns): the caller strides the whole warp
// over them (tail_rest) and reduces what the helper lanes computed back into lane 0.
template <typename Win>
__device__ __forceinline__ void tail_windows(int deg, int& col, int& k, int lane, Win win) {
  for (;;) {
    if (__popc(__ballot_sync(0xffffffffu, deg > k)) <= 1) return;
    int a[TW];
    bool on[TW];
#pragma unroll
    for (int j = 0; j < TW; ++j) {
      on[j] = deg > k + j;
      a[j] = col + lane;
      col += __popc(__ballot_sync(0xffffffffu, on[j]));
    }
    k += TW;
    win(a, on);
  }
}
__device__ __forceinline__ int tail_rest(int deg, int k) {
  const int r = __shfl_sync(0xffffffffu, deg, 0) - k;
  return r > 0 ? r : 0;
}

// online logsumexp pair (m, s): value = m + log(s); branch-free, finite floor instead of -inf
template <typename RT>
__device__ __forceinline__ void lse_push(RT& m, float& s, RT v) {
  const float d = static_cast<float>(v - m);
  const float e = ex2_approx(-fabsf(d) * kLog2e);
  const bool up = d > 0.0f;
  s = up ? fmaf(s, e, 1.0f) : s + e;
  m = up ? v : m;
}
template <typename RT>
__device__ __forceinline__ void lse_join(RT& m, float& s, RT m2, float s2) {
  const float d = static_cast<float>(m2 - m);
  const float e = ex2_approx(-fabsf(d) * kLog2e);
  const bool up = d > 0.0f;
  s = up ? fmaf(s, e, s2) : fmaf(s2, e, s);
  m = up ? m2 : m;
}

// =====================================================================================
// pull pass
// =====================================================================================
// NT_MAX: largest block the instantiation may be launched with.  128: 72 registers (7 blocks per SM: 1024 lattices
// fit the 148 SMs in ONE wave; at 64 registers ptxas spills the loop-carried slice descriptors and every reload
// is a local-memory round trip on the slice's critical path -- measured 1.39 vs 0.65 ms); 256: 80 registers
template <bool TROP, bool SC, bool TH, bool COND, typename OT, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, NT_MAX == 128 ? SELL_PULL_MIN_BLOCKS_128 : NT_MAX == 256 ? SELL_PULL_MIN_BLOCKS_256 : 1)
    sell_pull_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int W, int max_levels,
                     int theta_smem, int far, int stage_off, const float* __restrict__ arc_scores, const float* __restrict__ theta,
                     OT* beta, OT* __restrict__ logz, float* __restrict__ cond, float* delta,
                     int32_t* __restrict__ backptr, float* __restrict__ vit_score) {
  using RT = typename std::conditional<TROP, float, double>::type;
  using RingT = typename std::conditional<TROP, float, OT>::type;  // ring precision = the state dtype
  RingT* const ring = reinterpret_cast<RingT*>(sell_smem);
  const int lvl_words = (max_levels + 2 + 3) & ~3;
  int* const lvl = reinterpret_cast<int*>(sell_smem + static_cast<size_t>(W) * sizeof(RingT));
  int* const lsl = lvl + lvl_words;
  const float* th = theta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int lo = L.level_off[b];
  const int n_levels = L.level_off[b + 1] - lo - 1;
  for (int i = tid; i <= n_levels; i += blockDim.x) {
    lvl[i] = L.level_ptr[lo + i];
    lsl[i] = L.sell_lvl_slice[lo + i];
  }
  const int4* __restrict__ desc = reinterpret_cast<const int4*>(L.sell_desc);
  if (TH && theta_smem) {
    float* sth = reinterpret_cast<float*>(lsl + lvl_words);
    for (int i = tid; i < L.vocab; i += blockDim.x) sth[i] = theta[i];
    th = sth;
  }
  // per-warp stage at byte offset stage_off (computed by the host: the compiler would re-derive it from W, the
  // level count and the vocabulary on every slice): two meta buffers, then [dst | scores (SC) | labels (TH)]
  int32_t* const stage = reinterpret_cast<int32_t*>(sell_smem + stage_off) + warp * ((1 + SC + TH) * SA + 2 * META) + 2 * META;
  __syncthreads();
  const int mask = W - 1;
  const int start = L.start_state[b];
  const uint8_t* __restrict__ deg8 = L.out_deg8;
  const int32_t* __restrict__ dst_out = L.dst_out;
  const int32_t* __restrict__ label_out = L.label_out;
  // DP value of an arc's destination.  States of the last level are final: beta = 1 (scorers.py:720),
  // delta = 0 -- no load.  Otherwise from the ring while dst < lim (= first state of the current level + W:
  // the state that would overwrite dst's slot has not been reached), else from global memory.
  const int last0 = lvl[n_levels - 1];
  int lim = 0x7fffffff;
  int lim2 = last0;
  // slow part of the rule (d >= lim2 = min(lim, last0), set per level below)
  auto nbr_slow = [&](int d) -> RT {
    if (d >= last0) return static_cast<RT>(0);
    return TROP ? static_cast<RT>(far_load(delta, d)) : static_cast<RT>(far_load(beta, d));
  };
  auto nbr = [&](int d) -> RT {
    if (d < lim2) return static_cast<RT>(ring[d & mask]);
    return nbr_slow(d);
  };
  auto arc_score = [&](int a) -> float {
    float w = SC ? arc_scores[a] : 0.0f;
    if (TH) w += th[label_out[a]];
    return w;
  };
  // values of one window of tail columns, loads first
  auto window_values = [&](const int (&a)[TW], const bool (&on)[TW], RT (&v)[TW], RT fill) {
    int d[TW];
    float w[TW];
#pragma unroll
    for (int j = 0; j < TW; ++j) {
      d[j] = 0;
      w[j] = 0.0f;
      if (on[j]) {
        d[j] = dst_out[a[j]];
        w[j] = arc_score(a[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < TW; ++j) v[j] = on[j] ? (TROP ? static_cast<RT>(__fadd_rn(w[j], static_cast<float>(nbr(d[j]))))
                                                      : static_cast<RT>(w[j]) + nbr(d[j]))
                                              : fill;
  };

  // ---- pipeline state: the slice being computed (its arc columns staged in shared memory), the next one
  // (descriptor + degree byte in registers, its stage copy issued once the current columns are read out) and
  // the one after that (descriptor + degree byte in flight).  None of these addresses depends on a DP value.
  const unsigned stage_s = smem_u32(stage);
  auto issue = [&](const int4& d) {
    stage_issue<true, SC && TH>(d, lane, L.n_arcs, stage_s, dst_out, SC ? static_cast<const void*>(arc_scores) : static_cast<const void*>(label_out),
                        SC ? static_cast<const void*>(label_out) : nullptr);
  };
  Pos pc, pn;
  pc.l = n_levels - 1; pc.j = -1; pc.ok = false;
  advance<true>(pc, lsl, n_levels, warp, nw);
  pn = pc;
  if (pn.ok) advance<true>(pn, lsl, n_levels, warp, nw);
  int32_t* const meta = stage - 2 * META;  // two meta buffers in front of the warp's arc stage
  const unsigned meta_s = smem_u32(meta);
  int par = 0;                             // parity of the warp's slice counter: meta buffer of the current slice
  if (pc.ok) meta_issue(pc, lvl, lsl, lane, warp, nw, L.n_states, deg8, desc, meta_s);
  if (pn.ok) meta_issue(pn, lvl, lsl, lane, warp, nw, L.n_states, deg8, desc, meta_s + META * 4);
  cp_async_wait_all();
  __syncwarp();
  if (pc.ok) issue(*reinterpret_cast<const int4*>(meta));
  const int lim_cap = last0;  // ring reads need dst < min(lim, last0)

#pragma unroll 1
  for (int l = n_levels - 1; l >= 0; --l) {
    if (far) {
      lim = lvl[l] + W;
      lim2 = min(lim, lim_cap);
    }
#pragma unroll 1
    while (pc.ok && pc.l == l) {
      Pos pa = pn;
      if (pa.ok) advance<true>(pa, lsl, n_levels, warp, nw);
      // everything issued in the previous iteration has landed: this slice's arcs, the next slice's meta data
      cp_async_wait_all();
      __syncwarp();
      const StA ac = meta_read(pc, lvl, lane, warp, nw, meta + par * META);

      // ---- the current slice's register window, out of the warp's stage ----
      const int deg = slice_degree(ac, L.out_ptr), s = ac.s;
      const int dmax = slice_dmax8(ac.d);  // 255 = "255 or more"
      int dstc[KU];
      float wc[KU];
      by_width<(TROP ? 4 : 2)>(dmax, [&](auto nc) {
        constexpr int NC = decltype(nc)::value;
        // branch-free: lanes without a k-th arc read whatever the stage holds at their position (always inside
        // the stage) and ignore the value -- cheaper than a reconvergence point, or even a select, per column
        const int32_t* const sd = stage + (ac.d.x & 3);
        const float* const sw = reinterpret_cast<const float*>(sd + SA);
        const int32_t* const sl = sd + (SC ? 2 : 1) * SA;
        for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int) {
          constexpr int k = decltype(kc)::value;
          const int o = col_start<k>(ac.d) + lane;  // lanes without a k-th arc read a neighbour's (ignored)
          dstc[k] = sd[o];
          wc[k] = SC ? sw[o] : 0.0f;
          if (TH) wc[k] += th[on ? sl[o] : 0];  // an empty slice stages nothing: no stale label may index theta
        });
      });
      __syncwarp();  // every lane has its columns and meta data: both buffers are free
      if (pn.ok) issue(*reinterpret_cast<const int4*>(meta + (par ^ 1) * META));
      if (pa.ok) meta_issue(pa, lvl, lsl, lane, warp, nw, L.n_states, deg8, desc, meta_s + par * (META * 4));
      by_width<(TROP ? 4 : 2)>(dmax, [&](auto nc) {
      constexpr int NC = decltype(nc)::value;
      // destinations' DP values: from the ring, straight-line; the rare ones beyond it are patched afterwards
      RingT rv[KU];
      bool any_slow = false;
      for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int) {
        constexpr int k = decltype(kc)::value;
        rv[k] = ring[dstc[k] & mask];
        any_slow |= on && dstc[k] >= lim2;
      });
      if (any_slow) {
        for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int) {
          constexpr int k = decltype(kc)::value;
          if (on && dstc[k] >= lim2) rv[k] = static_cast<RingT>(nbr_slow(dstc[k]));
        });
      }

      // ---- DP of the current slice ----
      if (!TROP) {
        // register window: two passes over registers (max, then exponentials)
        // t_k = w_k + beta[dst_k] is only needed as an offset from a reference arc's t0 (the first arc; every
        // state with arcs has one).  float32 state: the offset is formed as (beta_k - beta_0) + (w_k - w_0) --
        // two differences of float32 numbers, each rounded at its own small size (6e-8 relative), so nothing is
        // lost against the float64 sum, and the per-arc float64 converts and adds (the busiest pipe of this
        // kernel in ncu) are gone; t0 itself stays float64, once per slice.  float64 state: float64 throughout.
        float rw = wc[0];
        RingT rb = rv[0];
        if (deg > 0 && !(rw + static_cast<float>(rb) > kFloor)) {  // rare: the first arc scores -inf -- take the largest
          float tbest = kFloor;
          rw = kFloor;
          rb = static_cast<RingT>(0);
          for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int) {
            constexpr int k = decltype(kc)::value;
            const float t = wc[k] + static_cast<float>(rv[k]);
            if (on && t > tbest) {
              tbest = t;
              rw = wc[k];
              rb = rv[k];
            }
          });
        }
        float tf[KU];
        for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int) {
          constexpr int k = decltype(kc)::value;
          float off;
          if constexpr (sizeof(RingT) == 4) off = (static_cast<float>(rv[k]) - static_cast<float>(rb)) + (wc[k] - rw);
          else off = static_cast<float>(static_cast<RT>(wc[k]) + static_cast<RT>(rv[k]) - (static_cast<RT>(rw) + static_cast<RT>(rb)));
          tf[k] = on ? off : kFloor;
        });
        float mf = kFloor;
#pragma unroll
        for (int k = 0; k < NC; ++k) mf = fmaxf(mf, tf[k]);
        float e[KU];
        float sum = 0.0f;
        [[maybe_unused]] const float mfl = -mf * kLog2e;  // float32 state only
        for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool, int) {
          constexpr int k = decltype(kc)::value;
          // idle lanes / -inf arcs: exp(-huge) = 0.  float32 state: one fused multiply-add per arc; the rounding of
          // mf * log2(e) then scales every term of the state alike (cancels in cond, shifts beta by < 1e-7, far
          // below the float32 state's own rounding).  float64 state: exact difference first -- the shift would
          // add up over hundreds of levels.
          if constexpr (sizeof(RingT) == 4) e[k] = ex2_approx(fmaf(tf[k], kLog2e, mfl));
          else e[k] = ex2_approx((tf[k] - mf) * kLog2e);
          sum += e[k];
        });
        if (!(mf > kFloor)) sum = 0.0f;  // no finite arc: every term above was exp(0)
        // sinks: beta = 1 (scorers.py:720); a state whose arcs all score -inf: beta = -inf.
        // beta[s] = beta_0 + (w_0 + max offset + log sum): one rounding at beta's magnitude
        RingT bv = static_cast<RingT>(0);
        float inv_reg = sum > 0.0f ? rcp_approx(sum) : 0.0f;  // scale of the window's exponentials in cond[]
        if (deg > 0) {
          const float lg = lg2_approx(sum) * kLn2;
          if (!(sum > 0.0f)) bv = static_cast<RingT>(kNegInf);
          else if constexpr (sizeof(RingT) == 4) bv = rb + (rw + (mf + lg));
          else bv = rb + (static_cast<RT>(rw) + (static_cast<RT>(mf) + static_cast<RT>(lg)));
        }
        // further columns (warp-uniform branch, rare): merged online into (m, sum) in float64; what depends on
        // them is redone inside the branch, so that no float64 value lives outside it
        if (NC == KU && dmax > KU) {
          RT m = mf > kFloor ? (static_cast<RT>(rw) + static_cast<RT>(rb)) + static_cast<RT>(mf) : static_cast<RT>(kFloor);
          const RT m_reg = m;
          int col_tail = 0;
          int col = ac.d.x + col_start<KU - 1>(ac.d) + __popc(__ballot_sync(0xffffffffu, deg > KU - 1));
          col_tail = col;
          int k = KU;
          tail_windows(deg, col, k, lane, [&](const int (&a)[TW], const bool (&on)[TW]) {
            RT v[TW];
            window_values(a, on, v, static_cast<RT>(kFloor));
            RT mx = m;
#pragma unroll
            for (int j = 0; j < TW; ++j) mx = v[j] > mx ? v[j] : mx;
            float add = 0.0f;
#pragma unroll
            for (int j = 0; j < TW; ++j) add += ex2_approx(static_cast<float>(v[j] - mx) * kLog2e);
            sum = fmaf(sum, ex2_approx(static_cast<float>(m - mx) * kLog2e), add);
            m = mx;
          });
          const int rest = tail_rest(deg, k);
          if (rest > 0) {  // lane 0's state alone: every lane takes a share of its remaining arcs
            RT hm = static_cast<RT>(kFloor);
            float hs = 0.0f;
#pragma unroll 4
            for (int i = lane; i < rest; i += 32) {
              const int a = col + i;
              lse_push(hm, hs, static_cast<RT>(arc_score(a)) + nbr(dst_out[a]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const RT m2 = __shfl_xor_sync(0xffffffffu, hm, o);
              const float s2 = __shfl_xor_sync(0xffffffffu, hs, o);
              lse_join(hm, hs, m2, s2);
            }
            if (lane == 0) lse_join(m, sum, hm, hs);
          }
          if (deg > 0) bv = sum > 0.0f ? static_cast<RingT>(m + static_cast<RT>(lg2_approx(sum) * kLn2)) : static_cast<RingT>(kNegInf);
          const float inv = sum > 0.0f ? __frcp_rn(sum) : 0.0f;
          inv_reg = inv * ex2_approx(static_cast<float>(m_reg - m) * kLog2e);
          if (COND) {
            int col = col_tail, k = KU;
            tail_windows(deg, col, k, lane, [&](const int (&a)[TW], const bool (&on)[TW]) {
              RT v[TW];
              window_values(a, on, v, static_cast<RT>(kFloor));
#pragma unroll
              for (int j = 0; j < TW; ++j)
                if (on[j]) cond[a[j]] = ex2_approx(static_cast<float>(v[j] - m) * kLog2e) * inv;
            });
            const int rest = tail_rest(deg, k);
            if (rest > 0) {
              const RT m0 = __shfl_sync(0xffffffffu, m, 0);
              const float inv0 = __shfl_sync(0xffffffffu, inv, 0);
#pragma unroll 4
              for (int i = lane; i < rest; i += 32) {
                const int a = col + i;
                const RT v = static_cast<RT>(arc_score(a)) + nbr(dst_out[a]);
                cond[a] = ex2_approx(static_cast<float>(v - m0) * kLog2e) * inv0;
              }
            }
          }
        }
        if (COND) {
          for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int a) {
            constexpr int k = decltype(kc)::value;
            if (on) cond[a] = e[k] * inv_reg;
          });
        }
        if (s != 0x7fffffff) {
          // the ring slot of s is only read by shallower levels, i.e. after the level barrier
          ring[s & mask] = static_cast<RingT>(bv);
          if (beta) beta[s] = static_cast<OT>(bv);
          if (s == start && logz) logz[b] = static_cast<OT>(bv);
        }
      } else {
        float best = 0.0f;
        int arg = -1;
        for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int a) {
          constexpr int k = decltype(kc)::value;
          const float c = __fadd_rn(wc[k], rv[k]);
          if (on && (arg < 0 || c > best)) {  // strict: arcs of a state come in label order
            best = c;
            arg = a;
          }
        });
        if (NC == KU && dmax > KU) {
          int col = ac.d.x + col_start<KU - 1>(ac.d) + __popc(__ballot_sync(0xffffffffu, deg > KU - 1)), k = KU;
          tail_windows(deg, col, k, lane, [&](const int (&a)[TW], const bool (&on)[TW]) {
            RT v[TW];
            window_values(a, on, v, kNegInf);
#pragma unroll
            for (int j = 0; j < TW; ++j)
              if (on[j] && v[j] > best) {
                best = v[j];
                arg = a[j];
              }
          });
          const int rest = tail_rest(deg, k);
          if (rest > 0) {
            float hbest = kNegInf;
            int harg = 0x7fffffff;
#pragma unroll 4
            for (int i = lane; i < rest; i += 32) {
              const int a = col + i;
              const float c = __fadd_rn(arc_score(a), nbr(dst_out[a]));
              if (harg == 0x7fffffff || c > hbest) {  // a lane's arcs ascend: its first maximum stays
                hbest = c;
                harg = a;
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ob = __shfl_xor_sync(0xffffffffu, hbest, o);
              const int oa = __shfl_xor_sync(0xffffffffu, harg, o);
              if (oa != 0x7fffffff && (harg == 0x7fffffff || ob > hbest || (ob == hbest && oa < harg))) {
                hbest = ob;
                harg = oa;
              }
            }
            if (lane == 0 && harg != 0x7fffffff && hbest > best) {  // earlier columns win ties (smaller label)
              best = hbest;
              arg = harg;
            }
          }
        }
        if (s != 0x7fffffff) {
          ring[s & mask] = best;  // sinks: delta = 0, backpointer -1
          if (delta) delta[s] = best;
          backptr[s] = arg;
          if (s == start && vit_score) vit_score[b] = best;
        }
      }

      });  // by_width

      // ---- rotate the pipeline ----
      pc = pn;
      pn = pa;
      par ^= 1;
    }
    __syncthreads();
  }
}

// =====================================================================================
// flow pass
// =====================================================================================
template <bool DTH, bool ALPHA, typename OT, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, NT_MAX == 128 ? 7 : NT_MAX == 256 ? SELL_FLOW_MIN_BLOCKS : 1)
    sell_flow_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, int W, int max_levels,
                     int dtheta_smem, int far, int stage_off, const float* cond, const float* __restrict__ grad_logz, float* post,
                     const OT* __restrict__ beta, const OT* __restrict__ logz, OT* __restrict__ alpha,
                     float* __restrict__ dtheta, float* gamma_far) {
  float* const ring = reinterpret_cast<float*>(sell_smem);
  const int lvl_words = (max_levels + 2 + 3) & ~3;
  int* const lvl = reinterpret_cast<int*>(sell_smem + static_cast<size_t>(W) * sizeof(float));
  int* const lsl = lvl + lvl_words;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int lo = L.level_off[b];
  const int n_levels = L.level_off[b + 1] - lo - 1;
  for (int i = tid; i <= n_levels; i += blockDim.x) {
    lvl[i] = L.level_ptr[lo + i];
    lsl[i] = L.sell_lvl_slice[lo + i];
  }
  const int4* __restrict__ desc = reinterpret_cast<const int4*>(L.sell_desc);
  for (int i = tid; i < W; i += blockDim.x) ring[i] = 0.0f;
  float* hist = dtheta;
  if (DTH && dtheta_smem) {
    hist = reinterpret_cast<float*>(lsl + lvl_words);
    for (int i = tid; i < L.vocab; i += blockDim.x) hist[i] = 0.0f;
  }
  // per-warp stage at byte offset stage_off: two meta buffers, then [dst | cond | labels (DTH)]
  int32_t* const stage = reinterpret_cast<int32_t*>(sell_smem + stage_off) + warp * ((2 + DTH) * SA + 2 * META) + 2 * META;
  __syncthreads();
  const int mask = W - 1;
  const int start = L.start_state[b];
  const float g = grad_logz ? grad_logz[b] : 1.0f;
  if (tid == 0) ring[start & mask] = g;
  __syncthreads();
  const uint8_t* __restrict__ deg8 = L.out_deg8;
  const int32_t* __restrict__ dst_out = L.dst_out;
  const int32_t* __restrict__ label_out = L.label_out;
  const double lz = ALPHA ? static_cast<double>(logz[b]) : 0.0;
  const float inv_g = ALPHA ? 1.0f / g : 0.0f;
  // gamma[dst] += p.  States of the last level have no arcs: their gamma is only needed for alpha (then it
  // goes through gamma_far).  Otherwise in the ring while dst < lim (= first state of the current level + W:
  // the previous owner of the slot has been consumed), else in global memory (gamma_far, read back when dst
  // is reached).
  const int last0 = lvl[n_levels - 1];
  int lim = 0x7fffffff;
  auto push = [&](int d, float p) {
    if (d >= last0) {
      if (ALPHA) atomicAdd(gamma_far + d, p);
    } else if (d < lim) {
      atomicAdd(&ring[d & mask], p);
    } else {
      atomicAdd(gamma_far + d, p);
    }
  };
  const bool read_far = far || ALPHA;

  // pipeline as in the pull pass: current slice staged, next slice's copy in flight, descriptors two ahead.
  // cond may be the same buffer as post: an arc's conditional is staged before its own posterior is written
  // (the 16-byte superset may re-read a neighbour slice's arcs, whose values are ignored).
  const unsigned stage_s = smem_u32(stage);
  auto issue = [&](const int4& d) {
    stage_issue<true, DTH>(d, lane, L.n_arcs, stage_s, dst_out, cond, label_out);
  };
  Pos pc, pn;
  pc.l = 0; pc.j = -1; pc.ok = false;
  advance<false>(pc, lsl, n_levels, warp, nw);
  pn = pc;
  if (pn.ok) advance<false>(pn, lsl, n_levels, warp, nw);
  int32_t* const meta = stage - 2 * META;  // two meta buffers in front of the warp's arc stage
  const unsigned meta_s = smem_u32(meta);
  int par = 0;                             // parity of the warp's slice counter: meta buffer of the current slice
  if (pc.ok) meta_issue(pc, lvl, lsl, lane, warp, nw, L.n_states, deg8, desc, meta_s);
  if (pn.ok) meta_issue(pn, lvl, lsl, lane, warp, nw, L.n_states, deg8, desc, meta_s + META * 4);
  cp_async_wait_all();
  __syncwarp();
  if (pc.ok) issue(*reinterpret_cast<const int4*>(meta));

#pragma unroll 1
  for (int l = 0; l < n_levels; ++l) {
    if (far) lim = lvl[l] + W;
    const int lim2 = min(lim, last0);  // ring pushes need dst < lim2
#pragma unroll 1
    while (pc.ok && pc.l == l) {
      Pos pa = pn;
      if (pa.ok) advance<false>(pa, lsl, n_levels, warp, nw);
      // everything issued in the previous iteration has landed: this slice's arcs, the next slice's meta data
      cp_async_wait_all();
      __syncwarp();
      const StA ac = meta_read(pc, lvl, lane, warp, nw, meta + par * META);

      const int deg = slice_degree(ac, L.out_ptr), s = ac.s;
      const int dmax = slice_dmax8(ac.d);
      // flow that reached s through global memory (complete: earlier levels are done)
      float g_far = 0.0f;
      if (read_far && s != 0x7fffffff) g_far = *reinterpret_cast<volatile float*>(gamma_far + s);
      int dstc[KU], labc[KU];
      float cc[KU];
      by_width<4>(dmax, [&](auto nc) {
        constexpr int NC = decltype(nc)::value;
        const int32_t* const sd = stage + (ac.d.x & 3);
        const float* const sc = reinterpret_cast<const float*>(sd + SA);
        const int32_t* const sl = sd + 2 * SA;
        for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int) {  // branch-free, see the pull pass
          constexpr int k = decltype(kc)::value;
          const int o = col_start<k>(ac.d) + lane;  // lanes without a k-th arc read a neighbour's (ignored)
          dstc[k] = sd[o];
          cc[k] = sc[o];
          labc[k] = DTH ? sl[o] : 0;
        });
      });
      __syncwarp();  // every lane has its columns and meta data: both buffers are free
      if (pn.ok) issue(*reinterpret_cast<const int4*>(meta + (par ^ 1) * META));
      if (pa.ok) meta_issue(pa, lvl, lsl, lane, warp, nw, L.n_states, deg8, desc, meta_s + par * (META * 4));
      float gs = 0.0f;
      if (s != 0x7fffffff) {
        // every arc into s comes from a shallower level: gamma[s] is final; free the slot
        gs = ring[s & mask] + g_far;
        ring[s & mask] = 0.0f;
        if (ALPHA) {
          const float r = gs * inv_g;
          alpha[s] = r > 0.0f ? static_cast<OT>(static_cast<double>(lg2_approx(r) * kLn2) + lz -
                                                static_cast<double>(beta[s]))
                              : static_cast<OT>(kNegInf);
        }
      }
      by_width<4>(dmax, [&](auto nc) {
      constexpr int NC = decltype(nc)::value;
      bool any_slow = false;
      for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int a) {
        constexpr int k = decltype(kc)::value;
        const float p = gs * cc[k];
        if (on) {
          post[a] = p;
          if (dstc[k] < lim2) atomicAdd(&ring[dstc[k] & mask], p);
          else any_slow = true;
          if (DTH) atomicAdd(&hist[labc[k]], p);
        }
      });
      if (any_slow) {  // rare: flow into the last level (alpha only) or beyond the ring
        for_columns<0, NC>(ac.d, deg, lane, [&](auto kc, bool on, int) {
          constexpr int k = decltype(kc)::value;
          if (on && dstc[k] >= lim2) push(dstc[k], gs * cc[k]);
        });
      }
      });  // by_width
      if (dmax > KU) {
        int col = ac.d.x + col_start<KU - 1>(ac.d) + __popc(__ballot_sync(0xffffffffu, deg > KU - 1)), k = KU;
        tail_windows(deg, col, k, lane, [&](const int (&a)[TW], const bool (&on)[TW]) {
          int d[TW];
          float c[TW];
          int lab[TW];
#pragma unroll
          for (int j = 0; j < TW; ++j) {
            d[j] = 0;
            c[j] = 0.0f;
            lab[j] = 0;
            if (on[j]) {
              d[j] = dst_out[a[j]];
              c[j] = cond[a[j]];
              if (DTH) lab[j] = label_out[a[j]];
            }
          }
#pragma unroll
          for (int j = 0; j < TW; ++j)
            if (on[j]) {
              const float p = gs * c[j];
              post[a[j]] = p;
              push(d[j], p);
              if (DTH) atomicAdd(&hist[lab[j]], p);
            }
        });
        const int rest = tail_rest(deg, k);
        if (rest > 0) {
          const float gs0 = __shfl_sync(0xffffffffu, gs, 0);
#pragma unroll 4
          for (int i = lane; i < rest; i += 32) {
            const int a = col + i;
            const float p = gs0 * cond[a];
            post[a] = p;
            push(dst_out[a], p);
            if (DTH) atomicAdd(&hist[label_out[a]], p);
          }
        }
      }

      pc = pn;
      pn = pa;
      par ^= 1;
    }
    __syncthreads();
  }
  if (DTH && dtheta_smem) {
    for (int i = tid; i < L.vocab; i += blockDim.x) {
      const float v = hist[i];
      if (v != 0.0f) atomicAdd(dtheta + i, v);
    }
  }
}

// byte offset of the per-warp stages (ring, level tables, theta / dtheta table in front of them)
size_t pull_stage_off(int W, int max_levels, int vocab, int ring_bytes, bool theta_smem) {
  size_t o = static_cast<size_t>(W) * ring_bytes;
  o += 2 * static_cast<size_t>((max_levels + 2 + 3) & ~3) * 4;
  if (theta_smem) o += static_cast<size_t>(vocab) * 4;
  return (o + 15) & ~static_cast<size_t>(15);
}
size_t flow_stage_off(int W, int max_levels, int vocab, bool dtheta_smem) { return pull_stage_off(W, max_levels, vocab, 4, dtheta_smem); }
size_t stage_bytes(int threads, int n_staged) { return static_cast<size_t>(threads / 32) * (n_staged * SA + 2 * META) * 4; }
size_t pull_smem(int W, int max_levels, int vocab, int ring_bytes, bool theta_smem, int threads, int n_staged) {
  return pull_stage_off(W, max_levels, vocab, ring_bytes, theta_smem) + stage_bytes(threads, n_staged);
}
size_t flow_smem(int W, int max_levels, int vocab, bool dtheta_smem, int threads, int n_staged) {
  return flow_stage_off(W, max_levels, vocab, dtheta_smem) + stage_bytes(threads, n_staged);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename K>
int prepare(K kernel, size_t smem) {
  if (smem > 48 * 1024) {
    SELL_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel),
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  return 0;
}

// blocks above 256 threads need the 64-register instantiation (NFST_SELL_REGS64=1 forces it, for experiments)
bool big_blocks(const nfst_launch_t* launch) {
  static const bool force = getenv("NFST_SELL_REGS64") != nullptr && getenv("NFST_SELL_REGS64")[0] == '1';
  return force || launch->block_threads > 256;
}

int check_launch(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch) {
  if (!lat || !launch) return nfst_fail_msg(NFST_ERR_BAD_ARG, "null argument");
  if (!launch->sell) return nfst_fail_msg(NFST_ERR_BAD_ARG, "launch group is not a sliced-column group");
  if (!lat->out_deg8 || !lat->sell_desc || !lat->sell_lvl_slice)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "packed lattices carry no sliced-column arrays (out_deg8, sell_desc, sell_lvl_slice)");
  const int W = launch->window_states;
  if (W < 32 || (W & (W - 1))) return nfst_fail_msg(NFST_ERR_BAD_ARG, "window_states must be a power of two >= 32");
  const int nt = launch->block_threads;
  if (nt < 32 || nt > 1024 || (nt & 31)) return nfst_fail_msg(NFST_ERR_BAD_ARG, "block_threads must be 32..1024, a multiple of 32");
  if (launch->n_levels <= 0) return nfst_fail_msg(NFST_ERR_BAD_ARG, "n_levels (largest level count of the group) is required");
  return 0;
}

template <bool TROP, bool COND, typename OT>
int launch_pull(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* sc, OT* beta,
                OT* logz, float* cond, float* delta, int32_t* backptr, float* vit, cudaStream_t stream) {
  const bool has_sc = sc->arc_scores != nullptr, has_th = sc->theta != nullptr;
  const bool th_smem = has_th && lat->vocab <= NFST_THETA_SMEM_MAX;
  const size_t smem = pull_smem(launch->window_states, launch->n_levels, lat->vocab, TROP ? 4 : static_cast<int>(sizeof(OT)), th_smem,
                                launch->block_threads, 1 + (has_sc ? 1 : 0) + (has_th ? 1 : 0));
  if (!aligned16(lat->dst_out) || !aligned16(sc->arc_scores) || (has_th && !aligned16(lat->label_out)))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "dst_out, label_out and arc_scores must be 16-byte aligned (they are staged with 16-byte copies)");
  if (smem > 227 * 1024) return nfst_fail_msg(NFST_ERR_TOO_LARGE, "sliced-column window needs %zu bytes of shared memory", smem);
#define SELL_PULL_NT(SCv, THv, NTv)                                                                               \
  do {                                                                                                            \
    auto k = sell_pull_kernel<TROP, SCv, THv, COND, OT, NTv>;                                                     \
    if (int rc = prepare(k, smem)) return rc;                                                                     \
    k<<<launch->n_ids, launch->block_threads, smem, stream>>>(                                                    \
        *lat, launch->lattice_ids, launch->window_states, launch->n_levels, th_smem ? 1 : 0, launch->sell_far,    \
        static_cast<int>(pull_stage_off(launch->window_states, launch->n_levels, lat->vocab,                      \
                                        TROP ? 4 : static_cast<int>(sizeof(OT)), th_smem)),                       \
        sc->arc_scores, sc->theta, beta, logz, cond, delta, backptr, vit);                                        \
  } while (0)
#define SELL_PULL(SCv, THv)                                                                                       \
  do {                                                                                                            \
    if (big_blocks(launch)) SELL_PULL_NT(SCv, THv, 1024);                                                         \
    else if (launch->block_threads <= 128) SELL_PULL_NT(SCv, THv, 128);                                           \
    else SELL_PULL_NT(SCv, THv, 256);                                                                             \
  } while (0)
  if (has_sc && has_th) SELL_PULL(true, true);
  else if (has_th) SELL_PULL(false, true);
  else SELL_PULL(true, false);  // no scores at all: SC with a null pointer is rejected by the caller
#undef SELL_PULL
#undef SELL_PULL_NT
  SELL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

size_t nfst_sell_smem_bytes(const nfst_launch_t* launch, int32_t vocab, int pass, int with_trop, int with_table) {
  if (!launch) return 0;
  const bool tab = with_table && vocab <= NFST_THETA_SMEM_MAX;
  // upper bound: three staged arrays per warp (dst, scores or conditionals, labels)
  return pass == 0 ? pull_smem(launch->window_states, launch->n_levels, vocab, (with_trop || !launch->state_f64) ? 4 : 8, tab,
                               launch->block_threads, 3)
                   : flow_smem(launch->window_states, launch->n_levels, vocab, tab, launch->block_threads, 3);
}

int nfst_sell_pull_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                       void* beta, void* logz_bwd, float* cond, float* delta, int32_t* backptr, float* vit_score,
                       void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!scores || (!scores->arc_scores && !scores->theta))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "arc_scores or theta is required");
  if (launch->n_ids == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const bool logs = beta || logz_bwd || cond;
  const bool trop = delta || backptr || vit_score;
  if (trop && !backptr) return nfst_fail_msg(NFST_ERR_BAD_ARG, "the tropical pass needs backptr[S]");
  if (launch->sell_far && ((logs && !beta) || (trop && !delta)))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "a group with arcs longer than its ring (sell_far) needs beta[S] / delta[S]");
  if (logs) {
    int rc;
    if (launch->state_f64) {
      rc = cond ? launch_pull<false, true, double>(lat, launch, scores, static_cast<double*>(beta),
                                                   static_cast<double*>(logz_bwd), cond, nullptr, nullptr, nullptr, stream)
                : launch_pull<false, false, double>(lat, launch, scores, static_cast<double*>(beta),
                                                    static_cast<double*>(logz_bwd), nullptr, nullptr, nullptr, nullptr, stream);
    } else {
      rc = cond ? launch_pull<false, true, float>(lat, launch, scores, static_cast<float*>(beta),
                                                  static_cast<float*>(logz_bwd), cond, nullptr, nullptr, nullptr, stream)
                : launch_pull<false, false, float>(lat, launch, scores, static_cast<float*>(beta),
                                                   static_cast<float*>(logz_bwd), nullptr, nullptr, nullptr, nullptr, stream);
    }
    if (rc) return rc;
  }
  if (trop) {
    if (int rc = launch_pull<true, false, float>(lat, launch, scores, nullptr, nullptr, nullptr, delta, backptr,
                                                 vit_score, stream))
      return rc;
  }
  return 0;
}

int nfst_sell_flow_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const float* cond,
                       const float* grad_logz, float* post, const void* beta, const void* logz, void* alpha,
                       float* dtheta, float* gamma_far, void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!cond || !post) return nfst_fail_msg(NFST_ERR_BAD_ARG, "cond[A] and post[A] are required");
  if ((launch->sell_far || alpha) && !gamma_far)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "gamma_far[S] is required for groups with arcs longer than their ring "
                                           "(sell_far) and whenever alpha is requested");
  if (alpha && (!beta || !logz)) return nfst_fail_msg(NFST_ERR_BAD_ARG, "alpha needs beta[S] and logz[B]");
  if (launch->n_ids == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const bool dth_smem = dtheta && lat->vocab <= NFST_THETA_SMEM_MAX;
  const size_t smem = flow_smem(launch->window_states, launch->n_levels, lat->vocab, dth_smem, launch->block_threads, dtheta ? 3 : 2);
  if (!aligned16(lat->dst_out) || !aligned16(cond) || (dtheta && !aligned16(lat->label_out)))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "dst_out, label_out and cond must be 16-byte aligned (they are staged with 16-byte copies)");
  if (smem > 227 * 1024) return nfst_fail_msg(NFST_ERR_TOO_LARGE, "sliced-column window needs %zu bytes of shared memory", smem);
#define SELL_FLOW(DTHv, ALv, OT)                                                                                  \
  do {                                                                                                            \
    if (big_blocks(launch)) SELL_FLOW_NT(DTHv, ALv, OT, 1024);                                                    \
    else if (launch->block_threads <= 128) SELL_FLOW_NT(DTHv, ALv, OT, 128);                                      \
    else SELL_FLOW_NT(DTHv, ALv, OT, 256);                                                                        \
  } while (0)
#define SELL_FLOW_NT(DTHv, ALv, OT, NTv)                                                                          \
  do {                                                                                                            \
    auto k = sell_flow_kernel<DTHv, ALv, OT, NTv>;                                                                \
    if (int rc = prepare(k, smem)) return rc;                                                                     \
    k<<<launch->n_ids, launch->block_threads, smem, stream>>>(                                                    \
        *lat, launch->lattice_ids, launch->window_states, launch->n_levels, dth_smem ? 1 : 0, launch->sell_far,   \
        static_cast<int>(flow_stage_off(launch->window_states, launch->n_levels, lat->vocab, dth_smem)),          \
        cond, grad_logz,                                                                                          \
        post, static_cast<const OT*>(beta), static_cast<const OT*>(logz), static_cast<OT*>(alpha), dtheta,        \
        gamma_far);                                                                                               \
  } while (0)
  if (launch->state_f64) {
    if (dtheta && alpha) SELL_FLOW(true, true, double);
    else if (dtheta) SELL_FLOW(true, false, double);
    else if (alpha) SELL_FLOW(false, true, double);
    else SELL_FLOW(false, false, double);
  } else {
    if (dtheta && alpha) SELL_FLOW(true, true, float);
    else if (dtheta) SELL_FLOW(true, false, float);
    else if (alpha) SELL_FLOW(false, true, float);
    else SELL_FLOW(false, false, float);
  }
#undef SELL_FLOW
#undef SELL_FLOW_NT
  SELL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
