// nfst_tiles.cu -- "tile stream" execution model of the lattice DP (sm_100a).
//
// Same recurrence as nfst_kernels.cu / nfst_sell.cu (the reference's FSAGRUScorer.compute_beta_per_sample,
// src/modules/scorers.py:692-751, with Wh = 0, in log space; arc scores as in WFSTScorer, scorers.py:1663-1687).
// What is different is who decides the order of work: the PACKER (nfst_b200/tiles.py) deals every lattice to the
// nw warps of its block, level by level, and lays states and arcs out so that what one warp does in one level is
// a contiguous run of both.  A warp therefore walks a private list of TILES (about 384 arcs: a few 32-state
// slices, column-major without padding) and fetches each with ONE bulk copy per array -- cp.async.bulk (TMA,
// SASS UBLKCP) issued by lane 0, completing on a per-stage mbarrier (SYNCS) -- into a ring of shared-memory stages
// that runs several tiles ahead.  No lane computes an address for staging, no warp derives its position in the
// lattice: the tile's header (which arrives with the tile) says which states and ring slots it covers.
//
//   * destinations are 16-bit RING SLOTS computed at pack time: beta[dst] is ring[code], no masking, no range
//     test; slot W holds the constant of the last level (beta = 0 / delta = 0); a destination that has left the
//     ring when one of its sources is processed also owns a slot of a small far table behind the ring (W+1 ...,
//     written once, never recycled), so its arcs read ring[code] like all others;
//   * slices are processed from registers by a code path specialised for the slice's column count (1..8);
//     deeper columns (states with 9..32 arcs) loop over the staged tile, states with more arcs are HEAVY:
//     a slice of their own, possibly cut into several tiles, the whole warp striding over the arcs;
//   * pull pass (deepest level first): beta[s] = logsumexp_k(w_k + beta[dst_k]) with per-arc terms as float32
//     offsets from a reference arc, cond[a] = exp(w + beta[dst] - beta[s]); or delta / backpointer with one fp32
//     add per arc and first-maximum-in-label-order ties (bit exact);
//   * flow pass (start level first): post[a] = g * gamma[src] * cond[a]; gamma[dst] += gamma[src] * cond[a] in
//     32-bit fixed point with native integer shared-memory atomics (ATOMS.ADD; a float add there is a
//     compare-and-swap loop, measured 2.3-5x slower, tools/microbench): deterministic, absolute error < 1e-9;
//   * one block barrier per level when nw > 1, none when a warp owns the whole lattice (nw = 1).
#include "nfst_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

int nfst_fail_msg(int code, const char* fmt, ...);  // nfst_kernels.cu (sets the thread's last error)

namespace {

#define TILE_CUDA_OK(expr)                                                                                  \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) return nfst_fail_msg(NFST_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

constexpr int KU = 8;  // arc columns of a slice held in registers
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kFloor = -1.0e30f;
constexpr float kNegInf = -__builtin_huge_valf();
constexpr float kFix = 2147483648.0f;  // 2^31: fixed-point unit of the flow pass
constexpr int FLAG_FAR_IN = 2, FLAG_HEAVY = 4, FLAG_HEAVY_FIRST = 8, FLAG_HEAVY_LAST = 16;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// ---- mbarrier + bulk copy (TMA) ----
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void block_bar() { asm volatile("bar.sync 0;" ::: "memory"); }

extern __shared__ __align__(128) unsigned char tile_smem[];

// loads / stores on 32-bit shared-window addresses: the hot loops keep addresses, not generic pointers
template <typename T>
__device__ __forceinline__ T lds(unsigned a) {
  T v;
  if constexpr (sizeof(T) == 8) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(unsigned a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
template <typename T>
__device__ __forceinline__ void sts(unsigned a, T v) {
  if constexpr (sizeof(T) == 8) asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
  else asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// -DNFST_TILE_DEBUG: range checks on the data-dependent indices; the first violation is recorded (code, three
// values, block, thread) and the access is made harmless.  Read back with nfst_tile_debug_read.
#ifdef NFST_TILE_DEBUG
__device__ int tile_dbg[8];
#define TILE_CHECK(ok, code, a, b, c)                                                        \
  do {                                                                                        \
    if (!(ok) && atomicCAS(&tile_dbg[0], 0, (code)) == 0) {                                   \
      tile_dbg[1] = (a); tile_dbg[2] = (b); tile_dbg[3] = (c);                                \
      tile_dbg[4] = blockIdx.x; tile_dbg[5] = threadIdx.x;                                    \
    }                                                                                         \
  } while (0)
#else
#define TILE_CHECK(ok, code, a, b, c) do { } while (0)
#endif

// launch geometry (host-computed; byte offsets into dynamic shared memory)
struct TP {
  int stages;       // depth of every warp's stage ring
  int cap_bytes;    // stream part of a stage (largest tile of the group, 16-byte multiple)
  int arr_bytes;    // one staged 4-byte-per-arc array (scores / conditionals / labels), 16-byte multiple
  int stage_bytes;  // cap_bytes + staged arrays
  int bar_off, table_off, stage_off;
  int table;        // theta / dtheta table lives in shared memory
  float hist_scale, hist_inv;  // fixed-point unit of the dtheta histogram
};

// online logsumexp pair (m, s): value = m + log(s); branch-free, finite floor instead of -inf
__device__ __forceinline__ void lse_push(float& m, float& s, float v) {
  const float d = v - m;
  const float e = ex2_approx(-fabsf(d) * kLog2e);
  const bool up = d > 0.0f;
  s = up ? fmaf(s, e, 1.0f) : s + e;
  m = up ? v : m;
}
__device__ __forceinline__ void lse_join(float& m, float& s, float m2, float s2) {
  const float d = m2 - m;
  const float e = ex2_approx(-fabsf(d) * kLog2e);
  const bool up = d > 0.0f;
  s = up ? fmaf(s, e, s2) : fmaf(s2, e, s);
  m = up ? m2 : m;
}

// states with n_k..: byte k of the two header words
__device__ __forceinline__ int col_count(const int4& h, int k) {
  return (static_cast<unsigned>(k < 4 ? h.x : h.y) >> (8 * (k & 3))) & 0xff;
}

// everything a segment needs to find its data in the landed stage (flow pass: pointers)
struct Seg {
  const uint16_t* codes;  // ring slots of the tile's arcs (tile-relative arc index)
  const float* vals;      // staged scores (pull) / conditionals (flow), tile-relative
  const int32_t* labs;    // labels of the tile's arcs, tile-relative (staged, or global memory: LG)
  const unsigned char* st;  // the stage (tile header at 0)
  int arc0;               // canonical id of the tile's first arc
};
// the same as 32-bit shared-window addresses (pull pass)
struct SegA {
  unsigned codes, vals, labs, st;
  int arc0;
};

// offset of x from the reference, in float32: (beta - beta_ref) + (w - w_ref), each difference rounded at its own
// small size (float64 ring: the exact float64 sum first)
template <typename RingT>
__device__ __forceinline__ float rel_term(float w, RingT v, float rw, RingT rb) {
  if constexpr (sizeof(RingT) == 4) return (static_cast<float>(v) - static_cast<float>(rb)) + (w - rw);
  else return static_cast<float>((static_cast<double>(w) + static_cast<double>(v)) - (static_cast<double>(rw) + static_cast<double>(rb)));
}

// =====================================================================================
// pull pass
// =====================================================================================
// byte k (0..7) of a segment header's column counts: one PRMT
__device__ __forceinline__ int nk_byte(const int4& h, int k) {
  return static_cast<int>(__byte_perm(static_cast<unsigned>(k < 4 ? h.x : h.y), 0u, 0x4440u | static_cast<unsigned>(k & 3)));
}
// block barriers, `n` of them (n > 0), in a loop the compiler cannot mistake for divergent
__device__ __forceinline__ void block_bars(int n) {
  asm volatile("{\n\t.reg .pred p;\n\tLB_%=:\n\tbar.sync 0;\n\tadd.s32 %0, %0, -1;\n\tsetp.gt.s32 p, %0, 0;\n\t@p bra.uni LB_%=;\n\t}" : "+r"(n)::"memory");
}

// OT: type of beta[] / logz[] in global memory; RT: type of the DP ring (float32 for lattices of at most 96 levels even
// when the batch's state vectors are float64: their values are what a float32 launch computes, stored as doubles)
template <bool TROP, bool SC, bool TH, typename OT, typename RT = OT>
struct PullCtx {
  using RingT = typename std::conditional<TROP, float, RT>::type;
  unsigned ring_s;  // shared-window address of the ring
  const float* th;
  OT* beta;
  float* cond;
  float* delta;
  int32_t* __restrict__ backptr;
  int lane;
#ifdef NFST_TILE_DEBUG
  int W, ring_total, a_lo, a_hi, s_lo, s_hi;  // the lattice's slot / arc / state ranges
#endif

  __device__ __forceinline__ RingT ring_at(unsigned code, int where) const {
#ifdef NFST_TILE_DEBUG
    TILE_CHECK(static_cast<int>(code) < ring_total, 1, static_cast<int>(code), ring_total, where);
    if (static_cast<int>(code) >= ring_total) code = W;
#endif
    return lds<RingT>(ring_s + code * static_cast<unsigned>(sizeof(RingT)));
  }
  __device__ __forceinline__ void cond_store(float* p, float v, int where) const {
#ifdef NFST_TILE_DEBUG
    const long long a = p - cond;
    TILE_CHECK(a >= a_lo && a < a_hi, 2, static_cast<int>(a), a_hi, where);
    if (a < a_lo || a >= a_hi) return;
#endif
    *p = v;
  }
  // score of the arc at tile-relative position e
  __device__ __forceinline__ float score(const SegA& g, int e, bool on) const {
    float w = SC ? lds<float>(g.vals + 4u * e) : 0.0f;
    if (TH) w += th[on ? lds_s32(g.labs + 4u * e) : 0];  // lanes without an arc read a neighbour's entry: no stale label may index theta
    return w;
  }
  __device__ __forceinline__ RingT value(const SegA& g, int e, int where) const { return ring_at(lds_u16(g.codes + 2u * e), where); }
  // the DP value of state s: ring slot, the far table (far = its slot there, 0: none), global memory
  __device__ __forceinline__ void store_state(int s, int slot, int far, RingT v, int arg) const {
#ifdef NFST_TILE_DEBUG
    const bool ok = s >= s_lo && s < s_hi && slot >= 0 && slot < W && (far == 0 || (far > W && far < ring_total));
    TILE_CHECK(ok, 3, s, slot, far);
    if (!ok) return;
#endif
    sts<RingT>(ring_s + slot * static_cast<unsigned>(sizeof(RingT)), v);
    if (far) sts<RingT>(ring_s + far * static_cast<unsigned>(sizeof(RingT)), v);
    if (!TROP) {
      if (beta) beta[s] = static_cast<OT>(v);
    } else {
      if (delta) delta[s] = static_cast<float>(v);
      backptr[s] = arg;
    }
  }

  // ---- a regular segment with NC register columns (NC = min(largest degree, KU) >= 1) ----
  // The columns are read through three running addresses (ring slots, scores, labels) that advance by the
  // column's entry count; lanes without an arc in a column read a neighbour's entry (always inside the stage: the
  // slot region ends with 32 constant slots, the staged arrays with 32 spare elements) and ignore it.
  template <int NC>
  __device__ __forceinline__ void regular(const SegA& g, const int4 h, int s0, int vslot) const {
    const int arc_rel = h.z & 0xffff, nst = (h.z >> 16) & 0xff, dmax = static_cast<unsigned>(h.z) >> 24;
    const int flags = static_cast<unsigned>(h.w) >> 16;
    const int e0 = arc_rel + lane;
    unsigned pc = g.codes + 2u * e0, pv = g.vals + 4u * e0, pl = g.labs + 4u * e0;
    float wc[NC];
    RingT rv[NC];
    bool on[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int n = nk_byte(h, k);
      on[k] = lane < n;
      float w = SC ? lds<float>(pv) : 0.0f;
      if (TH) w += th[on[k] ? lds_s32(pl) : 0];
      wc[k] = w;
      rv[k] = ring_at(lds_u16(pc), 10 + k);
      pc += 2u * n;
      if (SC) pv += 4u * n;
      if (TH) pl += 4u * n;
    }
    const int e_tail = static_cast<int>(pc - g.codes) >> 1;  // this lane's entry in column KU (when the slice has one)
    const unsigned nk_ext = g.st + (h.w & 0xffff);            // n_8, n_9, ... (only read when dmax > KU)
    // f(on, e, w, v) over the columns k >= KU of this lane (states with 9..32 arcs: a few per level)
    auto tail = [&](auto&& f) {
      int et = e_tail;
      for (int k = KU; k < dmax; ++k) {
        const int n = (lds_s32((nk_ext + k - KU) & ~3u) >> (8 * ((nk_ext + k - KU) & 3u))) & 0xff;
        const bool o = lane < n;
        f(o, et, score(g, et, o), value(g, et, 30 + k));
        et += n;
      }
    };
    if constexpr (!TROP) {
      // reference arc: the state's first (every state with arcs has one); (0, 0) if that one scores -inf
      float rw = wc[0];
      RingT rb = rv[0];
      const bool fin = rw + static_cast<float>(rb) > kFloor;
      rw = fin ? rw : 0.0f;
      rb = fin ? rb : static_cast<RingT>(0);
      float tf[NC];
      float mf = kFloor;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        tf[k] = on[k] ? rel_term<RingT>(wc[k], rv[k], rw, rb) : kFloor;
        mf = fmaxf(mf, tf[k]);  // never NaN: fmaxf drops it
      }
      if (NC == KU && dmax > KU) tail([&](bool o, int, float w, RingT v) { if (o) mf = fmaxf(mf, rel_term<RingT>(w, v, rw, rb)); });
      float ex[NC];
      float sum = 0.0f;
      const float mfl = -mf * kLog2e;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        // idle lanes / -inf arcs: exp(-huge) = 0
        if constexpr (sizeof(RingT) == 4) ex[k] = ex2_approx(fmaf(tf[k], kLog2e, mfl));
        else ex[k] = ex2_approx((tf[k] - mf) * kLog2e);
        sum += ex[k];
      }
      if (NC == KU && dmax > KU)
        tail([&](bool o, int, float w, RingT v) { if (o) sum += ex2_approx((fmaxf(rel_term<RingT>(w, v, rw, rb), kFloor) - mf) * kLog2e); });
      const bool any = mf > kFloor;  // some finite arc (else every term above was exp(0))
      const float inv = any ? rcp_approx(sum) : 0.0f;
      // states without arcs: beta = 1 (scorers.py:720); all arcs -inf: beta = -inf.  Branch-free.
      const float lg = lg2_approx(sum) * kLn2;
      RingT bv;
      if constexpr (sizeof(RingT) == 4) bv = rb + (rw + (mf + lg));
      else bv = rb + (static_cast<double>(rw) + (static_cast<double>(mf) + static_cast<double>(lg)));
      bv = any ? bv : static_cast<RingT>(kNegInf);
      bv = on[0] ? bv : static_cast<RingT>(0);
      if (cond) {
        float* pq = cond + g.arc0 + e0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          if (on[k]) cond_store(pq, ex[k] * inv, 100 + k);
          pq += nk_byte(h, k);
        }
        if (NC == KU && dmax > KU)
          tail([&](bool o, int et, float w, RingT v) {
            if (o) cond_store(cond + g.arc0 + et, ex2_approx((fmaxf(rel_term<RingT>(w, v, rw, rb), kFloor) - mf) * kLog2e) * inv, 200);
          });
      }
      if (lane < nst) store_state(s0 + lane, vslot + lane, 0, bv, 0);
      if (flags & FLAG_FAR_IN) {  // warp-uniform, rare: some state of the segment also lives in the far table
        const unsigned far = lds_u16(nk_ext + (dmax > KU ? 32u : 0u) + 2u * lane);
        if (far) sts<RingT>(ring_s + far * static_cast<unsigned>(sizeof(RingT)), bv);
      }
    } else {
      float best = 0.0f;
      int arg = -1;
      int a = g.arc0 + e0;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const float c = __fadd_rn(wc[k], rv[k]);
        if (on[k] && (arg < 0 || c > best)) {  // strict: the columns of a state follow its labels
          best = c;
          arg = a;
        }
        a += nk_byte(h, k);
      }
      if (NC == KU && dmax > KU)
        tail([&](bool o, int et, float w, RingT v) {
          const float c = __fadd_rn(w, v);
          if (o && c > best) {
            best = c;
            arg = g.arc0 + et;
          }
        });
      if (lane < nst) store_state(s0 + lane, vslot + lane, 0, best, arg);  // states without arcs: delta = 0, backpointer -1
      if (flags & FLAG_FAR_IN) {
        const unsigned far = lds_u16(nk_ext + (dmax > KU ? 32u : 0u) + 2u * lane);
        if (far) sts<float>(ring_s + far * 4u, best);
      }
    }
  }
};

// heavy state carried across the pieces (tiles) of one warp
struct HeavyLog {
  float m, s, rw;
  double rb;  // reference arc (holds either ring precision)
};
struct HeavyTrop {
  float best;
  int arg;
};

template <bool TROP, bool SC, bool TH, typename OT, int NT_MAX, int MINB, typename RT = OT>
__global__ void __launch_bounds__(NT_MAX, MINB)
    tile_pull_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, const TP P,
                     const float* __restrict__ arc_scores, const float* __restrict__ theta, OT* beta, OT* __restrict__ logz,
                     float* cond, float* delta, int32_t* __restrict__ backptr, float* __restrict__ vit_score) {
  using Ctx = PullCtx<TROP, SC, TH, OT, RT>;
  using RingT = typename Ctx::RingT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int4 info = __ldg(reinterpret_cast<const int4*>(L.tile_lat_info) + b);
  const int s_base = L.state_off[b];
  const int a_base = info.w;
  const int W = info.y;
  const int n_levels = L.level_off[b + 1] - L.level_off[b] - 1;
  Ctx c;
  c.ring_s = smem_u32(tile_smem);
  c.th = theta;
  c.beta = beta; c.cond = cond; c.delta = delta; c.backptr = backptr;
  c.lane = lane;
#ifdef NFST_TILE_DEBUG
  c.W = W; c.ring_total = info.z; c.a_lo = a_base; c.a_hi = L.n_arcs; c.s_lo = s_base; c.s_hi = L.state_off[b + 1];
#endif
  if (TH && P.table) {
    float* sth = reinterpret_cast<float*>(tile_smem + P.table_off);
    for (int i = tid; i < L.vocab; i += blockDim.x) sth[i] = theta[i];
    c.th = sth;
  }
  if (tid == 0) reinterpret_cast<RingT*>(tile_smem)[W] = static_cast<RingT>(0);  // the last level: beta = 1 (scorers.py:720), delta = 0
  const int D = P.stages;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(tile_smem + P.bar_off) + warp * D;
  const unsigned bars_s = smem_u32(bars);
  if (lane == 0)
    for (int d = 0; d < D; ++d) mbar_init(bars_s + 8 * d, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  unsigned char* const my_stage = tile_smem + P.stage_off + static_cast<size_t>(warp) * D * P.stage_bytes;
  const unsigned stage_s = smem_u32(my_stage);
  const int lw = info.x + warp;
  const int t_lo = L.tile_lw_off[lw], t_hi = L.tile_lw_off[lw + 1];
  const int n_t = t_hi - t_lo;
  const int4* __restrict__ tab = reinterpret_cast<const int4*>(L.tile_tab);
  const unsigned char* __restrict__ stream = L.tile_stream;
  const int32_t* __restrict__ label_out = L.label_out;
  // lane 0: one expect_tx + one bulk copy per array; the tile's scores / labels are the 16-byte-aligned superset
  auto issue = [&](int d, const int4 e) {
    const unsigned st = stage_s + d * P.stage_bytes, bar = bars_s + 8 * d;
    const int arc0 = e.x, n_arcs = e.z & 0xffff;
    const unsigned sbytes = (static_cast<unsigned>(e.w) >> 16) << 4;
    const int a_lo = arc0 & ~3;
    const unsigned abytes = n_arcs ? static_cast<unsigned>(((arc0 + n_arcs + 3) & ~3) - a_lo) << 2 : 0u;
    mbar_expect_tx(bar, sbytes + (SC ? abytes : 0u) + (TH ? abytes : 0u));
    bulk_g2s(st, stream + static_cast<size_t>(e.y) * 16, sbytes, bar);
    if (abytes) {
      if (SC) bulk_g2s(st + P.cap_bytes, arc_scores + a_lo, abytes, bar);
      if (TH) bulk_g2s(st + P.cap_bytes + (SC ? P.arr_bytes : 0), label_out + a_lo, abytes, bar);
    }
  };
  // the pull pass walks the warp's tiles from the deepest level up: i-th tile = t_hi - 1 - i
  int4 nxt = make_int4(0, 0, 0, 0);
  if (lane == 0) {
    for (int i = 0; i < D && i < n_t; ++i) issue(i, __ldg(tab + (t_hi - 1 - i)));
    if (D < n_t) nxt = __ldg(tab + (t_hi - 1 - D));
  }
  HeavyLog hl = {kFloor, 0.0f, 0.0f, 0.0};
  HeavyTrop ht = {0.0f, -1};
  int cur = n_levels - 1;
  int d = 0;
  unsigned phase = 0;
#pragma unroll 1
  for (int i = 0; i < n_t; ++i) {
    mbar_wait(bars_s + 8 * d, phase);
    const unsigned char* const st = my_stage + d * P.stage_bytes;
    const int4 hd = *reinterpret_cast<const int4*>(st);
    const int level = hd.w & 0xffff, nseg = static_cast<unsigned>(hd.w) >> 16;
    TILE_CHECK(level <= cur && nseg >= 1 && nseg <= 64 && 16 + 16 * nseg <= P.cap_bytes, 5, level, cur, nseg);
    // every deeper level is complete before this tile reads the ring (one block barrier per level)
    if (cur > level) {
      if (nw > 1) block_bars(cur - level);
      else __syncwarp();
      cur = level;
    }
    SegA g;
    g.st = stage_s + d * P.stage_bytes;
    g.arc0 = hd.y + a_base;
    g.codes = g.st + (static_cast<unsigned>(hd.z) >> 16);
    const unsigned shift = 4u * (g.arc0 & 3);
    g.vals = g.st + P.cap_bytes + shift;
    g.labs = g.st + P.cap_bytes + (SC ? P.arr_bytes : 0) + shift;
    int s0 = hd.x + s_base;
    int vslot = hd.z & 0xffff;
    int4 h = *reinterpret_cast<const int4*>(st + 16);
#pragma unroll 1
    for (int sg = 0; sg < nseg; ++sg) {
      const int4 hn = *reinterpret_cast<const int4*>(st + 32 + 16 * sg);  // next header (or whatever follows the last)
      const int flags = static_cast<unsigned>(h.w) >> 16;
      const int dmax = static_cast<unsigned>(h.z) >> 24;
      if (flags & FLAG_HEAVY) {
        // ---- a piece of a heavy state: the warp strides over its arcs.  The pull pass walks a warp's tiles
        // backwards: the LAST piece of the state comes first, the FIRST one ends it ----
        const int n = h.w & 0xffff;        // arcs of this piece
        const int far_slot = h.z & 0xffff;  // the state's slot in the far table (0: none)
        if (!TROP) {
          if (flags & FLAG_HEAVY_LAST) {
            float w0 = c.score(g, 0, true);
            RingT v0 = c.value(g, 0, 50);
            if (!(w0 + static_cast<float>(v0) > kFloor)) { w0 = 0.0f; v0 = static_cast<RingT>(0); }
            hl.m = kFloor; hl.s = 0.0f; hl.rw = w0; hl.rb = static_cast<double>(v0);
          }
          const RingT rb = static_cast<RingT>(hl.rb);
#pragma unroll 2
          for (int e = lane; e < n; e += 32) {
            const float w = c.score(g, e, true);
            const RingT v = c.value(g, e, 51);
            const float t = fmaxf(rel_term<RingT>(w, v, hl.rw, rb), kFloor);
            lse_push(hl.m, hl.s, t);
            if (cond) c.cond_store(cond + g.arc0 + e, t, 300);  // provisional: the offset; rescaled below once beta is known
          }
          if (flags & FLAG_HEAVY_FIRST) {
            float m = hl.m, s = hl.s;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
              const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
              lse_join(m, s, m2, s2);
            }
            const bool finite = m > kFloor && s > 0.0f;
            RingT bv = static_cast<RingT>(kNegInf);
            if (finite) {
              const float lg = lg2_approx(s) * kLn2;
              if constexpr (sizeof(RingT) == 4) bv = rb + (hl.rw + (m + lg));
              else bv = rb + (static_cast<double>(hl.rw) + (static_cast<double>(m) + static_cast<double>(lg)));
            }
            if (lane == 0) c.store_state(s0, vslot, far_slot, bv, 0);
            if (cond) {
              // the state's arcs start with this piece; pieces are multiples of 32 arcs, so every lane re-reads
              // exactly what it wrote itself
              const int total = h.x;
              const float inv = finite ? rcp_approx(s) : 0.0f;
              for (int e = lane; e < total; e += 32) {
                const float t = cond[g.arc0 + e];
                c.cond_store(cond + g.arc0 + e, finite ? ex2_approx((t - m) * kLog2e) * inv : 0.0f, 400);
              }
            }
          }
        } else {
          if (flags & FLAG_HEAVY_LAST) { ht.best = 0.0f; ht.arg = -1; }
#pragma unroll 2
          for (int e = lane; e < n; e += 32) {
            const float w = c.score(g, e, true);
            const RingT v = c.value(g, e, 52);
            const float cnd = __fadd_rn(w, v);
            // later pieces come first: among equal candidates the smaller arc id (= smaller label) wins
            if (ht.arg < 0 || cnd > ht.best || (cnd == ht.best && g.arc0 + e < ht.arg)) {
              ht.best = cnd;
              ht.arg = g.arc0 + e;
            }
          }
          if (flags & FLAG_HEAVY_FIRST) {
            float best = ht.best;
            int arg = ht.arg;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ob = __shfl_xor_sync(0xffffffffu, best, o);
              const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
              if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) {
                best = ob;
                arg = oa;
              }
            }
            if (lane == 0) c.store_state(s0, vslot, far_slot, best, arg);
          }
        }
      } else if (dmax <= 4) {
        if (dmax <= 2) {
          if (dmax == 2) c.template regular<2>(g, h, s0, vslot);
          else if (dmax == 1) c.template regular<1>(g, h, s0, vslot);
          else {  // a slice of arc-less states (dead ends; the last level)
            int far = 0;
            if (flags & FLAG_FAR_IN) far = static_cast<int>(lds_u16(g.st + (h.w & 0xffff) + 2u * lane));
            if (lane < ((h.z >> 16) & 0xff)) c.store_state(s0 + lane, vslot + lane, far, static_cast<RingT>(0), -1);
          }
        } else if (dmax == 3) c.template regular<3>(g, h, s0, vslot);
        else c.template regular<4>(g, h, s0, vslot);
      } else if (dmax <= 6) {
        if (dmax == 5) c.template regular<5>(g, h, s0, vslot);
        else c.template regular<6>(g, h, s0, vslot);
      } else if (dmax == 7) c.template regular<7>(g, h, s0, vslot);
      else c.template regular<8>(g, h, s0, vslot);
      h = hn;
      s0 += 32;
      vslot += 32;
      if (vslot >= W) vslot -= W;
    }
    __syncwarp();  // every lane is done with the stage: refill it
    if (level == 0 && lane == 0) {  // the start state is the lattice's first: level 0, ring slot 0
      if (!TROP && logz) logz[b] = static_cast<OT>(lds<RingT>(c.ring_s));
      if (TROP && vit_score) vit_score[b] = static_cast<float>(lds<RingT>(c.ring_s));
    }
    if (lane == 0 && i + D < n_t) {
      issue(d, nxt);
      if (i + D + 1 < n_t) nxt = __ldg(tab + (t_hi - 1 - (i + D + 1)));
    }
    if (++d == D) {
      d = 0;
      phase ^= 1;
    }
  }
  // warps that run out of tiles keep the level barriers company
  if (nw > 1 && cur > 0) block_bars(cur);
}

// =====================================================================================
// flow pass
// =====================================================================================
// GT: the fixed-point accumulator of gamma -- unsigned (2^-31 units: absolute error of a posterior < 1e-9) or
// unsigned long long (2^-62 units: what is left is the float32 rounding of every contribution, ~6e-8 relative)
// LG: the labels of the dtheta histogram are not staged but read from global memory (coalesced: canonical arc order
// is column-major) after a bulk L2 prefetch issued with the tile -- the fallback for launches whose DP ring leaves no
// room for a second staged array (1M-arc lattices); measured slower than staging wherever both fit
template <bool DTH, bool POST, typename GT, int NT_MAX, int MINB, bool LG = false>
__global__ void __launch_bounds__(NT_MAX, MINB)
    tile_flow_kernel(const nfst_packed_lattices_t L, const int32_t* __restrict__ ids, const TP P, const float* cond,
                     const float* __restrict__ grad_logz, float* post, float* __restrict__ dtheta) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int b = ids ? ids[blockIdx.x] : blockIdx.x;
  const int4 info = __ldg(reinterpret_cast<const int4*>(L.tile_lat_info) + b);
  const int s_base = L.state_off[b];
  const int a_base = info.w;
  const int n_levels = L.level_off[b + 1] - L.level_off[b] - 1;
  const int W = info.y, ring_total = info.z;  // ring, the dump slot W (arcs into the last level), the far table
  constexpr bool WIDE = sizeof(GT) == 8;
  constexpr float kUnit = WIDE ? 4611686018427387904.0f : kFix;  // 2^62 : 2^31
  GT* const ring = reinterpret_cast<GT*>(tile_smem);
  unsigned* hist = nullptr;
  for (int i = tid; i < ring_total; i += blockDim.x) ring[i] = 0;
  if (DTH && P.table) {
    hist = reinterpret_cast<unsigned*>(tile_smem + P.table_off);
    for (int i = tid; i < L.vocab; i += blockDim.x) hist[i] = 0u;
  }
  const int D = P.stages;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(tile_smem + P.bar_off) + warp * D;
  const unsigned bars_s = smem_u32(bars);
  if (lane == 0)
    for (int d = 0; d < D; ++d) mbar_init(bars_s + 8 * d, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid == 0) ring[0] = static_cast<GT>(1) << (WIDE ? 62 : 31);  // gamma[start] = 1: the start state is state 0 / slot 0 of its lattice
  __syncthreads();
  const float gl = grad_logz ? grad_logz[b] : 1.0f;
  const float unfix = gl * (1.0f / kUnit);  // fixed-point gamma -> gradient-scaled posterior mass

  unsigned char* const my_stage = tile_smem + P.stage_off + static_cast<size_t>(warp) * D * P.stage_bytes;
  const unsigned stage_s = smem_u32(my_stage);
  const int lw = info.x + warp;
  const int t_lo = L.tile_lw_off[lw], t_hi = L.tile_lw_off[lw + 1];
  const int n_t = t_hi - t_lo;
  const int4* __restrict__ tab = reinterpret_cast<const int4*>(L.tile_tab);
  const unsigned char* __restrict__ stream = L.tile_stream;
  const int32_t* __restrict__ label_out = L.label_out;
  auto issue = [&](int d, const int4 e) {
    const unsigned st = stage_s + d * P.stage_bytes, bar = bars_s + 8 * d;
    const int arc0 = e.x, n_arcs = e.z & 0xffff;
    const unsigned sbytes = (static_cast<unsigned>(e.w) >> 16) << 4;
    const int a_lo = arc0 & ~3;
    const unsigned abytes = n_arcs ? static_cast<unsigned>(((arc0 + n_arcs + 3) & ~3) - a_lo) << 2 : 0u;
    mbar_expect_tx(bar, sbytes + abytes + (DTH && !LG ? abytes : 0u));
    bulk_g2s(st, stream + static_cast<size_t>(e.y) * 16, sbytes, bar);
    if (abytes) {
      bulk_g2s(st + P.cap_bytes, cond + a_lo, abytes, bar);
      if (DTH && !LG) bulk_g2s(st + P.cap_bytes + P.arr_bytes, label_out + a_lo, abytes, bar);
      if (DTH && LG) bulk_prefetch_l2(label_out + a_lo, abytes);
    }
  };
  int4 nxt = make_int4(0, 0, 0, 0);
  if (lane == 0) {
    for (int i = 0; i < D && i < n_t; ++i) issue(i, __ldg(tab + t_lo + i));
    if (D < n_t) nxt = __ldg(tab + t_lo + D);
  }
  // one arc: posterior out, flow into the destination's slot (ring, far table, or the dump slot of the last level)
  auto push = [&](const Seg& g, int e, float gam) {
    const float cd = g.vals[e];
    const unsigned code = g.codes[e];
    const float pf = gam * cd;  // in fixed-point units (gam = gamma * 2^31 or 2^62)
    if (POST) post[g.arc0 + e] = pf * unfix;
    TILE_CHECK(static_cast<int>(code) < ring_total, 6, static_cast<int>(code), ring_total, e);
    if constexpr (WIDE) {
      // 64-bit add as two native 32-bit shared-memory atomics (a 64-bit atomicAdd there is a compare-and-swap
      // loop): exactly the adds whose low word wraps carry one into the high word, so the sum is exact whatever
      // the order; the words are read together only after the level barrier
      const unsigned long long x = __float2ull_rn(pf);
      const unsigned xl = static_cast<unsigned>(x);
      unsigned* const w = reinterpret_cast<unsigned*>(&ring[code]);
      const unsigned old = atomicAdd(w, xl);
      atomicAdd(w + 1, static_cast<unsigned>(x >> 32) + (old + xl < old ? 1u : 0u));
    } else {
      atomicAdd(&ring[code], __float2uint_rn(pf));
    }
    if (DTH) {
      const float pr = pf * (1.0f / kUnit);
      const int lab = LG ? __ldg(g.labs + e) : g.labs[e];
      if (hist) atomicAdd(&hist[lab], __float2uint_rn(pr * P.hist_scale));
      else atomicAdd(dtheta + lab, pr * gl);
    }
  };
  float heavy_gam = 0.0f;
  int cur = 0;
  int d = 0;
  unsigned phase = 0;
#pragma unroll 1
  for (int i = 0; i < n_t; ++i) {
    mbar_wait(bars_s + 8 * d, phase);
    const unsigned char* const st = my_stage + d * P.stage_bytes;
    const int4 hd = *reinterpret_cast<const int4*>(st);
    const int level = hd.w & 0xffff, nseg = static_cast<unsigned>(hd.w) >> 16;
    while (cur < level) {
      if (nw > 1) block_bar();
      else __syncwarp();
      ++cur;
    }
    Seg g;
    g.st = st;
    g.arc0 = hd.y + a_base;
    g.codes = reinterpret_cast<const uint16_t*>(st + (static_cast<unsigned>(hd.z) >> 16));
    const int shift = g.arc0 & 3;
    g.vals = reinterpret_cast<const float*>(st + P.cap_bytes) + shift;
    g.labs = LG ? label_out + g.arc0 : reinterpret_cast<const int32_t*>(st + P.cap_bytes + P.arr_bytes) + shift;
    int s0 = hd.x + s_base;
    int vslot = hd.z & 0xffff;
#pragma unroll 1
    for (int sg = 0; sg < nseg; ++sg) {
      const int4 h = *reinterpret_cast<const int4*>(st + 16 + 16 * sg);
      const int flags = static_cast<unsigned>(h.w) >> 16;
      if (flags & FLAG_HEAVY) {
        if (flags & FLAG_HEAVY_FIRST) {
          // every arc into the state comes from a shallower level: gamma is final; free the ring slot
          const int far_slot = h.z & 0xffff;
          heavy_gam = static_cast<float>(ring[vslot] + (far_slot ? ring[far_slot] : static_cast<GT>(0)));
          __syncwarp();
          if (lane == 0) ring[vslot] = 0;
        }
        const int n = h.w & 0xffff;
#pragma unroll 2
        for (int e = lane; e < n; e += 32) push(g, e, heavy_gam);
      } else {
        const int arc_rel = h.z & 0xffff, nst = (h.z >> 16) & 0xff, dmax = static_cast<unsigned>(h.z) >> 24;
        float gam = 0.0f;
        if (lane < nst) {
          GT gfix = ring[vslot + lane];
          ring[vslot + lane] = 0;
          if (flags & FLAG_FAR_IN) {  // flow that arrived through the far table
            const int far_slot = reinterpret_cast<const uint16_t*>(st + (h.w & 0xffff) + (dmax > KU ? 32 : 0))[lane];
            if (far_slot) gfix += ring[far_slot];
          }
          gam = static_cast<float>(gfix);
        }
        int e = arc_rel + lane;
        auto cols = [&](auto nc) {
          constexpr int NC = decltype(nc)::value;
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const int n = col_count(h, k);
            if (lane < n) push(g, e, gam);
            e += n;
          }
        };
        switch (dmax < KU ? dmax : KU) {
          case 0: break;
          case 1: cols(std::integral_constant<int, 1>{}); break;
          case 2: cols(std::integral_constant<int, 2>{}); break;
          case 3: cols(std::integral_constant<int, 3>{}); break;
          case 4: cols(std::integral_constant<int, 4>{}); break;
          case 5: cols(std::integral_constant<int, 5>{}); break;
          case 6: cols(std::integral_constant<int, 6>{}); break;
          case 7: cols(std::integral_constant<int, 7>{}); break;
          default: cols(std::integral_constant<int, 8>{}); break;
        }
        if (dmax > KU) {
          const unsigned char* nk_ext = st + (h.w & 0xffff);
          for (int k = KU; k < dmax; ++k) {
            const int n = nk_ext[k - KU];
            if (lane < n) push(g, e, gam);
            e += n;
          }
        }
      }
      s0 += 32;
      vslot += 32;
      if (vslot >= W) vslot -= W;
    }
    __syncwarp();
    if (lane == 0 && i + D < n_t) {
      issue(d, nxt);
      if (i + D + 1 < n_t) nxt = __ldg(tab + t_lo + i + D + 1);
    }
    if (++d == D) {
      d = 0;
      phase ^= 1;
    }
  }
  if (nw > 1)
    while (cur < n_levels - 1) {
      block_bar();
      ++cur;
    }
  if (DTH && hist) {
    __syncthreads();
    const float sc = P.hist_inv * gl;
    for (int i = tid; i < L.vocab; i += blockDim.x) {
      const unsigned v = hist[i];
      if (v) atomicAdd(dtheta + i, static_cast<float>(v) * sc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Geometry {
  TP p;
  size_t smem;
};

// shared memory of a launch: ring (+ constant slot), mbarriers, theta / dtheta table, the warps' stage rings
Geometry geometry(const nfst_launch_t* launch, int vocab, int ring_elem_bytes, int n_arrays, bool table, int stages) {
  Geometry g;
  const int nw = launch->block_threads / 32;
  size_t o = round_up(static_cast<size_t>(launch->tile_ring + 32) * ring_elem_bytes, 16);
  g.p.bar_off = static_cast<int>(o);
  o += round_up(static_cast<size_t>(nw) * stages * 8, 16);
  g.p.table = table ? 1 : 0;
  g.p.table_off = static_cast<int>(o);
  if (table) o += round_up(static_cast<size_t>(vocab) * 4, 16);
  o = round_up(o, 128);
  g.p.stage_off = static_cast<int>(o);
  g.p.stages = stages;
  g.p.cap_bytes = static_cast<int>(round_up(launch->tile_cap_bytes, 16));
  // a staged array holds the 16-byte-aligned superset of the tile's range (up to 6 more elements), and lanes without
  // an arc in a column read up to 31 elements past the tile's last arc (their values are ignored)
  g.p.arr_bytes = static_cast<int>(round_up(static_cast<size_t>(launch->tile_cap_arcs + 8 + 32) * 4, 16));
  g.p.stage_bytes = g.p.cap_bytes + n_arrays * g.p.arr_bytes;
  g.p.hist_scale = g.p.hist_inv = 1.0f;
  g.smem = o + static_cast<size_t>(nw) * stages * g.p.stage_bytes;
  return g;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// stage depth: as deep as the shared memory left per block allows when every lattice of the launch is to be
// resident at once (1024 lattices on 148 SMs: 7 blocks per SM), between 2 and 4 (NFST_TILE_STAGES overrides)
Geometry pick_geometry(const nfst_launch_t* launch, int vocab, int ring_elem_bytes, int n_arrays, bool table) {
  static const int forced = getenv("NFST_TILE_STAGES") ? atoi(getenv("NFST_TILE_STAGES")) : 0;
  int stages = launch->tile_stages > 0 ? launch->tile_stages : forced;
  if (stages > 0) return geometry(launch, vocab, ring_elem_bytes, n_arrays, table, stages < 2 ? 2 : stages > 8 ? 8 : stages);
  int want = (launch->n_ids + sm_count() - 1) / sm_count();
  const int by_threads = 2048 / launch->block_threads;
  if (want > by_threads) want = by_threads;
  if (want > 32) want = 32;
  // ... but no more blocks than fit an SM with the shallowest ring of stages (a large DP ring: one or two blocks)
  const Geometry g2 = geometry(launch, vocab, ring_elem_bytes, n_arrays, table, 2);
  const int fit = static_cast<int>((227u * 1024u) / (g2.smem + 1024u));
  if (want > fit) want = fit;
  if (want < 1) want = 1;
  const size_t budget = (227u * 1024u) / want - 1024u;
  for (int s = 4; s > 2; --s) {
    Geometry g = geometry(launch, vocab, ring_elem_bytes, n_arrays, table, s);
    if (g.smem <= budget) return g;
  }
  return g2;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename K>
int prepare(K kernel, size_t smem) {
  if (smem > 48 * 1024) {
    TILE_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  }
  return 0;
}

int check_launch(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch) {
  if (!lat || !launch) return nfst_fail_msg(NFST_ERR_BAD_ARG, "null argument");
  if (!launch->tiles) return nfst_fail_msg(NFST_ERR_BAD_ARG, "launch group is not a tile-stream group");
  if (!lat->tile_stream || !lat->tile_tab || !lat->tile_lw_off || !lat->tile_lat_info)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "packed lattices carry no tile-stream arrays");
  const int nt = launch->block_threads;
  if (nt < 32 || nt > 1024 || (nt & (nt - 1))) return nfst_fail_msg(NFST_ERR_BAD_ARG, "block_threads must be a power of two in 32..1024");
  if (launch->tile_ring < 33 || launch->tile_ring > 65535) return nfst_fail_msg(NFST_ERR_BAD_ARG, "tile_ring must be in 33..65535 slots");
  if (launch->tile_cap_bytes <= 0 || launch->tile_cap_arcs < 0) return nfst_fail_msg(NFST_ERR_BAD_ARG, "tile_cap_bytes / tile_cap_arcs are required");
  if (!aligned16(lat->tile_stream) || !aligned16(lat->tile_tab) || !aligned16(lat->tile_lat_info))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "tile_stream, tile_tab and tile_lat_info must be 16-byte aligned");
  return 0;
}

// register budgets: blocks of up to 64 threads may use 128 registers, 128-thread blocks 72 (7 blocks per SM),
// 256-thread blocks 80 (3 per SM; at 64 the log pull pass spills), 512-thread blocks 128 (one per SM: lattices
// whose DP ring fills most of the shared memory), 1024-thread blocks 64
#define TILE_BY_BLOCK(CALL)                              \
  do {                                                   \
    if (launch->block_threads <= 64) CALL(64, 7);        \
    else if (launch->block_threads <= 128) CALL(128, 7); \
    else if (launch->block_threads <= 256) CALL(256, 3); \
    else if (launch->block_threads <= 512) CALL(512, 1); \
    else CALL(1024, 1);                                  \
  } while (0)

template <bool TROP, typename OT, typename RT = OT>
int launch_pull(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* sc, OT* beta, OT* logz,
                float* cond, float* delta, int32_t* backptr, float* vit, cudaStream_t stream) {
  const bool has_sc = sc->arc_scores != nullptr, has_th = sc->theta != nullptr;
  const bool table = has_th && lat->vocab <= NFST_THETA_SMEM_MAX;
  const Geometry g = pick_geometry(launch, lat->vocab, TROP ? 4 : static_cast<int>(sizeof(RT)), (has_sc ? 1 : 0) + (has_th ? 1 : 0), table);
  if (g.smem > 227 * 1024) return nfst_fail_msg(NFST_ERR_TOO_LARGE, "tile-stream launch needs %zu bytes of shared memory", g.smem);
  if ((has_sc && !aligned16(sc->arc_scores)) || (has_th && !aligned16(lat->label_out)))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "arc_scores and label_out must be 16-byte aligned (they are fetched with bulk copies)");
#define PULL_NT(NTv, MINBv)                                                                                     \
  do {                                                                                                          \
    if (has_sc && has_th) {                                                                                     \
      auto k = tile_pull_kernel<TROP, true, true, OT, NTv, MINBv, RT>;                                            \
      if (int rc = prepare(k, g.smem)) return rc;                                                               \
      k<<<launch->n_ids, launch->block_threads, g.smem, stream>>>(*lat, launch->lattice_ids, g.p, sc->arc_scores, sc->theta, \
                                                                  beta, logz, cond, delta, backptr, vit);       \
    } else if (has_th) {                                                                                        \
      auto k = tile_pull_kernel<TROP, false, true, OT, NTv, MINBv, RT>;                                            \
      if (int rc = prepare(k, g.smem)) return rc;                                                               \
      k<<<launch->n_ids, launch->block_threads, g.smem, stream>>>(*lat, launch->lattice_ids, g.p, sc->arc_scores, sc->theta, \
                                                                  beta, logz, cond, delta, backptr, vit);       \
    } else {                                                                                                    \
      auto k = tile_pull_kernel<TROP, true, false, OT, NTv, MINBv, RT>;                                            \
      if (int rc = prepare(k, g.smem)) return rc;                                                               \
      k<<<launch->n_ids, launch->block_threads, g.smem, stream>>>(*lat, launch->lattice_ids, g.p, sc->arc_scores, sc->theta, \
                                                                  beta, logz, cond, delta, backptr, vit);       \
    }                                                                                                           \
  } while (0)
  TILE_BY_BLOCK(PULL_NT);
#undef PULL_NT
  TILE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

/* debug builds (-DNFST_TILE_DEBUG): the first range-check violation {code, a, b, c, block, thread, 0, 0}; else zeros */
int nfst_tile_debug_read(int32_t* out8) {
  for (int i = 0; i < 8; ++i) out8[i] = 0;
#ifdef NFST_TILE_DEBUG
  int zero[8] = {0};
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out8, tile_dbg, sizeof(int) * 8) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(tile_dbg, zero, sizeof(zero));
#endif
  return 0;
}

size_t nfst_tile_smem_bytes(const nfst_launch_t* launch, int32_t vocab, int pass, int n_f32_arrays, int with_table, int stages) {
  if (!launch || !launch->tiles) return 0;
  const bool table = with_table && vocab <= NFST_THETA_SMEM_MAX;
  const int elem = (pass == 0 && launch->state_f64 && launch->n_levels > NFST_TILE_F64_LEVELS) || (pass == 1 && launch->tile_flow_bits == 64) ? 8 : 4;
  if (stages > 0) return geometry(launch, vocab, elem, n_f32_arrays, table, stages).smem;
  const size_t need = pick_geometry(launch, vocab, elem, n_f32_arrays, table).smem;
  // the flow pass reads its labels from global memory when the second staged array does not fit
  if (pass == 1 && n_f32_arrays == 2 && need > 227 * 1024) return pick_geometry(launch, vocab, elem, 1, table).smem;
  return need;
}

int nfst_tile_pull_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores, void* beta,
                       void* logz_bwd, float* cond, float* delta, int32_t* backptr, float* vit_score, void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!scores || (!scores->arc_scores && !scores->theta)) return nfst_fail_msg(NFST_ERR_BAD_ARG, "arc_scores or theta is required");
  if (launch->n_ids == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const bool logs = beta || logz_bwd || cond;
  const bool trop = delta || backptr || vit_score;
  if (trop && !backptr) return nfst_fail_msg(NFST_ERR_BAD_ARG, "the tropical pass needs backptr[S]");
  if (logs) {
    // float64 state vectors, but a group whose lattices have at most NFST_TILE_F64_LEVELS levels keeps the float32 ring
    // it was packed for: the values a float32 launch computes, written as doubles (a batch is float64 as soon as ONE of
    // its lattices is deep; its shallow, wide lattices must not pay -- or be refused -- for that)
    const int rc = launch->state_f64 && launch->n_levels <= NFST_TILE_F64_LEVELS
                       ? launch_pull<false, double, float>(lat, launch, scores, static_cast<double*>(beta), static_cast<double*>(logz_bwd),
                                                           cond, nullptr, nullptr, nullptr, stream)
                   : launch->state_f64
                       ? launch_pull<false, double>(lat, launch, scores, static_cast<double*>(beta), static_cast<double*>(logz_bwd), cond,
                                                    nullptr, nullptr, nullptr, stream)
                       : launch_pull<false, float>(lat, launch, scores, static_cast<float*>(beta), static_cast<float*>(logz_bwd), cond,
                                                   nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
  }
  if (trop) {
    if (int rc = launch_pull<true, float>(lat, launch, scores, nullptr, nullptr, nullptr, delta, backptr, vit_score, stream)) return rc;
  }
  return 0;
}

int nfst_tile_flow_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const float* cond, const float* grad_logz,
                       float* post, float* dtheta, void* cuda_stream) {
  if (int rc = check_launch(lat, launch)) return rc;
  if (!cond || (!post && !dtheta)) return nfst_fail_msg(NFST_ERR_BAD_ARG, "cond[A] and one of post[A] / dtheta[V] are required");
  if (launch->n_ids == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const bool table = dtheta && lat->vocab <= NFST_THETA_SMEM_MAX;
  const bool wide = launch->tile_flow_bits == 64;
  if (launch->tile_flow_bits != 0 && launch->tile_flow_bits != 32 && !wide)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "tile_flow_bits must be 0, 32 or 64");
  Geometry g = pick_geometry(launch, lat->vocab, wide ? 8 : 4, dtheta ? 2 : 1, table);
  // no room for the staged labels next to a large DP ring: read them from global memory (LG kernels)
  const bool lg = dtheta && g.smem > 227 * 1024;
  if (lg) g = pick_geometry(launch, lat->vocab, wide ? 8 : 4, 1, table);
  // dtheta histogram unit: an expected label count is at most the number of levels
  int lv = 1;
  while (lv < launch->n_levels) lv <<= 1;
  g.p.hist_scale = kFix / static_cast<float>(lv);
  g.p.hist_inv = static_cast<float>(lv) / kFix;
  if (g.smem > 227 * 1024) return nfst_fail_msg(NFST_ERR_TOO_LARGE, "tile-stream launch needs %zu bytes of shared memory", g.smem);
  if (!aligned16(cond) || (dtheta && !aligned16(lat->label_out)))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "cond and label_out must be 16-byte aligned (they are fetched with bulk copies)");
#define FLOW_GO(DTHv, POSTv, GTv, NTv, MINBv, LGv)                                                                        \
  do {                                                                                                                    \
    auto k = tile_flow_kernel<DTHv, POSTv, GTv, NTv, MINBv, LGv>;                                                         \
    if (int rc = prepare(k, g.smem)) return rc;                                                                           \
    k<<<launch->n_ids, launch->block_threads, g.smem, stream>>>(*lat, launch->lattice_ids, g.p, cond, grad_logz, post, dtheta); \
  } while (0)
#define FLOW_DTH(POSTv, GTv, NTv, MINBv)                      \
  do {                                                        \
    if (lg) FLOW_GO(true, POSTv, GTv, NTv, MINBv, true);      \
    else FLOW_GO(true, POSTv, GTv, NTv, MINBv, false);        \
  } while (0)
#define FLOW_NT(NTv, MINBv)                                                                  \
  do {                                                                                       \
    if (wide) {                                                                              \
      if (dtheta && post) FLOW_DTH(true, unsigned long long, NTv, MINBv);                    \
      else if (dtheta) FLOW_DTH(false, unsigned long long, NTv, MINBv);                      \
      else FLOW_GO(false, true, unsigned long long, NTv, MINBv, false);                      \
    } else {                                                                                 \
      if (dtheta && post) FLOW_DTH(true, unsigned, NTv, MINBv);                              \
      else if (dtheta) FLOW_DTH(false, unsigned, NTv, MINBv); /* no per-arc output */        \
      else FLOW_GO(false, true, unsigned, NTv, MINBv, false);                                \
    }                                                                                        \
  } while (0)
  TILE_BY_BLOCK(FLOW_NT);
#undef FLOW_NT
#undef FLOW_DTH
#undef FLOW_GO
  TILE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
