// nfst_pack.cu -- device packer for the lattices nFST itself builds (sm_100a).
//
// What it replaces: the per-call Python graph build at the top of FSAGRUScorer.compute_beta_per_sample /
// compute_beta_parallel (src/modules/scorers.py:704-716, :764-776: scan the dense transition table, keep cell
// (i, j) iff t != 0 and t != i, build adjacency lists) and the host-side packer of this repository
// (nfst_b200/pack.py, ~10^3 eager tensor ops) for the lattices that fit one SM's shared memory -- the size
// nFST's composed mark lattices have (hundreds to a few thousand states).
//
// One thread block per lattice, three launches for the whole batch:
//   pack_level_kernel   states the start state reaches (one pass over the raw arcs for every state's arc range, then a
//                       breadth-first search over the reachable states' arcs); the arcs out of those states -> shared
//                       memory; longest distance from the start by relaxation sweeps
//                       (shared-memory atomicMax, until nothing changes; more sweeps than states = a cycle);
//                       states unreachable from the start -- every row collate() padding adds
//                       (util/dataset_reader.py:175-186) -- are trimmed; per-lattice counts
//   pack_scan_kernel    one block: exclusive scans of the counts -> state / arc / level / sink offsets, totals
//   pack_build_kernel   renumber the states by (level, original id), emit CSR by source in (state, label) order --
//                       the reference's scan order, which is what Viterbi's first-label tie rule needs -- and CSR
//                       by destination in (state, canonical id) order, level pointers, sinks, original ids
// The caller reads the totals once (the only host synchronisation of a pack) to size the views it hands out.
#include "nfst_b200.h"

#include <cuda_runtime.h>

#include <cstdint>

int nfst_fail_msg(int code, const char* fmt, ...);  // nfst_kernels.cu (sets the thread's last error)

namespace {

#define PACK_CUDA_OK(expr)                                                                                  \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) return nfst_fail_msg(NFST_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

constexpr int kThreads = 256;
extern __shared__ __align__(16) int pack_smem[];

struct RawArcs {
  const int32_t* src;    // raw source: local id, or global row (src_global)
  const int32_t* dst;    // raw destination, local id
  const int32_t* label;
  const int32_t* state_off;  // [B+1] raw (untrimmed) state offsets
  const int32_t* arc_off;    // [B+1] raw arc offsets (arcs grouped by lattice, sorted by (src, label))
  int src_global;
  int start;  // local id of the start state (the reference's row 0, scorers.py:1005)
};

// exclusive scan of v[0..n) in place over the block (n arbitrary); returns the total.  tmp: kThreads ints.
__device__ int block_excl_scan(int* v, int n, int* tmp) {
  const int tid = threadIdx.x;
  const int per = (n + kThreads - 1) / kThreads;
  const int lo = tid * per, hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += v[i];
  tmp[tid] = sum;
  __syncthreads();
  for (int o = 1; o < kThreads; o <<= 1) {  // Hillis-Steele over the per-thread sums
    const int add = tid >= o ? tmp[tid - o] : 0;
    __syncthreads();
    tmp[tid] += add;
    __syncthreads();
  }
  int run = tid ? tmp[tid - 1] : 0;
  const int total = tmp[kThreads - 1];
  for (int i = lo; i < hi; ++i) {
    const int x = v[i];
    v[i] = run;
    run += x;
  }
  __syncthreads();
  return total;
}

// ---------------------------------------------------------------------------------------------------------
// reachability, kept arcs, levels, counts
// ---------------------------------------------------------------------------------------------------------
// every state's range in the raw arc list (sorted by source) + endpoint validation: grid (chunks, lattices)
__global__ void __launch_bounds__(kThreads)
    pack_ranges_kernel(const RawArcs R, int32_t* __restrict__ first_g, int32_t* __restrict__ end_g, int32_t* __restrict__ totals) {
  if (totals[4]) return;  // e.g. an arc list that overflowed its capacity (nfst_pack_dense)
  const int b = blockIdx.y;
  const int s0 = R.state_off[b], S0 = R.state_off[b + 1] - s0;
  const int a0 = R.arc_off[b], A0 = R.arc_off[b + 1] - a0;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= A0) return;
  const int base = R.src_global ? s0 : 0;
  const int s = R.src[a0 + i] - base, d = R.dst[a0 + i];
  if (s < 0 || s >= S0 || d < 0 || d >= S0) {
    if (atomicCAS(&totals[4], 0, 2) == 0) totals[5] = b;
    return;
  }
  const int sp = i ? R.src[a0 + i - 1] - base : -1;
  if (sp != s) {
    first_g[s0 + s] = i;
    if (sp >= 0 && sp < S0) end_g[s0 + sp] = i;
  }
  if (i == A0 - 1) end_g[s0 + s] = A0;
}

struct Kept {  // the arcs that survive trimming, per lattice at its raw arc offset (workspace)
  int32_t* pack;   // src | dst << 16 (raw local ids)
  int32_t* label;
  int32_t* raw;    // position in the lattice's raw arc list
};

__global__ void __launch_bounds__(kThreads)
    pack_level_kernel(const RawArcs R, int capS, int capA, int32_t* __restrict__ level_g, const int32_t* __restrict__ first_g,
                      const int32_t* __restrict__ end_g, const Kept K, int32_t* __restrict__ stats,
                      int32_t* __restrict__ totals) {
  if (totals[4]) return;  // pack_ranges_kernel met an arc that points outside its lattice
  const int b = blockIdx.x, tid = threadIdx.x;
  const int s0 = R.state_off[b], S0 = R.state_off[b + 1] - s0;
  const int a0 = R.arc_off[b], A0 = R.arc_off[b + 1] - a0;
  int* const level = pack_smem;             // [capS]
  int* const has_out = level + capS;        // [capS]
  int* const lvl_out = has_out + capS;      // [capS] arcs out of / into every level
  int* const lvl_in = lvl_out + capS;
  int* const arcs = lvl_in + capS;          // [capA] kept arcs: src | dst << 16
  __shared__ int changed, n_states, n_sinks, max_level, width_arcs;
  for (int s = tid; s < S0; s += kThreads) {
    level[s] = -1;
    has_out[s] = 0;
    lvl_out[s] = 0;
    lvl_in[s] = 0;
  }
  if (tid == 0) {
    changed = 0; n_states = 0; n_sinks = 0; max_level = 0; width_arcs = 0;
  }
  __syncthreads();
  if (tid == 0 && R.start >= 0 && R.start < S0) level[R.start] = 0;
  __syncthreads();
  // 1. which states the start state reaches.  A collate()-padded table carries V arcs per pad row -- far more than
  //    the lattice's own -- so the raw arcs never enter shared memory: pack_ranges_kernel found every state's arc
  //    range (they are sorted by source) and validated the endpoints; a breadth-first search from the start
  //    touches the arcs of reachable states only (a queue in shared memory, one thread per newly reached state).
  int* const raw_first = lvl_out;  // [S0] first raw arc of a state / one past its last (aliases: re-zeroed below)
  int* const raw_end = lvl_in;
  int* const queue = has_out;
  __shared__ int q_tail;
  for (int s = tid; s < S0; s += kThreads) {
    raw_first[s] = first_g[s0 + s];
    raw_end[s] = end_g[s0 + s];
  }
  __syncthreads();
  if (R.start < 0 || R.start >= S0) {
    if (tid == 0 && atomicCAS(&totals[4], 0, 2) == 0) totals[5] = b;
    return;
  }
  if (tid == 0) {
    queue[0] = R.start;
    q_tail = 1;
  }
  __syncthreads();
  for (int head = 0;;) {
    const int tail = q_tail;  // everything below [head, tail) was reached in the previous round
    __syncthreads();
    if (head >= tail) break;
    for (int q = head + tid; q < tail; q += kThreads) {
      const int s = queue[q];
      for (int i = raw_first[s]; i < raw_end[s]; ++i) {
        const int d = R.dst[a0 + i];
        if (atomicCAS(&level[d], -1, 0) == -1) queue[atomicAdd(&q_tail, 1)] = d;
      }
    }
    head = tail;
    __syncthreads();
  }
  // 2. the kept arcs (source reachable), in raw order: every reachable state's raw range, placed by an exclusive scan
  //    over the states (the raw list is sorted by source) -- shared memory for the level sweeps, workspace for the build
  int* const koff = queue;  // the search is over
  __shared__ int scan_tmp[kThreads];
  for (int s = tid; s < S0; s += kThreads) koff[s] = level[s] >= 0 ? raw_end[s] - raw_first[s] : 0;
  __syncthreads();
  const int Ak = block_excl_scan(koff, S0, scan_tmp);
  if (Ak > A0) {  // the ranges overlap: the raw list was not sorted by source
    if (tid == 0 && atomicCAS(&totals[4], 0, 2) == 0) totals[5] = b;
    return;
  }
  if (Ak > capA) {  // more arcs than the shared memory this launch was sized for: the caller packs this batch elsewhere
    if (tid == 0 && atomicCAS(&totals[4], 0, 4) == 0) totals[5] = b;
    return;
  }
  for (int s = tid; s < S0; s += kThreads) {
    if (level[s] < 0) continue;
    const int first = raw_first[s], n = raw_end[s] - first, at = koff[s];
    for (int q = 0; q < n; ++q) {
      const int d = R.dst[a0 + first + q];
      const int pk = static_cast<int>((static_cast<unsigned>(s) & 0xffffu) | (static_cast<unsigned>(d) << 16));
      arcs[at + q] = pk;
      K.pack[a0 + at + q] = pk;
      K.label[a0 + at + q] = R.label[a0 + first + q];
      K.raw[a0 + at + q] = first + q;
    }
  }
  __syncthreads();
  for (int s = tid; s < S0; s += kThreads) {  // give the aliased arrays back
    has_out[s] = 0;
    lvl_out[s] = 0;
    lvl_in[s] = 0;
  }
  __syncthreads();
  // 3. longest distance from the start over the kept arcs: sweep until a sweep changes nothing
  for (int s = tid; s < S0; s += kThreads) level[s] = (level[s] >= 0) ? (s == R.start ? 0 : -2) : -1;  // -2: reachable, not yet placed
  __syncthreads();
  for (int sweep = 0;; ++sweep) {
    for (int i = tid; i < Ak; i += kThreads) {
      const int a = arcs[i];
      const int ls = level[a & 0xffff];
      if (ls >= 0 && atomicMax(&level[static_cast<unsigned>(a) >> 16], ls + 1) < ls + 1) changed = 1;
    }
    __syncthreads();
    const int again = changed;
    __syncthreads();
    if (!again) break;
    if (tid == 0) changed = 0;
    if (sweep > S0) {  // a path longer than the number of states: a cycle
      if (tid == 0 && atomicCAS(&totals[4], 0, 1) == 0) totals[5] = b;
      return;
    }
    __syncthreads();
  }
  for (int i = tid; i < Ak; i += kThreads) {
    const int a = arcs[i];
    has_out[a & 0xffff] = 1;
    atomicAdd(&lvl_out[level[a & 0xffff]], 1);
    atomicAdd(&lvl_in[level[static_cast<unsigned>(a) >> 16]], 1);
  }
  __syncthreads();
  // widest level, in arcs (either direction): sizes the thread block of the DP kernels
  for (int l = tid; l < S0; l += kThreads) atomicMax(&width_arcs, max(lvl_out[l], lvl_in[l]));
  int cs = 0, ck = 0, ml = 0;
  for (int s = tid; s < S0; s += kThreads) {
    const int l = level[s];
    level_g[s0 + s] = l;
    if (l >= 0) {
      ++cs;
      ck += has_out[s] ? 0 : 1;
      ml = max(ml, l);
    }
  }
  atomicAdd(&n_states, cs);
  atomicAdd(&n_sinks, ck);
  atomicMax(&max_level, ml);
  __syncthreads();
  if (tid == 0) {
    stats[8 * b + 0] = n_states;
    stats[8 * b + 1] = Ak;
    stats[8 * b + 2] = max_level + 1;
    stats[8 * b + 3] = n_sinks;
    stats[8 * b + 4] = width_arcs;
  }
}

// ---------------------------------------------------------------------------------------------------------
// offsets (one block)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    pack_scan_kernel(int B, const int32_t* __restrict__ stats, int32_t* __restrict__ state_off, int32_t* __restrict__ arc_off,
                     int32_t* __restrict__ level_off, int32_t* __restrict__ sink_off, int32_t* __restrict__ n_levels,
                     int32_t* __restrict__ totals) {
  __shared__ long long carry[4];
  __shared__ int tmp[4][kThreads];
  __shared__ int mx[kThreads];
  const int tid = threadIdx.x;
  if (tid < 4) carry[tid] = 0;
  int my_max = 0;
  __syncthreads();
  for (int base = 0; base < B; base += kThreads) {
    const int b = base + tid;
    int v[4] = {0, 0, 0, 0};
    if (b < B) {
      v[0] = stats[8 * b + 0];
      v[1] = stats[8 * b + 1];
      v[2] = stats[8 * b + 2] + 1;  // a lattice owns levels + 1 entries of level_ptr
      v[3] = stats[8 * b + 3];
      n_levels[b] = stats[8 * b + 2];
      my_max = max(my_max, stats[8 * b + 2]);
    }
    for (int q = 0; q < 4; ++q) tmp[q][tid] = v[q];
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1) {
      int add[4];
      for (int q = 0; q < 4; ++q) add[q] = tid >= o ? tmp[q][tid - o] : 0;
      __syncthreads();
      for (int q = 0; q < 4; ++q) tmp[q][tid] += add[q];
      __syncthreads();
    }
    if (b < B) {
      state_off[b] = static_cast<int32_t>(carry[0] + tmp[0][tid] - v[0]);
      arc_off[b] = static_cast<int32_t>(carry[1] + tmp[1][tid] - v[1]);
      level_off[b] = static_cast<int32_t>(carry[2] + tmp[2][tid] - v[2]);
      sink_off[b] = static_cast<int32_t>(carry[3] + tmp[3][tid] - v[3]);
    }
    __syncthreads();
    if (tid < 4) carry[tid] += tmp[tid][kThreads - 1];
    __syncthreads();
  }
  mx[tid] = my_max;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (tid < o) mx[tid] = max(mx[tid], mx[tid + o]);
    __syncthreads();
  }
  if (tid == 0) {
    state_off[B] = static_cast<int32_t>(carry[0]);
    arc_off[B] = static_cast<int32_t>(carry[1]);
    level_off[B] = static_cast<int32_t>(carry[2]);
    sink_off[B] = static_cast<int32_t>(carry[3]);
    if (carry[0] >= 0x7fffffffLL || carry[1] >= 0x7fffffffLL) atomicCAS(&totals[4], 0, 3);
    totals[0] = static_cast<int32_t>(carry[0]);
    totals[1] = static_cast<int32_t>(carry[1]);
    totals[2] = static_cast<int32_t>(carry[2]);
    totals[3] = static_cast<int32_t>(carry[3]);
    totals[6] = mx[0];
  }
}

// ---------------------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    pack_build_kernel(const RawArcs R, int capS, int capA, const int32_t* __restrict__ level_g, const Kept K,
                      const nfst_pack_out_t O, const int32_t* __restrict__ totals) {
  if (totals[4]) return;  // an invalid lattice somewhere: the caller raises, nothing is used
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int s0 = R.state_off[b], S0 = R.state_off[b + 1] - s0;
  const int a0 = R.arc_off[b];
  const int s_lo = O.state_off[b], Sb = O.state_off[b + 1] - s_lo;
  const int a_lo = O.arc_off[b], Ab = O.arc_off[b + 1] - a_lo;
  const int A0 = Ab;  // the kept arcs, in raw order (pack_level_kernel left them in the workspace)
  const int l_lo = O.level_off[b], nlev = O.level_off[b + 1] - l_lo - 1;
  const int k_lo = O.sink_off[b];
  int* p = pack_smem;
  int* const level = p;      p += capS;      // [S0] raw state -> level (-1: trimmed)
  int* const new_id = p;     p += capS;      // [S0] raw state -> packed local id
  int* const first_raw = p;  p += capS;      // [S0] raw state -> index of its first raw arc
  int* const lvl = p;        p += capS + 2;  // [nlev + 1] level start (packed local ids), then a cursor copy
  int* const cur = p;        p += capS + 2;
  int* const out_ptr = p;    p += capS + 1;  // [Sb + 1]
  int* const in_ptr = p;     p += capS + 1;
  int* const in_cur = p;     p += capS + 1;
  int* const raw = p;        p += capA;      // [A0] src | dst << 16 (raw local ids)
  unsigned short* const can_src = reinterpret_cast<unsigned short*>(p);  // [Ab] packed local source of each canonical arc
  unsigned short* const in_tmp = can_src + capA;                           // [Ab] canonical arc (local) per in-order position
  __shared__ int tmp[kThreads];
  __shared__ int sink_cursor;

  for (int s = tid; s < S0; s += kThreads) {
    level[s] = level_g[s0 + s];
    first_raw[s] = 0;
  }
  for (int i = tid; i <= nlev + 1; i += kThreads) lvl[i] = 0;
  for (int i = tid; i <= Sb; i += kThreads) {
    out_ptr[i] = 0;
    in_ptr[i] = 0;
    in_cur[i] = 0;
  }
  if (tid == 0) sink_cursor = 0;
  for (int i = tid; i < A0; i += kThreads) raw[i] = K.pack[a0 + i];
  __syncthreads();
  // states per level -> level starts
  for (int s = tid; s < S0; s += kThreads)
    if (level[s] >= 0) atomicAdd(&lvl[level[s]], 1);
  __syncthreads();
  block_excl_scan(lvl, nlev + 1, tmp);  // lvl[nlev] = Sb
  for (int i = tid; i <= nlev; i += kThreads) cur[i] = lvl[i];
  __syncthreads();
  // packed id = (level, original id) order: one warp walks the states in original order, 32 at a time; the states
  // of a batch that share a level take consecutive ids in lane order (match_any), so the numbering is deterministic
  if (tid < 32) {
    for (int base = 0; base < S0; base += 32) {
      const int s = base + lane;
      const int l = s < S0 ? level[s] : -1;
      const unsigned same = __match_any_sync(0xffffffffu, l);
      if (l >= 0) {
        const int rank = __popc(same & ((1u << lane) - 1u));
        new_id[s] = cur[l] + rank;
      } else if (s < S0) {
        new_id[s] = -1;
      }
      __syncwarp();
      if (l >= 0 && (same & ((1u << lane) - 1u)) == 0) cur[l] += __popc(same);
      __syncwarp();
    }
  }
  __syncthreads();
  // degrees of the kept states; first raw arc of every raw state (raw arcs are sorted by source)
  for (int i = tid; i < A0; i += kThreads) {
    const int a = raw[i];
    const int s = a & 0xffff, d = static_cast<unsigned>(a) >> 16;
    if (i == 0 || (raw[i - 1] & 0xffff) != s) first_raw[s] = i;
    atomicAdd(&out_ptr[new_id[s]], 1);
    atomicAdd(&in_ptr[new_id[d]], 1);
  }
  __syncthreads();
  // sinks (kept states without arcs), in packed order: states of the deepest level and dead ends
  for (int base = 0; base < Sb; base += kThreads) {
    const int s = base + tid;
    const bool is_sink = s < Sb && out_ptr[s] == 0;
    tmp[tid] = is_sink ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
      int run = sink_cursor;
      for (int i = 0; i < kThreads; ++i) {
        const int x = tmp[i];
        tmp[i] = run;
        run += x;
      }
      sink_cursor = run;
    }
    __syncthreads();
    if (is_sink) O.sinks[k_lo + tmp[tid]] = s_lo + s;
    __syncthreads();
  }
  block_excl_scan(out_ptr, Sb + 1, tmp);
  block_excl_scan(in_ptr, Sb + 1, tmp);
  // canonical (out) order: arcs of one source are contiguous in the raw list and sorted by label already
  for (int i = tid; i < A0; i += kThreads) {
    const int a = raw[i];
    const int s = a & 0xffff, d = static_cast<unsigned>(a) >> 16;
    const int ns = new_id[s], nd = new_id[d];
    const int pos = out_ptr[ns] + (i - first_raw[s]);
    O.dst_out[a_lo + pos] = s_lo + nd;
    O.label_out[a_lo + pos] = K.label[a0 + i];
    O.arc_origin[a_lo + pos] = a0 + K.raw[a0 + i];
    can_src[pos] = static_cast<unsigned short>(ns);
    in_tmp[in_ptr[nd] + atomicAdd(&in_cur[nd], 1)] = static_cast<unsigned short>(pos);
  }
  __syncthreads();
  // in order: every destination's arcs sorted by canonical id (insertion sort per state: in-degrees are small),
  // which makes the layout independent of the order the atomics above were served in
  for (int s = tid; s < Sb; s += kThreads) {
    const int lo = in_ptr[s], hi = in_ptr[s + 1];
    for (int i = lo + 1; i < hi; ++i) {
      const unsigned short x = in_tmp[i];
      int j = i - 1;
      while (j >= lo && in_tmp[j] > x) {
        in_tmp[j + 1] = in_tmp[j];
        --j;
      }
      in_tmp[j + 1] = x;
    }
  }
  __syncthreads();
  for (int i = tid; i < Ab; i += kThreads) {
    const int c = in_tmp[i];
    O.in2out[a_lo + i] = a_lo + c;
    O.src_in[a_lo + i] = s_lo + can_src[c];
    O.label_in[a_lo + i] = O.label_out[a_lo + c];  // written above by this block
    O.src_out[a_lo + i] = s_lo + can_src[i];
  }
  for (int s = tid; s < Sb; s += kThreads) {
    O.out_ptr[s_lo + s] = a_lo + out_ptr[s];
    O.in_ptr[s_lo + s] = a_lo + in_ptr[s];
    if (O.out_deg8) O.out_deg8[s_lo + s] = static_cast<uint8_t>(min(out_ptr[s + 1] - out_ptr[s], 255));
  }
  for (int s = tid; s < S0; s += kThreads)
    if (level[s] >= 0) O.orig_state[s_lo + new_id[s]] = s;
  for (int i = tid; i <= nlev; i += kThreads) O.level_ptr[l_lo + i] = s_lo + lvl[i];
  if (tid == 0) {
    O.start_state[b] = s_lo + new_id[R.start];
    if (b == gridDim.x - 1) {  // closing entries of the CSR pointers
      O.out_ptr[s_lo + Sb] = a_lo + Ab;
      O.in_ptr[s_lo + Sb] = a_lo + Ab;
    }
  }
}

size_t level_smem(int capS, int capA) { return static_cast<size_t>(4 * capS + capA) * 4; }
size_t build_smem(int capS, int capA) { return static_cast<size_t>(8 * capS + 8 + capA) * 4 + static_cast<size_t>(capA) * 4; }

}  // namespace

extern "C" {

size_t nfst_pack_small_smem_bytes(int32_t max_states, int32_t max_arcs) {
  const size_t a = level_smem(max_states, max_arcs), b = build_smem(max_states, max_arcs);
  return a > b ? a : b;
}

size_t nfst_pack_small_workspace_bytes(int64_t n_states_raw, int64_t n_arcs_raw) {
  // level of every raw state, its range in the raw arc list; the kept arcs {src | dst << 16, label, raw position}
  return (3 * static_cast<size_t>(n_states_raw) + 3 * static_cast<size_t>(n_arcs_raw) + 16) * 4;
}

int nfst_pack_small(int32_t n_lattices, const int32_t* raw_state_off, const int32_t* raw_arc_off, const int32_t* raw_src,
                    const int32_t* raw_dst, const int32_t* raw_label, int32_t src_is_global, int32_t start_state,
                    int32_t max_states, int32_t max_arcs, int32_t max_raw_arcs, const nfst_pack_out_t* out, void* workspace,
                    size_t workspace_bytes, int64_t n_states_raw, int64_t n_arcs_raw, int32_t phases, void* cuda_stream) {
  if (n_lattices <= 0 || !raw_state_off || !raw_arc_off || !out || !workspace || !(phases & 3))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_pack_small: null argument, empty batch or no phase");
  if (n_lattices > 65535)
    return nfst_fail_msg(NFST_ERR_TOO_LARGE, "nfst_pack_small: at most 65535 lattices per call (got %d)", n_lattices);
  if (max_states < 1 || max_states > 65535 || max_arcs < 0 || max_arcs > 65535 || max_raw_arcs < 0)
    return nfst_fail_msg(NFST_ERR_TOO_LARGE, "nfst_pack_small: at most 65535 states and arcs per lattice (got %d, %d)", max_states,
                         max_arcs);
  if (workspace_bytes < nfst_pack_small_workspace_bytes(n_states_raw, n_arcs_raw))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_pack_small: workspace too small");
  const size_t smem = nfst_pack_small_smem_bytes(max_states, max_arcs);
  if (smem > 227 * 1024)
    return nfst_fail_msg(NFST_ERR_TOO_LARGE, "nfst_pack_small: a lattice of %d states / %d arcs needs %zu bytes of shared memory",
                         max_states, max_arcs, smem);
  if (!out->state_off || !out->arc_off || !out->level_off || !out->sink_off || !out->n_levels || !out->lattice_stats || !out->totals)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_pack_small: the per-lattice outputs are required in every phase");
  if ((phases & 2) && (!out->start_state || !out->level_ptr || !out->sinks || !out->orig_state || !out->in_ptr || !out->out_ptr ||
                       !out->src_in || !out->label_in || !out->in2out || !out->dst_out || !out->label_out || !out->src_out ||
                       !out->arc_origin))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_pack_small: the build phase needs every per-state / per-arc output");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  int32_t* level_g = static_cast<int32_t*>(workspace);
  int32_t* first_g = level_g + n_states_raw;
  int32_t* end_g = first_g + n_states_raw;
  int32_t* kept0 = end_g + n_states_raw;
  Kept K{kept0, kept0 + n_arcs_raw, kept0 + 2 * n_arcs_raw};
  int32_t* stats = out->lattice_stats;
  RawArcs R{raw_src, raw_dst, raw_label, raw_state_off, raw_arc_off, src_is_global, start_state};
  if (smem > 48 * 1024) {
    PACK_CUDA_OK(cudaFuncSetAttribute(pack_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    PACK_CUDA_OK(cudaFuncSetAttribute(pack_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  if (phases & 1) {
    if (!(phases & 4)) PACK_CUDA_OK(cudaMemsetAsync(out->totals, 0, 8 * sizeof(int32_t), st));
    PACK_CUDA_OK(cudaMemsetAsync(first_g, 0, 2 * static_cast<size_t>(n_states_raw) * sizeof(int32_t), st));
    if (max_raw_arcs > 0) {
      const dim3 grid((max_raw_arcs + kThreads - 1) / kThreads, n_lattices);
      pack_ranges_kernel<<<grid, kThreads, 0, st>>>(R, first_g, end_g, out->totals);
    }
    pack_level_kernel<<<n_lattices, kThreads, level_smem(max_states, max_arcs), st>>>(R, max_states, max_arcs, level_g, first_g,
                                                                                    end_g, K, stats, out->totals);
    pack_scan_kernel<<<1, kThreads, 0, st>>>(n_lattices, stats, out->state_off, out->arc_off, out->level_off, out->sink_off,
                                            out->n_levels, out->totals);
  }
  if (phases & 2)
    pack_build_kernel<<<n_lattices, kThreads, build_smem(max_states, max_arcs), st>>>(R, max_states, max_arcs, level_g, K, *out,
                                                                                    out->totals);
  PACK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// =========================================================================================================
// On-device construction of the transliteration lattices
// =========================================================================================================
// What the reference does offline with OpenFst for every training pair (src/preprocess/tr.py:142-190:
// x o T o y, then OwnAST.mfst_weight_projection, src/modules/path_semiring.py:120-180, which replaces every arc
// by the chain of its marks): T is the one-state edit machine of src/fsm/tr.py:321-390 --
//   insertion of y_j        marks [output-mark, y_j]
//   deletion of x_i         marks [input-mark, x_i]
//   substitution x_i : y_j  marks [insertion-mark, input-mark, x_i, output-mark, y_j]   (add_sub)
// -- so x o T o y is the (|x|+1) x (|y|+1) edit grid and the mark lattice is that grid with every arc expanded
// into a chain, plus the bos arc in front and the eos arc behind (Preprocess.composed_to_matrices,
// src/preprocess/preprocess.py:51-174).  This kernel writes that lattice's arc list straight from the id
// strings, grouped by lattice and sorted by (source, label) -- the input format of nfst_pack_small.
// (The reference additionally runs pynini's optimize(); that renumbers and may merge states but keeps the set
// of mark strings, which is all the dynamic programme sees.)
namespace {

struct EditMarks {
  int bos, eos, input_mark, output_mark, sub_mark;
  int rank_del, rank_ins, rank_sub;  // position of each edit's first arc among a grid state's arcs (label order)
  int add_sub;
};

__global__ void __launch_bounds__(kThreads)
    edit_arcs_kernel(const int32_t* __restrict__ x, const int32_t* __restrict__ x_len, int x_stride,
                     const int32_t* __restrict__ y, const int32_t* __restrict__ y_len, int y_stride, const EditMarks M,
                     const int32_t* __restrict__ arc_off, int32_t* __restrict__ src, int32_t* __restrict__ dst,
                     int32_t* __restrict__ label) {
  const int b = blockIdx.x;
  const int n = x_len[b], m = y_len[b];
  const int32_t* xs = x + static_cast<size_t>(b) * x_stride;
  const int32_t* ys = y + static_cast<size_t>(b) * y_stride;
  const int c = 2 + (M.add_sub ? 1 : 0);  // arcs of an interior grid state
  const int G = (n + 1) * (m + 1);
  const int id_del = 1 + G, id_ins = id_del + n * (m + 1), id_sub = id_ins + (n + 1) * m;
  const int sink = id_sub + (M.add_sub ? 4 * n * m : 0);
  const int a_grid = 1, a_del = a_grid + n * (c * m + 1) + m + 1, a_ins = a_del + n * (m + 1), a_sub = a_ins + (n + 1) * m;
  const int a0 = arc_off[b];
  auto put = [&](int pos, int s, int l, int d) {
    src[a0 + pos] = s;
    label[a0 + pos] = l;
    dst[a0 + pos] = d;
  };
  auto grid = [&](int i, int j) { return 1 + i * (m + 1) + j; };
  if (threadIdx.x == 0) put(0, 0, M.bos, grid(0, 0));
  for (int cell = threadIdx.x; cell < G; cell += kThreads) {
    const int i = cell / (m + 1), j = cell % (m + 1);
    const int g = grid(i, j);
    const bool can_del = i < n, can_ins = j < m, can_sub = M.add_sub && can_del && can_ins;
    const int base = a_grid + (i < n ? i * (c * m + 1) + c * j : n * (c * m + 1) + j);
    if (can_del && can_ins) {  // interior: all edits, in label order
      put(base + M.rank_del, g, M.input_mark, id_del + i * (m + 1) + j);
      put(base + M.rank_ins, g, M.output_mark, id_ins + i * m + j);
      if (can_sub) put(base + M.rank_sub, g, M.sub_mark, id_sub + 4 * (i * m + j));
    } else if (can_del) {
      put(base, g, M.input_mark, id_del + i * (m + 1) + j);
    } else if (can_ins) {
      put(base, g, M.output_mark, id_ins + i * m + j);
    } else {
      put(base, g, M.eos, sink);
    }
    if (can_del) put(a_del + i * (m + 1) + j, id_del + i * (m + 1) + j, xs[i], grid(i + 1, j));
    if (can_ins) put(a_ins + i * m + j, id_ins + i * m + j, ys[j], grid(i, j + 1));
    if (can_sub) {
      const int k = id_sub + 4 * (i * m + j), a = a_sub + 4 * (i * m + j);
      put(a, k, M.input_mark, k + 1);
      put(a + 1, k + 1, xs[i], k + 2);
      put(a + 2, k + 2, M.output_mark, k + 3);
      put(a + 3, k + 3, ys[j], grid(i + 1, j + 1));
    }
  }
}

}  // namespace

extern "C" {

/* states and arcs of the lattice of one (|x|, |y|) pair */
void nfst_edit_lattice_size(int32_t n, int32_t m, int32_t add_sub, int64_t* n_states, int64_t* n_arcs) {
  const int64_t c = 2 + (add_sub ? 1 : 0), N = n, Mm = m;
  if (n_states) *n_states = 1 + (N + 1) * (Mm + 1) + N * (Mm + 1) + (N + 1) * Mm + (add_sub ? 4 * N * Mm : 0) + 1;
  if (n_arcs) *n_arcs = 1 + N * (c * Mm + 1) + Mm + 1 + N * (Mm + 1) + (N + 1) * Mm + (add_sub ? 4 * N * Mm : 0);
}

int nfst_edit_lattice_arcs(int32_t n_lattices, const int32_t* x, const int32_t* x_len, int32_t x_stride, const int32_t* y,
                           const int32_t* y_len, int32_t y_stride, int32_t bos, int32_t eos, int32_t input_mark,
                           int32_t output_mark, int32_t sub_mark, int32_t add_sub, const int32_t* raw_arc_off, int32_t* src,
                           int32_t* dst, int32_t* label, void* cuda_stream) {
  if (n_lattices <= 0 || !x || !x_len || !y || !y_len || !raw_arc_off || !src || !dst || !label)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_edit_lattice_arcs: null argument or empty batch");
  if (input_mark == output_mark || (add_sub && (sub_mark == input_mark || sub_mark == output_mark)))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_edit_lattice_arcs: the edit marks must be distinct labels");
  EditMarks M{bos, eos, input_mark, output_mark, sub_mark, 0, 0, 0, add_sub};
  // order of a grid state's arcs = order of their labels (the scan order of the reference's dense tables)
  M.rank_del = (output_mark < input_mark) + (add_sub && sub_mark < input_mark);
  M.rank_ins = (input_mark < output_mark) + (add_sub && sub_mark < output_mark);
  M.rank_sub = (input_mark < sub_mark) + (output_mark < sub_mark);
  edit_arcs_kernel<<<n_lattices, kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(x, x_len, x_stride, y, y_len, y_stride, M,
                                                                                     raw_arc_off, src, dst, label);
  PACK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// =========================================================================================================
// Level sweeps for the tensor-op packer (lattices too large for nfst_pack_small)
// =========================================================================================================
// level[dst] = max(level[dst], level[src] + 1) over all arcs, in place (atomicMax: updates are visible to the rest of
// the sweep, so a sweep usually settles several levels); *changed is set when anything moved.  The caller repeats
// until a sweep changes nothing -- the fixed point is the longest distance from the start states, whatever the order.
namespace {
__global__ void __launch_bounds__(256)
    level_sweep_kernel(const int64_t* __restrict__ gsrc, const int64_t* __restrict__ gdst, int64_t n_arcs, int32_t* level,
                       int32_t* __restrict__ changed) {
  bool moved = false;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_arcs;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ls = level[gsrc[i]];
    if (ls < 0) continue;
    // levels only grow: a plain read that already shows ls + 1 or more settles the arc without an atomic (after the
    // first sweeps that is nearly every arc); a stale smaller value only costs the atomic it would have cost anyway
    const int64_t d = gdst[i];
    if (level[d] <= ls && atomicMax(&level[d], ls + 1) < ls + 1) moved = true;
  }
  if (__any_sync(0xffffffffu, moved) && (threadIdx.x & 31) == 0) *changed = 1;
}
}  // namespace

extern "C" int nfst_level_sweeps(const int64_t* gsrc, const int64_t* gdst, int64_t n_arcs, int32_t* level, int32_t* changed,
                                 int32_t n_sweeps, void* cuda_stream) {
  if (!gsrc || !gdst || !level || !changed || n_arcs < 0 || n_sweeps < 1)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_level_sweeps: null argument");
  if (n_arcs == 0) return 0;
  const int64_t want = (n_arcs + 255) / 256;
  const int blocks = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
  for (int k = 0; k < n_sweeps; ++k)
    level_sweep_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(gsrc, gdst, n_arcs, level, changed);
  PACK_CUDA_OK(cudaGetLastError());
  return 0;
}

// =========================================================================================================
// nfst_pack_dense: collate()-padded dense tables -> packed lattices through the C ABI alone
// =========================================================================================================
// The entry SURVEY.md section 8(b) asks for: everything pack_dense() does on the device, in one call per phase --
// the edge rule over transition[B, S, V] (nfst_dense_count_arcs), the row offsets (one block scans the B*S row
// counts), the arc list (nfst_dense_extract_arcs), then nfst_pack_small.  The caller sizes the raw arc list by a
// capacity; more arcs than that is reported through totals[4] = 5.
namespace {

__global__ void __launch_bounds__(kThreads)
    dense_offsets_kernel(const int32_t* __restrict__ row_counts, int64_t n_rows, int S, int B, int64_t capacity,
                         int64_t* __restrict__ row_start, int32_t* __restrict__ raw_state_off, int32_t* __restrict__ raw_arc_off,
                         int32_t* __restrict__ totals) {
  __shared__ long long part[kThreads];
  const int tid = threadIdx.x;
  const int64_t per = (n_rows + kThreads - 1) / kThreads;
  const int64_t lo = tid * per, hi = lo + per < n_rows ? lo + per : n_rows;
  long long sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += row_counts[i];
  part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < kThreads; o <<= 1) {
    const long long add = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += add;
    __syncthreads();
  }
  long long run = tid ? part[tid - 1] : 0;
  for (int64_t i = lo; i < hi; ++i) {
    row_start[i] = run;
    run += row_counts[i];
  }
  __syncthreads();
  const long long total = part[kThreads - 1];
  if (tid == 0) {
    row_start[n_rows] = total;
    if (total > capacity || total >= 0x7fffffffLL) atomicCAS(&totals[4], 0, 5);
  }
  __syncthreads();
  for (int b = tid; b <= B; b += kThreads) {
    raw_state_off[b] = b * S;
    const long long at = b < B ? row_start[static_cast<int64_t>(b) * S] : total;
    raw_arc_off[b] = static_cast<int32_t>(at < 0x7fffffffLL ? at : 0x7fffffffLL);
  }
}

// arc_origin: index into the extracted arc list -> dense cell (b * S + s) * V + label
__global__ void __launch_bounds__(kThreads)
    origin_to_cell_kernel(int64_t* __restrict__ arc_origin, const int32_t* __restrict__ arc_row, const int32_t* __restrict__ arc_label,
                          int V, const int32_t* __restrict__ totals) {
  if (totals[4]) return;
  const int64_t A = totals[1];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < A; i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int64_t r = arc_origin[i];
    arc_origin[i] = static_cast<int64_t>(arc_row[r]) * V + arc_label[r];
  }
}

}  // namespace

extern "C" {


size_t nfst_pack_workspace_bytes(int32_t n_lattices, int32_t states_per_lattice, int64_t raw_arc_capacity) {
  const size_t rows = static_cast<size_t>(n_lattices) * states_per_lattice;
  // row counts, row offsets (int64), raw lattice offsets, the raw arc list {row, label, dst}, nfst_pack_small's workspace
  return rows * 4 + (rows + 1) * 8 + 2 * (static_cast<size_t>(n_lattices) + 1) * 4 + 3 * static_cast<size_t>(raw_arc_capacity) * 4 + 64 +
         nfst_pack_small_workspace_bytes(static_cast<int64_t>(rows), raw_arc_capacity);
}

int nfst_pack_dense(const int64_t* transition, int32_t n_lattices, int32_t states_per_lattice, int32_t vocab,
                    int64_t raw_arc_capacity, int32_t max_arcs, const nfst_pack_out_t* out, void* workspace, size_t workspace_bytes,
                    int32_t phases, void* cuda_stream) {
  if (!transition || !out || !workspace || n_lattices <= 0 || states_per_lattice < 1 || vocab < 1 || raw_arc_capacity < 0)
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_pack_dense: bad argument");
  if (workspace_bytes < nfst_pack_workspace_bytes(n_lattices, states_per_lattice, raw_arc_capacity))
    return nfst_fail_msg(NFST_ERR_BAD_ARG, "nfst_pack_dense: workspace too small");
  if (raw_arc_capacity >= 0x7fffffffLL || static_cast<int64_t>(n_lattices) * states_per_lattice >= 0x7fffffffLL)
    return nfst_fail_msg(NFST_ERR_TOO_LARGE, "nfst_pack_dense: batch beyond int32 indices");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int64_t rows = static_cast<int64_t>(n_lattices) * states_per_lattice;
  char* w = static_cast<char*>(workspace);
  int64_t* row_start = reinterpret_cast<int64_t*>(w);                        w += (rows + 1) * 8;
  int32_t* row_counts = reinterpret_cast<int32_t*>(w);                       w += rows * 4;
  int32_t* raw_state_off = reinterpret_cast<int32_t*>(w);                    w += (static_cast<size_t>(n_lattices) + 1) * 4;
  int32_t* raw_arc_off = reinterpret_cast<int32_t*>(w);                      w += (static_cast<size_t>(n_lattices) + 1) * 4;
  w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(w) + 15) & ~static_cast<uintptr_t>(15));
  int32_t* arc_row = reinterpret_cast<int32_t*>(w);                          w += raw_arc_capacity * 4;
  int32_t* arc_label = reinterpret_cast<int32_t*>(w);                        w += raw_arc_capacity * 4;
  int32_t* arc_dst = reinterpret_cast<int32_t*>(w);                          w += raw_arc_capacity * 4;
  w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(w) + 15) & ~static_cast<uintptr_t>(15));
  const size_t small_bytes = nfst_pack_small_workspace_bytes(rows, raw_arc_capacity);
  if (phases & 1) {
    PACK_CUDA_OK(cudaMemsetAsync(out->totals, 0, 8 * sizeof(int32_t), st));
    if (int rc = nfst_dense_count_arcs(transition, rows, states_per_lattice, vocab, row_counts, cuda_stream)) return rc;
    dense_offsets_kernel<<<1, kThreads, 0, st>>>(row_counts, rows, states_per_lattice, n_lattices, raw_arc_capacity, row_start,
                                                raw_state_off, raw_arc_off, out->totals);
    PACK_CUDA_OK(cudaGetLastError());
    // more arcs than the capacity: totals[4] = 5 (set above), the extract drops what does not fit, the pack kernels return
    if (int rc = nfst_dense_extract_arcs(transition, rows, states_per_lattice, vocab, row_start, raw_arc_capacity, arc_row, arc_label,
                                         arc_dst, cuda_stream))
      return rc;
  }
  const int64_t cell_cap = static_cast<int64_t>(states_per_lattice) * vocab;
  const int32_t max_raw = static_cast<int32_t>(cell_cap < raw_arc_capacity ? cell_cap : raw_arc_capacity);
  // nfst_pack_small's count phase would zero the totals: bit 2 of phases keeps them (the overflow code above)
  if (int rc = nfst_pack_small(n_lattices, raw_state_off, raw_arc_off, arc_row, arc_dst, arc_label, /*src_is_global=*/1, /*start=*/0,
                               states_per_lattice, max_arcs, max_raw, out, w, small_bytes, rows, raw_arc_capacity, phases | 4,
                               cuda_stream))
    return rc;
  if (phases & 2) {
    origin_to_cell_kernel<<<148 * 4, kThreads, 0, st>>>(out->arc_origin, arc_row, arc_label, vocab, out->totals);
    PACK_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
