/*
 * nfst_b200.h -- C ABI of the B200 (sm_100a) lattice dynamic-programming library.
 *
 * Drop-in boundary for the one hot path of steventan0110/nFST that this repository
 * accelerates: dynamic programming over batched, acyclic, CSR-packed mark lattices.
 * Plain C: raw device pointers, explicit sizes, a CUDA stream handle; no C++ or torch
 * types cross this boundary.  The library never allocates or frees device memory and
 * keeps no pointer after a call returns; everything is enqueued on the caller's stream
 * and nothing synchronises with the host.
 *
 * Reference interfaces replaced (file:line in steventan0110/nFST):
 *   nfst_fwd_f32            log-semiring forward pass (alpha, logZ).  No counterpart in
 *                           the reference; mirror image of the beta pass below.
 *   nfst_bwd_fused_f32      FSAGRUScorer.compute_beta_per_sample / compute_beta_parallel
 *                           (src/modules/scorers.py:692-751, :753-856) in the Wh=0 regime,
 *                           i.e. beta[c] = sum_{c-j->n} exp(theta_j) beta[n] in log space;
 *                           fused with arc posteriors (the autograd gradient of logZ that
 *                           the reference cannot produce, quirk Q7), the per-label
 *                           gradient d logZ / d theta (WFSTScorer, scorers.py:1663-1687)
 *                           and the tropical-semiring Viterbi recursion + backpointers.
 *   nfst_fwd_bwd_small_f32  the same two passes (and Viterbi) for lattices of the size nFST builds, whole lattice
 *                           in shared memory, one launch (scorers.py:692-751, :753-856 as above).
 *   nfst_tile_pull_f32 / nfst_tile_flow_f32
 *                           the same recurrence and the arc posteriors for wide lattices stored as a per-warp tile
 *                           stream (the bench workload); same reference functions as nfst_bwd_fused_f32.
 *   nfst_viterbi_f32 / nfst_viterbi_paths_f32
 *                           tropical-semiring pass (+ backtrace in the same call): the exact counterpart of
 *                           "the sample with the highest weight" (src/modules/lightning.py:474-479); tie rule =
 *                           torch.argmax's first index over the dense row, as the reference's consumer gathers it
 *                           (scorers.py:586-590).
 *   nfst_compact_paths / nfst_pad_paths
 *                           the best path in ragged form / in the reference's padded int64 sample layout
 *                           (what JointProb.forward(return_samples=True) returns, lightning.py:474-479).
 *   nfst_level_sweeps       topological levels of an arc list: the order the reference gets implicitly from its
 *                           per-state message counters (scorers.py:741-749, :846-852).
 *   nfst_backtrace          best-path read-out; replaces best-of-k-samples selection
 *                           (src/modules/lightning.py:474-479) reached from
 *                           src/decode/decoder.py:77-79.
 *   nfst_beta_hat_level_f32 the same recurrence with Wh != 0 (the beta-hat messages,
 *                           scorers.py:732-747), one topological level per call.
 *   nfst_walk_step_f32      one time step of Sampler.stateful_sample (src/modules/samplers.py:243-297):
 *                           update_fsa_state (scorers.py:683-690), mask_out_invalid (:1037-1054), the
 *                           beta look-ahead (:583-592) and the Categorical sample / log_prob, fused.
 *   nfst_sell_pull_f32 / nfst_sell_flow_f32
 *                           the beta recurrence and the arc posteriors for wide lattices stored as
 *                           column-major 32-state slices (same reference functions as nfst_bwd_fused_f32).
 *   nfst_beta_to_dense      layout of compute_beta()'s return value, real-space
 *                           beta[B*k, S] (scorers.py:854, :858-875).
 *   nfst_pack_small         the per-call graph build at the top of compute_beta_per_sample / compute_beta_parallel
 *                           (scorers.py:704-716, :764-776: adjacency lists from the dense table) plus the
 *                           topological ordering the reference does implicitly with its message counters
 *                           (:741-749); trims the rows collate() padding adds (util/dataset_reader.py:175-186).
 *   nfst_edit_lattice_arcs  the offline lattice construction x o T o y + mark expansion for the transliteration edit
 *                           machine (src/preprocess/tr.py:142-190, src/fsm/tr.py:321-390,
 *                           src/modules/path_semiring.py:120-180), emitted on the device from id strings.
 *   nfst_dense_count_arcs / nfst_dense_extract_arcs
 *                           the dense-table edge rule `t != 0 and t != i`
 *                           (scorers.py:704-716, :764-776) over collate()-padded
 *                           transition[B,S,V] int64 tables (util/dataset_reader.py:175-186).
 *
 *   nfst_abi_version, nfst_last_error_string, nfst_device_info, nfst_*_smem_bytes, nfst_*_workspace_bytes,
 *   nfst_tile_debug_read    housekeeping of the boundary itself (version, error text -- the reference raises Python
 *                           exceptions, e.g. scorers.py:719,878-879,1005,1030 --, device check, buffer sizes the
 *                           caller must provide, range-check counters of a debug build); no reference counterpart.
 *
 * Every entry point returns 0 on success or a negative nfst_status; a human-readable
 * message for the last failure on the calling thread is at nfst_last_error_string().
 */
#ifndef NFST_B200_H_
#define NFST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFST_ABI_VERSION 18

typedef enum nfst_status {
  NFST_OK = 0,
  NFST_ERR_BAD_ARG = -1,
  NFST_ERR_CUDA = -2,
  NFST_ERR_UNSUPPORTED_DEVICE = -3,
  NFST_ERR_TOO_LARGE = -4,
  NFST_ERR_CYCLIC = -5,      /* pack: a lattice has a cycle (the DP is defined for acyclic lattices only) */
  NFST_ERR_BAD_LATTICE = -6  /* pack: arc endpoint or start state out of range */
} nfst_status;

/* One unit of work of a lattice: a run of consecutive states of ONE topological level
 * together with their (contiguous) CSR arc range.  16 bytes. */
typedef struct nfst_chunk {
  int32_t arc_begin, arc_end;     /* in-order positions (forward) / canonical ids (backward) */
  int32_t state_begin, state_end; /* packed state ids */
} nfst_chunk_t;

/*
 * A batch of packed lattices (all arrays are DEVICE pointers, int32 unless noted).
 *
 * States of every lattice are renumbered so that they are contiguous per lattice and,
 * inside a lattice, sorted by topological level (longest distance from the start
 * state); ids below are these global packed ids.  Arcs are stored twice:
 *   "out" order (canonical arc id): sorted by (source state, label) -- CSR by source;
 *        this is the dense-table scan order of the reference (state, then label).
 *   "in" order: sorted by (destination state, canonical id) -- CSR by destination.
 * Lattice b has levels 0..L_b-1; level l spans packed states
 *   [ level_ptr[level_off[b]+l], level_ptr[level_off[b]+l+1] ).
 * The kernels walk each lattice chunk by chunk: forward chunks in ascending state order,
 * backward chunks in descending state order; a chunk never crosses a level boundary, so
 * everything a chunk reads was produced by earlier chunks.
 * Per-arc arrays must be readable up to the next multiple of 4 elements (128-bit loads).
 */
typedef struct nfst_packed_lattices {
  int32_t n_lattices; /* B */
  int32_t n_states;   /* S: total packed states */
  int32_t n_arcs;     /* A: total packed arcs   */
  int32_t vocab;      /* V: labels are in [0, V) */
  const int32_t* state_off;   /* [B+1] */
  const int32_t* level_off;   /* [B+1] offsets into level_ptr (lattice b owns L_b+1 entries) */
  const int32_t* level_ptr;   /* [sum_b (L_b+1)] */
  const int32_t* start_state; /* [B] packed id of the start state (reference row 0) */
  const int32_t* sink_off;    /* [B+1] */
  const int32_t* sinks;       /* packed ids of states without outgoing arcs, grouped by lattice */
  const int32_t* in_ptr;      /* [S+1] */
  const int32_t* src_in;      /* [A] source state of the arc at each in-order position */
  const int32_t* label_in;    /* [A] */
  const int32_t* in2out;      /* [A] canonical arc id of each in-order position */
  const int32_t* out_ptr;     /* [S+1] */
  const int32_t* dst_out;     /* [A] destination state of each canonical arc */
  const int32_t* label_out;   /* [A] */
  const uint8_t* lanes_in_log2;  /* [B] log2 of lanes cooperating on one state, forward  */
  const uint8_t* lanes_out_log2; /* [B] same, backward */
  const int32_t* fwd_chunk_off;     /* [B+1] */
  const nfst_chunk_t* fwd_chunks;   /* ascending */
  const int32_t* bwd_chunk_off;     /* [B+1] */
  const nfst_chunk_t* bwd_chunks;   /* descending */
  const int32_t* fwd_gather;        /* [n_fwd_chunks][2]: canonical-id range [lo, hi) the chunk's arcs gather from */
  const int32_t* bwd_order;         /* [S] states in backward processing order: a permutation inside every
                                       backward chunk's state range, sorted by out-degree */
  /* Sliced-column lattices (launch groups with nfst_launch_t.sell != 0; see nfst_sell_pull_f32).  The states
   * of every level are sorted by out-degree, descending, and cut into SLICES of 32 consecutive states; the
   * arcs of a slice occupy the canonical ids [out_ptr[first], out_ptr[end of slice]) like in CSR, but
   * COLUMN-MAJOR without padding: first the label-wise 1st arc of every state of the slice that has one (in
   * state order), then the 2nd arcs, ...  Lane i of a warp owns state i, so each column is one coalesced load.
   * out_ptr[s+1] - out_ptr[s] is still the out-degree; out_deg8[s] = min(out-degree, 255) (255 = look at
   * out_ptr).  Once a single state of a slice is left, its remaining arcs are contiguous. */
  const uint8_t* out_deg8;          /* [S] */
  /* One 16-byte descriptor per slice of the sliced-column lattices, in (lattice, level, position) order:
   *   [0] canonical id of the slice's first arc, [1] id behind its last arc,
   *   [2] bytes: start of column 1, 2, 3, 4 relative to [0]   (column 0 starts at 0),
   *   [3] bytes: start of column 5, 6, 7; top byte = min(largest out-degree of the slice, 255).
   * sell_lvl_slice is aligned with level_ptr: entry level_off[b]+l is the index of the first slice of level l
   * of lattice b (the spare entry of each lattice closes its last level); CSR lattices own no slices. */
  const int32_t* sell_desc;         /* [n_slices][4] */
  const int32_t* sell_lvl_slice;    /* [sum_b (L_b+1)] */
  /* Tile-stream lattices (launch groups with nfst_launch_t.tiles != 0; see nfst_tile_pull_f32).  The lattice is
   * dealt to the nw warps of its block at pack time: the states of a level are sorted by out-degree
   * (descending), cut into slices of 32 (a state with more than 32 arcs is a slice of its own: "heavy"), the
   * slices are dealt to the warps in serpentine rounds, and states and arcs are numbered so that what one warp
   * does in one level is a contiguous run of both.  The arcs of a slice are column-major without padding;
   * consecutive slices of one (level, warp) form TILES, the unit a warp fetches with one bulk copy per array.
   *   tile_tab     one int32[4] per tile, per (lattice, warp) in level order:
   *                  [0] canonical id of the tile's first arc, [1] offset of the tile in tile_stream / 16,
   *                  [2] arcs | segments << 16 | extension blocks << 24, [3] level | (stream bytes / 16) << 16
   *   tile_lw_off  [n_(lattice,warp) + 1] offsets into tile_tab
   *   tile_lat_info one int32[4] per lattice: [0] index of its warp 0 in tile_lw_off, [1] W = slots of its DP
   *                ring (a multiple of 32), [2] slots in all: ring, the constant slot W, the far table,
   *                [3] canonical id of its first arc
   *   tile_stream  per tile, 16-byte aligned: a 16-byte tile header {first state and first arc (both
   *                lattice-relative), ring slot of the first slice | offset of the slot region << 16,
   *                level | segments << 16}; one 16-byte header per segment {bytes n_0..n_7 = states with more
   *                than k arcs, arc offset in the tile (16 bit), states (8), largest degree (8), offset of the
   *                segment's extension region (16), flags (16)}; extension regions {32 bytes n_8..n_39 when the
   *                largest degree exceeds 8; 64 bytes = the 32 lanes' far-table slots (0 = none) when flag 2 is
   *                set}; the arcs' destinations as 16-bit slots: the ring slot of the destination while it is
   *                resident, W = "a constant" for the last level (beta = 0), or the destination's slot in the FAR
   *                TABLE (W+1 ...) -- states that outlive their ring slot are additionally kept there, written
   *                once and never recycled, so every arc reads ring[slot]; 64 bytes of slot W close the region.
   *                Heavy segments: header words {arcs of the state, arcs before this piece}, arc-offset field =
   *                the state's far-table slot, extension field = arcs of the piece, flags 4 (heavy) | 8 (first
   *                piece) | 16 (last piece). */
  const uint8_t* tile_stream;
  const int32_t* tile_tab;
  const int32_t* tile_lw_off;
  const int32_t* tile_lat_info;
  /* Column-major lattices (sliced-column and tile-stream groups) do not keep a state's arcs together.  out_arc
   * [A] lists them: out_arc[out_ptr[s] + k] = canonical id of the k-th arc of state s in label order (for CSR
   * lattices that is the identity).  NULL when the batch has no column-major lattice.  Read by the consumers that
   * visit the arcs of ONE state: the sampling-loop kernels and the beta-hat recurrence. */
  const int32_t* out_arc;
} nfst_packed_lattices_t;

/*
 * One kernel launch = one thread block per lattice in `lattice_ids` (NULL = lattices
 * 0..n_ids-1).  The arc arrays of a chunk (up to `chunk_cap` arcs and states) are staged in
 * shared memory with cp.async, multi-buffered; larger chunks (a single state whose degree
 * exceeds that) take a slower block-wide path.
 * `window_states` (a power of two, >= 32) is the number of most recent per-state DP
 * values kept in shared memory; older ones are re-read from global memory.
 * `state_f64` != 0 keeps the log-semiring state vectors (alpha, beta, logZ) in float64:
 * needed for posteriors within 1e-5 on deep lattices, where |alpha| is in the hundreds
 * or thousands and an fp32 ulp is no longer small against 1e-5.
 */
/* tile-stream groups deeper than this keep float64 DP rings when the state vectors are float64 (see tile_flow_bits) */
#define NFST_TILE_F64_LEVELS 96
typedef struct nfst_launch {
  const int32_t* lattice_ids;
  int32_t n_ids;
  int32_t block_threads; /* consumer threads: 32, 64, 128 or 256 (one producer warp is added) */
  int32_t window_states;
  int32_t state_f64;
  int32_t chunk_cap; /* multiple of 8; >= the largest staged chunk of the launch, in arcs and in states */
  /* Level-major execution (for lattices with wide levels): when fwd_level_chunks / bwd_level_chunks
   * are non-NULL the pass runs as ONE KERNEL LAUNCH PER TOPOLOGICAL LEVEL, one thread block per
   * chunk, over the chunks of all lattices of the group, instead of one block per lattice.
   * *_level_chunks: DEVICE arrays of the group's chunks sorted by level; *_level_off: HOST arrays
   * [n_levels+1] of offsets into them; bwd_level_lat: DEVICE lattice id of every backward chunk. */
  int32_t n_levels;
  const nfst_chunk_t* fwd_level_chunks;
  const int32_t* fwd_level_off;
  const nfst_chunk_t* bwd_level_chunks;
  const int32_t* bwd_level_off;
  const int32_t* bwd_level_lat;
  /* Small-lattice execution: when small_max_arcs > 0 every lattice of the group fits in shared
   * memory (at most small_max_states states, small_max_arcs arcs, small_max_levels levels) and is
   * processed there in one piece (nfst_small_kernel). */
  int32_t small_max_states, small_max_arcs, small_max_levels;
  /* Sliced-column execution (sell != 0): one thread block of block_threads (32..1024) per lattice, one
   * barrier per level, arcs read column by column straight into registers; window_states (a power of two) is
   * the size of the shared-memory ring of DP values (>= the widest level of the group).  An arc is served from the ring iff
   * dst < (first state of src's level) + window_states; the others (sell_far != 0 announces that some exist)
   * go through global memory: the pull pass re-reads beta / delta, the flow pass adds into gamma_far.
   * n_levels = largest level count of the group. */
  int32_t sell;
  int32_t sell_far;
  /* Tile-stream execution (tiles != 0): one block of block_threads = 32 * nw threads per lattice (nw = the
   * warps the lattices were dealt to), every warp streams its own tiles through a tile_stages-deep ring of
   * shared-memory stages filled by bulk copies (TMA) that complete on per-stage mbarriers.  tile_ring = largest
   * DP ring of the group in slots (ring + constant slot + far table), tile_cap_arcs / tile_cap_bytes = largest
   * tile of the group (arcs / stream bytes), tile_far = largest far table of the group (informational),
   * tile_stages: 0 = chosen by the library from the shared memory the launch leaves per block.
   * tile_flow_bits: width of the flow pass's fixed-point gamma accumulator -- 32 (or 0): 2^-31 units, posteriors
   * carry an ABSOLUTE error below 1e-9 (relative 1e-5 down to posteriors of 1e-4); 64: 2^-62 units, the error is
   * the float32 rounding of each contribution (relative ~1e-6 at any size), twice the ring's shared memory.
   * With state_f64 the DP ring holds doubles only when n_levels > NFST_TILE_F64_LEVELS (the packer sizes such
   * lattices' rings for 8-byte slots); shallower groups keep the float32 ring and write beta / logZ as doubles. */
  int32_t tiles;
  int32_t tile_ring;
  int32_t tile_far;
  int32_t tile_cap_arcs;
  int32_t tile_cap_bytes;
  int32_t tile_stages;
  int32_t tile_flow_bits;
} nfst_launch_t;

/* Arc scores: w(a) = (arc_scores ? arc_scores[a] : 0) + (theta ? theta[label(a)] : 0);
 * arc_scores is indexed by canonical arc id (readable up to the next multiple of 4). */
typedef struct nfst_scores {
  const float* arc_scores; /* [A] or NULL */
  const float* theta;      /* [V] or NULL */
} nfst_scores_t;

int nfst_abi_version(void);
const char* nfst_last_error_string(void);

/* Fails with NFST_ERR_UNSUPPORTED_DEVICE unless `device` is compute capability 10.x.
 * Any of the out pointers may be NULL. */
int nfst_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* max_smem_optin);

/* theta / dtheta are staged in shared memory when V <= NFST_THETA_SMEM_MAX. */
#define NFST_THETA_SMEM_MAX 4096
/* Dynamic shared memory (bytes) of a forward (pass = 0) or backward (pass = 1) launch.
 * with_log / with_trop select the semirings of the backward pass. */
size_t nfst_launch_smem_bytes(const nfst_launch_t* launch, int32_t vocab, int pass, int with_log, int with_trop,
                              int with_post, int with_scores, int with_theta, int with_dtheta);

/* alpha[S] (log space), logz[B] = logsumexp over the lattice's sinks of alpha.  alpha and
 * logz are float32, or float64 when launch->state_f64. */
int nfst_fwd_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                 void* alpha, void* logz, void* cuda_stream);

/*
 * Fused backward.  Any output may be NULL, with these constraints:
 *   log semiring  : runs when beta, logz_bwd, post or dtheta is given.  beta[S] (log space),
 *                   logz_bwd[B] = beta[start]  (float32, or float64 when state_f64; beta is
 *                   required as the pass's working vector).
 *       post[A]   : float32; needs alpha and logz from nfst_fwd_f32 (same state dtype);
 *                   post[a] = g_b * exp(alpha[src] + w + beta[dst] - logz[b]) with
 *                   g_b = grad_logz[b] (float32; 1 when grad_logz == NULL).
 *       dtheta[V] : float32; needs alpha/logz; atomically accumulates post[a] by label
 *                   (caller zero-fills).
 *   tropical      : runs when delta/backptr is given (both required; always float32).
 *                   delta[S], backptr[S] (canonical arc id, -1 at sinks), vit_score[B] =
 *                   delta[start].  delta[s] = max_a fl32(w_a + delta[dst_a]); ties ->
 *                   smallest canonical arc id (= smallest label).
 */
int nfst_bwd_fused_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                       const void* alpha, const void* logz, const float* grad_logz, void* beta, void* logz_bwd,
                       float* post, float* dtheta, float* delta, int32_t* backptr, float* vit_score,
                       void* cuda_stream);

/* Forward + fused backward in ONE launch for a small-lattice group (launch->small_max_arcs > 0):
 * alpha never leaves the SM.  logz is required; alpha, beta, logz_bwd, post, dtheta optional. */
int nfst_fwd_bwd_small_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                           const float* grad_logz, void* alpha, void* logz, void* beta, void* logz_bwd, float* post,
                           float* dtheta, void* cuda_stream);

/* Tropical pass only (thin wrapper over nfst_bwd_fused_f32). */
int nfst_viterbi_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                     float* delta, int32_t* backptr, float* vit_score, void* cuda_stream);

/* Follow backptr from each start state.  path_arcs[path_off[b] .. path_off[b]+path_len[b])
 * receives canonical arc ids; capacity of lattice b is path_off[b+1]-path_off[b] (L_b-1
 * always suffices). */
int nfst_backtrace(const nfst_packed_lattices_t* lat, const int32_t* backptr, const int32_t* path_off,
                   int32_t* path_arcs, int32_t* path_len, void* cuda_stream);

/* Viterbi recursion + best-path read-out for one launch group: as nfst_viterbi_f32 followed by
 * nfst_backtrace restricted to the group's lattices; small-lattice groups follow the
 * backpointers inside the kernel, from shared memory (delta may then be NULL). */
int nfst_viterbi_paths_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                           float* delta, int32_t* backptr, float* vit_score, const int32_t* path_off,
                           int32_t* path_arcs, int32_t* path_len, void* cuda_stream);

/* Ragged result: copies every lattice's path_len[b] arcs from its slot path_buf[path_off[b]..] to
 * out_arcs[out_off[b]..] and (if out_labels != NULL) their labels to out_labels (path_labels of
 * lattice_viterbi; replaces the best-sample read-out of lightning.py:474-479). */
int nfst_compact_paths(const nfst_packed_lattices_t* lat, const int32_t* path_off, const int32_t* path_len,
                       const int32_t* path_buf, const int64_t* out_off, int32_t* out_arcs, int32_t* out_labels,
                       void* cuda_stream);

/* Padded result, no host read: row b of out_labels[B, row_len] (int64, the reference's sample dtype) = the labels of
 * lattice b's path -- without a leading skip_label (>= 0: the reference's samples never hold bos, scorers.py:230-231)
 * -- followed by pad_label; out_len[b] (may be NULL) = labels written.  The best-sample read-out of
 * lightning.py:474-479 as one [B, T] tensor. */
int nfst_pad_paths(const nfst_packed_lattices_t* lat, const int32_t* path_off, const int32_t* path_len, const int32_t* path_buf,
                   int32_t skip_label, int64_t pad_label, int32_t row_len, int64_t* out_labels, int32_t* out_len,
                   void* cuda_stream);

/* out[(b*k+j)*dense_states + orig_state[s]] = exp(beta[s]) for j < k (float32 out; beta
 * float32 or float64 per beta_f64); `out` must be zero-filled by the caller (states that
 * were trimmed at pack time keep 0). */
int nfst_beta_to_dense(const nfst_packed_lattices_t* lat, const void* beta, int beta_f64, const int32_t* orig_state,
                       int32_t k, int32_t dense_states, float* out, void* cuda_stream);

/* Dense-table edge rule.  transition is int64 [n_rows = B*S, V] row-major, row r belongs
 * to state r % S.  count: row_counts[r] = #cells with t != 0 && t != r % S.  extract:
 * given exclusive prefix sums row_start[r], writes (row, label, dst) in scan order. */
int nfst_dense_count_arcs(const int64_t* transition, int64_t n_rows, int32_t states_per_lattice, int32_t vocab,
                          int32_t* row_counts, void* cuda_stream);
int nfst_dense_extract_arcs(const int64_t* transition, int64_t n_rows, int32_t states_per_lattice, int32_t vocab,
                            const int64_t* row_start, int64_t arc_capacity, int32_t* arc_row, int32_t* arc_label,
                            int32_t* arc_dst, void* cuda_stream);

/*
 * Sliced-column passes (launch->sell != 0).
 *
 * nfst_sell_pull_f32 -- deepest level first.  Log semiring when beta, logz_bwd or cond is given:
 *   beta[S] / logz_bwd[B] (float32, or float64 when launch->state_f64; per-arc terms are float32 offsets from
 *   a reference arc of the state, float64 with a float64 state), and cond[A] (float32, real space):
 *   cond[a] = exp(w_a + beta[dst_a] - beta[src_a]), the probability of arc a given its source state.
 *   Tropical semiring when backptr is given: delta[S] (optional), backptr[S], vit_score[B] (optional),
 *   same rule as nfst_bwd_fused_f32.  Groups with sell_far need beta (resp. delta): arcs longer than the
 *   ring re-read them from global memory.  Replaces FSAGRUScorer.compute_beta_* (scorers.py:692-856, Wh = 0).
 * nfst_sell_flow_f32 -- start level first: gamma[start] = grad_logz[b] (1 if NULL),
 *   post[a] = gamma[src_a] * cond[a], gamma[dst_a] += post[a]; post may alias cond (in place).
 *   post[a] = grad * d logZ / d w_a, the arc posterior.  Optional: dtheta[V] += post by label (caller
 *   zero-fills); alpha[S] = log(gamma[s] / grad) + logz[b] - beta[s] (needs beta, logz in the launch's state
 *   dtype; -inf where the state posterior underflows fp32).  gamma is accumulated with shared-memory
 *   atomics: posteriors are reproducible to rounding (~1e-7 relative), not bit for bit.
 *   gamma_far[S] (float32, zero-filled by the caller) is required for groups with sell_far and whenever alpha
 *   is requested: it receives the flow of arcs longer than the ring (and, for alpha, into the last level).
 * Both passes stage the arc arrays of every slice in shared memory with 16-byte cp.async copies:
 *   dst_out, label_out, scores->arc_scores and cond must be 16-byte aligned (NFST_ERR_BAD_ARG otherwise; their
 *   lengths need no padding), out_deg8 4-byte aligned and readable up to the next multiple of 4 bytes.
 */
size_t nfst_sell_smem_bytes(const nfst_launch_t* launch, int32_t vocab, int pass, int with_trop, int with_table);
int nfst_sell_pull_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                       void* beta, void* logz_bwd, float* cond, float* delta, int32_t* backptr, float* vit_score,
                       void* cuda_stream);
int nfst_sell_flow_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const float* cond,
                       const float* grad_logz, float* post, const void* beta, const void* logz, void* alpha,
                       float* dtheta, float* gamma_far, void* cuda_stream);

/*
 * Tile-stream passes (launch->tiles != 0); same mathematics as the sliced-column passes above.
 *
 * nfst_tile_pull_f32 -- deepest level first.  Log semiring when beta, logz_bwd or cond is given: beta[S] /
 *   logz_bwd[B] (float32, or float64 when launch->state_f64) and cond[A] = exp(w_a + beta[dst_a] - beta[src_a]);
 *   tropical semiring when backptr is given: delta[S] (optional), backptr[S] (canonical arc id, -1 at sinks),
 *   vit_score[B] (optional); first maximum in label order wins.
 * nfst_tile_flow_f32 -- start level first: post[a] = grad_logz[b] * gamma[src_a] * cond[a] with gamma[start] = 1,
 *   gamma[dst_a] += gamma[src_a] * cond[a] accumulated in 32-bit FIXED POINT (2^-31 units) with native integer
 *   shared-memory atomics: the state posteriors carry an absolute error below 1e-9 and the result is bit
 *   reproducible (integer addition commutes).  post may alias cond.  dtheta[V] += post by label (caller
 *   zero-fills; accumulated per lattice in fixed point, then added with one float atomic per label).  post may be
 *   NULL when dtheta is given: the pass then writes nothing per arc (the theta-mode training step).
 * Arc arrays are fetched with bulk copies of the 16-byte-aligned superset of a tile's range: arc_scores, cond
 *   and label_out must be 16-byte aligned and readable up to the next multiple of 4 elements.
 */
size_t nfst_tile_smem_bytes(const nfst_launch_t* launch, int32_t vocab, int pass, int n_f32_arrays, int with_table,
                            int stages);
int nfst_tile_pull_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const nfst_scores_t* scores,
                       void* beta, void* logz_bwd, float* cond, float* delta, int32_t* backptr, float* vit_score,
                       void* cuda_stream);
int nfst_tile_flow_f32(const nfst_packed_lattices_t* lat, const nfst_launch_t* launch, const float* cond,
                       const float* grad_logz, float* post, float* dtheta, void* cuda_stream);
/* Diagnostic: builds with -DNFST_TILE_DEBUG range-check every data-dependent index of the tile-stream kernels;
 * this synchronises the device and returns the first violation {code, a, b, c, block, thread, 0, 0} (and clears
 * it).  Regular builds return zeros. */
int nfst_tile_debug_read(int32_t* out8);

/*
 * One time step of the lattice-constrained sampling / scoring loop (Sampler.stateful_sample,
 * src/modules/samplers.py:243-297) for n_rows = B * rows_per_lattice rows; replaces, in one launch over the
 * out-arcs of each row's current state, the dense-row gathers of FSAGRUScorer.update_fsa_state
 * (scorers.py:683-690) and mask_out_invalid (:1037-1054), the beta look-ahead of
 * GRUScorer.actual_left_to_right_score (:583-592), pad_masking (:182-187) and the Categorical
 * log_prob / sample / logsumexp of the loop body (samplers.py:251-283).
 *   state[n_rows]       packed id of each row's current state
 *   look_state[n_rows]  NULL: beta look-ahead = beta of each arc's own destination.  Otherwise the reference's
 *                       behaviour (quirk: scorers.py:584 reads the state before :679 advances it): beta of the
 *                       label's successor from look_state (the state BEFORE the previous symbol), or beta of
 *                       the start state where look_state has no arc with that label (transition = 0)
 *   prefix[n_rows, V]   the network's score of every label (beta_scorer output)
 *   base_mask[n_rows,V] vocabulary mask of the scorer base class (scorers.py:314-338), or NULL
 *   beta_real[S]        REAL-space beta per packed state (the reference adds beta itself, not its log)
 *   arc_static[A]       log-weights of weighted emission tables, or NULL (boolean tables)
 *   masked logit of label l at state s = ((l == pad ? 0 : beta[dst] + prefix[l]) + base_mask[l] + static) / T
 *   for the labels on an arc out of s, -inf for all others; a state without arcs is the absorbing sink and
 *   emits pad with probability 1.
 *   given_sym[n_rows]   score these symbols (evaluate_only), or
 *   uniform[n_rows]     sample by inverse CDF over the arcs in label order -- exactly one of the two.
 * Outputs: sym_out, logp_out = log softmax at the symbol (-inf: symbol not allowed; the state then stays),
 * next_state (packed), logz_out = logsumexp of the masked logits (optional).
 */
int nfst_walk_step_f32(const nfst_packed_lattices_t* lat, int32_t n_rows, int32_t rows_per_lattice,
                       const int32_t* state, const int32_t* look_state, const float* prefix, const float* base_mask,
                       const float* beta_real,
                       const float* arc_static, float temperature, int32_t pad_id, const int32_t* given_sym,
                       const float* uniform, int32_t* sym_out, float* logp_out, int32_t* next_state, float* logz_out,
                       void* cuda_stream);

/*
 * Sampler.stripping_pad (src/modules/samplers.py:162-180): left-compacts every row of sequences[n_rows, seq_len]
 * (int64), dropping symbol id 0, exactly as the reference's column loop does -- including its stop at the first
 * column that holds pad in every row and the zero it leaves behind a row's last symbol when the row ends in zeros.
 * out[n_rows, seq_len] (int64; columns >= *width_out are not written), col_flags[seq_len] (int32 scratch),
 * width_out[1] (int32, device): the width of the reference's result.
 */
int nfst_strip_pad(const int64_t* sequences, int32_t n_rows, int32_t seq_len, int64_t pad_id, int64_t* out,
                   int32_t* col_flags, int32_t* width_out, void* cuda_stream);

/*
 * The whole sampling loop for an arc-factored proposal (WFSTScorer, scorers.py:1663-1687): n_rows =
 * B * rows_per_lattice walks from the start state, next arc drawn with probability
 * exp(w_a + beta[dst_a] - beta[s]) (inverse CDF over the state's arcs in label order, uniform[n_rows, max_len]),
 * until a state without arcs.  beta[S]: log-space beta of the same scores (nfst_bwd_fused_f32), float32 or
 * float64 per beta_f64.  The paths are exact posterior samples: log_q[r] = score(path) - logZ, so every
 * importance weight of Estimators.iwae (estimatros.py:10-44) equals logZ.  labels[n_rows, max_len] (padded
 * with pad_id), arcs[n_rows, max_len] canonical arc ids (-1 padded; may be NULL), length[n_rows], log_q[n_rows].
 * max_len >= the deepest lattice's level count - 1.
 */
int nfst_sample_paths_f32(const nfst_packed_lattices_t* lat, int32_t n_rows, int32_t rows_per_lattice, int32_t max_len,
                          const nfst_scores_t* scores, const void* beta, int beta_f64, const float* uniform,
                          int32_t pad_id, int32_t* labels, int32_t* arcs, int32_t* length, float* log_q,
                          void* cuda_stream);

/* One level of the beta-hat recurrence (FSAGRUScorer.compute_beta_per_sample with Wh != 0,
 * scorers.py:732-747), for the `n_states` packed state ids in `states` (all of one topological
 * level, any lattices; call for the deepest level first -- every arc's destination must be done):
 *   m_hat = tanh(label_proj[label] + h_proj[dst]),  log m = w . m_hat + log_beta[dst],
 *   log_beta[s] = logsumexp(log m),  beta_hat[s] = sum softmax(log m) m_hat,  h_proj[s] = Wh beta_hat[s].
 * label_proj [V, hidden] = Wx e_l + b;  wh_t [hidden, hidden] = Wh transposed;  w [hidden];
 * log_beta [S];  beta_hat, h_proj [S, hidden].  hidden <= 1024.  Sinks get beta = 1, beta_hat = 0. */
int nfst_beta_hat_level_f32(const nfst_packed_lattices_t* lat, const int32_t* states, int32_t n_states, int32_t hidden,
                            const float* label_proj, const float* wh_t, const float* w, float* log_beta,
                            float* beta_hat, float* h_proj, void* cuda_stream);

/*
 * Device packer for lattices that fit one SM's shared memory (the size nFST's composed mark lattices have).
 * Input: a raw arc list grouped by lattice and sorted by (source state, label) -- the scan order of the dense
 * tables, which nfst_dense_extract_arcs produces -- with LOCAL destination ids and either local source ids or
 * global rows (src_is_global: raw_state_off[b] + local id).  Output: the arrays of nfst_packed_lattices_t that
 * the small-lattice kernels, the sampling-loop kernels and the best-path read-out use (everything except the
 * chunk lists, the sliced-column descriptors and the tile stream), written into caller-allocated buffers sized
 * for the RAW counts (trimming only shrinks): per-state arrays [n_states_raw (+1)], per-arc arrays [n_arcs_raw],
 * level_ptr [n_states_raw + B], offsets [B + 1].  phases: 1 = count (levels, trimming, per-lattice counts, offsets,
 * totals: two launches; only the per-lattice outputs are touched), 2 = build (one launch; same workspace, same
 * inputs; max_arcs may shrink to the largest kept count), 3 = both -- a caller that reads the totals between the
 * phases can allocate the per-state / per-arc outputs at their exact sizes.  On the caller's stream, no host synchronisation:
 * the caller reads totals[8] = {S, A, level_ptr entries, sinks, error, lattice of the error, largest level count,
 * 0} (error 4: a lattice keeps more than max_arcs arcs -- max_arcs bounds the KEPT arcs of a lattice, max_raw_arcs
 * its raw list, which may be far longer: collate() padding) and lattice_stats[B][8] = {states, arcs, levels, sinks, arcs of the widest level, 0, 0, 0} (zero-filled by the
 * caller) when it needs the sizes.  error: 0 ok, 1 = cyclic
 * lattice, 2 = arc endpoint / start state out of range, 3 = batch beyond int32 indices.
 */
typedef struct nfst_pack_out {
  int32_t *state_off, *arc_off, *level_off, *sink_off; /* [B+1] */
  int32_t *n_levels, *start_state;                     /* [B]   */
  int32_t *level_ptr, *sinks, *orig_state;             /* per state */
  int32_t *in_ptr, *out_ptr;                           /* [S+1] */
  int32_t *src_in, *label_in, *in2out, *dst_out, *label_out, *src_out; /* per arc */
  int64_t *arc_origin;                                 /* per arc: index into the raw arc list */
  uint8_t *out_deg8;                                   /* per state: min(out-degree, 255); may be NULL */
  int32_t *lattice_stats;                              /* [B][8] */
  int32_t *totals;                                     /* [8] */
} nfst_pack_out_t;
size_t nfst_pack_small_smem_bytes(int32_t max_states, int32_t max_arcs);
size_t nfst_pack_small_workspace_bytes(int64_t n_states_raw, int64_t n_arcs_raw);
int nfst_pack_small(int32_t n_lattices, const int32_t* raw_state_off, const int32_t* raw_arc_off, const int32_t* raw_src,
                    const int32_t* raw_dst, const int32_t* raw_label, int32_t src_is_global, int32_t start_state,
                    int32_t max_states, int32_t max_arcs, int32_t max_raw_arcs, const nfst_pack_out_t* out, void* workspace,
                    size_t workspace_bytes, int64_t n_states_raw, int64_t n_arcs_raw, int32_t phases, void* cuda_stream);

/*
 * collate()-padded dense tables -> packed lattices in one call per phase: the edge rule over transition[B, S, V]
 * (int64, scorers.py:704-716), row offsets, the arc list, then nfst_pack_small (same outputs, same phases, same
 * totals / lattice_stats; the start state is row 0, scorers.py:1005).  raw_arc_capacity bounds the arc list the
 * tables produce INCLUDING the arcs of collate() padding rows (V per pad row); more than that sets totals[4] = 5 and
 * nothing is built.  max_arcs bounds the arcs one lattice keeps (shared memory; see nfst_pack_small_smem_bytes).
 * arc_origin of the result is the dense cell of every arc, (b * S + s) * V + label.  The workspace must survive
 * from the count phase to the build phase.
 */
size_t nfst_pack_workspace_bytes(int32_t n_lattices, int32_t states_per_lattice, int64_t raw_arc_capacity);
int nfst_pack_dense(const int64_t* transition, int32_t n_lattices, int32_t states_per_lattice, int32_t vocab,
                    int64_t raw_arc_capacity, int32_t max_arcs, const nfst_pack_out_t* out, void* workspace, size_t workspace_bytes,
                    int32_t phases, void* cuda_stream);

/*
 * On-device construction of the transliteration lattices (what src/preprocess/tr.py:142-190 builds offline with
 * OpenFst: x o T o y for the one-state edit machine of src/fsm/tr.py:321-390, every arc expanded into the chain
 * of its marks, src/modules/path_semiring.py:120-180, bos in front and eos behind).  Writes the arc list of B
 * lattices from the id strings x[B][x_stride] / y[B][y_stride] (lengths x_len / y_len): grouped by lattice at
 * raw_arc_off[b], sorted by (source, label) -- the input nfst_pack_small takes (local source ids).  State
 * numbering: 0 = start, then the (|x|+1) x (|y|+1) grid row-major, the intermediate states of the deletion /
 * insertion / substitution chains, the sink last.  Marks: deletion [input_mark, x_i], insertion
 * [output_mark, y_j], substitution (add_sub) [sub_mark, input_mark, x_i, output_mark, y_j].
 */
void nfst_edit_lattice_size(int32_t n, int32_t m, int32_t add_sub, int64_t* n_states, int64_t* n_arcs);
int nfst_edit_lattice_arcs(int32_t n_lattices, const int32_t* x, const int32_t* x_len, int32_t x_stride, const int32_t* y,
                           const int32_t* y_len, int32_t y_stride, int32_t bos, int32_t eos, int32_t input_mark,
                           int32_t output_mark, int32_t sub_mark, int32_t add_sub, const int32_t* raw_arc_off, int32_t* src,
                           int32_t* dst, int32_t* label, void* cuda_stream);

/*
 * n_sweeps relaxation sweeps of level[dst] = max(level[dst], level[src] + 1) over an arc list with GLOBAL state ids
 * (int64), in place; level[] = 0 at the start states and -1 elsewhere on entry.  *changed (device, caller-zeroed) is
 * set when a sweep moved anything: the caller repeats until it stays 0 -- the longest distance from the start states
 * (the topological levels of the tensor-op packer; more sweeps than states = a cycle).
 */
int nfst_level_sweeps(const int64_t* gsrc, const int64_t* gdst, int64_t n_arcs, int32_t* level, int32_t* changed,
                      int32_t n_sweeps, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* NFST_B200_H_ */
