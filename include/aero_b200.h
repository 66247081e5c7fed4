/*
 * aero_b200.h -- C ABI of the B200-native LDE / row-commitment / DEEP / FRI core.
 *
 * This is the drop-in boundary for the data-parallel hot path of the Miden/Winterfell prover that
 * starkoracles/Aero drives.  The reference has no FFI for this path (it is all Rust); each entry
 * point below names the Rust interface it replaces (paths relative to the Aero checkout, the
 * `winterfell/` prefix omitted for winterfell crates).  INTEGRATION.md shows the Rust `extern "C"`
 * shim a maintainer would add.
 *
 * Conventions
 *  - Every call returns an aero_status; non-zero means failure and aero_last_error(ctx) holds a
 *    message.  Nothing aborts or throws across the ABI (the reference asserts/panics instead:
 *    prover/src/matrix.rs:42-61, math/src/fft/mod.rs:179-199).
 *  - Field elements crossing the ABI are u64 in the form selected by aero_ctx_set_form():
 *    AERO_FORM_MONTGOMERY (default; the in-memory image of math::fields::f64::BaseElement,
 *    math/src/field/f64/mod.rs:59-61) or AERO_FORM_CANONICAL.  Outputs documented as "canonical"
 *    are always canonical little-endian integers, as on the proof wire (f64/mod.rs:532-535).
 *  - Digests are raw 32-byte BLAKE2s-256 values (crypto/src/hash/blake2s/mod.rs:33-46).
 *  - Matrices are column-major like prover::Matrix (prover/src/matrix.rs:26-28).
 *  - Caller owns host buffers; the library owns aero_* handles until the matching *_destroy.
 *  - A context drives one GPU and one proof at a time; contexts are independent.
 *  - There is no CPU fallback: every entry point fails with AERO_ERR_CUDA if no device is usable.
 */
#ifndef AERO_B200_H
#define AERO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int aero_status;
enum {
    AERO_OK = 0,
    AERO_ERR_INVALID = 1,     /* bad argument (non power of two, zero width, out-of-range index...) */
    AERO_ERR_CUDA = 2,        /* CUDA runtime error or no device */
    AERO_ERR_NOMEM = 3,
    AERO_ERR_STATE = 4,       /* call sequence violated (e.g. fold before commit) */
    AERO_ERR_UNSUPPORTED = 5, /* shape outside what the kernels implement */
    AERO_ERR_BUFFER = 6       /* output buffer too small; required size is stored in *len */
};
enum { AERO_FORM_MONTGOMERY = 0, AERO_FORM_CANONICAL = 1 };

typedef struct aero_ctx aero_ctx;
typedef struct aero_segment aero_segment;
typedef struct aero_fri aero_fri;
typedef struct aero_upload aero_upload;

/* (x^a - b) / prod_k (x - exemptions[k]) : air::ConstraintDivisor (air/src/air/divisor.rs:14-17) */
typedef struct aero_divisor {
    uint64_t a;
    uint64_t b;              /* field element, ABI form */
    uint32_t n_exemptions;   /* <= 8 */
    uint64_t exemptions[8];  /* field elements, ABI form */
} aero_divisor;

/* ---- context --------------------------------------------------------------------------------- */
/* One context drives one GPU: n_devices must be 1 (device_ids == NULL selects the current device).
 * Several GPUs work on one proof as one context per GPU joined by an exchange window: see the
 * multi-GPU section below, and aero_group_create (include/aero_prover.h) for the one-process form. */
aero_status aero_ctx_create(const int *device_ids, int n_devices, aero_ctx **out);
void aero_ctx_destroy(aero_ctx *ctx);
const char *aero_last_error(aero_ctx *ctx);
/* Launch everything on this cudaStream_t (default: the legacy default stream). */
aero_status aero_ctx_set_stream(aero_ctx *ctx, void *cuda_stream);
aero_status aero_ctx_set_form(aero_ctx *ctx, int form);
int aero_ctx_get_form(aero_ctx *ctx);
/* Execution knobs (results never change, only scheduling):
 *   "overlap_hash"    0 (default) / 1: hash the rows of column batch k on a second stream while batch
 *                     k+1 is extended (commit_to_rows, matrix.rs:222, interleaved with
 *                     evaluate_columns_over, :189) -- 0 serialises everything on the context's stream
 *                     (faster on B200);
 *   "hash_blocks_per_sm" grid cap (blocks per SM) of an overlapped row-hash launch (default 2);
 *   "lde_batch_bytes" NTT scratch budget per column batch (default 1 GiB);
 *   "ntt_table_max_bytes" largest full inter-pass twiddle table (8 bytes per output element of one
 *                     column) a two-pass transform plan may keep (default 1 GiB; 0 = never, the factor
 *                     is then advanced by a running product); read when a plan is first built;
 *   "own_stream"      1: create a non-blocking stream for this context and launch on it (contexts that
 *                     share a device or a process must not meet on the legacy default stream);
 *   "force_host_sync" 1: sharded proofs keep host-synchronised barriers even for warm shapes (tests);
 *   "push_parts"      1..4 (default 4): copy streams a rank's coefficient block is split over when it is sent
 *                     to a peer (one copy engine does not fill NVLink);
 *   "ntt_outer_log"   -1 (default): transforms above 2^20 points split off a third factor 2^(logn-20) (three-pass
 *                     plans); 0: two passes only; 1..4: force that factor wherever the
 *                     transform has at least 2^(12+k) points (tests); read when a plan is first built;
 *   "hash_early_batches" 0..16 (default 2): host-buffer commits of >= 32 columns hash this many upload
 *                     batches right after their extension (chained row hash) so that the stream does not
 *                     wait for later batches;
 *   "fri_fused"       0 (default) / 1 / 2: fold a FRI layer and hash the next layer's leaves in one kernel --
 *                     never / layers of <= 2^14 folded leaves / all layers (measured slower on B200);
 *   "air_blocks_per_sm" 0 (default): the AIR evaluator runs one thread per evaluation step; n > 0: n resident blocks
 *                     per SM, each thread looping over steps (keeps the value slots in L2; measured slower).
 * Environment hooks for experiments (read once): AERO_NTT_BULK=0 (LDGSTS instead of TMA tile loads),
 * AERO_NTT_COLFAST=0 (tile-major pass-1 grid), AERO_NTT_OUTER, AERO_HASH_EARLY, AERO_FRI_FUSED (defaults of the
 * options above), AERO_HOST_TIMING=1 (host-side section marks of aero_prove on stderr). */
aero_status aero_ctx_set_option(aero_ctx *ctx, const char *key, long long value);
/* Used by the host driver layered above this ABI to report its own failures through aero_last_error. */
void aero_ctx_set_error(aero_ctx *ctx, const char *msg);
/* Per-phase CUDA-event timing (mirrors the reference's debug! timers, prover/src/lib.rs:228-630).
 * aero_ctx_profile_read writes a JSON object {"phase": [calls, total_ms], ...} and resets. */
aero_status aero_ctx_profile_enable(aero_ctx *ctx, int enable);
/* Restricts the timing to phases whose name starts with `prefix` (NULL or "" = all phases).  Every
 * timed phase costs two event records on the stream (~4 us of device time): a benchmark that needs
 * one kernel's duration inside its timed region filters on that phase and reads the rest in a
 * separate pass. */
aero_status aero_ctx_profile_filter(aero_ctx *ctx, const char *prefix);
aero_status aero_ctx_profile_read(aero_ctx *ctx, char *json_out, size_t *len);
/* Number of CUDA kernels this library has launched in this process. */
uint64_t aero_launch_count(void);
const char *aero_version(void);

/* ---- trace / constraint segment: iNTT + coset LDE + row hashes + Merkle tree ------------------- */
/* Replaces Prover::build_trace_commitment (prover/src/lib.rs:551-589) =
 * Matrix::interpolate_columns (matrix.rs:151) + evaluate_columns_over (:189) + commit_to_rows
 * (:222), and with input_is_coeffs != 0 Prover::build_constraint_commitment (lib.rs:599-632).
 * cols: n_cols host pointers to n_rows elements each.  root receives MerkleTree::root(). */
aero_status aero_segment_commit(aero_ctx *ctx, const uint64_t *const *cols, uint32_t n_cols, uint64_t n_rows,
                                uint32_t blowup, int input_is_coeffs, aero_segment **out, uint8_t root[32]);
/* root may be NULL in every commit call: nothing is synchronised then and the root is collected later with
 * aero_segments_roots -- one host round trip for several commitments whose roots the caller does not need
 * in between (trace segments that do not depend on each other's randomness).
 * Same with the matrix already resident in device memory (column c at d_cols + c*col_stride). */
aero_status aero_segment_commit_device(aero_ctx *ctx, const uint64_t *d_cols, size_t col_stride, uint32_t n_cols,
                                       uint64_t n_rows, uint32_t blowup, int input_is_coeffs, aero_segment **out,
                                       uint8_t root[32]);
/* ---- multi-GPU: ONE proof over G GPUs (one context = one rank) -----------------
 * After aero_ctx_set_shard(rank, G) (G a power of two dividing the blowup, G <= 8) the same call
 * sequence on every rank produces one proof, byte-identical to the single-GPU one:
 *   - trace columns are interpolated by the rank that owns them (columns [rank*w/G, (rank+1)*w/G)),
 *     which is also the only rank that uploads them, and the coefficients are stored into every peer's
 *     copy of the coefficient matrix as they are produced;
 *   - every rank extends and row-hashes LDE cosets [rank*B/G, (rank+1)*B/G) of ALL columns -- complete
 *     rows k with k mod B in that range, so no LDE value ever moves;
 *   - the row-hash kernel stores each leaf digest straight into the rank that owns the leaf's block
 *     [r*N/G, (r+1)*N/G); each rank builds the subtree over its block (the reference's concurrent
 *     build_merkle_nodes splits the tree the same way, crypto/src/merkle/concurrent.rs:21-70), the G
 *     sub-roots are exchanged and every rank finishes the top log2(G) levels;
 *   - aero_deep_compose extends the DEEP polynomial over the own cosets and stores them into the peers;
 *     composition, OOD, DEEP coefficients and the one-column FRI are replicated so every rank's
 *     Fiat-Shamir coin stays in lock-step without a broadcast;
 *   - openings: the rank that stores an opened row / tree node writes it into every rank's result buffer.
 * All of this goes through the EXCHANGE WINDOW: one device allocation per rank, the same size
 * everywhere, mapped by every peer -- over CUDA IPC between processes (aero_ctx_window_attach; the
 * caller all-gathers the 64-byte handles with any transport), directly inside one process
 * (aero_ctx_window_attach_local; include/aero_prover.h wraps this as aero_group_*).  Buffers other ranks
 * write into are bump-allocated inside the window, at the same offset on every rank.  Size for a proof
 * of n rows (N = B*n): 8*n*(all trace columns) for the coefficients + 3 * 32*N/G for the leaf blocks +
 * 8*N combined constraint evaluations + 24*n DEEP accumulators + 8*N DEEP evaluations + 16 MiB
 * (aero_b200/sharded.py window_bytes).
 * Barriers between the ranks are stream-ordered flag barriers on the device (a rank that never arrives
 * is reported as AERO_ERR_STATE after ~4 s instead of hanging).  The first proof of a shape on a context
 * still calls cudaMalloc, which can block on a peer that already spins in such a barrier (the caveat
 * NCCL documents for its own kernels); it therefore synchronises on the host instead, through the
 * rendezvous the caller provides with aero_ctx_set_host_barrier.  aero_ctx_shard_begin / _end bracket a
 * proof: begin(shape_key) selects the barrier kind (device-side once `shape_key` has completed before on
 * this context), end(ok) records the outcome.  Outside such a bracket barriers are host-synchronised. */
typedef aero_status (*aero_host_barrier_fn)(void *user);
aero_status aero_ctx_set_shard(aero_ctx *ctx, int rank, int world);
/* ipc_handle_out may be NULL when the window is only attached inside this process. */
aero_status aero_ctx_window_create(aero_ctx *ctx, size_t bytes, uint8_t ipc_handle_out[64]);
aero_status aero_ctx_window_attach(aero_ctx *ctx, int n_ranks, const uint8_t *ipc_handles /* n_ranks x 64 */);
/* Same-process variant: ranks[r] is the context of rank r (ranks[own rank] == ctx); peer access between
 * the devices is enabled as needed.  Contexts may share a device (used by the single-GPU tests). */
aero_status aero_ctx_window_attach_local(aero_ctx *ctx, int n_ranks, aero_ctx *const *ranks);
int aero_ctx_window_ranks(aero_ctx *ctx);
aero_status aero_ctx_set_host_barrier(aero_ctx *ctx, aero_host_barrier_fn barrier, void *user);
aero_status aero_ctx_shard_begin(aero_ctx *ctx, const char *shape_key);
aero_status aero_ctx_shard_end(aero_ctx *ctx, int ok);
aero_status aero_window_barrier(aero_ctx *ctx);
/* MerkleTree::root() of n_segs committed segments (roots_out: n_segs x 32 bytes), one synchronisation. */
aero_status aero_segments_roots(aero_ctx *ctx, aero_segment *const *segs, uint32_t n_segs, uint8_t *roots_out);
void aero_segment_destroy(aero_segment *seg);
aero_status aero_segment_info(aero_segment *seg, uint32_t *n_cols, uint64_t *n_rows, uint32_t *blowup);
/* Natural-order LDE columns (lde[c][k] = poly_c(7 * g_N^k), matrix.rs:189-201) for the host-side
 * AIR evaluator (TraceLde, prover/src/trace/trace_lde.rs:15-111). */
aero_status aero_segment_download_lde(aero_segment *seg, uint64_t *const *cols_out);
/* Coefficient columns (TracePolyTable, prover/src/trace/poly_table.rs:21-57). */
aero_status aero_segment_download_polys(aero_segment *seg, uint64_t *const *cols_out);
/* Leaf digests in natural order (MerkleTree::leaves, crypto/src/merkle/mod.rs:141-143). */
aero_status aero_segment_download_leaves(aero_segment *seg, uint8_t *leaves_out);
/* Queries: rows at `positions` (canonical, row-major n_pos x n_cols; trace/commitment.rs:115-140,
 * constraints/commitment.rs:54-70) and the batch Merkle proof in BatchMerkleProof::serialize_nodes
 * format (crypto/src/merkle/mod.rs:188-250, merkle/proofs.rs:421-439).  *len: in = capacity of
 * batch_nodes_out, out = bytes written / required. */
aero_status aero_segment_open(aero_segment *seg, const uint64_t *positions, uint32_t n_pos, uint64_t *rows_out,
                              uint8_t *batch_nodes_out, size_t *len);

/* ---- constraint evaluations -> composition polynomial columns ---------------------------------- */
/* Replaces ConstraintEvaluationTable::into_poly + CompositionPoly::new
 * (prover/src/constraints/evaluation_table.rs:166-190,330-419; composition_poly.rs:21-49,111-128).
 * eval_cols: n_div host columns of ce_domain_size merged evaluations.  The result is a segment that
 * holds ce_domain_size/trace_len coefficient columns; commit it with aero_segment_commit_polys. */
aero_status aero_constraints_into_poly(aero_ctx *ctx, const uint64_t *const *eval_cols, const aero_divisor *divs,
                                       uint32_t n_div, uint64_t ce_domain_size, uint64_t trace_len,
                                       aero_segment **composition_polys);
/* Same with the evaluation columns resident in device memory (column d at d_eval_cols + d*col_stride). */
aero_status aero_constraints_into_poly_device(aero_ctx *ctx, const uint64_t *d_eval_cols, size_t col_stride,
                                              const aero_divisor *divs, uint32_t n_div, uint64_t ce_domain_size,
                                              uint64_t trace_len, aero_segment **composition_polys);
/* ---- AIR constraint evaluation on the device ------------------------------------- */
/* Replaces ConstraintEvaluator::evaluate (prover/src/constraints/evaluator.rs:74-230; boundary.rs:59-100,
 * 255-275; air/src/air/transition/mod.rs:224-283 merge_evaluations; domain.rs:99-117) for an AIR whose
 * transition constraints are handed over as an arithmetic program over the evaluation frame, so that the
 * trace LDE never leaves the device (the callback route downloads all of it: 5.4 GB at 2^20 rows).
 *
 * The program is a list of nodes in evaluation order; node k may refer to nodes < k:
 *   AERO_AIR_CUR / AERO_AIR_NEXT  a = trace column (segments concatenated: main, then auxiliary) of the
 *                                  current / next row of the frame (EvaluationFrame, air/src/air/mod.rs)
 *   AERO_AIR_CONST                a = index into consts (ABI form: public inputs, constants of the AIR, the
 *                                  auxiliary segment's random elements -- consts is read when the evaluator
 *                                  runs, so inside aero_prove a caller stores the random elements from its
 *                                  aux_builder callback, which has returned by then)
 *   AERO_AIR_PERIODIC             a = periodic column (Air::get_periodic_column_values,
 *                                  air/src/air/mod.rs:246-248): the value periodic_values[a] of
 *                                  Air::evaluate_transition at this step
 *   AERO_AIR_ADD / SUB / MUL      a, b = operand nodes
 * A Rust caller records it once per AIR by running Air::evaluate_transition over a symbolic element type.
 * transition_out[i] = node holding constraint i, transition_adj[i] = its group's degree adjustment
 * (TransitionConstraintGroup::degree_adjustment); boundary constraint j is the single-value assertion
 * column boundary_col[j] == boundary_value[j] (ABI form) with degree adjustment boundary_adj[j], merged
 * into evaluation column boundary_div[j] >= 1 (column 0 is the transition divisor's, evaluator.rs:66-67).
 * coeffs: the drawn composition coefficients in ABI form, a pair per transition constraint, then a pair
 * per boundary constraint in the order given (Air::get_constraint_composition_coefficients,
 * air/src/air/mod.rs:511-533).
 * Periodic columns are handed over as their cycle values (column k: periodic_len[k] values, a power of two
 * in 2..trace_len, concatenated in periodic_values, ABI form); the library builds PeriodicValueTable
 * (prover/src/constraints/periodic_table.rs:25-90): each column interpolated over its cycle
 * (Air::get_periodic_column_polys, air/src/air/mod.rs:310-344) and evaluated over the coset
 * offset^(trace_len / cycle) * <w_(cycle * ce_blowup)>, looked up by step modulo cycle * ce_blowup.  The degree
 * adjustments the caller passes already reflect the cycles (TransitionConstraintDegree::with_cycles).
 * Sequence and periodic ASSERTIONS are not covered (Miden's ProcessorAir makes single-value assertions only,
 * miden/air/src/lib.rs:124-160).
 * Output: n_div columns of trace_len * ce_blowup merged evaluations in natural order of the constraint
 * evaluation domain (ABI form, column d at d_eval_cols + d * col_stride): the input of
 * aero_constraints_into_poly_device.  On a sharded context only the steps s with s mod ce_blowup among the
 * rank's cosets are written (aero_constraints_evaluate_into_poly combines and exchanges them).  At most 65536 nodes of which at most 1024 values alive at once (slots
 * are assigned by liveness), 32 distinct degree adjustments. */
enum { AERO_AIR_CUR = 0, AERO_AIR_NEXT = 1, AERO_AIR_CONST = 2, AERO_AIR_ADD = 3, AERO_AIR_SUB = 4, AERO_AIR_MUL = 5,
       AERO_AIR_PERIODIC = 6 };
typedef struct aero_air_node {
    uint32_t op, a, b;
} aero_air_node;
typedef struct aero_air_program {
    const aero_air_node *nodes;
    uint32_t n_nodes;
    const uint64_t *consts;
    uint32_t n_consts;
    uint32_t n_transition;
    const uint32_t *transition_out;
    const uint64_t *transition_adj;
    uint32_t n_boundary;
    const uint32_t *boundary_col;
    const uint64_t *boundary_value;
    const uint64_t *boundary_adj;
    const uint32_t *boundary_div;
    uint32_t n_periodic;             /* 0 = none (the three fields below may then be 0 / NULL) */
    const uint32_t *periodic_len;    /* cycle length of each periodic column */
    const uint64_t *periodic_values; /* sum(periodic_len) values, column after column */
} aero_air_program;
aero_status aero_constraints_evaluate_device(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs,
                                             const aero_air_program *program, const uint64_t *coeffs, uint32_t n_coeffs,
                                             uint32_t ce_blowup, uint32_t n_div, uint64_t *d_eval_cols, size_t col_stride);
/* ConstraintEvaluator::evaluate followed by ConstraintEvaluationTable::into_poly, the evaluation table kept in
 * device scratch: the route aero_prove takes for aero_prove_inputs.air_program. */
aero_status aero_constraints_evaluate_into_poly(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs,
                                                const aero_air_program *program, const uint64_t *coeffs, uint32_t n_coeffs,
                                                uint32_t ce_blowup, const aero_divisor *divs, uint32_t n_div,
                                                aero_segment **composition_polys);
/* One column of PeriodicValueTable::new (prover/src/constraints/periodic_table.rs:25-75) as the evaluator above
 * builds it: cycle_len CANONICAL values in, cycle_len * ce_blowup canonical values out (the column polynomial of
 * Air::get_periodic_column_polys over offset^(trace_len / cycle_len) * <w_(cycle_len * ce_blowup)>; the value
 * at constraint-evaluation step s is out[s % (cycle_len * ce_blowup)]).  Host arithmetic only: needs no context
 * and no GPU. */
aero_status aero_periodic_column_table(const uint64_t *cycle_values, uint64_t cycle_len, uint64_t trace_len,
                                       uint32_t ce_blowup, uint64_t *out);

/* Extends + commits a coefficient-only segment: CompositionPoly::evaluate + commit_to_rows
 * (prover/src/lib.rs:599-632). */
aero_status aero_segment_commit_polys(aero_segment *seg, uint32_t blowup, uint8_t root[32]);

/* ---- out-of-domain evaluation and DEEP composition --------------------------------------------- */
/* TracePolyTable::get_ood_frame (prover/src/trace/poly_table.rs:69-72) and
 * CompositionPoly::evaluate_at (constraints/composition_poly.rs:93-96).  out_trace: 2*W values,
 * row z (segments in order) then row z*g_n; out_comp: m values at z^m.  comp may be NULL. */
aero_status aero_ood_eval(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs, aero_segment *comp,
                          uint64_t z, uint64_t *out_trace, uint64_t *out_comp);
/* DeepCompositionPoly::{add_trace_polys, add_composition_poly, adjust_degree, evaluate}
 * (prover/src/composer/mod.rs:71-252).  ood_trace: 2*W, ood_comp: m, cc: W triples, then m, then 2
 * (air::DeepCompositionCoefficients, air/src/air/mod.rs:537-561).  The result is a FRI prover
 * primed with the DEEP evaluations over the LDE domain. */
aero_status aero_deep_compose(aero_ctx *ctx, aero_segment *const *trace_segs, uint32_t n_trace_segs,
                              aero_segment *comp, uint64_t z, const uint64_t *ood_trace, const uint64_t *ood_comp,
                              const uint64_t *cc, aero_fri **out);
/* DEEP polynomial coefficients / evaluations (natural order) for inspection and tests. */
aero_status aero_fri_download_evaluations(aero_fri *fri, uint64_t *out, uint64_t *count);

/* ---- FRI (folding factor 8), stepwise so Fiat-Shamir stays with the caller ---------------------- */
/* FriProver::build_layer split at the channel round trip (fri/src/prover/mod.rs:197-218):
 * commit = transpose_slice + hash_values + MerkleTree::new; fold = apply_drp with offset 7. */
aero_status aero_fri_from_evaluations(aero_ctx *ctx, const uint64_t *evaluations, uint64_t count, aero_fri **out);
aero_status aero_fri_commit_layer(aero_fri *fri, uint8_t root[32]);
aero_status aero_fri_fold(aero_fri *fri, uint64_t alpha);
/* FriProver::build_layers (fri/src/prover/mod.rs:166-191) in one call and ONE host round trip: all
 * num_layers + 1 commitments (the last one is the remainder) and num_layers folds, with the channel's
 * part -- commit_fri_layer = coin.reseed(root), draw_fri_alpha = coin.draw()
 * (prover/src/channel.rs:172-184, crypto/src/random/mod.rs:105-108,179-196) -- evaluated on the device
 * between the kernels.  coin_seed: the public coin's seed when FRI starts.  roots_out:
 * (num_layers + 1) x 32 bytes; alphas_out: the num_layers + 1 challenges drawn (ABI form), so that
 * the caller replays its own channel (reseed + draw per layer) and checks it agrees. */
aero_status aero_fri_build_layers(aero_fri *fri, const uint8_t coin_seed[32], uint32_t num_layers, uint8_t *roots_out,
                                  uint64_t *alphas_out);
/* The same plus ProverChannel::grind_query_seed (prover/src/channel.rs:151-167) in that round trip: the
 * coin's seed after the last layer commitment is already on the device, so the search for the minimum
 * nonce (aero_pow_min_nonce) is queued right behind the layers. */
aero_status aero_fri_build_layers_grind(aero_fri *fri, const uint8_t coin_seed[32], uint32_t num_layers, uint32_t grinding_bits,
                                        uint8_t *roots_out, uint64_t *alphas_out, uint64_t *nonce_out);
/* FriProver::build_proof (fri/src/prover/mod.rs:231-275) in FriProof::write_into format
 * (fri/src/proof.rs:201-214,351-359); the last committed layer is the remainder. */
aero_status aero_fri_open(aero_fri *fri, const uint64_t *positions, uint32_t n_pos, uint8_t *fri_proof_bytes,
                          size_t *len);
void aero_fri_destroy(aero_fri *fri);
/* The whole query phase of Prover::prove_after_constraint_eval (prover/src/lib.rs:518-539:
 * fri_prover.build_proof, trace_commitment.query, constraint_commitment.query) in one call and ONE
 * host round trip: aero_fri_open (fri may be NULL) plus aero_segment_open for each of `segs`, all at
 * the same `positions`.  rows_out[i] / batch_nodes_out[i] / batch_len[i] are segment i's
 * aero_segment_open arguments; on AERO_ERR_BUFFER every *len holds the size its buffer needs. */
aero_status aero_open_queries(aero_ctx *ctx, aero_fri *fri, aero_segment *const *segs, uint32_t n_segs,
                              const uint64_t *positions, uint32_t n_pos, uint8_t *fri_proof_bytes, size_t *fri_len,
                              uint64_t *const *rows_out, uint8_t *const *batch_nodes_out, size_t *batch_len);

/* ---- auxiliary-segment construction: running-product columns ------------------------------------- */
/* The data-parallel core of Trace::build_aux_segment for Miden (miden/processor/src/trace/mod.rs:188-249):
 * every auxiliary column is a running product over the table updates of the main trace, built by
 * build_aux_column (processor/src/trace/utils.rs:153-199):
 *     col[0] = init[c],   col[i + 1] = col[i] * multiplicands[c][i]     (multiplicand 1 on rows without an update)
 * -- an exclusive prefix product, a blocked scan here.  multiplicands[c] has n_rows - 1 entries that matter
 * (ABI form).  Which rows update which table, and the row values mixed from the random elements, are Miden
 * VM logic and stay with the caller; aero_batch_inverse serves the inversions of those row values
 * (build_lookup_table_row_values, utils.rs; math::batch_inversion semantics: zero maps to zero).
 * The *_device form takes and leaves the columns in device memory (column c at base + c*stride), so the
 * result can go straight into aero_segment_commit_device without crossing PCIe again. */
aero_status aero_running_product_columns(aero_ctx *ctx, const uint64_t *const *multiplicands, const uint64_t *init,
                                         uint32_t n_cols, uint64_t n_rows, uint64_t *const *cols_out);
aero_status aero_running_product_columns_device(aero_ctx *ctx, const uint64_t *d_multiplicands, size_t m_stride,
                                                const uint64_t *init, uint32_t n_cols, uint64_t n_rows, uint64_t *d_out,
                                                size_t out_stride);
aero_status aero_batch_inverse(aero_ctx *ctx, const uint64_t *values, uint64_t count, uint64_t *out);

/* ---- grinding ---------------------------------------------------------------------------------- */
/* ProverChannel::grind_query_seed, serial build (prover/src/channel.rs:151-167): smallest nonce >= 1
 * with trailing_zeros(LE64(merge_with_int(seed, nonce)[0..8])) >= grinding_bits. */
aero_status aero_pow_min_nonce(aero_ctx *ctx, const uint8_t seed[32], uint32_t grinding_bits, uint64_t *nonce);

/* ---- standalone primitives (microbenchmarks, tests) --------------------------------------------- */
/* Matrix::commit_to_rows on a natural-order device matrix (column c at d_m + c*col_stride). */
aero_status aero_commit_rows_device(aero_ctx *ctx, const uint64_t *d_m, size_t col_stride, uint32_t n_cols,
                                    uint64_t n_rows, uint8_t root[32]);
/* Self-test of the device Goldilocks arithmetic (math/src/field/f64/mod.rs:273-330): for n operand
 * pairs (any u64; reduced mod p first) writes out[0..n) = a*b, out[n..2n) = a+b, out[2n..3n) = a-b
 * (canonical), out[3n..4n) = a*b computed from the unreduced operands, out[(4+k)n..(5+k)n) =
 * a * 2^(12(k+1)) for k = 0..6 from the unreduced a (the power-of-two twiddles inside the NTT
 * rounds), and out[11n..12n) = a+b through the canonical-sum adder; out holds 12n words. */
aero_status aero_test_field_ops(aero_ctx *ctx, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);
/* Roofline denominator of this path, measured on the spot: issue rate of the ALU pipe
 * (LOP3/SHF/PRMT/IADD3, what BLAKE2s and the butterflies' carry chains saturate) in lane-operations
 * per second, from a ~1 ms stream of independent LOP3 at 8 warps per scheduler. */
aero_status aero_measure_alu_peak(aero_ctx *ctx, double *lane_ops_per_s);
/* Device scratch helpers so callers without a CUDA runtime binding can stage data. */
aero_status aero_device_alloc(aero_ctx *ctx, size_t bytes, void **d_ptr);
aero_status aero_device_free(aero_ctx *ctx, void *d_ptr);
aero_status aero_device_upload(aero_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
aero_status aero_device_download(aero_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
aero_status aero_device_sync(aero_ctx *ctx);
/* Prefetched host->device upload of a column-major matrix (n_cols host pointers, ABI form) on the
 * context's copy stream, so that it lands while earlier phases compute: e.g. the auxiliary segment and
 * the constraint evaluations of Prover::prove travel under the main segment's NTTs.  defer != 0 queues
 * the copies behind the uploads of the next aero_segment_commit instead of ahead of them.
 * shard_mode: what a sharded context copies -- AERO_UPLOAD_OWN_COLUMNS: only the columns it interpolates itself
 * (for a trace matrix consumed by aero_segment_commit_device, which reads exactly those);
 * AERO_UPLOAD_OWN_ROWS: only its block of rows [rank*n_rows/G, (rank+1)*n_rows/G) of every column (for the
 * constraint evaluations consumed by aero_constraints_into_poly_device, likewise); the rest of the device
 * matrix stays unwritten.  AERO_UPLOAD_ALL (and any mode on an unsharded context) copies everything.
 * aero_upload_wait orders the context's stream after the copy and returns the device matrix (column c
 * at d_cols + c*n_rows) for the *_device entry points.  The host columns must stay valid until
 * aero_upload_free, which also releases the device block. */
enum { AERO_UPLOAD_ALL = 0, AERO_UPLOAD_OWN_COLUMNS = 1, AERO_UPLOAD_OWN_ROWS = 2 };
aero_status aero_upload_start(aero_ctx *ctx, const uint64_t *const *cols, uint32_t n_cols, uint64_t n_rows, int defer,
                              int shard_mode, aero_upload **out);
aero_status aero_upload_wait(aero_upload *up, const uint64_t **d_cols);
void aero_upload_free(aero_upload *up);

#ifdef __cplusplus
}
#endif
#endif /* AERO_B200_H */
