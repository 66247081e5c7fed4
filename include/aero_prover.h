/*
 * aero_prover.h -- C entry to the host-side prover driver built above the device C ABI.
 *
 * The driver (aero_b200/host/prover.{hpp,cpp}) mirrors winter-prover's orchestration
 * (prover/src/lib.rs:203-540 Prover::generate_proof and its three public stages, and
 * prover/src/channel.rs ProverChannel) with the hot path delegated to include/aero_b200.h.  The two
 * pieces the north star keeps on the reference Rust path -- Miden aux-segment construction
 * (miden/processor/src/trace/mod.rs:188) and AIR constraint evaluation
 * (prover/src/constraints/evaluator.rs:74) -- enter through callbacks, or as precomputed matrices
 * for synthetic workloads.
 */
#ifndef AERO_PROVER_H
#define AERO_PROVER_H

#include "aero_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* air::ProofOptions (air/src/options.rs:20-95); Miden preset = ProofOptions::with_96_bit_security
 * (miden/air/src/options.rs:29-39): 27, 8, 16, Blake2s_256, None, 8, 256. */
typedef struct aero_proof_options {
    uint8_t num_queries;
    uint8_t blowup_factor;
    uint8_t grinding_factor;
    uint8_t hash_fn;            /* 4 = Blake2s_256 (only supported value) */
    uint8_t field_extension;    /* 1 = FieldExtension::None (only supported value) */
    uint8_t fri_folding_factor; /* 8 (only supported value) */
    uint16_t fri_max_remainder_size;
} aero_proof_options;

/* Called after the main-segment commitment with the drawn random elements (ABI form); must fill
 * aux_cols_out[aux_width] with host pointers that stay valid until aero_prove returns.
 * Mirrors Trace::build_aux_segment (prover/src/lib.rs:316-318). */
typedef aero_status (*aero_aux_builder)(void *user, const uint64_t *rand_elements, uint32_t n_rand,
                                        const uint64_t **aux_cols_out);
/* Called after all trace commitments with the natural-order LDE of every trace column (ABI form;
 * main then aux) and the drawn composition coefficients; must fill eval_cols_out[n_div] (host
 * pointers, ce_domain_size entries each).  Mirrors ConstraintEvaluator::evaluate
 * (prover/src/lib.rs:350-382). */
typedef aero_status (*aero_constraint_evaluator)(void *user, const uint64_t *const *trace_lde, uint32_t width,
                                                 uint64_t lde_size, const uint64_t *coeffs, uint32_t n_coeffs,
                                                 const uint64_t **eval_cols_out);

typedef struct aero_prove_inputs {
    aero_proof_options options;
    uint64_t trace_len;
    uint32_t main_width, aux_width, aux_rands;
    /* inputs_on_device == 0: arrays of host column pointers.  != 0: main_cols[0], aux_cols[0] and
     * ce_cols[0] are device pointers to contiguous column-major matrices. */
    int inputs_on_device;
    const uint64_t *const *main_cols;
    const uint64_t *const *aux_cols;   /* used when aux_builder == NULL and aux_width > 0 */
    const uint64_t *const *ce_cols;    /* used when constraint_evaluator == NULL */
    const aero_divisor *divisors;
    uint32_t n_div;
    uint32_t n_constraint_coeffs;      /* field elements drawn for constraint composition */
    /* Constraint evaluation domain = trace_len * ce_blowup (AirContext::ce_blowup_factor,
     * air/src/air/context.rs:124-137; a power of two <= options.blowup_factor); 0 = the LDE domain, as
     * for Miden.  ce_cols / the evaluator's columns hold that many evaluations each (over
     * offset * <g_ce>), and the composition polynomial gets ce_blowup columns. */
    uint32_t ce_blowup;
    aero_aux_builder aux_builder;
    aero_constraint_evaluator constraint_evaluator;
    void *user;
    const uint8_t *pub_inputs_bytes;   /* coin seed input (ProverChannel::new, channel.rs:49-68) */
    size_t pub_inputs_len;
    const uint8_t *trace_meta;         /* TraceInfo meta bytes for the proof context (may be NULL) */
    uint16_t trace_meta_len;
    /* Third source of the constraint evaluations (used when constraint_evaluator == NULL and ce_cols == NULL):
     * the AIR's transition constraints as a program, evaluated on the device from the resident trace LDE
     * (aero_constraints_evaluate_device, include/aero_b200.h) -- nothing is downloaded.  Works with host and
     * with device inputs, and on a sharded context: step s = i * ce_blowup + r of the evaluation domain reads
     * LDE coset r * blowup / ce_blowup only, so every rank evaluates and combines the cosets it holds and
     * only the combined column (8 bytes per step) crosses the exchange window. */
    const aero_air_program *air_program;
} aero_prove_inputs;

/* Prover::prove: writes StarkProof::to_bytes (air/src/proof/mod.rs:122-132) into proof_out.
 * *len: in = capacity, out = bytes written / required (AERO_ERR_BUFFER).  Options are checked like
 * ProofOptions::new (air/src/options.rs:120-160) and rejected with AERO_ERR_INVALID.
 * On a sharded context (aero_ctx_set_shard + exchange window) every rank calls this with the same inputs
 * and gets the same bytes.  aux_builder is called by EVERY rank there (each has drawn the same random
 * elements; the ranks of an aero_group take turns) and must return the same columns each time;
 * constraint_evaluator is not available on a sharded context (AERO_ERR_UNSUPPORTED): the constraint
 * evaluations come from air_program or precomputed. */
aero_status aero_prove(aero_ctx *ctx, const aero_prove_inputs *in, uint8_t *proof_out, size_t *len);

/* ONE proof on several GPUs of this process: n_ranks contexts (one per entry of device_ids; the same
 * device may appear more than once, which is how the single-GPU tests run the sharded path) joined by an
 * exchange window of window_bytes each (size: include/aero_b200.h, multi-GPU section).  aero_group_prove
 * runs aero_prove on every rank, one host thread per rank, and returns the common proof.  in: one
 * aero_prove_inputs shared by all ranks (host inputs: every rank uploads only the columns it
 * interpolates) or n_ranks entries (n_inputs == n_ranks: e.g. device-resident inputs per GPU).
 * This is the reference's rayon "concurrent" feature across GPUs: Prover::prove stays one call. */
typedef struct aero_group aero_group;
aero_status aero_group_create(const int *device_ids, int n_ranks, size_t window_bytes, aero_group **out);
void aero_group_destroy(aero_group *g);
int aero_group_size(aero_group *g);
aero_ctx *aero_group_ctx(aero_group *g, int rank);
const char *aero_group_last_error(aero_group *g);
aero_status aero_group_prove(aero_group *g, const aero_prove_inputs *in, int n_inputs, uint8_t *proof_out, size_t *len);

/* Host Fiat-Shamir primitives, exported for tests (crypto/src/random/mod.rs:73-306,
 * crypto/src/hash/blake2s/mod.rs:33-77).  Elements canonical. */
void aero_host_blake2s(const uint8_t *data, size_t len, uint8_t out[32]);
void aero_host_hash_elements(const uint64_t *canonical_elements, size_t count, uint8_t out[32]);
typedef struct aero_coin aero_coin;
aero_coin *aero_coin_new(const uint8_t *seed_bytes, size_t len);
void aero_coin_free(aero_coin *c);
void aero_coin_reseed(aero_coin *c, const uint8_t digest[32]);
void aero_coin_reseed_with_int(aero_coin *c, uint64_t v);
aero_status aero_coin_draw(aero_coin *c, uint64_t *canonical_out);
aero_status aero_coin_draw_integers(aero_coin *c, uint32_t num_values, uint64_t domain_size, uint64_t *out);
uint32_t aero_coin_leading_zeros(aero_coin *c);
uint32_t aero_coin_check_leading_zeros(aero_coin *c, uint64_t v);
void aero_coin_seed(aero_coin *c, uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif /* AERO_PROVER_H */
