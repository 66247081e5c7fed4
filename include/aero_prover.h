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

/* Multi-GPU exchange hooks (NULL on a single GPU).  all_gather_cosets completes a device buffer that
 * every rank filled for its own LDE cosets: interleaved != 0 -> layout [outer][B][inner_bytes]
 * (leaf digests), else [B][outer][inner_bytes] (DEEP evaluations).  sum_rows adds the ranks'
 * disjoint host row matrices (aero_segment_open fills foreign rows with zeros). */
typedef aero_status (*aero_all_gather_cosets)(void *user, void *d_buf, uint64_t outer, uint32_t n_cosets,
                                              uint32_t inner_bytes, int interleaved, uint32_t coset_begin,
                                              uint32_t coset_count);
typedef aero_status (*aero_sum_rows)(void *user, uint64_t *host_rows, uint64_t count);

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
    aero_aux_builder aux_builder;
    aero_constraint_evaluator constraint_evaluator;
    void *user;
    const uint8_t *pub_inputs_bytes;   /* coin seed input (ProverChannel::new, channel.rs:49-68) */
    size_t pub_inputs_len;
    const uint8_t *trace_meta;         /* TraceInfo meta bytes for the proof context (may be NULL) */
    uint16_t trace_meta_len;
    aero_all_gather_cosets all_gather_cosets;
    aero_sum_rows sum_rows;
} aero_prove_inputs;

/* Prover::prove: writes StarkProof::to_bytes (air/src/proof/mod.rs:122-132) into proof_out.
 * *len: in = capacity, out = bytes written / required (AERO_ERR_BUFFER). */
aero_status aero_prove(aero_ctx *ctx, const aero_prove_inputs *in, uint8_t *proof_out, size_t *len);

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
