// Host-side prover driver above the device C ABI (include/aero_b200.h).
//
// Mirrors the orchestration layer of winter-prover with the same names and call order so that the
// parity tests read like the reference's own:
//   RandomCoin      <- winterfell/crypto/src/random/mod.rs:60-306
//   ProverChannel   <- winterfell/prover/src/channel.rs:22-205
//   Prover          <- winterfell/prover/src/lib.rs:124-632 (generate_proof and its three stages)
//   StarkProof      <- winterfell/air/src/proof/mod.rs:51-168 (+ context/commitments/queries/ood_frame)
// Fiat-Shamir (a handful of single-block BLAKE2s calls per proof) stays on the host, as the north
// star prescribes; everything sized by the trace runs on the GPU through the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/aero_b200.h"
#include "../../include/aero_prover.h"

namespace aero {
namespace host {

using Digest = std::vector<uint8_t>;  // 32 bytes

Digest blake2s(const uint8_t *data, size_t len);
Digest hash_elements(const std::vector<uint64_t> &canonical);  // blake2s/mod.rs:52-77
Digest merge(const Digest &a, const Digest &b);                // blake2s/mod.rs:37-39
Digest merge_with_int(const Digest &seed, uint64_t v);         // blake2s/mod.rs:41-46

class RandomCoin {
   public:
    explicit RandomCoin(const uint8_t *seed, size_t len);  // random/mod.rs:73-80
    void reseed(const Digest &data);                       // :105-108
    void reseed_with_int(uint64_t value);                  // :131-134
    uint32_t leading_zeros() const;                        // :156-160
    uint32_t check_leading_zeros(uint64_t value) const;    // :164-169
    bool draw(uint64_t *canonical);                        // :179-196 (false after 1000 rejections)
    bool draw_integers(size_t num_values, uint64_t domain_size, std::vector<uint64_t> *out);  // :252-297
    const Digest &seed() const { return seed_; }

   private:
    Digest next();  // :303-306
    Digest seed_;
    uint64_t counter_ = 0;
};

struct ProofOptions {
    aero_proof_options o;
    std::vector<uint8_t> to_bytes() const;        // air/src/options.rs:231-239
    size_t num_fri_layers(uint64_t domain) const;  // fri/src/options.rs:96-103
};

struct Queries {  // air/src/proof/queries.rs:50-153
    std::vector<uint8_t> values, paths;
    void write_into(std::vector<uint8_t> &out) const;
};

struct StarkProof {
    std::vector<uint8_t> context, commitments, ood_trace_states, ood_evaluations, fri_proof;
    std::vector<Queries> trace_queries;
    Queries constraint_queries;
    uint64_t pow_nonce = 0;
    std::vector<uint8_t> to_bytes() const;  // air/src/proof/mod.rs:122-132
};

// Field elements inside the channel are canonical; the ABI form is applied at the C boundary.
class ProverChannel {
   public:
    ProverChannel(aero_ctx *ctx, const aero_prove_inputs &in);
    void commit_trace(const Digest &root);                                       // channel.rs:73-76
    void commit_constraints(const Digest &root);                                 // :79-82
    void send_ood_trace_states(const std::vector<std::vector<uint64_t>> &rows);  // :86-91
    void send_ood_constraint_evaluations(const std::vector<uint64_t> &evals);    // :95-98
    bool draw_elements(size_t n, std::vector<uint64_t> *out);                    // :104-134
    void commit_fri_layer(const Digest &root);                                   // :200-203
    aero_status grind_query_seed();                                              // :151-167 (GPU search)
    aero_status set_pow_nonce(uint64_t nonce);                                   // same, nonce found with the FRI layers
    bool get_query_positions(std::vector<uint64_t> *out);                        // :140-146
    StarkProof build_proof(std::vector<Queries> trace_queries, Queries constraint_queries,
                           std::vector<uint8_t> fri_proof);                      // :173-194
    RandomCoin &coin() { return coin_; }
    uint64_t pow_nonce() const { return pow_nonce_; }

   private:
    aero_ctx *ctx_;
    ProofOptions options_;
    uint64_t lde_domain_size_;
    RandomCoin coin_;
    std::vector<uint8_t> context_, commitments_, ood_trace_, ood_evals_;
    uint64_t pow_nonce_ = 0;
};

// Prover::generate_proof (prover/src/lib.rs:203-267).
aero_status prove(aero_ctx *ctx, const aero_prove_inputs &in, std::vector<uint8_t> *proof_bytes, std::string *err);

}  // namespace host
}  // namespace aero
