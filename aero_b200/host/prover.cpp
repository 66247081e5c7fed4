// Host-side prover driver; see prover.hpp for the mapping to winter-prover.
#include "prover.hpp"

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

#include "../csrc/blake2s.cuh"
#include "../csrc/gl.cuh"

namespace aero {
namespace host {

// ---------------------------------------------------------------------------------------------
// hashing (host copies of the single-block helpers used by Fiat-Shamir)
// ---------------------------------------------------------------------------------------------
static Digest to_digest(const uint32_t h[8]) {
    Digest d(32);
    memcpy(d.data(), h, 32);
    return d;
}
Digest blake2s(const uint8_t *data, size_t len) {
    uint32_t h[8], m[16];
    b2s::init(h);
    size_t off = 0;
    while (len - off > 64) {
        memcpy(m, data + off, 64);
        off += 64;
        b2s::compress(h, m, (uint32_t)off, false);
    }
    uint8_t last[64] = {0};
    if (len - off) memcpy(last, data + off, len - off);
    memcpy(m, last, 64);
    b2s::compress(h, m, (uint32_t)len, true);
    return to_digest(h);
}
Digest hash_elements(const std::vector<uint64_t> &e) {
    if (e.empty()) return blake2s(nullptr, 0);  // BLAKE2s of the empty message, not the bare IV
    uint32_t h[8];
    b2s::init(h);
    const size_t nblocks = (e.size() + 1) / 2;
    for (size_t b = 0; b < nblocks; b++) {
        const bool last = b + 1 == nblocks;
        const uint64_t e1 = 2 * b + 1 < e.size() ? e[2 * b + 1] : 0;
        b2s::compress_pair(h, e[2 * b], e1, last ? (uint32_t)(32 * e.size()) : (uint32_t)(64 * (b + 1)), last);
    }
    return to_digest(h);
}
Digest merge(const Digest &a, const Digest &b) {
    uint32_t x[8], y[8], o[8];
    memcpy(x, a.data(), 32);
    memcpy(y, b.data(), 32);
    b2s::merge(x, y, o);
    return to_digest(o);
}
Digest merge_with_int(const Digest &seed, uint64_t v) {
    uint32_t s[8], o[8];
    memcpy(s, seed.data(), 32);
    b2s::merge_with_int(s, v, o);
    return to_digest(o);
}
static uint64_t head64(const Digest &d) {
    uint64_t v;
    memcpy(&v, d.data(), 8);
    return v;
}
static uint32_t tz64(uint64_t x) { return x ? (uint32_t)__builtin_ctzll(x) : 64u; }

// ---------------------------------------------------------------------------------------------
// RandomCoin
// ---------------------------------------------------------------------------------------------
RandomCoin::RandomCoin(const uint8_t *seed, size_t len) : seed_(blake2s(seed, len)) {}
void RandomCoin::reseed(const Digest &data) {
    seed_ = merge(seed_, data);
    counter_ = 0;
}
void RandomCoin::reseed_with_int(uint64_t value) {
    seed_ = merge_with_int(seed_, value);
    counter_ = 0;
}
uint32_t RandomCoin::leading_zeros() const { return tz64(head64(seed_)); }
uint32_t RandomCoin::check_leading_zeros(uint64_t value) const { return tz64(head64(merge_with_int(seed_, value))); }
Digest RandomCoin::next() {
    counter_ += 1;
    return merge_with_int(seed_, counter_);
}
bool RandomCoin::draw(uint64_t *out) {
    for (int i = 0; i < 1000; i++) {
        const uint64_t v = head64(next());
        if (v < gl::P) {  // from_random_bytes rejects non-canonical values (f64/mod.rs:438-454)
            *out = v;
            return true;
        }
    }
    return false;
}
bool RandomCoin::draw_integers(size_t num_values, uint64_t domain_size, std::vector<uint64_t> *out) {
    out->clear();
    if (domain_size == 0 || (domain_size & (domain_size - 1)) || num_values >= domain_size) return false;
    const uint64_t mask = domain_size - 1;
    for (int i = 0; i < 1000; i++) {
        const uint64_t v = head64(next()) & mask;
        if (std::find(out->begin(), out->end(), v) != out->end()) continue;
        out->push_back(v);
        if (out->size() == num_values) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// wire format
// ---------------------------------------------------------------------------------------------
template <typename T>
static void put(std::vector<uint8_t> &o, T v) {
    const uint8_t *p = (const uint8_t *)&v;
    o.insert(o.end(), p, p + sizeof(T));
}
static void put_bytes(std::vector<uint8_t> &o, const std::vector<uint8_t> &b) { o.insert(o.end(), b.begin(), b.end()); }
static void put_elems(std::vector<uint8_t> &o, const std::vector<uint64_t> &e) {
    for (uint64_t v : e) put<uint64_t>(o, v);
}
static int ilog2(uint64_t x) {
    int l = 0;
    while ((1ULL << l) < x) l++;
    return l;
}
std::vector<uint8_t> ProofOptions::to_bytes() const {
    return {o.num_queries, o.blowup_factor, o.grinding_factor, o.hash_fn, o.field_extension, o.fri_folding_factor,
            (uint8_t)ilog2(o.fri_max_remainder_size)};
}
size_t ProofOptions::num_fri_layers(uint64_t domain) const {
    size_t r = 0;
    while (domain > o.fri_max_remainder_size) {
        domain /= o.fri_folding_factor;
        r++;
    }
    return r;
}
void Queries::write_into(std::vector<uint8_t> &out) const {
    put<uint32_t>(out, (uint32_t)values.size());
    put_bytes(out, values);
    put<uint32_t>(out, (uint32_t)paths.size());
    put_bytes(out, paths);
}
std::vector<uint8_t> StarkProof::to_bytes() const {
    std::vector<uint8_t> r;
    put_bytes(r, context);
    put<uint16_t>(r, (uint16_t)commitments.size());  // commitments.rs:84-90
    put_bytes(r, commitments);
    for (auto &q : trace_queries) q.write_into(r);
    constraint_queries.write_into(r);
    put<uint16_t>(r, (uint16_t)ood_trace_states.size());  // ood_frame.rs:113-122
    put_bytes(r, ood_trace_states);
    put<uint16_t>(r, (uint16_t)ood_evaluations.size());
    put_bytes(r, ood_evaluations);
    put_bytes(r, fri_proof);
    put<uint64_t>(r, pow_nonce);
    return r;
}

// ---------------------------------------------------------------------------------------------
// ProverChannel
// ---------------------------------------------------------------------------------------------
ProverChannel::ProverChannel(aero_ctx *ctx, const aero_prove_inputs &in)
    : ctx_(ctx), options_{in.options}, lde_domain_size_(in.trace_len * in.options.blowup_factor),
      coin_(in.pub_inputs_bytes, in.pub_inputs_len) {
    // Context::write_into (air/src/proof/context.rs:98-107) + TraceLayout (trace_info.rs:274-290)
    context_.push_back((uint8_t)in.main_width);
    context_.push_back((uint8_t)in.aux_width);
    context_.push_back((uint8_t)(in.aux_width ? in.aux_rands : 0));
    context_.push_back((uint8_t)ilog2(in.trace_len));
    put<uint16_t>(context_, in.trace_meta_len);
    if (in.trace_meta_len) context_.insert(context_.end(), in.trace_meta, in.trace_meta + in.trace_meta_len);
    context_.push_back(8);
    put<uint64_t>(context_, gl::P);
    put_bytes(context_, options_.to_bytes());
}
void ProverChannel::commit_trace(const Digest &root) {
    put_bytes(commitments_, root);
    coin_.reseed(root);
}
void ProverChannel::commit_constraints(const Digest &root) {
    put_bytes(commitments_, root);
    coin_.reseed(root);
}
void ProverChannel::send_ood_trace_states(const std::vector<std::vector<uint64_t>> &rows) {
    for (auto &row : rows) {
        put_elems(ood_trace_, row);
        coin_.reseed(hash_elements(row));
    }
}
void ProverChannel::send_ood_constraint_evaluations(const std::vector<uint64_t> &evals) {
    put_elems(ood_evals_, evals);
    coin_.reseed(hash_elements(evals));
}
bool ProverChannel::draw_elements(size_t n, std::vector<uint64_t> *out) {
    out->resize(n);
    for (size_t i = 0; i < n; i++)
        if (!coin_.draw(&(*out)[i])) return false;
    return true;
}
void ProverChannel::commit_fri_layer(const Digest &root) {
    put_bytes(commitments_, root);
    coin_.reseed(root);
}
aero_status ProverChannel::set_pow_nonce(uint64_t nonce) {
    // the nonce came from the device-side search seeded with the device coin: it must satisfy the host coin too
    if (coin_.check_leading_zeros(nonce) < options_.o.grinding_factor) return AERO_ERR_STATE;
    pow_nonce_ = nonce;
    coin_.reseed_with_int(nonce);
    return AERO_OK;
}
aero_status ProverChannel::grind_query_seed() {
    uint64_t nonce = 0;
    aero_status st = aero_pow_min_nonce(ctx_, coin_.seed().data(), options_.o.grinding_factor, &nonce);
    if (st != AERO_OK) return st;
    pow_nonce_ = nonce;
    coin_.reseed_with_int(nonce);
    return AERO_OK;
}
bool ProverChannel::get_query_positions(std::vector<uint64_t> *out) {
    return coin_.draw_integers(options_.o.num_queries, lde_domain_size_, out);
}
StarkProof ProverChannel::build_proof(std::vector<Queries> trace_queries, Queries constraint_queries,
                                      std::vector<uint8_t> fri_proof) {
    StarkProof p;
    p.context = context_;
    p.commitments = commitments_;
    p.ood_trace_states = ood_trace_;
    p.ood_evaluations = ood_evals_;
    p.trace_queries = std::move(trace_queries);
    p.constraint_queries = std::move(constraint_queries);
    p.fri_proof = std::move(fri_proof);
    p.pow_nonce = pow_nonce_;
    return p;
}

// ---------------------------------------------------------------------------------------------
// Prover
// ---------------------------------------------------------------------------------------------
namespace {
struct Handles {  // RAII for the device handles of one proof
    std::vector<aero_segment *> segs;
    aero_fri *fri = nullptr;
    aero_upload *up_aux = nullptr, *up_ce = nullptr;
    ~Handles() {
        aero_upload_free(up_aux);
        aero_upload_free(up_ce);
        aero_fri_destroy(fri);
        for (auto s : segs) aero_segment_destroy(s);
    }
};
#define P_TRY(expr)                                                      \
    do {                                                                 \
        aero_status _s = (expr);                                         \
        if (_s != AERO_OK) {                                             \
            if (err) *err = std::string(#expr) + ": " + aero_last_error(ctx); \
            return _s;                                                   \
        }                                                                \
    } while (0)
#define P_FAIL(code, msg)          \
    do {                           \
        if (err) *err = (msg);     \
        return (code);             \
    } while (0)
}  // namespace

extern "C" int aero_ctx_get_form(aero_ctx *ctx);

static aero_status prove_inner(aero_ctx *ctx, const aero_prove_inputs &in, std::vector<uint8_t> *proof_bytes, std::string *err);

// ProofOptions::new (air/src/options.rs:120-160) panics on these; here they are AERO_ERR_INVALID.
static const char *check_options(const aero_proof_options &o) {
    auto pow2 = [](uint32_t x) { return x && !(x & (x - 1)); };
    if (o.num_queries == 0) return "number of queries must be greater than 0";
    if (o.num_queries > 128) return "number of queries cannot be greater than 128";
    if (!pow2(o.blowup_factor)) return "blowup factor must be a power of 2";
    if (o.blowup_factor < 2) return "blowup factor cannot be smaller than 2";
    if (o.blowup_factor > 128) return "blowup factor cannot be greater than 128";
    if (o.grinding_factor > 32) return "grinding factor cannot be greater than 32";
    if (!pow2(o.fri_folding_factor)) return "FRI folding factor must be a power of 2";
    if (o.fri_folding_factor < 4) return "FRI folding factor cannot be smaller than 4";
    if (o.fri_folding_factor > 16) return "FRI folding factor cannot be greater than 16";
    if (!pow2(o.fri_max_remainder_size)) return "FRI max remainder size must be a power of 2";
    if (o.fri_max_remainder_size < 32) return "FRI max remainder size cannot be smaller than 32";
    if (o.fri_max_remainder_size > 1024) return "FRI max remainder size cannot be greater than 1024";
    return nullptr;
}

// Brackets a (possibly sharded) proof: the first proof of a shape on a sharded context synchronises its
// rank barriers on the host, later ones on the device (include/aero_b200.h, multi-GPU section).
aero_status prove(aero_ctx *ctx, const aero_prove_inputs &in, std::vector<uint8_t> *proof_bytes, std::string *err) {
    if (const char *msg = check_options(in.options)) P_FAIL(AERO_ERR_INVALID, msg);
    // everything that decides which device blocks a proof asks for (the first proof of a key may call cudaMalloc)
    char key[224];
    const bool device_air = in.air_program && !in.constraint_evaluator && !in.ce_cols;
    unsigned long long prog_size = 0;
    if (device_air) {
        prog_size = (unsigned long long)in.air_program->n_nodes * 131 + in.air_program->n_consts * 17 + in.air_program->n_transition * 5 +
                    in.air_program->n_boundary;
        for (uint32_t k = 0; k < in.air_program->n_periodic && in.air_program->periodic_len; k++) prog_size = prog_size * 31 + in.air_program->periodic_len[k];
    }
    snprintf(key, sizeof key, "prove/%llu/%u/%u/%u/%d/q%u/b%u/g%u/r%u/c%u/m%d%d/p%llu", (unsigned long long)in.trace_len, in.main_width,
             in.aux_width, in.n_div, in.inputs_on_device, in.options.num_queries, in.options.blowup_factor, in.options.grinding_factor,
             in.options.fri_max_remainder_size, in.ce_blowup, device_air ? 1 : 0, in.aux_builder ? 1 : 0, prog_size);
    P_TRY(aero_ctx_shard_begin(ctx, key));
    aero_status st = prove_inner(ctx, in, proof_bytes, err);
    aero_ctx_shard_end(ctx, st == AERO_OK);
    return st;
}

// Diagnostic (AERO_HOST_TIMING=1): wall-clock marks of the host-side sections of a proof, printed to stderr.
struct HostMarks {
    bool on = getenv("AERO_HOST_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    std::string log;
    void mark(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof buf, "  %-28s +%8.1f us  (at %9.1f us)\n", what,
                 std::chrono::duration<double, std::micro>(now - last).count(),
                 std::chrono::duration<double, std::micro>(now - t0).count());
        log += buf;
        last = now;
    }
    ~HostMarks() {
        if (on) fprintf(stderr, "aero_prove host marks:\n%s", log.c_str());
    }
};

static aero_status prove_inner(aero_ctx *ctx, const aero_prove_inputs &in, std::vector<uint8_t> *proof_bytes, std::string *err) {
    HostMarks hm;
    const aero_proof_options &o = in.options;
    if (o.hash_fn != 4) P_FAIL(AERO_ERR_UNSUPPORTED, "only Blake2s_256 (hash_fn = 4) is supported");
    if (o.field_extension != 1) P_FAIL(AERO_ERR_UNSUPPORTED, "only FieldExtension::None is supported");
    if (o.fri_folding_factor != 8) P_FAIL(AERO_ERR_UNSUPPORTED, "only FRI folding factor 8 is supported");
    if (!in.main_cols || in.main_width == 0) P_FAIL(AERO_ERR_INVALID, "main trace segment is required");
    if (in.aux_width && !in.aux_builder && !in.aux_cols) P_FAIL(AERO_ERR_INVALID, "auxiliary segment columns are required");
    if (!in.constraint_evaluator && !in.ce_cols && !in.air_program) P_FAIL(AERO_ERR_INVALID, "constraint evaluations are required");
    const bool device_air = in.air_program && !in.constraint_evaluator && !in.ce_cols;
    if (in.inputs_on_device && (in.aux_builder || in.constraint_evaluator)) P_FAIL(AERO_ERR_INVALID, "callbacks need host inputs");
    const bool sharded = aero_ctx_window_ranks(ctx) > 1;
    // A sharded proof evaluates the AIR on the device (each rank the cosets it holds) or takes precomputed
    // evaluations; the evaluator callback would need the whole LDE on the host of one rank.
    if (sharded && in.constraint_evaluator) P_FAIL(AERO_ERR_UNSUPPORTED, "sharded proofs take an AIR program or precomputed constraint evaluations, not the evaluator callback");
    if (in.trace_len < 2 || (in.trace_len & (in.trace_len - 1))) P_FAIL(AERO_ERR_INVALID, "trace length must be a power of two >= 2");
    const bool mont = aero_ctx_get_form(ctx) == AERO_FORM_MONTGOMERY;
    auto to_abi = [&](uint64_t x) { return mont ? gl::canon_to_mont(x) : x; };
    auto from_abi = [&](uint64_t x) { return mont ? gl::mont_to_canon(x) : gl::canon(x); };
    const uint64_t n = in.trace_len, N = n * o.blowup_factor;
    const uint32_t W = in.main_width + in.aux_width;
    const uint32_t ce_blowup = in.ce_blowup ? in.ce_blowup : o.blowup_factor;
    if ((ce_blowup & (ce_blowup - 1)) || ce_blowup < 2 || ce_blowup > o.blowup_factor)
        P_FAIL(AERO_ERR_INVALID, "constraint evaluation blowup must be a power of two in 2..blowup_factor");
    const uint64_t CE = n * ce_blowup;  // constraint evaluation domain (StarkDomain::ce_domain_size)

    Handles H;
    ProverChannel channel(ctx, in);
    uint8_t root[32];
    // Nothing on the device depends on a trace root before the OOD point is drawn unless a callback has to
    // see the randomness in between: then the three commitments are queued back to back and their roots
    // collected in ONE host round trip (aero_segments_roots), with the transcript replayed in order.
    const bool defer_roots = !in.aux_builder && !in.constraint_evaluator && !device_air;
    uint8_t *root_now = defer_roots ? nullptr : root;

    // Host inputs that are already known (no callback produces them) start travelling now: their copies
    // queue behind the main segment's own and land while its columns are being extended and hashed.
    if (!in.inputs_on_device) {
        // (a sharded proof uploads only what this rank reads: its trace columns, its rows of the constraint evaluations)
        if (in.aux_width && !in.aux_builder) P_TRY(aero_upload_start(ctx, in.aux_cols, in.aux_width, n, 1, AERO_UPLOAD_OWN_COLUMNS, &H.up_aux));
        if (!in.constraint_evaluator && !device_air) P_TRY(aero_upload_start(ctx, in.ce_cols, in.n_div, CE, 1, AERO_UPLOAD_OWN_ROWS, &H.up_ce));
    }

    // 1 ----- commit to the execution trace (lib.rs:239-248, 269-348)
    aero_segment *main_seg = nullptr;
    if (in.inputs_on_device)
        P_TRY(aero_segment_commit_device(ctx, in.main_cols[0], n, in.main_width, n, o.blowup_factor, 0, &main_seg, root_now));
    else
        P_TRY(aero_segment_commit(ctx, in.main_cols, in.main_width, n, o.blowup_factor, 0, &main_seg, root_now));
    H.segs.push_back(main_seg);
    if (!defer_roots) channel.commit_trace(Digest(root, root + 32));

    aero_segment *aux_seg = nullptr;
    if (in.aux_width) {
        std::vector<uint64_t> rand_elements;
        if (!defer_roots && !channel.draw_elements(in.aux_rands, &rand_elements)) P_FAIL(AERO_ERR_STATE, "failed to draw random elements");
        std::vector<const uint64_t *> aux_ptrs(in.aux_width);
        const uint64_t *const *aux_cols = in.aux_cols;
        if (in.aux_builder) {
            std::vector<uint64_t> abi(rand_elements.size());
            for (size_t i = 0; i < abi.size(); i++) abi[i] = to_abi(rand_elements[i]);
            // Every rank of a sharded proof has drawn the same elements and calls the builder itself (ranks in
            // separate processes have no other way; the ranks of an aero_group take turns, so the callback
            // never runs concurrently with itself); it must return the same columns every time.
            aero_status st;
            {
                static std::mutex callback_mu;
                std::lock_guard<std::mutex> lock(callback_mu);
                st = in.aux_builder(in.user, abi.data(), (uint32_t)abi.size(), aux_ptrs.data());
            }
            if (st != AERO_OK) P_FAIL(st, "aux_builder callback failed");
            aux_cols = aux_ptrs.data();
        }
        if (H.up_aux) {
            const uint64_t *d_aux = nullptr;
            P_TRY(aero_upload_wait(H.up_aux, &d_aux));
            P_TRY(aero_segment_commit_device(ctx, d_aux, n, in.aux_width, n, o.blowup_factor, 0, &aux_seg, root_now));
        } else if (in.inputs_on_device)
            P_TRY(aero_segment_commit_device(ctx, aux_cols[0], n, in.aux_width, n, o.blowup_factor, 0, &aux_seg, root_now));
        else
            P_TRY(aero_segment_commit(ctx, aux_cols, in.aux_width, n, o.blowup_factor, 0, &aux_seg, root_now));
        H.segs.push_back(aux_seg);
        if (!defer_roots) channel.commit_trace(Digest(root, root + 32));
    }

    // 2 ----- evaluate constraints (lib.rs:350-382): coefficients are drawn here; evaluation itself
    // stays with the caller (reference Rust path) or is supplied precomputed.
    std::vector<uint64_t> coeffs;
    if (!defer_roots && !channel.draw_elements(in.n_constraint_coeffs, &coeffs)) P_FAIL(AERO_ERR_STATE, "failed to draw composition coefficients");
    std::vector<const uint64_t *> ce_ptrs(in.n_div);
    const uint64_t *const *ce_cols = in.ce_cols;
    std::vector<std::vector<uint64_t>> lde_host;
    if (in.constraint_evaluator) {
        lde_host.resize(W);
        std::vector<uint64_t *> ptrs(W);
        for (uint32_t c = 0; c < W; c++) {
            lde_host[c].resize(N);
            ptrs[c] = lde_host[c].data();
        }
        P_TRY(aero_segment_download_lde(main_seg, ptrs.data()));
        if (aux_seg) P_TRY(aero_segment_download_lde(aux_seg, ptrs.data() + in.main_width));
        std::vector<uint64_t> abi(coeffs.size());
        for (size_t i = 0; i < abi.size(); i++) abi[i] = to_abi(coeffs[i]);
        aero_status st = in.constraint_evaluator(in.user, (const uint64_t *const *)ptrs.data(), W, N, abi.data(),
                                                 (uint32_t)abi.size(), ce_ptrs.data());
        if (st != AERO_OK) P_FAIL(st, "constraint_evaluator callback failed");
        ce_cols = ce_ptrs.data();
    }

    // 3 ----- commit to constraint evaluations (lib.rs:396-419)
    aero_segment *comp_seg = nullptr;
    if (device_air) {
        // ConstraintEvaluator::evaluate on the device: the frame is read from the resident LDE
        std::vector<uint64_t> abi(coeffs.size());
        for (size_t i = 0; i < abi.size(); i++) abi[i] = to_abi(coeffs[i]);
        std::vector<aero_segment *> tsegs = {main_seg};
        if (aux_seg) tsegs.push_back(aux_seg);
        P_TRY(aero_constraints_evaluate_into_poly(ctx, tsegs.data(), (uint32_t)tsegs.size(), in.air_program, abi.data(),
                                                  (uint32_t)abi.size(), ce_blowup, in.divisors, in.n_div, &comp_seg));
    } else if (H.up_ce) {
        const uint64_t *d_ce = nullptr;
        P_TRY(aero_upload_wait(H.up_ce, &d_ce));
        P_TRY(aero_constraints_into_poly_device(ctx, d_ce, CE, in.divisors, in.n_div, CE, n, &comp_seg));
    } else if (in.inputs_on_device)
        P_TRY(aero_constraints_into_poly_device(ctx, ce_cols[0], CE, in.divisors, in.n_div, CE, n, &comp_seg));
    else
        P_TRY(aero_constraints_into_poly(ctx, ce_cols, in.divisors, in.n_div, CE, n, &comp_seg));
    H.segs.push_back(comp_seg);
    lde_host.clear();
    P_TRY(aero_segment_commit_polys(comp_seg, o.blowup_factor, root_now));
    if (defer_roots) {  // replay: commit_trace, draw aux randomness, commit_trace, draw coefficients (draws never move the seed)
        std::vector<uint8_t> roots(H.segs.size() * 32);
        P_TRY(aero_segments_roots(ctx, H.segs.data(), (uint32_t)H.segs.size(), roots.data()));
        std::vector<uint64_t> unused;
        for (size_t i = 0; i + 1 < H.segs.size(); i++) {
            if (i == 1 && !channel.draw_elements(in.aux_rands, &unused)) P_FAIL(AERO_ERR_STATE, "failed to draw random elements");
            channel.commit_trace(Digest(roots.begin() + i * 32, roots.begin() + (i + 1) * 32));
        }
        if (!channel.draw_elements(in.n_constraint_coeffs, &unused)) P_FAIL(AERO_ERR_STATE, "failed to draw composition coefficients");
        memcpy(root, roots.data() + (H.segs.size() - 1) * 32, 32);
    }
    channel.commit_constraints(Digest(root, root + 32));
    hm.mark("commitments + roots");

    // 4 ----- OOD frame + DEEP composition polynomial (lib.rs:421-467)
    uint64_t z;
    if (!channel.coin().draw(&z)) P_FAIL(AERO_ERR_STATE, "failed to draw OOD point");
    uint32_t m = 0;
    P_TRY(aero_segment_info(comp_seg, &m, nullptr, nullptr));
    std::vector<uint64_t> ood_trace(2 * W), ood_comp(m);
    std::vector<aero_segment *> trace_segs = {main_seg};
    if (aux_seg) trace_segs.push_back(aux_seg);
    P_TRY(aero_ood_eval(ctx, trace_segs.data(), (uint32_t)trace_segs.size(), comp_seg, to_abi(z), ood_trace.data(), ood_comp.data()));
    {
        std::vector<std::vector<uint64_t>> rows(2, std::vector<uint64_t>(W));
        for (uint32_t i = 0; i < W; i++) {
            rows[0][i] = from_abi(ood_trace[i]);
            rows[1][i] = from_abi(ood_trace[W + i]);
        }
        channel.send_ood_trace_states(rows);
        std::vector<uint64_t> ev(m);
        for (uint32_t i = 0; i < m; i++) ev[i] = from_abi(ood_comp[i]);
        channel.send_ood_constraint_evaluations(ev);
    }
    hm.mark("ood frame");
    // get_deep_composition_coefficients (air/src/air/mod.rs:537-561): W triples, m singles, one pair
    std::vector<uint64_t> cc;
    if (!channel.draw_elements((size_t)3 * W + m + 2, &cc)) P_FAIL(AERO_ERR_STATE, "failed to draw DEEP coefficients");
    for (auto &v : cc) v = to_abi(v);
    P_TRY(aero_deep_compose(ctx, trace_segs.data(), (uint32_t)trace_segs.size(), comp_seg, to_abi(z), ood_trace.data(),
                            ood_comp.data(), cc.data(), &H.fri));

    hm.mark("deep_compose queued");
    // 6 ----- FRI layers (fri/src/prover/mod.rs:166-191)
    const size_t num_layers = ProofOptions{o}.num_fri_layers(N);
    {
        // all layers in one host round trip: the coin's reseed/draw per layer runs on the device
        // between the kernels; the channel is replayed from the returned roots and must agree
        std::vector<uint8_t> roots((num_layers + 1) * 32);
        std::vector<uint64_t> alphas(num_layers + 1);
        uint64_t nonce = 0;
        // ... and the grinding search (channel.rs:151-167) queued right behind them, on the device coin's seed
        P_TRY(aero_fri_build_layers_grind(H.fri, channel.coin().seed().data(), (uint32_t)num_layers, o.grinding_factor,
                                          roots.data(), alphas.data(), &nonce));
        hm.mark("fri layers + grind (sync)");
        for (size_t l = 0; l < num_layers + 1; l++) {
            channel.commit_fri_layer(Digest(roots.begin() + l * 32, roots.begin() + (l + 1) * 32));
            uint64_t alpha;
            if (!channel.coin().draw(&alpha)) P_FAIL(AERO_ERR_STATE, "failed to draw FRI alpha");
            if (to_abi(alpha) != alphas[l]) P_FAIL(AERO_ERR_STATE, "device coin diverged from the channel's coin");
        }
        // 7 ----- query positions (lib.rs:502-516)
        if (channel.set_pow_nonce(nonce) != AERO_OK) P_FAIL(AERO_ERR_STATE, "device grinding nonce does not satisfy the channel's coin");
    }

    hm.mark("channel replay + nonce");
    std::vector<uint64_t> positions;
    if (!channel.get_query_positions(&positions)) P_FAIL(AERO_ERR_STATE, "failed to draw query positions");
    hm.mark("query positions");

    // 8 ----- proof object (lib.rs:518-539)
    // FriProver::build_proof + build_segment_queries (prover/src/trace/commitment.rs:115-140) for every
    // segment: one batched gather, one host round trip
    std::vector<aero_segment *> open_segs = trace_segs;
    open_segs.push_back(comp_seg);
    std::vector<uint32_t> widths = {in.main_width};
    if (aux_seg) widths.push_back(in.aux_width);
    widths.push_back(m);
    const size_t ns = open_segs.size(), np = positions.size();
    std::vector<std::vector<uint64_t>> rows(ns);
    std::vector<std::vector<uint8_t>> paths(ns);
    std::vector<uint64_t *> rows_ptr(ns);
    std::vector<uint8_t *> paths_ptr(ns);
    std::vector<size_t> paths_len(ns);
    for (size_t i = 0; i < ns; i++) {
        rows[i].resize(np * widths[i]);
        paths[i].resize(1 + np * (1 + 32 * 40));
        rows_ptr[i] = rows[i].data();
        paths_ptr[i] = paths[i].data();
        paths_len[i] = paths[i].size();
    }
    std::vector<uint8_t> fri_bytes(256 << 10);
    size_t flen = fri_bytes.size();
    hm.mark("opening buffers");
    aero_status ost = aero_open_queries(ctx, H.fri, open_segs.data(), (uint32_t)ns, positions.data(), (uint32_t)np,
                                        fri_bytes.data(), &flen, rows_ptr.data(), paths_ptr.data(), paths_len.data());
    if (ost == AERO_ERR_BUFFER) {  // sizes were reported: grow and repeat
        fri_bytes.resize(flen);
        for (size_t i = 0; i < ns; i++) {
            paths[i].resize(paths_len[i]);
            paths_ptr[i] = paths[i].data();
        }
        ost = aero_open_queries(ctx, H.fri, open_segs.data(), (uint32_t)ns, positions.data(), (uint32_t)np, fri_bytes.data(),
                                &flen, rows_ptr.data(), paths_ptr.data(), paths_len.data());
    }
    P_TRY(ost);
    hm.mark("open_queries (sync)");
    fri_bytes.resize(flen);
    std::vector<Queries> all_q(ns);
    for (size_t i = 0; i < ns; i++) {
        paths[i].resize(paths_len[i]);
        all_q[i].values.assign((uint8_t *)rows[i].data(), (uint8_t *)rows[i].data() + rows[i].size() * 8);
        all_q[i].paths = std::move(paths[i]);
    }
    Queries cq = std::move(all_q.back());
    all_q.pop_back();
    std::vector<Queries> tq = std::move(all_q);
    *proof_bytes = channel.build_proof(std::move(tq), std::move(cq), std::move(fri_bytes)).to_bytes();
    hm.mark("proof bytes");
    return AERO_OK;
}

}  // namespace host
}  // namespace aero

// ---------------------------------------------------------------------------------------------
// C entry points (include/aero_prover.h)
// ---------------------------------------------------------------------------------------------
using namespace aero::host;

// ---------------------------------------------------------------------------------------------
// aero_group: G contexts of one process working on one proof (include/aero_prover.h)
// ---------------------------------------------------------------------------------------------
struct aero_group {
    std::vector<aero_ctx *> ctxs;
    // host rendezvous of the ranks' threads (first proof of a shape); abort() releases waiters with an
    // error when a rank has failed
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long long generation = 0;
    bool broken = false;
    std::string err;
    aero_status wait() {
        std::unique_lock<std::mutex> lock(mu);
        if (broken) return AERO_ERR_STATE;
        const unsigned long long gen = generation;
        if (++waiting == (int)ctxs.size()) {
            waiting = 0;
            generation++;
            cv.notify_all();
            return AERO_OK;
        }
        cv.wait(lock, [&] { return generation != gen || broken; });
        return generation != gen ? AERO_OK : AERO_ERR_STATE;
    }
    void abort() {
        std::lock_guard<std::mutex> lock(mu);
        broken = true;
        cv.notify_all();
    }
};
static aero_status group_host_barrier(void *user) { return static_cast<aero_group *>(user)->wait(); }

extern "C" {
void aero_ctx_set_error(aero_ctx *ctx, const char *msg);

aero_status aero_group_create(const int *device_ids, int n_ranks, size_t window_bytes, aero_group **out) {
    if (!out || !device_ids) return AERO_ERR_INVALID;
    *out = nullptr;
    if (n_ranks < 1 || n_ranks > 8 || (n_ranks & (n_ranks - 1))) return AERO_ERR_INVALID;
    // Ranks that share a device (tests) must not share a hardware work queue: a copy queued behind
    // another rank's spinning barrier kernel would never start.  Only effective before CUDA initialises.
    for (int r = 1; r < n_ranks; r++)
        if (device_ids[r] == device_ids[0]) setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    aero_group *g = new aero_group();
    aero_status st = AERO_OK;
    for (int r = 0; r < n_ranks && st == AERO_OK; r++) {
        aero_ctx *c = nullptr;
        st = aero_ctx_create(&device_ids[r], 1, &c);
        if (st != AERO_OK) break;
        g->ctxs.push_back(c);
        st = aero_ctx_set_option(c, "own_stream", 1);
        bool shared_device = false;  // ranks on one device: few streams each, there are only 32 hardware queues
        for (int q = 0; q < n_ranks; q++) shared_device = shared_device || (q != r && device_ids[q] == device_ids[r]);
        if (st == AERO_OK && shared_device) st = aero_ctx_set_option(c, "push_parts", 1);
        if (st == AERO_OK) st = aero_ctx_set_shard(c, r, n_ranks);
        if (st == AERO_OK && n_ranks > 1) st = aero_ctx_window_create(c, window_bytes, nullptr);
        if (st == AERO_OK) st = aero_ctx_set_host_barrier(c, group_host_barrier, g);
    }
    for (int r = 0; r < n_ranks && st == AERO_OK && n_ranks > 1; r++) st = aero_ctx_window_attach_local(g->ctxs[r], n_ranks, g->ctxs.data());
    if (st != AERO_OK) {
        for (aero_ctx *c : g->ctxs) aero_ctx_destroy(c);
        delete g;
        return st;
    }
    *out = g;
    return AERO_OK;
}
void aero_group_destroy(aero_group *g) {
    if (!g) return;
    for (aero_ctx *c : g->ctxs) aero_ctx_destroy(c);
    delete g;
}
int aero_group_size(aero_group *g) { return g ? (int)g->ctxs.size() : 0; }
aero_ctx *aero_group_ctx(aero_group *g, int rank) { return (g && rank >= 0 && rank < (int)g->ctxs.size()) ? g->ctxs[rank] : nullptr; }
const char *aero_group_last_error(aero_group *g) { return g ? g->err.c_str() : "null group"; }

aero_status aero_group_prove(aero_group *g, const aero_prove_inputs *in, int n_inputs, uint8_t *proof_out, size_t *len) {
    if (!g || !in || !len) return AERO_ERR_INVALID;
    const int G = (int)g->ctxs.size();
    if (n_inputs != 1 && n_inputs != G) {
        g->err = "n_inputs must be 1 (shared host inputs) or the number of ranks";
        return AERO_ERR_INVALID;
    }
    {
        std::lock_guard<std::mutex> lock(g->mu);
        g->broken = false;
        g->waiting = 0;
    }
    std::vector<std::vector<uint8_t>> bytes(G);
    std::vector<std::string> errs(G);
    std::vector<aero_status> sts(G, AERO_OK);
    std::vector<std::thread> threads;
    for (int r = 0; r < G; r++)
        threads.emplace_back([&, r] {
            sts[r] = prove(g->ctxs[r], in[n_inputs == 1 ? 0 : r], &bytes[r], &errs[r]);
            if (sts[r] != AERO_OK) g->abort();  // do not leave the other ranks waiting for this one
        });
    for (auto &t : threads) t.join();
    for (int r = 0; r < G; r++)
        if (sts[r] != AERO_OK) {
            char b[64];
            snprintf(b, sizeof b, "rank %d: ", r);
            g->err = b + errs[r];
            aero_ctx_set_error(g->ctxs[r], errs[r].c_str());
            return sts[r];
        }
    for (int r = 1; r < G; r++)
        if (bytes[r] != bytes[0]) {
            g->err = "the ranks produced different proofs";
            return AERO_ERR_STATE;
        }
    if (!proof_out || *len < bytes[0].size()) {
        *len = bytes[0].size();
        g->err = "proof buffer too small";
        return AERO_ERR_BUFFER;
    }
    memcpy(proof_out, bytes[0].data(), bytes[0].size());
    *len = bytes[0].size();
    return AERO_OK;
}
}

struct aero_coin {
    RandomCoin c;
};
extern "C" {
void aero_ctx_set_error(aero_ctx *ctx, const char *msg);

aero_status aero_prove(aero_ctx *ctx, const aero_prove_inputs *in, uint8_t *proof_out, size_t *len) {
    if (!ctx || !in || !len) return AERO_ERR_INVALID;
    std::vector<uint8_t> bytes;
    std::string err;
    aero_status st = prove(ctx, *in, &bytes, &err);
    if (st != AERO_OK) {
        aero_ctx_set_error(ctx, err.c_str());
        return st;
    }
    if (!proof_out || *len < bytes.size()) {
        *len = bytes.size();
        aero_ctx_set_error(ctx, "proof buffer too small");
        return AERO_ERR_BUFFER;
    }
    memcpy(proof_out, bytes.data(), bytes.size());
    *len = bytes.size();
    return AERO_OK;
}
void aero_host_blake2s(const uint8_t *data, size_t len, uint8_t out[32]) { memcpy(out, blake2s(data, len).data(), 32); }
void aero_host_hash_elements(const uint64_t *e, size_t count, uint8_t out[32]) {
    memcpy(out, hash_elements(std::vector<uint64_t>(e, e + count)).data(), 32);
}
aero_coin *aero_coin_new(const uint8_t *seed_bytes, size_t len) { return new aero_coin{RandomCoin(seed_bytes, len)}; }
void aero_coin_free(aero_coin *c) { delete c; }
void aero_coin_reseed(aero_coin *c, const uint8_t digest[32]) { c->c.reseed(Digest(digest, digest + 32)); }
void aero_coin_reseed_with_int(aero_coin *c, uint64_t v) { c->c.reseed_with_int(v); }
aero_status aero_coin_draw(aero_coin *c, uint64_t *out) { return c->c.draw(out) ? AERO_OK : AERO_ERR_STATE; }
aero_status aero_coin_draw_integers(aero_coin *c, uint32_t num_values, uint64_t domain_size, uint64_t *out) {
    std::vector<uint64_t> v;
    if (!c->c.draw_integers(num_values, domain_size, &v)) return AERO_ERR_INVALID;
    memcpy(out, v.data(), v.size() * 8);
    return AERO_OK;
}
uint32_t aero_coin_leading_zeros(aero_coin *c) { return c->c.leading_zeros(); }
uint32_t aero_coin_check_leading_zeros(aero_coin *c, uint64_t v) { return c->c.check_leading_zeros(v); }
void aero_coin_seed(aero_coin *c, uint8_t out[32]) { memcpy(out, c->c.seed().data(), 32); }
}
