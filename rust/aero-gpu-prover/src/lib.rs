//! `GpuExecutionProver`: the Miden `ExecutionProver` with the trace-sized steps of
//! `winter_prover::Prover::prove` served by aero_b200 (hand-written sm_100a kernels behind a C ABI).
//!
//! What moves to the GPU: column interpolation and the blowup-8 coset
//! extension (`Matrix::interpolate_columns` / `evaluate_columns_over`), the blake2s row commitment
//! (`Matrix::commit_to_rows`, `MerkleTree::new`), `ConstraintEvaluationTable::into_poly`,
//! `CompositionPoly::evaluate`, the OOD frame, `DeepCompositionPoly`, the FRI layers and the query
//! openings.  What stays in Rust, as the north star asks: the Miden VM, auxiliary-column
//! construction, AIR constraint evaluation and the Fiat-Shamir channel (`ProverChannel`).
//!
//! The insertion points are the provided methods this fork made public for its wasm worker
//! (winterfell/prover/src/lib.rs:269 `commit_to_trace_and_validate`, :350 `evaluate_constraints`,
//! :384 `prove_after_constraint_eval`, :551 `build_trace_commitment`, :599
//! `build_constraint_commitment`).  Everything that crosses the boundary is a plain pointer and a
//! size; `BaseElement` is a `u64` in Montgomery form (math/src/field/f64/mod.rs:56-61), which is the
//! ABI's default element form, so `Matrix` columns are handed over without copy or conversion.
//!
//! Reference-side changes this crate needs are the five patches under `rust/patches/` (applied with
//! `patch -p1` at the root of the Aero checkout; `tests/test_host_cpu.py` checks that they still apply
//! to the reference sources and that every non-public item used below is one they export):
//!   0001  `pub use trace::{TraceCommitment, TracePolyTable}` in winter-prover (they appear in the
//!         signatures of the overridable provided methods but were crate-private);
//!   0002  `MerkleTree::from_root`, 0003 `Queries::from_raw_parts`,
//!   0004  `ConstraintEvaluationTable::into_raw_parts`,
//!   0005  the `gpu` / `dump-fixture` features of miden-proof-generator.
//!
//! This file has not been compiled in the aero_b200 build image (no Rust toolchain there); the C++
//! driver `aero_b200/host/prover.cpp` performs the identical call sequence and is what the parity
//! tests exercise (`tests/test_gpu_parity.py::test_prove_byte_identical`, and through the callbacks this
//! crate's route takes, `::test_prove_callback_route_byte_identical`).

pub mod dump;
pub mod ffi;

use std::ffi::CStr;
use std::ptr;

use miden_air::{Felt, ProcessorAir, PublicInputs};
use miden_processor::ExecutionTrace;
use miden_prover::ExecutionProver;
use winter_air::proof::{Queries, StarkProof};
use winter_air::{Air, ProofOptions};
use winter_crypto::{ElementHasher, MerkleTree};
use winter_fri::FriProof;
use winter_math::{FieldElement, StarkField};
// ProverChannel is re-exported at the crate root (winterfell/prover/src/lib.rs:97-98); TraceCommitment and
// TracePolyTable are exported by rust/patches/0001.
use winter_prover::{
    ConstraintEvaluationTable, Matrix, Prover, ProverChannel, ProverError, StarkDomain, TraceCommitment,
    TracePolyTable,
};
use winter_utils::{Deserializable, SliceReader};

/// A 32-byte GPU digest as the hasher's digest type (`Digest: Deserializable`,
/// winterfell/crypto/src/hash/mod.rs:67-79; `ByteDigest::read_from` copies the bytes, :152-156).
fn digest_from_bytes<D: Deserializable>(bytes: &[u8; 32]) -> D {
    D::read_from(&mut SliceReader::new(bytes)).expect("32-byte digest")
}

/// Error text of the last failed call on `ctx` (aero_last_error).
fn last_error(ctx: *mut ffi::aero_ctx) -> String {
    unsafe { CStr::from_ptr(ffi::aero_last_error(ctx)).to_string_lossy().into_owned() }
}

/// The reference panics on violated preconditions (prover/src/matrix.rs:42-61,
/// math/src/fft/mod.rs:179-199); the ABI reports them as status codes, which become panics with the
/// library's message here so that the observable behaviour is unchanged.
fn check(ctx: *mut ffi::aero_ctx, st: ffi::aero_status) {
    if st != ffi::AERO_OK {
        panic!("aero_b200 status {st}: {}", last_error(ctx));
    }
}

/// One GPU, one in-flight proof (the reference's `&self` prover methods are single-threaded too).
pub struct Context(pub(crate) *mut ffi::aero_ctx);

impl Context {
    pub fn new(device: i32) -> Self {
        let mut ctx = ptr::null_mut();
        let ids = [device];
        let st = unsafe { ffi::aero_ctx_create(ids.as_ptr(), 1, &mut ctx) };
        assert_eq!(st, ffi::AERO_OK, "no usable sm_100 CUDA device (aero_b200 has no CPU fallback)");
        Context(ctx)
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::aero_ctx_destroy(self.0) }
    }
}

/// A committed matrix on the GPU: coefficient columns, coset-major LDE, heap-layout Merkle tree.
pub struct Segment {
    ctx: *mut ffi::aero_ctx,
    h: *mut ffi::aero_segment,
    pub root: [u8; 32],
    pub width: usize,
}
impl Drop for Segment {
    fn drop(&mut self) {
        unsafe { ffi::aero_segment_destroy(self.h) }
    }
}

fn column_ptrs<E: FieldElement<BaseField = Felt>>(m: &Matrix<E>) -> Vec<*const u64> {
    // E == Felt on this path (FieldExtension::None, miden/air/src/options.rs:29-39)
    m.columns().map(|c| c.as_ptr() as *const u64).collect()
}

impl Segment {
    /// Prover::build_trace_commitment (lib.rs:551-589): interpolate, extend, hash rows, build tree.
    pub fn commit<E: FieldElement<BaseField = Felt>>(ctx: &Context, m: &Matrix<E>, blowup: usize, is_coeffs: bool) -> Self {
        let cols = column_ptrs(m);
        let (mut h, mut root) = (ptr::null_mut(), [0u8; 32]);
        check(ctx.0, unsafe {
            ffi::aero_segment_commit(ctx.0, cols.as_ptr(), cols.len() as u32, m.num_rows() as u64, blowup as u32,
                                     is_coeffs as i32, &mut h, root.as_mut_ptr())
        });
        Segment { ctx: ctx.0, h, root, width: m.num_cols() }
    }

    fn download<E: FieldElement<BaseField = Felt>>(&self, rows: usize,
        f: unsafe extern "C" fn(*mut ffi::aero_segment, *const *mut u64) -> ffi::aero_status) -> Matrix<E> {
        let mut cols: Vec<Vec<E>> = (0..self.width).map(|_| unsafe { winter_utils::uninit_vector(rows) }).collect();
        let ptrs: Vec<*mut u64> = cols.iter_mut().map(|c| c.as_mut_ptr() as *mut u64).collect();
        check(self.ctx, unsafe { f(self.h, ptrs.as_ptr()) });
        Matrix::new(cols)
    }
    /// The extended trace the Rust AIR evaluator walks row by row (constraints/evaluator.rs:74).
    pub fn lde<E: FieldElement<BaseField = Felt>>(&self, lde_size: usize) -> Matrix<E> {
        self.download(lde_size, ffi::aero_segment_download_lde)
    }
    pub fn polys<E: FieldElement<BaseField = Felt>>(&self, n: usize) -> Matrix<E> {
        self.download(n, ffi::aero_segment_download_polys)
    }

    /// TraceCommitment::query / ConstraintCommitment::query (trace/commitment.rs:115-140,
    /// constraints/commitment.rs:50-70): `values` is the canonical little-endian row image and
    /// `paths` is BatchMerkleProof::serialize_nodes (crypto/src/merkle/proofs.rs:421-439).
    pub fn query(&self, positions: &[usize]) -> Queries {
        let pos: Vec<u64> = positions.iter().map(|&p| p as u64).collect();
        let mut rows = vec![0u64; pos.len() * self.width];
        let mut paths = vec![0u8; 2 + pos.len() * (1 + 32 * 64)];
        let mut len = paths.len();
        check(self.ctx, unsafe {
            ffi::aero_segment_open(self.h, pos.as_ptr(), pos.len() as u32, rows.as_mut_ptr(), paths.as_mut_ptr(), &mut len)
        });
        paths.truncate(len);
        let values = rows.iter().flat_map(|v| v.to_le_bytes()).collect();
        Queries::from_raw_parts(values, paths) // 3-line constructor added to air/src/proof/queries.rs
    }
}

pub struct GpuExecutionProver {
    inner: ExecutionProver,
    ctx: Context,
    /// segments of the proof in flight, in commitment order (main, aux..., constraint)
    segs: std::cell::RefCell<Vec<Segment>>,
}

impl GpuExecutionProver {
    pub fn new(inner: ExecutionProver, device: i32) -> Self {
        Self { inner, ctx: Context::new(device), segs: Default::default() }
    }
}

impl Prover for GpuExecutionProver {
    type BaseField = Felt;
    type Air = ProcessorAir;
    type Trace = ExecutionTrace;

    fn get_pub_inputs(&self, trace: &ExecutionTrace) -> PublicInputs {
        self.inner.get_pub_inputs(trace)
    }
    fn options(&self) -> &ProofOptions {
        self.inner.options()
    }

    /// winterfell/prover/src/lib.rs:551 -- same return type as the default method; the tree handed
    /// back is a shell carrying the GPU root (MerkleTree::from_root, a 5-line constructor added to
    /// crypto/src/merkle/mod.rs next to `new`): openings are answered by `Segment::query`.
    fn build_trace_commitment<E, H>(&self, trace: &Matrix<E>, domain: &StarkDomain<Felt>) -> (Matrix<E>, MerkleTree<H>, Matrix<E>)
    where
        E: FieldElement<BaseField = Felt>,
        H: ElementHasher<BaseField = Felt>,
    {
        let seg = Segment::commit(&self.ctx, trace, domain.trace_to_lde_blowup(), false);
        let lde = seg.lde(domain.lde_domain_size());
        let polys = seg.polys(trace.num_rows());
        let tree = MerkleTree::<H>::from_root(digest_from_bytes(&seg.root), domain.lde_domain_size());
        self.segs.borrow_mut().push(seg);
        (lde, tree, polys)
    }

    /// winterfell/prover/src/lib.rs:384-540, restated over the C ABI in the order of the default
    /// method; every Fiat-Shamir step stays on the Rust `ProverChannel`.
    fn prove_after_constraint_eval<E, H>(
        &self,
        air: &ProcessorAir,
        mut channel: ProverChannel<ProcessorAir, E, H>,
        constraint_evaluations: ConstraintEvaluationTable<E>,
        _trace_polys: TracePolyTable<E>,
        _trace_commitment: TraceCommitment<E, H>,
    ) -> Result<StarkProof, ProverError>
    where
        E: FieldElement<BaseField = Felt>,
        H: ElementHasher<BaseField = Felt>,
    {
        let ctx = self.ctx.0;
        let (n, lde_size) = (air.trace_length(), air.lde_domain_size());
        let mut segs = self.segs.borrow_mut();

        // ConstraintEvaluationTable::into_poly (constraints/evaluation_table.rs:166-190) +
        // build_constraint_commitment (lib.rs:599-632)
        let (cols, divisors) = constraint_evaluations.into_raw_parts(); // accessor added beside into_poly
        let col_ptrs: Vec<*const u64> = cols.iter().map(|c| c.as_ptr() as *const u64).collect();
        let divs: Vec<ffi::aero_divisor> = divisors.iter().map(to_aero_divisor).collect();
        let (mut comp, mut root) = (ptr::null_mut(), [0u8; 32]);
        check(ctx, unsafe {
            ffi::aero_constraints_into_poly(ctx, col_ptrs.as_ptr(), divs.as_ptr(), divs.len() as u32, lde_size as u64, n as u64, &mut comp)
        });
        check(ctx, unsafe { ffi::aero_segment_commit_polys(comp, air.options().blowup_factor() as u32, root.as_mut_ptr()) });
        segs.push(Segment { ctx, h: comp, root, width: air.context().num_constraint_composition_columns() });
        channel.commit_constraints(digest_from_bytes(&root));

        // OOD frame (trace/poly_table.rs:59-72, composition_poly.rs:93-96)
        let z: E = channel.get_ood_point();
        let w: usize = segs[..segs.len() - 1].iter().map(|s| s.width).sum();
        let m = segs.last().unwrap().width;
        let trace_h: Vec<*mut ffi::aero_segment> = segs[..segs.len() - 1].iter().map(|s| s.h).collect();
        let (mut ood_trace, mut ood_comp) = (vec![0u64; 2 * w], vec![0u64; m]);
        check(ctx, unsafe {
            ffi::aero_ood_eval(ctx, trace_h.as_ptr(), trace_h.len() as u32, comp, as_u64(z), ood_trace.as_mut_ptr(), ood_comp.as_mut_ptr())
        });
        channel.send_ood_trace_states(&[from_u64_slice(&ood_trace[..w]), from_u64_slice(&ood_trace[w..])]);
        channel.send_ood_constraint_evaluations(&from_u64_slice::<E>(&ood_comp));

        // DEEP composition (composer/mod.rs:71-252): W triples, m singles, the degree-adjustment pair
        let cc = channel.get_deep_composition_coeffs();
        let flat: Vec<u64> = cc.trace.iter().flat_map(|t| [as_u64(t.0), as_u64(t.1), as_u64(t.2)])
            .chain(cc.constraints.iter().map(|&c| as_u64(c)))
            .chain([as_u64(cc.degree.0), as_u64(cc.degree.1)]).collect();
        let mut fri = ptr::null_mut();
        check(ctx, unsafe {
            ffi::aero_deep_compose(ctx, trace_h.as_ptr(), trace_h.len() as u32, comp, as_u64(z), ood_trace.as_ptr(), ood_comp.as_ptr(), flat.as_ptr(), &mut fri)
        });

        // FRI commit phase (fri/src/prover/mod.rs:166-218): one root out, one alpha in, per layer
        let layers = air.options().to_fri_options().num_fri_layers(lde_size);
        for _ in 0..=layers {
            check(ctx, unsafe { ffi::aero_fri_commit_layer(fri, root.as_mut_ptr()) });
            winter_fri::ProverChannel::commit_fri_layer(&mut channel, digest_from_bytes(&root));
            let alpha: E = winter_fri::ProverChannel::draw_fri_alpha(&mut channel);
            check(ctx, unsafe { ffi::aero_fri_fold(fri, as_u64(alpha)) });
        }

        // grinding stays on the channel (channel.rs:151-167); aero_pow_min_nonce returns the same
        // minimum nonce as the serial build and can replace it
        channel.grind_query_seed();
        let positions = channel.get_query_positions();
        let pos64: Vec<u64> = positions.iter().map(|&p| p as u64).collect();

        // query phase (prover/src/lib.rs:518-539): FriProver::build_proof bytes and the rows + batch
        // Merkle proof of every segment in ONE host round trip (aero_open_queries)
        let ns = segs.len();
        let mut fri_buf = vec![0u8; 1 << 22];
        let mut fri_len = fri_buf.len();
        let mut rows: Vec<Vec<u64>> = segs.iter().map(|s| vec![0u64; pos64.len() * s.width]).collect();
        let mut paths: Vec<Vec<u8>> = (0..ns).map(|_| vec![0u8; 2 + pos64.len() * (1 + 32 * 64)]).collect();
        let mut path_len: Vec<usize> = paths.iter().map(|p| p.len()).collect();
        let seg_h: Vec<*mut ffi::aero_segment> = segs.iter().map(|s| s.h).collect();
        let rows_p: Vec<*mut u64> = rows.iter_mut().map(|r| r.as_mut_ptr()).collect();
        let paths_p: Vec<*mut u8> = paths.iter_mut().map(|p| p.as_mut_ptr()).collect();
        check(ctx, unsafe {
            ffi::aero_open_queries(ctx, fri, seg_h.as_ptr(), ns as u32, pos64.as_ptr(), pos64.len() as u32, fri_buf.as_mut_ptr(),
                                   &mut fri_len, rows_p.as_ptr(), paths_p.as_ptr(), path_len.as_mut_ptr())
        });
        let fri_proof = FriProof::read_from(&mut SliceReader::new(&fri_buf[..fri_len])).expect("FRI proof bytes");
        unsafe { ffi::aero_fri_destroy(fri) };
        let mut queries: Vec<Queries> = (0..ns)
            .map(|i| {
                paths[i].truncate(path_len[i]);
                let values = rows[i].iter().flat_map(|v| v.to_le_bytes()).collect();
                Queries::from_raw_parts(values, std::mem::take(&mut paths[i]))
            })
            .collect();
        let constraint_queries = queries.pop().unwrap();
        let trace_queries = queries;
        segs.clear();
        Ok(channel.build_proof(trace_queries, constraint_queries, fri_proof))
    }
}

/// air::ConstraintDivisor (air/src/air/divisor.rs:14-17) with a single (x^a - b) numerator term,
/// which is all the Miden AIR uses.
fn to_aero_divisor(d: &winter_air::ConstraintDivisor<Felt>) -> ffi::aero_divisor {
    let (a, b) = d.numerator()[0];
    let mut out = ffi::aero_divisor { a: a as u64, b: as_u64(b), n_exemptions: d.exemptions().len() as u32, exemptions: [0; 8] };
    for (o, e) in out.exemptions.iter_mut().zip(d.exemptions()) {
        *o = as_u64(*e);
    }
    out
}

/// Montgomery-form memory image of a base-field element (no conversion: the ABI takes this form).
fn as_u64<E: FieldElement<BaseField = Felt>>(e: E) -> u64 {
    debug_assert_eq!(E::EXTENSION_DEGREE, 1);
    unsafe { *(&e as *const E as *const u64) }
}
fn from_u64_slice<E: FieldElement<BaseField = Felt>>(v: &[u64]) -> Vec<E> {
    v.iter().map(|x| unsafe { *(x as *const u64 as *const E) }).collect()
}
