//! `DumpingProver`: the UNMODIFIED reference CPU prover, additionally writing a trace-dump fixture
//! (format: `aero_b200/fixture.py` of the aero_b200 repository) with everything needed to replay the proof
//! through `aero_prove` on a B200 and compare bytes: the trace segments as the prover saw them, the merged
//! constraint evaluations and their divisors, the options, the public inputs and the proof.
//!
//! It overrides the same provided methods as `GpuExecutionProver` but only to observe their arguments:
//! each delegates to the inner `ExecutionProver`, i.e. to winter-prover's default implementation
//! (winterfell/prover/src/lib.rs:384-632).  Enable with `--features dump-fixture` of
//! miden-proof-generator (rust/patches/0005); `tests/test_fixture_replay.py` replays any `*.aerofix`
//! placed under `tests/golden/`.
//!
//! Not compiled in the aero_b200 build image (no Rust toolchain there).

use std::cell::RefCell;
use std::fs::File;
use std::io::Write;

use miden_air::{Felt, ProcessorAir, PublicInputs};
use miden_processor::ExecutionTrace;
use miden_prover::ExecutionProver;
use winter_air::proof::StarkProof;
use winter_air::{Air, ProofOptions};
use winter_crypto::{ElementHasher, MerkleTree};
use winter_math::{FieldElement, StarkField};
use winter_prover::{
    ConstraintEvaluationTable, Matrix, Prover, ProverChannel, ProverError, StarkDomain, Trace, TraceCommitment,
    TracePolyTable,
};
use winter_utils::Serializable;

#[derive(Default)]
struct Captured {
    segments: Vec<Vec<Vec<u64>>>, // trace segments in commitment order, column-major, canonical
    divisors: Vec<(u64, u64, Vec<u64>)>,
    ce_cols: Vec<Vec<u64>>,
    header: Vec<u32>, // log2 n, aux_rands, ce_blowup, n_constraint_coeffs
    options: Vec<u8>,
}

pub struct DumpingProver {
    inner: ExecutionProver,
    path: String,
    cap: RefCell<Captured>,
}

fn canonical<E: FieldElement<BaseField = Felt>>(col: &[E]) -> Vec<u64> {
    // E == Felt on this path (FieldExtension::None); as_int() is the canonical value (f64/mod.rs:234)
    col.iter().map(|e| e.base_element(0).as_int()).collect()
}

impl DumpingProver {
    pub fn new(inner: ExecutionProver, path: &str) -> Self {
        Self { inner, path: path.to_string(), cap: Default::default() }
    }

    /// `Prover::prove` + the fixture file.
    pub fn prove_and_dump(&self, trace: ExecutionTrace) -> Result<StarkProof, ProverError> {
        let mut pub_inputs_bytes = Vec::new();
        self.get_pub_inputs(&trace).write_into(&mut pub_inputs_bytes);
        let meta = trace.get_info().meta().to_vec();
        let proof = self.prove(trace)?;
        let cap = self.cap.borrow();
        let (main, aux) = (&cap.segments[0], cap.segments.get(1));
        let mut f = File::create(&self.path).expect("fixture file");
        let mut w = |b: &[u8]| f.write_all(b).unwrap();
        w(b"AEROFIX1");
        let aux_w = aux.map_or(0, |a| a.len()) as u32;
        for v in [cap.header[0], main.len() as u32, aux_w, cap.header[1], cap.divisors.len() as u32, cap.header[2], cap.header[3], 0] {
            w(&v.to_le_bytes());
        }
        w(&cap.options);
        w(&(pub_inputs_bytes.len() as u32).to_le_bytes());
        w(&pub_inputs_bytes);
        w(&(meta.len() as u32).to_le_bytes());
        w(&meta);
        for seg in cap.segments.iter() {
            for col in seg {
                for v in col {
                    w(&v.to_le_bytes());
                }
            }
        }
        for (a, b, ex) in cap.divisors.iter() {
            w(&a.to_le_bytes());
            w(&b.to_le_bytes());
            w(&(ex.len() as u32).to_le_bytes());
            w(&0u32.to_le_bytes());
            for k in 0..8 {
                w(&ex.get(k).copied().unwrap_or(0).to_le_bytes());
            }
        }
        for col in cap.ce_cols.iter() {
            for v in col {
                w(&v.to_le_bytes());
            }
        }
        let proof_bytes = proof.to_bytes();
        w(&(proof_bytes.len() as u32).to_le_bytes());
        w(&proof_bytes);
        Ok(proof)
    }
}

impl Prover for DumpingProver {
    type BaseField = Felt;
    type Air = ProcessorAir;
    type Trace = ExecutionTrace;

    fn get_pub_inputs(&self, trace: &ExecutionTrace) -> PublicInputs {
        self.inner.get_pub_inputs(trace)
    }
    fn options(&self) -> &ProofOptions {
        self.inner.options()
    }

    /// Called once per trace segment (main, then each auxiliary segment), winterfell/prover/src/lib.rs:239,328.
    fn build_trace_commitment<E, H>(&self, trace: &Matrix<E>, domain: &StarkDomain<Felt>) -> (Matrix<E>, MerkleTree<H>, Matrix<E>)
    where
        E: FieldElement<BaseField = Felt>,
        H: ElementHasher<BaseField = Felt>,
    {
        self.cap.borrow_mut().segments.push(trace.columns().map(|c| canonical(c)).collect());
        self.inner.build_trace_commitment::<E, H>(trace, domain)
    }

    fn prove_after_constraint_eval<E, H>(
        &self,
        air: &ProcessorAir,
        channel: ProverChannel<ProcessorAir, E, H>,
        constraint_evaluations: ConstraintEvaluationTable<E>,
        trace_polys: TracePolyTable<E>,
        trace_commitment: TraceCommitment<E, H>,
    ) -> Result<StarkProof, ProverError>
    where
        E: FieldElement<BaseField = Felt>,
        H: ElementHasher<BaseField = Felt>,
    {
        {
            let mut cap = self.cap.borrow_mut();
            // `evaluations` is a public field; the divisors need rust/patches/0004 -- read them through a
            // clone of the table's parts so that the table itself goes on to the reference's into_poly
            cap.ce_cols = constraint_evaluations.evaluations.iter().map(|c| canonical(c)).collect();
            cap.divisors = constraint_evaluations
                .divisors()
                .iter()
                .map(|d| {
                    let (a, b) = d.numerator()[0];
                    (a as u64, b.as_int(), d.exemptions().iter().map(|e| e.as_int()).collect())
                })
                .collect();
            let ctx = air.context();
            let aux_rands = if air.trace_info().layout().num_aux_segments() > 0 {
                air.trace_info().layout().get_aux_segment_rand_elements(0) as u32
            } else {
                0
            };
            cap.header = vec![
                air.trace_length().trailing_zeros(),
                aux_rands,
                ctx.ce_blowup_factor() as u32, // = number of composition columns
                2 * (ctx.num_transition_constraints() + ctx.num_assertions()) as u32, // air/src/air/mod.rs:511-533
            ];
            let o = air.options();
            cap.options = vec![
                o.num_queries() as u8,
                o.blowup_factor() as u8,
                o.grinding_factor() as u8,
                o.hash_fn() as u8,
                o.field_extension() as u8,
                o.to_fri_options().folding_factor() as u8,
            ];
            cap.options.extend_from_slice(&(o.to_fri_options().max_remainder_size() as u16).to_le_bytes());
        }
        self.inner.prove_after_constraint_eval::<E, H>(air, channel, constraint_evaluations, trace_polys, trace_commitment)
    }
}
