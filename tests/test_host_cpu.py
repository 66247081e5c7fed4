"""CPU tests of the product's host side: the shared library loads, exports every symbol the headers
declare, refuses to run without a GPU (no CPU fallback), and its host Fiat-Shamir code agrees with
the oracle / golden vectors.  No compute kernels are launched here."""
import hashlib
import os
import re

import pytest

import aero_b200
from oracle import stark_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("aero_b200.h", "aero_prover.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(aero_[a-z0-9_]+)\s*\(", text))
    return names - {"aero_aux_builder", "aero_constraint_evaluator", "aero_status"}  # function-pointer typedefs


def test_library_exports_every_declared_symbol():
    lib = aero_b200.load()
    declared = _declared_symbols()
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(lib, name), "libaero_b200.so does not export %s" % name
    # ...and the ctypes table covers exactly the declared ABI
    assert set(aero_b200._lib.PROTOTYPES) == declared


def test_library_is_built_for_sm_100a():
    so_path = aero_b200.build.LIB
    assert os.path.exists(so_path)
    blob = open(so_path, "rb").read()
    assert b"sm_100a" in blob, "no sm_100a cubin embedded"


def test_context_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(aero_b200.AeroError) as e:
        aero_b200.Context()
    assert e.value.status == aero_b200.AERO_ERR_CUDA


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (it is test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "aero_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("starkoracles", ""), "%s mentions the oracle" % f


def test_host_blake2s_and_hash_elements():
    data = bytes(range(256)) * 2
    for ln in (0, 1, 32, 40, 63, 64, 65, 128, 129, 500):
        assert aero_b200.host_blake2s(data[:ln]) == hashlib.blake2s(data[:ln]).digest()
    for elems in ([3], [1, 2], list(range(81)), [so.P - 1] * 8):
        assert aero_b200.host_hash_elements(elems) == so.hash_elements(elems)


def test_host_coin_known_answers_and_oracle_agreement():
    inp, _ = so.read_proof_file(os.path.join(ROOT, "tests", "golden", "fib.bin"))
    seed = so.miden_pub_inputs_seed(inp)
    c = aero_b200.RandomCoin(seed)
    assert c.draw() == 15636605459427237624  # tests/integration/test_verifier.cairo:104
    assert c.draw_integers(20, 64) == [55, 46, 17, 44, 61, 8, 43, 39, 19, 3, 26, 31, 30, 4, 37, 40, 49, 7, 56, 29]
    a, b = aero_b200.RandomCoin(b"x"), so.RandomCoin(b"x")
    for step in range(20):
        if step % 3 == 0:
            d = so.blake2s(b"%d" % step)
            a.reseed(d)
            b.reseed(d)
        elif step % 7 == 0:
            a.reseed_with_int(step)
            b.reseed_with_int(step)
        assert a.draw() == b.draw()
        assert a.seed == b.seed
        assert a.leading_zeros() == b.leading_zeros()
        assert a.check_leading_zeros(step) == b.check_leading_zeros(step)
    assert a.draw_integers(27, 8192) == b.draw_integers(27, 8192)
    with pytest.raises(RuntimeError):
        a.draw_integers(64, 64)  # num_values must be smaller than the domain (random/mod.rs:262-265)


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference: JSON line with the contract keys, from the oracle port, on the workload
    the line names (here shrunk with --log-rows), with the steps / warm-up that actually ran and all host
    threads even under torchrun's OMP_NUM_THREADS=1."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--log-rows", "11"], capture_output=True, text=True, check=True, env=env)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "trace_rows_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["log_rows"] == 11 and "2^11 rows" in line["config"]["workload"]
    assert line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]


def test_rust_shim_bindings_match_the_headers():
    """rust/aero-gpu-prover/src/ffi.rs is generated from include/*.h; it must be current and bind every
    symbol the shared object exports through those headers (the crate itself cannot be compiled here)."""
    import subprocess
    import sys
    from aero_b200 import _lib

    gen = os.path.join(ROOT, "tools", "gen_rust_ffi.py")
    assert subprocess.run([sys.executable, gen, "--check"]).returncode == 0, "run tools/gen_rust_ffi.py"
    ffi = open(os.path.join(ROOT, "rust", "aero-gpu-prover", "src", "ffi.rs")).read()
    for name in _lib.PROTOTYPES:
        assert "pub fn %s(" % name in ffi, "%s is not bound in ffi.rs" % name


def test_rust_patches_apply_and_export_what_the_shim_uses(tmp_path):
    """rust/patches/*.patch are the complete reference-side change set of the shim crate: they apply to the
    reference sources, and every item the crate imports from winter-prover or calls on a winterfell type is
    public after them (the crate itself cannot be compiled here: no Rust toolchain)."""
    import re
    import shutil
    import subprocess

    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "winterfell")):
        pytest.skip("reference checkout not present on this box")
    patches = sorted(os.listdir(os.path.join(ROOT, "rust", "patches")))
    assert [p[:4] for p in patches] == ["0001", "0002", "0003", "0004", "0005"]
    touched = set()
    for p in patches:
        for line in open(os.path.join(ROOT, "rust", "patches", p)):
            if line.startswith("+++ b/"):
                touched.add(line[6:].strip())
    for rel in touched:  # a private copy of exactly the files the patches touch
        os.makedirs(os.path.dirname(tmp_path / rel), exist_ok=True)
        shutil.copy(os.path.join(ref, rel), tmp_path / rel)
    for p in patches:
        r = subprocess.run(["patch", "-p1", "--forward", "-i", os.path.join(ROOT, "rust", "patches", p)], cwd=tmp_path,
                           capture_output=True, text=True)
        assert r.returncode == 0, "%s does not apply: %s" % (p, r.stdout + r.stderr)
    prover_lib = open(tmp_path / "winterfell/prover/src/lib.rs").read()
    exported = set()
    for m in re.finditer(r"pub use [\w:]+::\{([^}]*)\};|pub use [\w:]+::(\w+);", prover_lib):
        exported |= {x.strip() for x in (m.group(1) or m.group(2)).split(",") if x.strip()}
    exported |= set(re.findall(r"^pub (?:trait|struct|enum|fn) (\w+)", prover_lib, flags=re.M))  # defined in lib.rs itself
    shim = "".join(open(os.path.join(ROOT, "rust", "aero-gpu-prover", "src", f)).read() for f in ("lib.rs", "dump.rs"))
    imported = set()
    for m in re.finditer(r"use winter_prover::\{([^}]*)\};", shim):
        imported |= {x.strip() for x in m.group(1).split(",") if x.strip()}
    assert imported and "channel::ProverChannel" not in imported
    assert imported <= exported, "not exported by winter-prover: %s" % sorted(imported - exported)
    # constructors / accessors the shim calls that the reference does not have without the patches
    for name, rel in (("from_root", "winterfell/crypto/src/merkle/mod.rs"), ("from_raw_parts", "winterfell/air/src/proof/queries.rs"),
                      ("into_raw_parts", "winterfell/prover/src/constraints/evaluation_table.rs"),
                      ("divisors", "winterfell/prover/src/constraints/evaluation_table.rs")):
        assert re.search(r"pub fn %s\b" % name, open(tmp_path / rel).read()), name
        assert ("%s(" % name) in shim
        assert not re.search(r"pub fn %s\b" % name, open(os.path.join(ref, rel)).read()), "%s exists upstream: drop the patch" % name
    main_rs = open(tmp_path / "miden-proof-generator/src/main.rs").read()
    assert "GpuExecutionProver::new(inner, 0)" in main_rs and "DumpingProver::new(inner" in main_rs


def test_air_program_builder_marshals_the_c_struct():
    """AirProgramBuilder -> aero_air_program (include/aero_b200.h): node order, operand indices, the constant
    pool and the per-constraint arrays arrive in the C layout the kernel-side validation expects."""
    from aero_b200 import AirProgramBuilder
    from aero_b200 import _lib

    b = AirProgramBuilder()
    c0, n0, k = b.cur(0), b.next(0), b.const(7)
    t = b.sub(n0, b.mul(c0, k))
    b.transition(t, 5)
    b.assertion(0, 1, 3, 1)
    p, keep = b.finish()
    assert p.n_nodes == 5 and p.n_consts == 1 and p.n_transition == 1 and p.n_boundary == 1
    ops = [(p.nodes[i].op, p.nodes[i].a, p.nodes[i].b) for i in range(p.n_nodes)]
    assert ops == [(_lib.AERO_AIR_CUR, 0, 0), (_lib.AERO_AIR_NEXT, 0, 0), (_lib.AERO_AIR_CONST, 0, 0),
                   (_lib.AERO_AIR_MUL, c0, k), (_lib.AERO_AIR_SUB, n0, 3)]
    assert p.consts[0] == 7 and p.transition_out[0] == t and p.transition_adj[0] == 5
    assert (p.boundary_col[0], p.boundary_value[0], p.boundary_adj[0], p.boundary_div[0]) == (0, 1, 3, 1)
    # the header's enum and the ctypes constants agree
    import re, os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "aero_b200.h")).read()
    m = re.search(r"enum \{ AERO_AIR_CUR = (\d), AERO_AIR_NEXT = (\d), AERO_AIR_CONST = (\d), AERO_AIR_ADD = (\d), AERO_AIR_SUB = (\d), AERO_AIR_MUL = (\d),\s+AERO_AIR_PERIODIC = (\d) \}", hdr)
    assert m and [int(x) for x in m.groups()] == [_lib.AERO_AIR_CUR, _lib.AERO_AIR_NEXT, _lib.AERO_AIR_CONST,
                                                  _lib.AERO_AIR_ADD, _lib.AERO_AIR_SUB, _lib.AERO_AIR_MUL,
                                                  _lib.AERO_AIR_PERIODIC]
    # periodic columns: cycle values concatenated, one length per column
    b = AirProgramBuilder()
    k0, k1 = b.periodic_column([1, 2]), b.periodic_column([3, 4, 5, 6])
    b.transition(b.mul(b.periodic(k1), b.periodic(k0)), 1)
    p, keep = b.finish()
    assert p.n_periodic == 2 and [p.periodic_len[i] for i in range(2)] == [2, 4]
    assert [p.periodic_values[i] for i in range(6)] == [1, 2, 3, 4, 5, 6]
    assert (p.nodes[0].op, p.nodes[0].a) == (_lib.AERO_AIR_PERIODIC, k1)


def test_periodic_column_table_of_the_library_equals_the_oracle():
    """aero_periodic_column_table (the host arithmetic behind AERO_AIR_PERIODIC nodes: interpolation over the
    cycle, evaluation over offset^(n / cycle) * <w_(cycle * ce_blowup)>) against the oracle's restatement of
    PeriodicValueTable::new (prover/src/constraints/periodic_table.rs:25-75), on the reference's own test
    columns (:110-120), MaskedChainAir's, and random columns up to a cycle as long as the trace."""
    import numpy as np
    from aero_b200.prover import periodic_column_table
    from oracle.air import MaskedChainAir, SimpleAir, P

    def oracle_table(cols, n, ce_blowup):
        class A(SimpleAir):
            trace_width = 1
            transition_degrees = [ce_blowup + 1 if ce_blowup > 2 else 2]
            periodic_columns = cols

            def get_assertions(self):
                return []
        a = A(n, 0)
        assert a.ce_blowup == ce_blowup
        return a.periodic_value_table()

    rng = np.random.default_rng(7)
    cases = [([[1, 2], [3, 4, 5, 6]], 32, 2), (MaskedChainAir.periodic_columns, 64, 4),
             ([[int(v) % P for v in rng.integers(0, 2**63, c, dtype=np.uint64)] for c in (2, 16, 64)], 64, 8)]
    for cols, n, ceb in cases:
        want = oracle_table(cols, n, ceb)
        for col, w in zip(cols, want):
            got = periodic_column_table(col, n, ceb)
            assert [int(v) for v in got] == w
    # Air::get_periodic_column_polys' assertions (air/src/air/mod.rs:319-335)
    for bad in ([1], [1, 2, 3], [1] * 64):
        with pytest.raises(aero_b200.AeroError):
            periodic_column_table(bad, 32, 2)
    with pytest.raises(aero_b200.AeroError):
        periodic_column_table([1, P], 32, 2)


def test_copy_pool_stress(tmp_path):
    """aero_b200/host/copy_pool.hpp (the worker pool behind the pageable-column staging): many batches of
    different chunk sizes and four concurrent callers copy exactly, compiled with g++ and run here."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "copy_pool_stress")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-pthread", "-I", os.path.join(root, "aero_b200", "host"),
                           os.path.join(root, "tests", "cpp", "copy_pool_stress.cpp"), "-o", exe], env=env)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "copy pool ok" in r.stdout, r.stdout + r.stderr
