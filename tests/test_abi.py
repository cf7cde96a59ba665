"""CPU tests: the C-ABI shared library loads without a GPU and exports exactly what include/sumcheck_b200.h declares;
host-only entry points (the transcript RNG) agree with the oracle; compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sumcheck_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sumcheck_b200 import capi
    assert os.path.exists(capi.lib_path()), "libsumcheck_b200.so not built (run __graft_entry__.build())"
    L = C.CDLL(capi.lib_path())
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sumcheck_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == names, "ctypes table and header disagree"


def test_rng_struct_layout():
    from sumcheck_b200 import capi
    assert C.sizeof(capi.RngState) == 8 * 8 + 2 * 8 + 128 + 8


def test_host_transcript_matches_oracle(orc):
    """sc_rng_* (host code of the product) against the oracle's Blake2b512Rng, same feed/sample sequence as rng.rs:128-158."""
    import random
    from sumcheck_b200 import Blake2b512Rng, IPForMLSumcheck
    rnd = random.Random(4)
    a, b = Blake2b512Rng.setup(), orc.Rng()
    for step in range(12):
        msg = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 16, 127, 128, 129, 300])))
        a.feed(msg); b.feed_bytes(msg)
        for _ in range(rnd.randrange(3)):
            assert np.array_equal(IPForMLSumcheck.sample_round(a).randomness, b.sample_fr())
        n = rnd.choice([0, 8, 63, 64, 65, 127, 128, 777])
        assert a.fill_bytes(n) == b.fill_bytes(n)
        assert a.next_u64() == b.next_u64()


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sumcheck_b200 import IPForMLSumcheck, ListOfProductsOfPolynomials, SumcheckError
    poly = ListOfProductsOfPolynomials.new(2)
    t = np.zeros((4, 4), dtype=np.uint64)
    poly.add_product([t], np.zeros(4, dtype=np.uint64))
    with pytest.raises(SumcheckError) as e:
        IPForMLSumcheck.prover_init(poly)
    assert e.value.code in (-10, -11)


def test_nv0_panics_before_touching_the_device():
    from sumcheck_b200 import IPForMLSumcheck, ListOfProductsOfPolynomials, Panic
    poly = ListOfProductsOfPolynomials.new(0)
    poly.add_product([np.zeros((1, 4), dtype=np.uint64)], np.zeros(4, dtype=np.uint64))
    with pytest.raises(Panic) as e:   # zero_polynomial_should_error (ml_sumcheck/test.rs:187-204)
        IPForMLSumcheck.prover_init(poly)
    assert e.value.code == -1


def test_synth_generators_agree(orc):
    """numpy twin == C helper in the product library == the oracle's generator."""
    from sumcheck_b200.synth import synth_table, synth_table_fast
    for seed in (0x5C0000, 0x5C0301, 1):
        a, b, c = synth_table(5000, seed), synth_table_fast(5000, seed), orc.synth_table(5000, seed)
        assert np.array_equal(a, b) and np.array_equal(a, c)


def test_host_interpolation_matches_the_model():
    """sc_fr_interpolate (csrc/host_fr.h: the host-side scalar routine that finishes every prover round, P(1) =
    P_prev(r) - P(0)) against the big-int model's interpolate_uni_poly for every degree the reference's tests reach."""
    import random
    from oracle import pymodel as pm
    from sumcheck_b200 import capi
    L = capi.lib()
    rnd = random.Random(2024)
    for n in list(range(1, 15)) + [33]:
        for trial in range(4):
            evals = [rnd.randrange(pm.P) for _ in range(n)]
            r = [0, 1, pm.P - 1, rnd.randrange(pm.P)][trial] if n > 1 else rnd.randrange(pm.P)
            ev = np.array([pm.to_mont_limbs(v) for v in evals], dtype=np.uint64).reshape(n, 4)
            rr = np.array(pm.to_mont_limbs(r), dtype=np.uint64)
            out = np.zeros(4, dtype=np.uint64)
            assert L.sc_fr_interpolate(ev.ctypes.data_as(capi.U64P), n, rr.ctypes.data_as(capi.U64P), out.ctypes.data_as(capi.U64P)) == 0
            assert pm.from_mont_limbs(out) == pm.interpolate(evals, r)
    assert L.sc_fr_interpolate(None, 0, None, None) != 0


def test_header_is_plain_c_and_links(tmp_path):
    """include/sumcheck_b200.h is the drop-in boundary: it must compile as C99 and as C++, and a C program must link against
    the shared library (entry points that need no GPU are called; without a device the compute ones fail loudly)."""
    import shutil
    import subprocess
    from sumcheck_b200 import capi
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "use.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "sumcheck_b200.h"\n'
        "int main(void) {\n"
        "  sc_blake2b512_rng rng; unsigned char out[8]; uint64_t fr[4];\n"
        "  sc_rng_setup(&rng); sc_rng_feed_bytes(&rng, (const uint8_t*)\"abc\", 3); sc_rng_fill_bytes(&rng, out, 8);\n"
        "  sc_rng_sample_fr(&rng, fr);\n"
        "  uint64_t evals[8] = {1,0,0,0, 2,0,0,0}, r[4] = {0,0,0,0}, v[4];\n"
        "  if (sc_fr_interpolate(evals, 2, r, v) != 0) return 2;\n"
        "  if (memcmp(v, evals, 32) != 0) return 3;   /* the interpolant at 0 is evals[0] */\n"
        "  printf(\"%d\\n\", sc_device_count() >= 0 ? 0 : 1);\n"
        "  return 0;\n}\n")
    inc = os.path.join(ROOT, "include")
    exe = tmp_path / "use"
    lib = capi.lib_path()
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe), lib,
                           "-Wl,-rpath," + os.path.dirname(lib)])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
