"""Storage codecs and the def_w decimal round trip of the CPU oracle, against known-answer vectors produced by the
reference's own device converters (tests/golden/fp16c_codec.npz, made by oracle/make_golden.py) and IEEE binary16."""
import os
import subprocess
import numpy as np
import pytest
from helpers import OracleBackend, ROOT, REF_DIR

G = np.load(os.path.join(ROOT, "tests", "golden", "fp16c_codec.npz"))
orc = OracleBackend()


def test_fp16c_decode_all_codes():
    got = np.array([orc.lib.orc_fp16c_decode(int(c)) for c in range(65536)], dtype=np.float32).view(np.uint32)
    assert np.array_equal(got, G["decode_all_codes"])


def test_fp16c_encode_known_answers():
    xs = G["encode_in"].view(np.float32)
    got = np.array([orc.lib.orc_fp16c_encode(float(x)) for x in xs], dtype=np.uint16)
    assert np.array_equal(got, G["encode_out"])


def test_fp16c_format_ranges():
    # ranges quoted in the reference comment, src/kernel.cpp:848: +-1.99951168, +-6.10351562E-5, +-2.98023224E-8
    assert orc.lib.orc_fp16c_decode(0x7FFF) == pytest.approx(1.99951168, rel=1e-7)
    assert orc.lib.orc_fp16c_decode(0x0800) == pytest.approx(6.10351562e-5, rel=1e-7)
    assert orc.lib.orc_fp16c_decode(0x0001) == pytest.approx(2.98023224e-8, rel=1e-7)
    # device encode does not saturate: 2.0 wraps the 4-bit exponent (SURVEY section 7 hard part 2)
    assert orc.lib.orc_fp16c_encode(2.0) == 0x0000


def test_fp16c_roundtrip_is_identity_on_codes():
    for c in list(range(0, 65536, 7)) + [0x7FFF, 0xFFFF, 0x8000, 0x0800, 0x07FF]:
        assert orc.lib.orc_fp16c_encode(orc.lib.orc_fp16c_decode(c)) == c


def test_fp16s_matches_ieee_binary16():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-2.2, 2.2, 20000), rng.uniform(-1, 1, 20000) * 10.0 ** rng.uniform(-13, 0, 20000),
                         [0.0, -0.0, 1.99902344, 1.9995, 2.0, 1.86446416e-9, 1.81898936e-12, 9e-13]]).astype(np.float32)
    want = (xs * np.float32(32768.0)).astype(np.float16).view(np.uint16)
    got = np.array([orc.lib.orc_fp16s_encode(float(x)) for x in xs], dtype=np.uint16)
    assert np.array_equal(got, want)
    codes = np.arange(65536, dtype=np.uint16)
    finite = np.isfinite(codes.view(np.float16))
    want_d = (codes.view(np.float16).astype(np.float32) * np.float32(3.0517578E-5))
    got_d = np.array([orc.lib.orc_fp16s_decode(int(c)) for c in codes], dtype=np.float32)
    assert np.array_equal(got_d[finite].view(np.uint32), want_d[finite].view(np.uint32))
    # ranges quoted in src/lbm.cpp:411
    assert orc.lib.orc_fp16s_decode(0x7BFF) == pytest.approx(1.99902344, rel=1e-7)
    assert orc.lib.orc_fp16s_decode(0x0001) == pytest.approx(1.81898936e-12, rel=1e-7)


def test_w_decimal_round_trip_examples():
    # to_string(float) keeps 1+8 significant digits (src/utilities.hpp:2745-2754); strtof of that is what the device sees
    assert orc.float_to_string(1.0 / 3.5) == "2.85714296E-1"  # BENCHMARK scene: nu=1 -> def_w 2.85714296E-1f
    assert orc.float_to_string(1.0) == "1.00000000"
    for nu in (1.0, 1.0 / 6.0, 0.01, 0.02, 0.05, 1e-3, 3.3e-5, 0.1 * 62 / 1000):
        w = orc.w_from_nu(nu)
        exact = np.float32(1.0) / (np.float32(3.0) * np.float32(nu) + np.float32(0.5))
        assert abs(np.float32(w) - exact) <= 2 * np.spacing(exact)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "ref_to_string")), reason="oracle/_ref not built (needs the mounted reference tree)")
def test_float_to_string_matches_reference_formatter():
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(0.05, 2.0, 3000), 10.0 ** rng.uniform(-30, 30, 2000), -rng.uniform(0, 5, 200),
                         [0.0, 1.0, 10.0, 0.1, 1.0 / 3.5, 1.9999999, 9.9999999, 0.99999994, 123456.78]]).astype(np.float32)
    inp = "\n".join(f"{b:08x}" for b in xs.view(np.uint32)) + "\n"
    ref = subprocess.run([os.path.join(REF_DIR, "ref_to_string")], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")[:-1]
    got = [orc.float_to_string(float(x)) for x in xs]
    assert got == ref
