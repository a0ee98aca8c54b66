"""The numeric text of the device (csrc/vcfx_numfmt.cuh) compiled for the host, against the oracle
and glibc: AF in both modes, the HWE p-value op sequence, and both p-value formatters."""
import ctypes as C
import random
import struct
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "vcfx_b200" / "csrc" / "vcfx_numfmt_host.cpp"
SO = ROOT / "vcfx_b200" / "_numfmt_host.so"


@pytest.fixture(scope="module")
def host():
    hdr = ROOT / "vcfx_b200" / "csrc" / "vcfx_numfmt.cuh"
    if not SO.exists() or SO.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", "-o", str(SO), str(SRC)], check=True)
    h = C.CDLL(str(SO))
    h.vcfx_host_hwe_pvalue.restype = C.c_double
    h.vcfx_host_fmt_fixed.argtypes = [C.c_double, C.c_int, C.c_char_p]
    h.vcfx_host_fmt_p_file.argtypes = [C.c_double, C.c_char_p]
    h.vcfx_host_fmt_p_stdin.argtypes = [C.c_double, C.c_char_p]
    return h


def test_af_text_both_modes(host, oracle):
    rng = random.Random(1)
    buf = C.create_string_buffer(64)
    for _ in range(60000):
        t = rng.choice([rng.randrange(1, 5009), rng.randrange(1, 64), rng.randrange(1, 10 ** 6), 2 ** rng.randrange(1, 20)])
        a = rng.randrange(0, t + 1)
        n = host.vcfx_host_fmt_af_file(a, t, buf); g1 = buf.raw[:n]
        n = host.vcfx_host_fmt_af_stdin(a, t, buf); g2 = buf.raw[:n]
        assert g1 == oracle.fmt("af_file", a / t) and g2 == oracle.fmt("af_stdin", a / t), (a, t)
    n = host.vcfx_host_fmt_af_file(0, 0, buf)
    assert buf.raw[:n] == b"0.0000"


def test_hwe_value_and_text(host, oracle):
    rng = random.Random(2)
    buf = C.create_string_buffer(64)
    for _ in range(60000):
        N = rng.choice([5, 40, 100, 2504, 2504, 100000])
        hr = rng.randrange(0, N + 1); het = rng.randrange(0, N + 1 - hr)
        ha = rng.randrange(0, N + 1 - hr - het) if rng.random() < .5 else N - hr - het
        p1 = host.vcfx_host_hwe_pvalue(hr, het, ha); p2 = oracle.hwe_pvalue(hr, het, ha)
        assert p1 == p2, (hr, het, ha)
        n = host.vcfx_host_fmt_p_file(p2, buf); a1 = buf.raw[:n]
        n = host.vcfx_host_fmt_p_stdin(p2, buf); a2 = buf.raw[:n]
        assert a1 == oracle.fmt("p_file", p2) and a2 == oracle.fmt("p_stdin", p2), (hr, het, ha, p2)


def test_fixed_exact_is_printf(host):
    """fmt_fixed_exact is '%.Nf': exact binary value, round-half-even, incl. ties, tiny and subnormal."""
    rng = random.Random(3)
    buf = C.create_string_buffer(64)
    vals = [0.0, 1.0, 0.5, 0.00005, 0.00015, 0.00025, 0.03125, 0.0000005, 0.0000015, 1e-300, 5e-324, 999.9999995]
    for i in range(60000):
        if i % 3 == 0:
            vals.append(struct.unpack("<d", struct.pack("<Q", rng.getrandbits(62) % (0x3FF << 52)))[0])
        elif i % 3 == 1:
            vals.append(rng.random())
        else:
            vals.append(rng.randrange(0, 10 ** 6) / 10 ** rng.randrange(1, 7) + rng.choice([0, 5e-5, 5e-7, 1e-17]))
    for v in vals:
        for dg in (4, 6):
            n = host.vcfx_host_fmt_fixed(v, dg, buf)
            assert buf.raw[:n] == b"%.*f" % (dg, v), (v, dg)
