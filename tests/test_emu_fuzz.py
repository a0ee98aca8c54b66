"""Differential fuzzing without a GPU: seeded random inputs (adversarial text, the synthetic shapes, allele_counter's domain),
random tile and chunk sizes, either mode — every tool's Python twin over the kernels' own sources under the warp emulator
(tests/emu/) against the oracle.  A longer run of the same loop (`python tests/test_emu_fuzz.py SEED SECONDS`) found the
nonref_filter quick-check bug of round 2 (a three-byte column with GT not the first key)."""
import random
import sys
import time
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "emu"))
sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def one_case(api, O, rng, data=None):
    import vcfgen
    from vcfx_b200 import synth
    seed = rng.randrange(10 ** 9)
    kind = rng.random()
    if data is not None:
        d = data; kind = 0.7                  # (allele_counter runs as well)
    elif kind < 0.6:
        d = vcfgen.make_vcf(seed, n_lines=rng.randrange(1, 120), n_samples=rng.randrange(1, 40), crlf=rng.random() < 0.2,
                            final_newline=rng.random() < 0.7, header=rng.choice(["normal", "normal", "late", "none", "double"]))
    elif kind < 0.8:
        d = synth.make_vcf(rng.choice([1, 2, 3, 4]), rng.randrange(1, 60), rng.randrange(0, 400), seed=seed % 1000)
    else:
        d = vcfgen.make_vcf(seed, n_lines=rng.randrange(1, 60), n_samples=rng.randrange(1, 12), domain="ac", final_newline=rng.random() < 0.7,
                            header=rng.choice(["normal", "late", "double"]))
    kw = {}
    t = rng.choice([0, 512, 1024, 4096]); c = rng.choice([0, 0, 4096, 16384])
    if t:
        kw["tile_bytes"] = t
    if c and max(len(x) for x in d.split(b"\n")) < c - 64:
        kw["chunk_bytes"] = c
    mode = rng.choice([0, 1])
    tag = f"seed {seed} mode {mode} {kw}"
    r = api.allele_freq_calc(d, mode, **kw); o = O.allele_freq(d, mode); assert (r.out, r.rc) == (o.out, o.rc), f"af {tag}"
    r = api.hwe_tester(d, mode, **kw); o = O.hwe(d, mode); assert r.out == o.out, f"hwe {tag}"
    if not (mode == 0 and max(len(x) for x in d.split(b"\n")) > 60000):       # (the reference's file mode has a 64 KB line buffer)
        r = api.missing_detector(d, mode, **kw); o = O.missing(d, mode); assert r.out == o.out, f"md {tag}"
    st = rng.random() < 0.5
    r = api.variant_counter(d, mode, st, **kw); o = O.variant_count(d, mode, st); assert (r.out, r.rc) == (o.out, o.rc), f"vc {tag}"
    if kind >= 0.6:
        r = api.allele_counter(d, **kw); o = O.allele_counter(d); assert (r.out, r.rc) == (o.out, o.rc), f"ac {tag}"
        r = api.allele_counter(d, api.AC_UNIFIED, api.AC_AGGREGATE, **kw); o = O.allele_counter(d, O.AC_UNIFIED, O.AC_AGGREGATE); assert (r.out, r.rc) == (o.out, o.rc), f"ac -a {tag}"
        r = api.allele_counter(d, api.AC_STREAM, **kw); o = O.allele_counter(d, O.AC_STREAM); assert (r.out, r.rc) == (o.out, o.rc), f"ac stream {tag}"
    r = api.nonref_filter(d, mode, **kw); o = O.nonref_filter(d, mode); assert r.out == o.out, f"nr {tag}"
    r = api.indexer(d, mode, **kw); o = O.indexer(d, mode); assert r.out == o.out, f"ix {tag}"
    r = api.phase_checker(d, mode, **kw); o = O.phase_checker(d, mode); assert (r.out, r.err) == (o.out, O.phase_checker_stderr(d, mode)), f"pc {tag}"
    fl = rng.randrange(8)
    r = api.inbreeding_calculator(d, mode, bool(fl & 1), bool(fl & 2), bool(fl & 4), quiet=False, **kw); o = O.inbreeding(d, mode, fl)
    assert (r.out, r.err) == (o.out, O.IB_MESSAGES[o.warnings]), f"ib flags {fl} {tag}"
    q = rng.choice(["0/1", "1/1", "0|1", "0/0", "1/2", "./.", "0/x", "2/1"]); strict = rng.random() < 0.3
    r = api.genotype_query(d, q, mode, strict, **kw); o, e = O.genotype_query(d, q, mode, strict); assert (r.out, r.err) == (o.out, e), f"gq {q} {strict} {tag}"
    r = api.dosage_calculator(d, mode, **kw); o = O.dosage(d, mode); assert (r.out, r.rc) == (o.out, o.rc), f"ds {tag}"


def lattice_file(rng):
    """Lines of three-byte genotypes (the four-byte lattice the fast paths work on) of random widths and phases, with samples that
    leave the lattice at a random rate, GT first / with a second key / behind another key, CRLF, a final tab, odd lines between."""
    base = [b"0|0", b"0|1", b"1|0", b"1|1", b"0/0", b"0/1", b"1/1"]
    odd = [b"0", b".", b"./.", b".|.", b"0|1:3", b"10|1", b"2|1", b" 0|1", b"", b"0|1|1", b"1", b"0/", b"|1", b"1|2", b"00|1", b"0|.", b"./1", b"0|1:9:9", b"0\r"]
    S = rng.choice([rng.randrange(1, 40), rng.randrange(100, 300), rng.randrange(120, 135), rng.randrange(250, 262), rng.randrange(500, 1500)])
    fmt = rng.choice([b"GT", b"GT", b"GT", b"GT:DP", b"DP:GT"])
    p_odd = rng.choice([0, 0, 0.002, 0.02, 0.2])
    lines = []
    for _ in range(rng.randrange(1, 12)):
        if rng.random() < 0.05:
            lines.append(rng.choice([b"", b"#c\tx", b"1\t2\tx"])); continue
        gts = []
        for _i in range(S):
            g = rng.choice(base) if rng.random() >= p_odd else rng.choice(odd)
            if fmt == b"GT:DP" and rng.random() < 0.9:
                g += b":7"
            if fmt == b"DP:GT":
                g = b"7:" + g
            gts.append(g)
        ln = b"%d\t%d\t%s\tA\t%s\t.\tPASS\t%s\t%s\t" % (rng.randrange(1, 23), 10 ** rng.randrange(0, 9), b"r" * rng.randrange(0, 6),
                                                     rng.choice([b"G", b"G", b"G,T", b""]), b"x" * rng.randrange(0, 40), fmt) + b"\t".join(gts)
        lines.append(ln + (b"\t" if rng.random() < 0.05 else b""))
    eol = b"\r\n" if rng.random() < 0.15 else b"\n"
    hdr = b"##f" + eol + b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + b"\t".join(b"S%d" % i for i in range(S)) + eol
    return hdr + eol.join(lines) + (eol if rng.random() < 0.7 else b"")


@pytest.mark.parametrize("seed", range(4))
def test_random_lattice_lines(oracle, seed):
    import build_emu
    api = build_emu.load_api()
    rng = random.Random(2000 + seed)
    for _ in range(12):
        one_case(api, oracle, rng, lattice_file(rng))


@pytest.mark.parametrize("seed", range(6))
def test_random_inputs_and_geometries(oracle, seed):
    import build_emu
    api = build_emu.load_api()
    rng = random.Random(1000 + seed)
    for _ in range(25):
        one_case(api, oracle, rng)


if __name__ == "__main__":            # python tests/test_emu_fuzz.py [--gpu] SEED SECONDS   (--gpu: the real library on cuda:0)
    from oracle import oracle as O
    args = [a for a in sys.argv[1:] if a != "--gpu"]
    if "--gpu" in sys.argv:
        from vcfx_b200 import api
    else:
        import build_emu
        api = build_emu.load_api()
    rng = random.Random(int(args[0]) if args else 1)
    t0 = time.time(); n = 0
    while time.time() - t0 < float(args[1] if len(args) > 1 else 60):
        one_case(api, O, rng, lattice_file(rng) if n % 3 == 2 else None); n += 1
    print(f"{n} cases, no difference")
