"""Paths of the five executables that never reach the GPU — help, version, unopenable input, bad
options — against the compiled reference tools: same stdout, stderr and exit code.  (Everything that
parses data is in tests/test_gpu_cli.py.)"""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "vcfx_b200" / "bin"
REF = ROOT / "oracle" / "_ref"
TOOLS = ["allele_freq_calc", "hwe_tester", "missing_detector", "variant_counter", "allele_counter", "nonref_filter", "indexer", "phase_checker", "inbreeding_calculator", "genotype_query", "dosage_calculator"]


def run(exe, args):
    r = subprocess.run([str(exe), *args], input=b"", capture_output=True, timeout=60)
    # getopt prefixes its own messages with argv[0]: compare them without the directory
    err = r.stderr.replace(str(exe).encode(), exe.name.encode())
    return r.returncode, r.stdout, err


@pytest.fixture(scope="module", autouse=True)
def built():
    from vcfx_b200 import build
    build.build_cuda(); build.build_tools()
    if not all((REF / f"VCFX_{t}").exists() for t in TOOLS):
        pytest.skip("oracle/_ref reference tools not built (make -C oracle ref)")


@pytest.mark.parametrize("tool", TOOLS)
@pytest.mark.parametrize("args", [["--help"], ["-h"], ["--version"], ["-v"], ["-i", "/nonexistent/x.vcf"], ["/nonexistent/x.vcf"],
                                  ["--no-such-option"], ["-q", "-i", "/nonexistent/x.vcf"], ["--help", "--version"], ["-g", "0/1", "-i", "/nonexistent/x.vcf"]])
def test_same_as_reference_without_a_device(tool, args):
    mine = run(BIN / f"VCFX_{tool}", args)
    ref = run(REF / f"VCFX_{tool}", args)
    assert mine == ref, (tool, args, mine, ref)


def test_genotype_query_without_a_query_prints_its_usage_line():
    for args in ([], ["-q"], ["-i", "/nonexistent/x.vcf"], ["-g", ""]):
        mine = run(BIN / "VCFX_genotype_query", args)
        ref = run(REF / "VCFX_genotype_query", args)
        assert mine == ref, (args, mine, ref)
