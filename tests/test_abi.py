"""The C ABI without a GPU: the library builds for sm_100a, loads, exports every symbol that
include/vcfx_cuda.h declares, and refuses to work without a device instead of falling back."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from vcfx_b200 import api, build
    build.build_cuda()
    return C.CDLL(str(api.lib_path()))


def test_every_declared_symbol_is_exported(lib):
    from vcfx_b200 import api
    header = (ROOT / "include" / "vcfx_cuda.h").read_text()
    declared = sorted(set(re.findall(r"\b(vcfx_cuda_[a-z_]+)\s*\(", header)))
    assert declared, "no entry points found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/vcfx_cuda.h but not exported"
    assert sorted(api.EXPORTS) == declared, "vcfx_b200.api.EXPORTS is out of sync with the header"


def test_abi_version_and_strerror(lib):
    assert lib.vcfx_cuda_abi_version() == 3
    lib.vcfx_cuda_strerror.restype = C.c_char_p
    assert b"no CPU fallback" in lib.vcfx_cuda_strerror(-2)


def test_struct_layouts_match_the_header():
    """ctypes mirrors vs the C structs (sizes are what the C side was compiled with)."""
    from vcfx_b200 import api
    assert C.sizeof(api.ChunkInfo) == 32          # ABI 2: + file_offset, ABI 3: + format_cache_from
    assert C.sizeof(api.ChunkStats) == 12 * 8 + 8
    assert C.sizeof(api.Cfg) == 4 * 4 + 8 + 8 + 4 + 4 + 8 + 4 + 4 + 8 + 8 + 8


def test_no_device_means_error_not_fallback(lib):
    """In the build container there is no GPU: create() must fail with VCFX_E_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vcfx_b200 import api
    with pytest.raises(api.VcfxCudaError) as e:
        api.Context(api.OP_ALLELE_FREQ)
    assert e.value.code == -2
    with pytest.raises(api.VcfxCudaError):
        api.allele_freq_calc(b"#CHROM\n1\t1\t.\tA\tG\t.\t.\t.\tGT\t0/1\n")


def test_product_does_not_import_the_oracle():
    """Nothing under vcfx_b200/ may reference oracle/ (it is test infrastructure)."""
    for p in (ROOT / "vcfx_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h", ".c") and p.name != "build.py":
            txt = p.read_text(errors="ignore")
            assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, p
