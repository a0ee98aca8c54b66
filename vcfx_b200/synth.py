"""ctypes front-end of the synthetic VCF generator (``csrc/vcfx_synth.c``).

Shapes follow BASELINE.json ``configs`` / SURVEY.md §8(d): 1 = C1 (unphased biallelic GT),
2 = C2 (1000G chr21 shape, phased GT), 3 = C3 (C2 + missing/unphased/haploid),
4 = C4 (multi-allelic ``GT:AD:DP:GQ:PL``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build


class _Cfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("shape", C.c_uint32), ("n_samples", C.c_uint32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.SYNTH_LIB
        if not path.exists():
            _build.build_synth()
        lib = C.CDLL(str(path))
        lib.vcfx_synth_header.restype = C.c_size_t
        lib.vcfx_synth_header.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_size_t]
        lib.vcfx_synth_lines.restype = C.c_size_t
        lib.vcfx_synth_lines.argtypes = [C.POINTER(_Cfg), C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_int]
        lib.vcfx_synth_line_bound.restype = C.c_size_t
        lib.vcfx_synth_line_bound.argtypes = [C.POINTER(_Cfg)]
        _lib = lib
    return _lib


def header(shape: int, n_samples: int, seed: int = 1) -> bytes:
    lib = _load()
    cfg = _Cfg(seed, shape, n_samples)
    need = lib.vcfx_synth_header(C.byref(cfg), None, 0)
    buf = C.create_string_buffer(need)
    n = lib.vcfx_synth_header(C.byref(cfg), buf, need)
    return buf.raw[:n]


def line_bound(shape: int, n_samples: int) -> int:
    return _load().vcfx_synth_line_bound(C.byref(_Cfg(0, shape, n_samples)))


def lines_into(dst: np.ndarray, shape: int, n_samples: int, first: int, count: int,
               seed: int = 1, threads: int | None = None) -> int:
    """Write variants [first, first+count) into the uint8 array ``dst``; returns bytes written."""
    lib = _load()
    cfg = _Cfg(seed, shape, n_samples)
    if threads is None:
        threads = min(32, os.cpu_count() or 1)
    assert dst.dtype == np.uint8 and dst.flags["C_CONTIGUOUS"]
    n = lib.vcfx_synth_lines(C.byref(cfg), first, count, dst.ctypes.data, dst.size, threads)
    if n > dst.size:
        raise ValueError(f"synthetic buffer too small: need {n}, have {dst.size}")
    return n


def make_vcf(shape: int, n_variants: int, n_samples: int, seed: int = 1, first: int = 0,
             with_header: bool = True, threads: int | None = None) -> bytes:
    """A whole synthetic VCF as bytes (small/medium sizes; large runs use ``lines_into``)."""
    hdr = header(shape, n_samples, seed) if with_header else b""
    cap = line_bound(shape, n_samples) * max(n_variants, 1)
    buf = np.empty(cap, dtype=np.uint8)
    n = lines_into(buf, shape, n_samples, first, n_variants, seed, threads) if n_variants else 0
    return hdr + buf[:n].tobytes()
