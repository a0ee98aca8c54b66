"""ctypes front-end of the CPU restatement (``oracle/vcfx_oracle.c``) and helpers to run the
compiled reference tools (``oracle/_ref/VCFX_*``).

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` — never from ``vcfx_b200``.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
LIB_PATH = REF_DIR / "liboracle.so"

FILE, STDIN = 0, 1
AC_MT_TEXT, AC_STREAM, AC_UNIFIED = 0, 1, 2
AC_TEXT, AC_AGGREGATE, AC_BINARY = 0, 1, 2


class _Result(C.Structure):
    _fields_ = [("out", C.c_void_p), ("out_len", C.c_size_t), ("rc", C.c_int),
                ("data_lines", C.c_longlong), ("rows", C.c_longlong), ("flagged", C.c_longlong),
                ("warnings", C.c_longlong), ("first_bad_line", C.c_longlong)]


class Result:
    def __init__(self, r: _Result):
        self.out = C.string_at(r.out, r.out_len) if r.out else b""
        self.rc = r.rc
        self.data_lines = r.data_lines
        self.rows = r.rows
        self.flagged = r.flagged
        self.warnings = r.warnings
        self.first_bad_line = r.first_bad_line


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            subprocess.run(["make", "-C", str(HERE), "all"], check=True, capture_output=True)
        l = C.CDLL(str(LIB_PATH))
        P = C.POINTER(_Result)
        for name in ("oracle_allele_freq", "oracle_hwe", "oracle_missing", "oracle_nonref_filter", "oracle_indexer", "oracle_phase_checker", "oracle_dosage"):
            getattr(l, name).argtypes = [C.c_char_p, C.c_size_t, C.c_int, P]
        l.oracle_variant_count.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, P]
        l.oracle_inbreeding.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, P]
        l.oracle_allele_counter.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p, P]
        l.oracle_free.argtypes = [P]
        for name in ("oracle_fmt_af_file", "oracle_fmt_af_stdin", "oracle_fmt_p_file", "oracle_fmt_p_stdin"):
            getattr(l, name).argtypes = [C.c_double, C.c_char_p]
            getattr(l, name).restype = C.c_int
        l.oracle_hwe_pvalue.argtypes = [C.c_int, C.c_int, C.c_int]
        l.oracle_hwe_pvalue.restype = C.c_double
        l.oracle_hwe_pvalues.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        l.oracle_hwe_pvalues.restype = None
        l.oracle_hwe_class.argtypes = [C.c_char_p, C.c_size_t]
        l.oracle_gt_index.argtypes = [C.c_char_p, C.c_size_t]
        l.oracle_af_counts.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.oracle_ac_counts.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = l
    return _lib


def _call(fn, *args) -> Result:
    r = _Result()
    # (data, length, ...): the length as size_t — a bare Python int is passed as a C int and loses the bits above 2^31
    args = tuple(C.c_size_t(a) if i == 1 else a for i, a in enumerate(args))
    fn(*args, C.byref(r))
    res = Result(r)
    lib().oracle_free(C.byref(r))
    return res


def allele_freq(data: bytes, mode: int = FILE) -> Result:
    return _call(lib().oracle_allele_freq, data, len(data), mode)


def hwe(data: bytes, mode: int = FILE) -> Result:
    return _call(lib().oracle_hwe, data, len(data), mode)


def missing(data: bytes, mode: int = FILE) -> Result:
    return _call(lib().oracle_missing, data, len(data), mode)


def phase_checker(data: bytes, mode: int = FILE) -> Result:
    return _call(lib().oracle_phase_checker, data, len(data), mode)


def phase_checker_stderr(data: bytes, mode: int = FILE) -> bytes:
    """What VCFX_phase_checker prints on stderr without -q."""
    l = lib()
    l.oracle_phase_checker_err.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_Result), C.POINTER(_Result)]
    r, e = _Result(), _Result()
    l.oracle_phase_checker_err(data, len(data), mode, C.byref(r), C.byref(e))
    out = C.string_at(e.out, e.out_len) if e.out_len else b""
    l.oracle_free(C.byref(r)); l.oracle_free(C.byref(e))
    return out


DS_WARNING = b"Warning: Skipping VCF line with fewer than 10 fields.\n"
DS_ERROR = b"Error: VCF header (#CHROM) not found before variant records.\n"


def dosage(data: bytes, mode: int = FILE) -> Result:
    """VCFX_dosage_calculator; .warnings = number of DS_WARNING lines on stderr (file mode: not with -q), .first_bad_line = 1:
    DS_ERROR and nothing on stdout."""
    return _call(lib().oracle_dosage, data, len(data), mode)


def genotype_query(data: bytes, query: str, mode: int = FILE, strict: bool = False):
    """VCFX_genotype_query -g query [--strict]: (Result with stdout, stderr text of a run without -q)."""
    l = lib()
    l.oracle_genotype_query.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_int, C.POINTER(_Result), C.POINTER(_Result)]
    r, e = _Result(), _Result()
    l.oracle_genotype_query(data, len(data), mode, query.encode(), int(strict), C.byref(r), C.byref(e))
    res = Result(r)
    err = C.string_at(e.out, e.out_len) if e.out_len else b""
    l.oracle_free(C.byref(r)); l.oracle_free(C.byref(e))
    return res, err


IB_GLOBAL, IB_SKIP_BOUNDARY, IB_COUNT_BOUNDARY, IB_QUIET = 1, 2, 4, 8
IB_MESSAGES = {0: b"", 1: b"Error: Empty file.\n", 2: b"Error: No #CHROM line or no samples found.\n", 3: b"No biallelic variants found.\n",
               4: b"Error: No #CHROM line found.\n", 5: b"Error: No sample columns found.\n"}


def inbreeding(data: bytes, mode: int = FILE, flags: int = 0) -> Result:
    """VCFX_inbreeding_calculator; .warnings = key of IB_MESSAGES (what goes to stderr), .rows = sites used."""
    return _call(lib().oracle_inbreeding, data, len(data), mode, flags)


def indexer(data: bytes, mode: int = FILE) -> Result:
    return _call(lib().oracle_indexer, data, len(data), mode)


def nonref_filter(data: bytes, mode: int = FILE) -> Result:
    return _call(lib().oracle_nonref_filter, data, len(data), mode)


def variant_count(data: bytes, mode: int = FILE, strict: bool = False) -> Result:
    return _call(lib().oracle_variant_count, data, len(data), mode, int(strict))


def allele_counter(data: bytes, path: int = AC_MT_TEXT, fmt: int = AC_TEXT, limit: int = 0,
                   samples: str | None = None) -> Result:
    s = samples.encode() if samples else None
    return _call(lib().oracle_allele_counter, data, len(data), path, fmt, limit, s)


def fmt(kind: str, v: float) -> bytes:
    buf = C.create_string_buffer(512)
    n = getattr(lib(), f"oracle_fmt_{kind}")(v, buf)
    return buf.raw[:n]


def hwe_pvalue(hr: int, het: int, ha: int) -> float:
    return lib().oracle_hwe_pvalue(hr, het, ha)


def p_text_diffs(a, b):
    """(#pairs whose FILE-mode text differs, #pairs whose "%.6f" text differs) for two float64 arrays."""
    import numpy as np
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    fd, sd = C.c_size_t(0), C.c_size_t(0)
    l = lib()
    l.oracle_p_text_diffs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    l.oracle_p_text_diffs.restype = None
    l.oracle_p_text_diffs(a.ctypes.data, b.ctypes.data, len(a), C.byref(fd), C.byref(sd))
    return fd.value, sd.value


def hwe_pvalues(counts):
    """counts: int32 array [n, 3] (homRef, het, homAlt) -> float64 array [n], the reference's arithmetic (glibc exp)."""
    import numpy as np
    c = np.ascontiguousarray(counts, dtype=np.int32).reshape(-1, 3)
    out = np.empty(len(c), dtype=np.float64)
    lib().oracle_hwe_pvalues(c.ctypes.data, len(c), out.ctypes.data)
    return out


# ---------------------------------------------------------------- compiled reference tools
def ref_tool(name: str) -> Path | None:
    p = REF_DIR / f"VCFX_{name}"
    return p if p.exists() else None


def have_reference() -> bool:
    return all(ref_tool(t) for t in ("allele_freq_calc", "allele_counter", "missing_detector",
                                     "variant_counter", "hwe_tester"))


def run_ref(name: str, args: list[str], stdin: bytes | None = None, timeout: float = 20):
    """Run oracle/_ref/VCFX_<name>; returns (rc, stdout, stderr)."""
    exe = ref_tool(name)
    if exe is None:
        raise FileNotFoundError(f"oracle/_ref/VCFX_{name} not built (make -C oracle ref)")
    r = subprocess.run([str(exe), *args], input=stdin if stdin is not None else b"",
                       capture_output=True, timeout=timeout)
    return r.returncode, r.stdout, r.stderr
