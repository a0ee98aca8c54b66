"""The kernels' logic without a GPU: the product's own kernel and C-ABI sources, compiled with g++ against
the warp emulator in tests/emu/ (every CUDA thread a fiber, warp collectives resolved when all 32 lanes
arrive), run through the same parity cases as the `-m gpu` suite and compared byte for byte with the oracle.

This is a development check (divergent collectives, wrong tallies, stray writes show up here in seconds
instead of on a GPU box); it proves nothing about the device build — the `-m gpu` tests do that — and it is
not a CPU path of the product: only this test file loads the emulator library.
"""
import inspect
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "emu"))
import test_gpu_parity as G


@pytest.fixture(scope="module")
def emu_api():
    import build_emu
    return build_emu.load_api()


def _clone(fn):
    sig = inspect.signature(fn)
    params = [p.replace(name="emu_api") if p.name == "cuda_api" else p for p in sig.parameters.values()]

    def w(**kw):
        kw["cuda_api"] = kw.pop("emu_api")
        return fn(**kw)
    w.__signature__ = sig.replace(parameters=params)
    w.__name__ = fn.__name__ + "_emulated"
    w.__doc__ = fn.__doc__
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name != "gpu"]
    if marks:
        w.pytestmark = marks
    return w


# the heaviest cases (130 k-line files, 40,000-sample lines) stay on the GPU
_SKIP = {"test_more_rows_than_the_default_record_capacity", "test_phase_checker_more_dropped_lines_than_the_event_list"}
for _name, _fn in sorted(vars(G).items()):
    if _name.startswith("test_") and callable(_fn) and _name not in _SKIP:
        globals()[_name + "_emulated"] = _clone(_fn)


def test_skip_ahead_loop_through_the_bulk_copy_ring(emu_api, oracle, monkeypatch):
    """The emulator build compiles the bulk-copy ring of the skip-ahead loop in (the product build leaves it out: measured, did
    not pay): with VCFX_C4_BULK=1 the multi-key shapes go through it."""
    from vcfx_b200 import synth
    monkeypatch.setenv("VCFX_C4_BULK", "1")
    emu_api.close_cached_contexts()
    try:
        G.test_multikey_format_exceptions(cuda_api=emu_api, oracle=oracle)
        for V, S in ((60, 300), (12, 2504)):
            data = synth.make_vcf(4, V, S, seed=17)
            G.run_all(emu_api, oracle, data, f"ring shape4 {V}x{S}", tools=("af", "hwe"))
            G.run_all(emu_api, oracle, data, f"ring shape4 {V}x{S} tile512", tile_bytes=512, tools=("af", "hwe"))
    finally:
        emu_api.close_cached_contexts()
