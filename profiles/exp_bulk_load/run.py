#!/usr/bin/env python
"""Runs the bulk-load experiment (bulk_load_exp.cu, built with the nvcc line in profiles/README.md) on cuda:0 over ~1 GB
of the multi-key shape (BASELINE C4) and, for contrast, of the GT-only shape (C2); one JSON line per variant."""
import ctypes as C, json, sys
from pathlib import Path
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
import numpy as np, torch
from vcfx_b200 import synth
lib = C.CDLL(str(HERE / "libbulk_exp.so"))
lib.bulk_exp_run.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_ulonglong)]
names = {0: "128-bit loads, three windows ahead in registers (the product's way)", 1: "cp.async.bulk, 2 x 2 KiB per warp",
         2: "cp.async.bulk, 4 x 2 KiB per warp", 3: "cp.async.bulk, 2 x 4 KiB per warp"}
for shape, V in ((4, 15000), (2, 100000)):
    data = synth.make_vcf(shape, V, 2504, seed=3)
    n = len(data) // (64 << 10) * (64 << 10)
    d = torch.empty(n + 8192, dtype=torch.uint8, device="cuda:0")
    d[:n].copy_(torch.frombuffer(bytearray(data[:n]), dtype=torch.uint8)); torch.cuda.synchronize()
    ref = None
    for v in range(4):
        ms = C.c_float(); t = (C.c_ulonglong * 4)()
        rc = lib.bulk_exp_run(v, d.data_ptr(), n, 5, C.byref(ms), t)
        tal = list(t)
        if ref is None: ref = tal
        print(json.dumps({"shape": shape, "bytes": n, "variant": names[v], "rc": rc, "ms": ms.value, "GBps": n / ms.value / 1e6 if ms.value else None,
                          "tallies_equal_to_variant_0": tal == ref, "tabs": tal[0], "newlines": tal[3]}), flush=True)
    del d; torch.cuda.empty_cache()
