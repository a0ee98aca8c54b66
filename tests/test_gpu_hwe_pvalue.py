"""hwe_tester's p-value on the device against the reference's arithmetic (oracle, glibc libm): every operation is
bit-identical by construction (FMA-free, __d*_rn) except exp().  BASELINE.json: any difference must stay within 1e-12
relative and be reported — this test reports it (see also the `hwe_pvalue` object in bench.py's JSON line)."""
import numpy as np
import pytest

from hwe_triples import triples

pytestmark = pytest.mark.gpu


def test_hwe_pvalue_max_relative_difference(cuda_api, oracle):
    c = triples()
    assert len(c) >= 1_000_000
    dev = cuda_api.hwe_pvalues(c)
    ref = oracle.hwe_pvalues(c)
    nz = ref != 0
    assert np.array_equal(dev == 0, ref == 0)
    rel = np.abs(dev[nz] - ref[nz]) / np.abs(ref[nz])
    fd, sd = oracle.p_text_diffs(dev, ref)
    print(f"\nHWE p-value, device vs reference arithmetic over {len(c)} triples: max relative difference {rel.max():.3e}, "
          f"{int((dev != ref).sum())} values differ in the last bits, {fd} FILE-mode texts and {sd} stdin-mode texts differ")
    assert rel.max() <= 1e-12          # the tolerance BASELINE.json's north_star states
    assert fd == 0 and sd == 0         # and none of them changes a printed digit on this set
