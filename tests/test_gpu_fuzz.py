"""The seeded differential fuzz loop of tests/test_emu_fuzz.py on the real device: random inputs, random tile and chunk sizes,
either mode, all eleven tools against the oracle."""
import random

import pytest

import test_emu_fuzz as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(3))
def test_random_inputs_on_the_device(cuda_api, oracle, seed):
    rng = random.Random(3000 + seed)
    for i in range(30):
        F.one_case(cuda_api, oracle, rng, F.lattice_file(rng) if i % 3 == 2 else None)
