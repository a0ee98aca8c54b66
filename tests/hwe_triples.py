"""(homRef, het, homAlt) triples for the HWE p-value comparison: what a 2,504-sample cohort can produce, plus larger
and degenerate cohorts.  Deterministic."""
import numpy as np


def triples(n_random: int = 1_000_000, seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = []
    # every split of small cohorts (exhaustive up to N = 40)
    small = [(a, b, c) for a in range(41) for b in range(41 - a) for c in range(41 - a - b)]
    out.append(np.array(small, dtype=np.int32))
    # cohorts of 2,504 under Hardy-Weinberg with a random allele frequency, and far from it
    q = rng.beta(0.3, 1.5, size=n_random // 2)
    N = 2504
    ha = rng.binomial(N, q * q); het = rng.binomial(N - ha, np.clip(2 * q * (1 - q) / np.maximum(1 - q * q, 1e-12), 0, 1))
    out.append(np.stack([N - ha - het, het, ha], axis=1).astype(np.int32))
    r = rng.integers(0, 2505, size=(n_random // 4, 3))
    out.append(r.astype(np.int32))
    big = rng.integers(0, 200_000, size=(n_random // 4, 3))
    out.append(big.astype(np.int32))
    return np.concatenate(out, axis=0)
