import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O  # test infrastructure only
    O.build()
    return O


@pytest.fixture(scope="session")
def hg():
    import hypergen_b200
    return hypergen_b200


@pytest.fixture(scope="session")
def ctx(hg):
    """A live hg_ctx on cuda:0 — fails loudly (no fallback) when there is no GPU."""
    c = hg.Context(0)
    yield c
    c.close()


def random_dna(rng, n, p_n=0.0, p_lower=0.0, p_junk=0.0):
    """ASCII DNA with optional N runs, lower-case stretches and arbitrary bytes."""
    s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if p_lower:
        m = rng.random(n) < p_lower
        s[m] |= 0x20
    if p_n:
        m = rng.random(n) < p_n
        s[m] = ord("N")
    if p_junk:
        m = rng.random(n) < p_junk
        s[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
    return s
