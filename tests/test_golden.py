"""Golden vectors generated from the compiled reference (tests/golden/make_golden.py) checked against
(a) the plain-C oracle on CPU and (b) the CUDA library on the GPU.  Bit-exact."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import make_cases, OracleBackend, GpuBackend, load_golden  # noqa: E402
from make_golden import digest  # noqa: E402

BD = [(0, 8), (1, 10)]


def compare(backend, hbd, bd):
    gold = load_golden(hbd)
    cases = make_cases(hbd, bd)
    assert len(gold) == len(cases)
    kinds = set()
    for i, ((kind, p, x), g) in enumerate(zip(cases, gold)):
        assert (digest(x) == g["_digest"]).all(), "case %d inputs drifted from the fixture" % i
        out = backend.run(kind, p, x)
        for k, v in out.items():
            assert np.array_equal(np.asarray(v), g[k]), (i, kind, p, k)
        kinds.add(kind)
    assert kinds == {"sad", "interp", "txfm", "intra", "cfl", "me", "filters"}


@pytest.mark.parametrize("hbd,bd", BD)
def test_oracle_matches_reference_golden(hbd, bd):
    compare(OracleBackend(hbd, bd), hbd, bd)


@pytest.mark.gpu
@pytest.mark.parametrize("hbd,bd", BD)
def test_cuda_matches_reference_golden(hbd, bd):
    compare(GpuBackend(hbd, bd), hbd, bd)
