"""The strip Gauss-Jordan algebra of kb_chainfac.cu (NumPy statement) inverts."""
import numpy as np
import pytest

from strip_gj_model import strip_invert


@pytest.mark.parametrize("b,w", [(7, 3), (20, 4), (33, 9), (64, 8), (50, 1), (9, 9)])
def test_strip_invert(b, w):
    rng = np.random.default_rng(b * 100 + w)
    S = rng.standard_normal((b, b)) + 1j * rng.standard_normal((b, b))
    S[np.arange(b), np.arange(b)] *= 1e-8  # the diagonal is never an acceptable pivot
    M = strip_invert(S, w)
    assert np.abs(M @ S - np.eye(b)).max() < 1e-10 * np.linalg.cond(S)
