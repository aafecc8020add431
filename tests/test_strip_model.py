"""The strip Gauss-Jordan algebra of kb_chainfac.cu (NumPy statement) inverts."""
import numpy as np
import pytest

from strip_gj_model import strip_invert, strip_invert_stream


@pytest.mark.parametrize("b,w", [(7, 3), (20, 4), (33, 9), (64, 8), (50, 1), (9, 9)])
def test_strip_invert(b, w):
    rng = np.random.default_rng(b * 100 + w)
    S = rng.standard_normal((b, b)) + 1j * rng.standard_normal((b, b))
    S[np.arange(b), np.arange(b)] *= 1e-8  # the diagonal is never an acceptable pivot
    M = strip_invert(S, w)
    assert np.abs(M @ S - np.eye(b)).max() < 1e-10 * np.linalg.cond(S)


@pytest.mark.parametrize("b,w", [(7, 3), (20, 4), (33, 9), (64, 8), (9, 9)])
def test_column_stream_equals_composite_transforms(b, w):
    # the elementary transforms streamed column by column (default path of kb_chainfac.cu) and the
    # composite transform of a whole strip (KB_CHAINFAC_NOSTREAM=1) are the same elimination
    rng = np.random.default_rng(b * 100 + w + 1)
    S = rng.standard_normal((b, b)) + 1j * rng.standard_normal((b, b))
    S[np.arange(b), np.arange(b)] *= 1e-8
    M1, M2 = strip_invert(S, w), strip_invert_stream(S, w)
    assert np.abs(M2 @ S - np.eye(b)).max() < 1e-10 * np.linalg.cond(S)
    assert np.abs(M1 - M2).max() < 1e-10 * np.abs(M1).max() * np.linalg.cond(S)
