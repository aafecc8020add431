"""CPU: kb_savetxt (the library's threaded writer of Kore's *.field / eigenvalues0.dat files)
produces the bytes np.savetxt produces with its defaults (solve.py:275-311)."""
import io

import numpy as np
import pytest


def np_bytes(X):
    f = io.BytesIO()
    np.savetxt(f, X)
    return f.getvalue()


@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (1000, 10), (4097, 2), (5, 1)])
def test_real_and_imag_parts_match_numpy(lib, tmp_path, shape):
    rng = np.random.default_rng(shape[0])
    X = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * 10.0 ** rng.integers(-300, 300, shape)
    X = np.asfortranarray(X)  # the layout kb_eigs returns
    for part in ("real", "imag"):
        p = tmp_path / ("x_%s.field" % part)
        lib.savetxt(str(p), X, part)
        assert p.read_bytes() == np_bytes(getattr(X, part))
    # row slices of a column-major block (the field split of solve.py:163-190) without a copy
    p = tmp_path / "slice.field"
    lib.savetxt(str(p), X[shape[0] // 3:, :], "imag", nthreads=3)
    assert p.read_bytes() == np_bytes(X[shape[0] // 3:, :].imag)


def test_special_values_and_float_input(lib, tmp_path):
    X = np.array([[0.0, -0.0, 1.0], [np.inf, -np.inf, np.nan], [5e-324, 1.7976931348623157e308, -1e-310],
                  [1 / 3, 2 / 3, 1e22]])
    p = tmp_path / "s.dat"
    lib.savetxt(str(p), X)
    assert p.read_bytes() == np_bytes(X)
    # C-ordered, 1-D (one value per line, like np.savetxt) and append mode (timing.dat)
    v = np.linspace(-1, 1, 11)
    lib.savetxt(str(p), v)
    assert p.read_bytes() == np_bytes(v)
    lib.savetxt(str(p), np.array([12.5]), append=True)
    assert p.read_bytes() == np_bytes(v) + np_bytes(np.array([12.5]))
    # what np.loadtxt reads back is what was written (spin_doctor.py:27-61 reads the files this way)
    Y = np.random.default_rng(0).standard_normal((50, 4))
    lib.savetxt(str(p), Y)
    assert np.array_equal(np.loadtxt(str(p)), Y)


def test_empty_and_errors(lib, tmp_path):
    p = tmp_path / "e.dat"
    lib.savetxt(str(p), np.zeros((0, 3)))
    assert p.read_bytes() == b""
    with pytest.raises(lib.KoreB200Error):
        lib.savetxt(str(tmp_path / "no_such_dir" / "x.dat"), np.zeros((2, 2)))
