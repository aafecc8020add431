"""CPU: host-side chain (l-major) layout logic."""
import numpy as np
import pytest

from conftest import load_case
from kore_b200 import chain, synthetic

ALL = ["spinover", "dormy", "jones", "magnetic_small", "forced_small", "forced_small_eig", "m0_small"]


@pytest.mark.parametrize("name", ALL)
def test_params_chain_is_block_tridiagonal(name):
    c = load_case(name)
    assert sorted(c.perm.tolist()) == list(range(c.n))
    assert chain.check_block_tridiagonal(c.A.indptr, c.A.indices, c.perm, c.nodeptr)
    if c.B is not None:
        assert chain.check_block_tridiagonal(c.B.indptr, c.B.indices, c.perm, c.nodeptr)
    sizes = set(np.diff(c.nodeptr).tolist())
    N1 = c.meta["N1"]
    assert sizes <= {N1, 2 * N1, 3 * N1}


@pytest.mark.parametrize("name", ["spinover", "dormy", "magnetic_small", "m0_small"])
def test_pattern_chain_is_block_tridiagonal(name):
    c = load_case(name)
    perm, nodeptr = chain.chain_from_pattern(c.A.indptr, c.A.indices, c.n, c.meta["N1"])
    assert sorted(perm.tolist()) == list(range(c.n))
    assert chain.check_block_tridiagonal(c.A.indptr, c.A.indices, perm, nodeptr)


def test_scrambled_chain_is_rejected_by_checker():
    c = load_case("m0_small")
    rng = np.random.default_rng(0)
    perm = rng.permutation(c.n)
    assert not chain.check_block_tridiagonal(c.A.indptr, c.A.indices, perm, c.nodeptr)


def test_ell_matches_reference_lists():
    # utils.py:174-183 probes quoted in SURVEY.md A.1
    top, bot, ll = chain.ell(1, 64, -1)
    assert top[:3].tolist() == [2, 4, 6] and bot[:3].tolist() == [1, 3, 5]
    top, bot, ll = chain.ell(9, 156, 1)
    assert top[:2].tolist() == [9, 11] and bot[:2].tolist() == [10, 12]
    top, bot, ll = chain.ell(0, 39, 1)
    assert ll[0] == 1 and ll[-1] == 40


def test_split_ranges_cover_chain():
    for P, G in [(64, 1), (64, 8), (149, 4), (5, 8)]:
        r = chain.split_ranges(P, G)
        assert r[0][0] == 0 and r[-1][1] == P
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1


def test_synthetic_has_kore_structure():
    A, B, perm, nodeptr = synthetic.synthetic_pencil(12, 24)
    n = A.shape[0]
    assert A.dtype == np.complex128 and B.dtype == np.float64
    assert chain.check_block_tridiagonal(A.indptr, A.indices, perm, nodeptr)
    # B: block diagonal with empty boundary rows (singular), Frobenius norm 1
    assert abs(np.sqrt((B.data ** 2).sum()) - 1.0) < 1e-12
    assert (np.diff(B.indptr) == 0).sum() == 6 * 4 + 6 * 2
    # deterministic
    A2, B2, _, _ = synthetic.synthetic_pencil(12, 24)
    assert (A != A2).nnz == 0 and (B != B2).nnz == 0


def test_sweep_deal_covers_all_items_once():
    from kore_b200 import sweep
    items = list(range(256))
    for world in (1, 2, 4, 8):
        parts = [sweep.deal(items, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == items
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
