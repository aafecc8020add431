"""CPU: the C-ABI library builds, loads, exports every symbol include/kore_b200.h declares,
fails loudly without a GPU, and its host-side projected-problem kernels are right."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "kore_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    L = lib.load()
    syms = header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(lib.EXPORTS) == syms


def test_no_cpu_fallback(lib):
    if have_gpu():
        pytest.skip("GPU present")
    with pytest.raises(lib.KoreB200Error) as e:
        lib.Solver(0)
    assert e.value.code == lib.KB_ENODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kore_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "kore_oracle" not in src and "oracle/" not in src.replace("tests/", ""), fn


@pytest.mark.parametrize("m", [1, 2, 5, 19, 25, 40])
def test_host_schur_matches_numpy(lib, m):
    import ctypes as C
    L = lib.load()
    rng = np.random.default_rng(m)
    H = np.asfortranarray(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    T = np.zeros((m, m), dtype=np.complex128, order="F")
    Q = np.zeros((m, m), dtype=np.complex128, order="F")
    sig = np.array([1j]); tau = np.array([1j])
    rc = L.kb_dbg_schur(m, H.ctypes.data, -1, sig.ctypes.data, tau.ctypes.data, T.ctypes.data, Q.ctypes.data)
    assert rc == 0
    assert np.allclose(np.tril(T, -1), 0)
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(m)) < 1e-12 * m
    assert np.linalg.norm(Q @ T @ Q.conj().T - H) < 1e-12 * m * np.linalg.norm(H)
    ev = np.sort_complex(np.linalg.eigvals(H))
    assert np.allclose(np.sort_complex(np.diag(T)), ev, atol=1e-10 * np.abs(ev).max())


@pytest.mark.parametrize("which", ["TM", "TR", "TI", "LM", "SM", "LR", "SR", "LI", "SI"])
def test_host_schur_ordering(lib, which):
    import kore_oracle as ko
    L = lib.load()
    m = 19
    rng = np.random.default_rng(7)
    H = np.asfortranarray(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    T = np.zeros((m, m), dtype=np.complex128, order="F")
    Q = np.zeros((m, m), dtype=np.complex128, order="F")
    sigma, tau = 0.2 + 1j, 0.1 + 0.9j
    sig = np.array([sigma]); ta = np.array([tau])
    rc = L.kb_dbg_schur(m, H.ctypes.data, lib.WHICH[which], sig.ctypes.data, ta.ctypes.data, T.ctypes.data,
                        Q.ctypes.data)
    assert rc == 0
    assert np.linalg.norm(Q @ T @ Q.conj().T - H) < 1e-11 * m * np.linalg.norm(H)
    lam = sigma + 1.0 / np.diag(T)
    key = ko.which_key(lam, which, tau)
    assert np.all(np.diff(key) >= -1e-9 * np.abs(key).max())


def test_defective_and_repeated_eigenvalues(lib):
    L = lib.load()
    m = 6
    H = np.asfortranarray(np.diag(np.ones(m - 1), 1).astype(np.complex128) + 2.0 * np.eye(m))  # Jordan block
    T = np.zeros((m, m), dtype=np.complex128, order="F")
    Q = np.zeros((m, m), dtype=np.complex128, order="F")
    z = np.array([0j])
    assert L.kb_dbg_schur(m, H.ctypes.data, -1, z.ctypes.data, z.ctypes.data, T.ctypes.data, Q.ctypes.data) == 0
    assert np.linalg.norm(Q @ T @ Q.conj().T - H) < 1e-12


def test_struct_layouts_match_the_header(lib, tmp_path):
    # the ctypes mirrors of kb_stats / kb_asm_program / kb_diag_params against the header as a C compiler lays
    # them out (gcc: sizeof and offsetof of every field)
    import ctypes as C
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"kb_stats": lib.KbStats, "kb_asm_program": lib.KbAsmProgram, "kb_diag_params": lib.KbDiagParams}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "kore_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f, _ in cls._fields_:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    src.append("return 0; }")
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(cfile), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, f)]) == getattr(cls, f).offset, (cname, f)
