"""NumPy model of the assembly kernels (kore_b200/csrc/kb_assemble.cu): evaluates an
`AsmProgram` into a canonical CSR with exactly the operations, in exactly the order, the kernel
performs (separate IEEE multiplications and additions, no fused multiply-add).  Test
infrastructure: lets the CPU suite check the program builder of kore_b200/assembly.py against
the reference-assembled fixtures bit for bit, and gives the GPU test its expected output."""
import numpy as np
import scipy.sparse as sp


def _block_values(prog, blk):
    """(re, im) band arrays [N1, W] of block `blk`."""
    N1, W = prog.N1, 2 * prog.H + 1
    acc = [None, None]
    for g in range(prog.blk_grp[blk], prog.blk_grp[blk + 1]):
        lin = None
        for t in range(prog.grp_term[g], prog.grp_term[g + 1]):
            p = prog.term_coef[t] * prog.ops[prog.term_op[t]]
            lin = p if lin is None else lin + p
        for k in range(prog.grp_nsc[g]):
            lin = prog.grp_sc[g, k] * lin
        if prog.grp_sign[g] < 0:
            lin = -lin
        part = prog.grp_part[g]
        acc[part] = lin if acc[part] is None else acc[part] + lin
    z = np.zeros((N1, W))
    return (z if acc[0] is None else acc[0]), (z if acc[1] is None else acc[1])


def evaluate(prog):
    """The CSR the kernel writes: rows ascending, columns ascending, exact zeros dropped."""
    N1, H = prog.N1, prog.H
    W = 2 * H + 1
    rows, cols, vre, vim = [], [], [], []
    i_idx = np.repeat(np.arange(N1), W).reshape(N1, W)
    j_idx = i_idx + np.arange(W)[None, :] - H
    inside = (j_idx >= 0) & (j_idx < N1)
    fs = prog.final_scale if prog.use_final else None
    for br in range(prog.nblockrows):
        chop = prog.br_chop[br]
        for blk in range(prog.blk_ptr[br], prog.blk_ptr[br + 1]):
            bc = prog.blk_col[blk]
            re, im = _block_values(prog, blk)
            if fs is not None:
                re, im = re * fs, im * fs
            keep = inside & ((re != 0) | (im != 0))
            keep[:chop, :] = False
            rows.append(br * N1 + i_idx[keep])
            cols.append(bc * N1 + j_idx[keep])
            vre.append(re[keep])
            vim.append(im[keep])
        if chop > 0:
            dense = prog.bc[prog.br_bc[br]:prog.br_bc[br] + chop]
            if fs is not None:
                dense = dense * fs
            q, j = np.nonzero(dense)
            rows.append(br * N1 + q)
            cols.append(br * N1 + j)
            vre.append(dense[q, j])
            vim.append(np.zeros(len(q)))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    vre, vim = np.concatenate(vre), np.concatenate(vim)
    if prog.is_complex:
        data = np.empty(len(vre), dtype=np.complex128)
        data.real, data.imag = vre, vim
    else:
        data = vre
    n = prog.n
    M = sp.csr_matrix((data, (rows, cols)), shape=(n, n))
    M.sort_indices()
    return M
