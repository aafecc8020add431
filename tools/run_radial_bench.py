#!/usr/bin/env python3
"""Time kore_b200/radial.py against the UNMODIFIED bin/submatrices.py (build container: needs /root/reference
for the second column) on the truncations of BASELINE.json's metric -- Kore's rule for E = 1e-7, 1e-8, 1e-9 and the
nominal N = 600 -- and compare the operators entry by entry.  Host work on both sides: ours one core, the
reference `ncpus` processes.  Usage: tools/run_radial_bench.py [out.json] [ncpus]"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kore_b200 import assembly as asm, radial  # noqa: E402
from make_case import patch_params, REF, STANDIN  # noqa: E402

CASES = [("hydro E=1e-7 (Kore rule)", "tests/spinover/params.spinover", dict(N=428, lmax=424, Ek=1e-7)),
         ("hydro E=1e-8 nominal", "tests/spinover/params.spinover", dict(N=600, lmax=600, Ek=1e-8)),
         ("hydro E=1e-8 (Kore rule)", "tests/spinover/params.spinover", dict(N=676, lmax=672, Ek=1e-8)),
         ("hydro E=1e-9 (Kore rule)", "tests/spinover/params.spinover", dict(N=1072, lmax=1072, Ek=1e-9)),
         ("thermal + dipole field, N=300", "tests/dormy2004/params.dormy04", dict(N=300, lmax=308, magnetic=1, B0="'dipole'", forcing=0))]


def reference_run(params, ov, ncpus):
    work = tempfile.mkdtemp(prefix="radbench_")
    shutil.copytree(os.path.join(REF, "bin"), os.path.join(work, "bin"))
    subprocess.check_call(["chmod", "-R", "u+w", work])
    txt = patch_params(open(os.path.join(REF, params)).read(), {k: str(v) for k, v in ov.items()})
    open(os.path.join(work, "bin", "parameters.py"), "w").write(txt)
    env = dict(os.environ, PYTHONPATH=os.path.abspath(STANDIN), PYTHONWARNINGS="ignore")
    if ov.get("magnetic"):
        subprocess.check_call([sys.executable, "bin/compute_profiles.py"], cwd=work, env=env, stdout=subprocess.DEVNULL)
    t = time.time()
    subprocess.check_call([sys.executable, "bin/submatrices.py", str(ncpus)], cwd=work, env=env, stdout=subprocess.DEVNULL)
    dt = time.time() - t
    probe = subprocess.check_output(
        [sys.executable, "-c", "import sys; sys.path.insert(0,'bin'); import parameters as p, utils as u, json;"
         "d={k:getattr(p,k) for k in %r if hasattr(p,k)}; d['rcmb']=u.rcmb;"
         "print(json.dumps({k:(v.item() if hasattr(v,'item') else v) for k,v in d.items()}))"
         % (list(asm.PhysicsParams.__dataclass_fields__),)], cwd=work, env=env).decode().strip().splitlines()[-1]
    ops = asm.load_operators(work)
    shutil.rmtree(work)
    return dt, asm.PhysicsParams.from_dict(json.loads(probe)), ops


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "radial_bench.json"
    ncpus = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rows = []
    for name, params, ov in CASES:
        t_ref, pp, ref = reference_run(params, ov, ncpus)
        radial._table_cache.clear()
        t = time.time()
        mine = radial.radial_operators(pp)
        t_blas = time.time() - t
        radial._table_cache.clear()
        t = time.time()
        radial.radial_operators(pp, dot="ordered")
        t_ord = time.time() - t
        same = sorted(mine) == sorted(ref) and all(np.array_equal(mine[k].toarray(), ref[k].toarray()) for k in ref)
        rows.append(dict(case=name, N=pp.N, operators=len(ref), reference_s=round(t_ref, 2), reference_processes=ncpus,
                         radial_s=round(t_blas, 3), radial_ordered_s=round(t_ord, 3), radial_cores=1,
                         bit_identical=bool(same), speedup_wall=round(t_ref / t_blas, 1)))
        print(json.dumps(rows[-1]), flush=True)
    json.dump(dict(host_cpus=os.cpu_count(), rows=rows), open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
