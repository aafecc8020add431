#!/usr/bin/env python3
"""Produce the matrices the reference would hand to SLEPc, in THIS container.

Runs the UNMODIFIED reference stages (bin/submatrices.py, bin/assemble.py and,
for magnetic runs, bin/compute_profiles.py) from /root/reference in a scratch
directory, with a parameters file derived from one of the reference's own
params files plus `key=value` overrides, and a single-rank mpi4py stand-in
(tests/fixtures/mpi4py) on PYTHONPATH.  Nothing from the reference is copied
into this repository: only the resulting A.npz / B.npz / B_forced.npz (and a
meta.json describing the chain layout inputs) are written to --out.

Usage:
  tools/make_case.py --params tests/spinover/params.spinover --out /tmp/c1 [k=v ...]

Reference pipeline: runKore.sh:14-16; tests/koretest.py:8-43.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

REF = os.environ.get("KORE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
STANDIN = os.path.join(HERE, "..", "tests", "fixtures")


def patch_params(text, overrides):
    """Replace top-level `name = value` assignments (the same thing the
    reference's own drivers do with sed, tests/dormy2004/find_Rac.py:31,44)."""
    for k, v in overrides.items():
        pat = re.compile(r"^(%s)\s*=.*$" % re.escape(k), re.M)
        if not pat.search(text):
            raise SystemExit("parameter %s not found in params file" % k)
        # replace only the LAST active assignment (later ones win in Python)
        matches = list(pat.finditer(text))
        m = matches[-1]
        text = text[: m.start()] + "%s = %s" % (k, v) + text[m.end():]
    return text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--params", default="bin/parameters.py",
                    help="params file relative to the reference root")
    ap.add_argument("--out", required=True)
    ap.add_argument("--ncpus", type=int, default=8)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--asm", action="store_true",
                    help="also store the radial operators (operators.npz) and the physics parameters "
                         "(asm_params.json) the device-side assembly needs (kore_b200/assembly.py)")
    ap.add_argument("--profiles", default=None,
                    help="Python source appended to the scratch copy of bin/radial_profiles.py: the run's own "
                         "background profiles (the file is the user's to edit in the reference), e.g. "
                         "'def conductivity(r): return 1 + 0.5*r**2'")
    ap.add_argument("--append-params", default=None,
                    help="Python source appended to the derived parameters.py (e.g. 'OmgTau = 1', which every shipped "
                         "params file leaves commented out although compositional runs need it)")
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args()

    ov = {}
    for o in a.overrides:
        k, v = o.split("=", 1)
        ov[k] = v

    work = tempfile.mkdtemp(prefix="korecase_")
    shutil.copytree(os.path.join(REF, "bin"), os.path.join(work, "bin"))
    subprocess.check_call(["chmod", "-R", "u+w", work])
    with open(os.path.join(REF, a.params)) as f:
        ptxt = f.read()
    ptxt = patch_params(ptxt, ov)
    if a.append_params:
        ptxt += "\n# --- appended by tools/make_case.py --append-params\n" + a.append_params.replace("\\n", "\n") + "\n"
    with open(os.path.join(work, "bin", "parameters.py"), "w") as f:
        f.write(ptxt)

    if a.profiles:
        with open(os.path.join(work, "bin", "radial_profiles.py"), "a") as f:
            f.write("\n\n# --- appended by tools/make_case.py --profiles\n" + a.profiles.replace("\\n", "\n") + "\n")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.abspath(STANDIN) + os.pathsep + env.get("PYTHONPATH", "")
    env["PYTHONWARNINGS"] = "ignore"

    def run(*cmd):
        subprocess.check_call(list(cmd), cwd=work, env=env,
                              stdout=subprocess.DEVNULL)

    # parameters needed by the driver afterwards
    probe = subprocess.check_output(
        [sys.executable, "-c",
         "import sys; sys.path.insert(0,'bin'); import parameters as p, utils as u, json;"
         "print(json.dumps(dict(hydro=p.hydro,magnetic=p.magnetic,thermal=p.thermal,"
         "compositional=p.compositional,m=p.m,lmax=p.lmax,N=p.N,symm=p.symm,ricb=p.ricb,"
         "forcing=p.forcing,nev=p.nev,maxit=p.maxit,tol=p.tol,which_eigenpairs=p.which_eigenpairs,"
         "rtau=p.tau.real,itau=p.tau.imag,B0=p.B0,Ek=p.Ek,N1=u.N1,n=u.n,sizmat=u.sizmat,"
         "symmB0=u.symmB0,anelastic=p.anelastic,forcing_frequency=p.forcing_frequency)))"],
        cwd=work, env=env).decode().strip().splitlines()[-1]
    meta = json.loads(probe)

    if meta["magnetic"] == 1 or meta["anelastic"] == 1 or "heating='two zone'" in a.overrides or "heating='user defined'" in a.overrides:
        run(sys.executable, "bin/compute_profiles.py")
    run(sys.executable, "bin/submatrices.py", str(a.ncpus))
    run(sys.executable, "bin/assemble.py")

    os.makedirs(a.out, exist_ok=True)
    for fn in ("A.npz", "B.npz", "B_forced.npz"):
        src = os.path.join(work, fn)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(a.out, fn))
    if a.profiles:
        meta["profiles_appended"] = a.profiles
    with open(os.path.join(a.out, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    with open(os.path.join(a.out, "parameters.py"), "w") as f:
        f.write(ptxt)
    if a.asm:
        sys.path.insert(0, os.path.join(HERE, ".."))
        from kore_b200 import assembly as asm
        asm.save_operators_npz(os.path.join(a.out, "operators.npz"), asm.load_operators(work))
        fields = list(asm.PhysicsParams.__dataclass_fields__)
        probe = subprocess.check_output(
            [sys.executable, "-c",
             "import sys; sys.path.insert(0,'bin'); import parameters as p, utils as u, json;"
             "d={k:getattr(p,k) for k in %r if hasattr(p,k)}; d['rcmb']=u.rcmb;"
             "import numpy as np;"
             "exec('import bc_variables as bv, radial_profiles as rap\\n"
             "lho=u.chebco_f(rap.log_density,p.N,p.ricb,u.rcmb,1e-9)\\n"
             "d[\\'lho1_icb\\']=float(np.dot(lho,bv.Ta[:,1])); d[\\'lho1_cmb\\']=float(np.dot(lho,bv.Tb[:,1]))') if p.anelastic else None;"
             "print(json.dumps({k:(v.item() if hasattr(v,'item') else v) for k,v in d.items()}))" % (fields,)],
            cwd=work, env=env).decode().strip().splitlines()[-1]
        with open(os.path.join(a.out, "asm_params.json"), "w") as f:
            json.dump(json.loads(probe), f, indent=1, sort_keys=True)
    if a.asm and os.path.exists(os.path.join(work, "radProfs.mat")):
        # the tables compute_profiles.py handed to submatrices.py (kore_b200/radial.py takes them as `radprofs`)
        import numpy as np
        import scipy.io as sio
        prof = {k: v for k, v in sio.loadmat(os.path.join(work, "radProfs.mat")).items() if k.startswith("cd_")}
        np.savez_compressed(os.path.join(a.out, "radprofs.npz"), **prof)
    if a.keep:
        print("scratch kept at", work)
    else:
        shutil.rmtree(work)
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
