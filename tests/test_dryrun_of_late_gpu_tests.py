"""The GPU tests written after the round's GPU budget was spent (tests/test_zz_assembly_extensions.py) cannot be
run here; their Python side can: tools/dryrun_gpu_tests.py calls each of them with `lib.Solver` replaced by CPU
stand-ins (assembly programs through the NumPy model of the kernel, solves through the oracle).  In a subprocess,
because the stand-in is patched into kore_b200.lib."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_late_gpu_tests_pass_on_cpu_stand_ins():
    r = subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "dryrun_gpu_tests.py")], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.endswith("0 failed") and int(last.split()[0]) >= 34, last
