"""CPU: bench.py's driver contract on the reference arm (the only arm that runs without a GPU):
exactly ONE JSON line on stdout -- library chatter written to file descriptor 1 behind Python's
back (NCCL's version banner) is diverted to stderr -- with the keys the driver reads."""
import json
import os
import subprocess
import sys
import textwrap

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    # (a small chain: the arm always runs one WHOLE step of the CPU path on the size it is given)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20",
                        "--warmup", "5", "--P", "12", "--b", "80"], capture_output=True, text=True, cwd=ROOT,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "eigenpairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["steps_requested"] == 20 and d["warmup_requested"] == 5 and d["steps_measured"] == 1
    # the claimed time is the time spent: ms_per_step x steps fits the run
    assert d["ms_per_step"] * d["steps"] / 1e3 <= d["wall_s_including_input_generation"]
    assert d["cpu_baseline"]["max_residual"] < 1e-10 and d["config"]["n"] == 12 * 80
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_fd_level_chatter_stays_off_stdout():
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import bench
        sys.stdout.flush(); bench._Out.real = os.dup(1); os.dup2(2, 1)
        os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')
        bench.emit({'metric': 'x', 'value': 1})
        os.write(1, b'late chatter\\n')
    """ % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": "x", "value": 1}\n'
    assert "NCCL version" in r.stderr and "late chatter" in r.stderr
