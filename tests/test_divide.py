"""gca_div_const_f32 (the MUFU-free division the observation writer uses) equals IEEE division."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constant_division_is_exact():
    src = os.path.join(ROOT, "tests", "test_divide.c")
    inc = os.path.join(ROOT, "gym-guidance-collision-avoidance-single_b200", "csrc")
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-I" + inc, src, "-o", exe, "-lm"])
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    lines = out.stdout.strip().splitlines()
    assert "bad_one_step=0 is_exact1=1" in lines[0] and "bad_one_step=0 is_exact1=1" in lines[1]   # default divisors
    assert lines[-1].startswith("f64") and lines[-1].endswith("bad=0")      # gca_div_const_f64 == IEEE division
