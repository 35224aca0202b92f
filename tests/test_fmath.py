"""The short FP64 exp / log / tanh sequences of the kernels (fortnet_b200/csrc/fmath.cuh compiles as plain C++ too) against
libm: tests/cpp/fmath_check.cpp prints its worst relative errors and returns non-zero above the documented bounds
(exp / log 1e-14, tanh 5e-13, the expm1-form tanh of the DMMA kernels 1e-14, table log 4e-16 absolute)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fmath_against_libm(tmp_path):
    exe = str(tmp_path / "fmath_check")
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-I", os.path.join(ROOT, "fortnet_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "fmath_check.cpp"), "-o", exe, "-lm"], env=env)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("FMATH_OK"), out.stdout
