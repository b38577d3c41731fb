"""The reference's example programs, rewritten against the header-only C++ front (examples/*.cpp), compiled with g++ and run on
the GPU: this is the C++-level drop-in check (the reference's only executable checks are these examples)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    import lbfgspp_b200 as lb
    lb.build_all()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples")], check=True)


def run(name, *args):
    return subprocess.run([os.path.join(BIN, name)] + list(args), capture_output=True, text=True, timeout=600)


def test_rosenbrock_host_functor_float():
    """example-rosenbrock.cpp: float, host functor through the compatibility mode, dense B and H (B*H = I)."""
    r = run("rosenbrock_host_functor")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "iterations" in r.stdout and "max |B*H - I|" in r.stdout


def test_quadratic_free_function():
    r = run("quadratic_free_function")
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("2 iterations")


def test_rosenbrock_box_device():
    """example-rosenbrock-box.cpp: 13 iterations, f = 360.2835855511515 with the reference headers."""
    r = run("rosenbrock_box_device")
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("13 iterations")


def test_line_search_comparison_self_check():
    """example-rosenbrock-comparison.cpp / -bracketing.cpp: all four line searches reach |x - 1|_inf <= 1e-4."""
    r = run("line_search_comparison", "12")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAILED" not in r.stdout and r.stdout.count("LineSearchMoreThuente") == 12
