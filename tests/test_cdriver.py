"""A plain-C host (tests/cdriver/julia_ccall_replay.c) compiled against include/gmrf_b200.h alone: replays the ccall sequence
of julia/B200Backend.jl argument for argument with 1-based arrays and checks every answer against a dense Cholesky it
computes itself. CPU: it must compile and link against the shared library (every symbol the glue binds exists with the
declared signature). GPU: it must run to completion."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cdriver", "julia_ccall_replay.c")
LIBDIR = os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200", "lib")


def _build(tmp_path):
    import __graft_entry__ as g
    g.build(quiet=True)
    exe = str(tmp_path / "julia_ccall_replay")
    r = subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-o", exe, SRC, "-I", os.path.join(ROOT, "include"), "-L", LIBDIR,
                        "-lgmrf_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_compiles_and_links_against_the_header(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_c_host_replays_the_julia_ccall_sequence(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
