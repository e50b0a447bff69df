"""A plain C program written against include/coupe.h the way coupe-ffi users
write one (cf. coupe-ffi/examples/rcb.c) compiles and links against the CUDA
library unchanged; on a GPU box it runs and prints the reference's answers."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "rcb_host.c")


def build_example(tmp_path, src=SRC, name="rcb_host"):
    from coupe_b200 import _lib

    _lib.build()
    exe = str(tmp_path / name)
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           src, "-o", exe, "-L", libdir, "-lcoupe_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_c_host_of_the_tools_header(tmp_path):
    """include/coupe_b200_tools.h is plain C; its host-only entry points (spec parser, file codecs) need no GPU."""
    exe = build_example(tmp_path, os.path.join(ROOT, "tests", "c", "tools_host.c"), "tools_host")
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stderr)
    assert out.stdout.strip() == "ok 10 0.001"


def test_c_host_compiles_and_links(tmp_path):
    assert os.path.exists(build_example(tmp_path))


@pytest.mark.gpu
def test_c_host_runs(tmp_path):
    out = subprocess.run([build_example(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split("\n")[:3] == ["0 0 1 1", "0 1 2 3", "0 1 2 3"]
