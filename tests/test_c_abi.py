"""The C ABI driven from plain C (tests/c_abi/abi_check.c): compiled with gcc against include/clr_b200.h and linked to
libclr_b200.so -- the view a cgo / ccall / JNI binding has.  Without a device the program must report the loud
clr_create failure (exit code 3: no CPU fallback); on a B200 it runs its structural checks (exit code 0)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "closedloopreachability.jl_b200")
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    if not os.path.exists(os.path.join(PKG, "libclr_b200.so")):
        pytest.skip("libclr_b200.so not built")
    exe = str(tmp_path / "abi_check")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", PKG, "-lclr_b200", "-lm",
                           "-o", exe])
    return exe


def _run(exe):
    env = dict(os.environ, LD_LIBRARY_PATH=PKG + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)


def test_c_program_links_and_fails_loudly_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu test")
    r = _run(_build(tmp_path))
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_program_on_the_device(tmp_path):
    r = _run(_build(tmp_path))
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi_check ok" in r.stdout
