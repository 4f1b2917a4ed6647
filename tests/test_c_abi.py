"""The drop-in boundary is a C ABI: compile a plain-C program against include/dmel.h with gcc (C99, -Wall -Werror),
link it to libdmel_b200.so and run it.  No GPU needed: it exercises version, sizing, planning and argument validation."""
import os
import shutil
import subprocess

import pytest

import dmel_b200 as dm
from conftest import ROOT


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not installed")
def test_plain_c_program_links_and_calls_the_library(tmp_path):
    lib_path = dm.build()
    lib_dir = os.path.dirname(lib_path)
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", lib_dir, "-l:libdmel_b200.so",
           f"-Wl,-rpath,{lib_dir}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "abi v" in run.stdout
