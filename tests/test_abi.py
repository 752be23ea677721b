"""The C-ABI shared library loads and exports every symbol include/vx3b200.h declares; compute entry points fail loudly
(no CPU fallback) when no GPU is usable."""
import ctypes as C
import os
import re
import subprocess

import pytest

import gpuvoxels_b200 as gv
from gpuvoxels_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vx3b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx3b_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    names = declared_symbols()
    assert len(names) >= 29
    assert sorted(_abi.SYMBOLS) == names
    lib = _abi.load()
    for n in names:
        assert getattr(lib, n) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vx3b_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_struct_sizes_match_the_header():
    # compile a tiny C program against the header and compare sizeof with the ctypes mirrors
    src = r'''
#include <stdio.h>
#include "vx3b200.h"
int main(){printf("%zu %zu %zu %zu %zu\n", sizeof(VX3B_Token), sizeof(VX3B_VoxelMaterial), sizeof(VX3B_LinkMaterial), sizeof(VX3B_WorldDesc), sizeof(VX3B_WorldResult));return 0;}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_abi.Token), C.sizeof(_abi.VoxelMaterial), C.sizeof(_abi.LinkMaterial), C.sizeof(_abi.WorldDesc),
                     C.sizeof(_abi.WorldResult)]


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_engine_fails_loudly_without_a_gpu():
    with pytest.raises(gv.VX3Error) as e:
        gv.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    # nothing under gpuvoxels_b200/ may import, link or reference oracle/
    pkg = os.path.join(ROOT, "gpuvoxels_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "vx3o_" not in text and "libvx3oracle" not in text and "oracle_lib" not in text, f
    ldd = subprocess.run(["ldd", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
