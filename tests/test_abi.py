"""The C-ABI shared library loads and exports exactly what include/gusto_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from util import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "gusto_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gusto_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    fns = header_functions()
    for need in ("gusto_create", "gusto_destroy", "gusto_set_problems", "gusto_set_trajectory", "gusto_linearize",
                 "gusto_solve_subproblem", "gusto_evaluate", "gusto_accept", "gusto_get_trajectory", "gusto_last_error"):
        assert need in fns


def test_library_exports_every_declared_symbol(host):
    lib = host.load_library()
    for fn in header_functions():
        assert hasattr(lib, fn), f"{fn} declared in include/gusto_b200.h but not exported"
    assert lib.gusto_version() >= 100


def test_library_is_a_native_sm100a_binary():
    lib = os.path.join(ROOT, "gusto.jl_b200", "libgusto_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z16linearize_kernelILi2EEvPKN5gusto9BatchDescENS0_9BatchPtrsE", lib],
                          capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "linearize kernel must stage states with TMA bulk copies"


def test_config_struct_layout_matches_the_header(host):
    """ctypes mirror and C struct agree on size (caught by a tiny C program compiled against the header)."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "gusto_b200.h"
int main(void) { printf("%zu %zu %zu %zu\n", sizeof(gusto_config), offsetof(gusto_config, robot_params),
                        offsetof(gusto_config, goal_type), offsetof(gusto_config, ipm_tol)); return 0; }
'''
    d = os.path.join(ROOT, "tests", "hostsim", "_build")
    os.makedirs(d, exist_ok=True)
    src, exe = os.path.join(d, "abi_probe.c"), os.path.join(d, "abi_probe")
    open(src, "w").write(prog)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
    size, o_rp, o_gt, o_tol = map(int, subprocess.check_output([exe]).split())
    C = host.GustoConfig
    assert ctypes.sizeof(C) == size and C.robot_params.offset == o_rp and C.goal_type.offset == o_gt and C.ipm_tol.offset == o_tol


def test_no_gpu_means_loud_failure_not_cpu_fallback(host, pkg):
    """In the GPU-less container gusto_create must fail with GUSTO_E_NODEVICE; on a GPU box it must succeed."""
    import numpy as np
    bp = pkg.problems.config_dubins(B=1, N=30)
    try:
        e = host.Engine(bp)
    except host.GustoError as ex:
        assert "no CUDA device" in str(ex)
    else:
        e.close()
