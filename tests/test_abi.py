"""The C ABI: every symbol include/pft_b200.h declares is exported, the library refuses to run
without a GPU (no CPU fallback), and the solver-module structs mirror the reference header."""
import ctypes as C
import os
import re
import subprocess

import pytest

from porousfreezethaw_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:pft|RK_MPI_SA)_\w+)\s*\(", src)) - {"pft_service_fn"})


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "porousfreezethaw_b200", "csrc"), "all"], check=True, stdout=subprocess.DEVNULL)


def test_pft_library_exports_every_declared_symbol():
    _build()
    lib = _lib.load()
    names = _declared("pft_b200.h")
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_dropin_library_exports_solver_module_and_plugin_symbols():
    _build()
    lib = C.CDLL(_lib.RK_LIB_PATH)
    for n in _declared("RK_MPI_SAsolver_b200.h") + _declared("pft_equation_b200.h"):
        assert hasattr(lib, n), n


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="box has a GPU")
def test_no_cpu_fallback():
    """without a device pft_create fails loudly with PFT_ENODEVICE and says there is no CPU path"""
    _build()
    lib = _lib.load()
    case = _lib.PftCase(8, 8, 8, 8, 0, 0, 1, 0, 1.0, 1.0, 1.0)
    h = C.c_void_p()
    rc = lib.pft_create(C.byref(case), 0, 0, C.byref(h))
    assert rc == -19 and not h.value
    assert b"no CPU path" in lib.pft_last_error()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")
def test_struct_layout_matches_reference_header(tmp_path):
    """sizeof/offsetof of RK_MEM_DIST and RK_MPI_S_SOLUTION: our mirror header vs the reference's own"""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include HEADER
int main(void){
 printf("%zu %zu %zu %zu %zu\n", sizeof(RK_MEM_DIST), offsetof(RK_MEM_DIST,n_chunks), offsetof(RK_MEM_DIST,chunk_start), offsetof(RK_MEM_DIST,chunk_size), offsetof(RK_MEM_DIST,chunk_eps_mult));
 printf("%zu", sizeof(RK_MPI_S_SOLUTION));
#define O(m) printf(" %zu", offsetof(RK_MPI_S_SOLUTION,m));
 O(n) O(t) O(x) O(meta_f) O(h) O(h_min) O(delta) O(delta_mode) O(DDLBF_Rearrange) O(Service_Callback) O(steps) O(steps_total)
 printf("\n%d %d %d %d %d %d\n", RKA_CMD_h_TOO_SMALL, RKA_CMD_NAN, RKA_CMD_UPDATE, RKA_CMD_FINISHED, RKA_CMD_NEXTFINISH, RKA_CMD_BREAK);
 return 0; }'''
    (tmp_path / "p.c").write_text(prog)
    outs = []
    for hdr, inc in (('"RK_MPI_SAsolver_b200.h"', [os.path.join(ROOT, "include")]),
                     ('"RK_MPI_SAsolver.h"', ["/root/reference/include"])):
        exe = tmp_path / ("a" + str(len(outs)))
        cmd = ["gcc", "-std=c99", "-w", f"-DHEADER={hdr}", "-I" + os.path.join(ROOT, "oracle", "shim")] + ["-I" + i for i in inc] + \
              [str(tmp_path / "p.c"), "-o", str(exe)]
        subprocess.run(cmd, check=True)
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1]
