"""csrc/pft_pow.h: the twin of the host C library's pow() that lets the step-size controller run on the
device.  Here (no GPU): the HOST instance of the twin, compiled from the same header, against libm --
bit for bit.  The DEVICE instance is held against libm by tests/test_gpu_parity.py
(test_device_pow_twin_matches_host_libm) and, implicitly, by every adaptive parity test."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "pft_pow.h"
#include <omp.h>
int main(int argc, char **argv) {
	pft::PowTables T; char err[256];
	if (pft::pft_pow_load_tables(&T, err, sizeof err)) { printf("LOAD FAILED: %s\n", err); return 2; }
	long n = atol(argv[1]), bad = 0;
	#pragma omp parallel for reduction(+:bad)
	for (int t = 0; t < 64; t++) bad += pft::pft_pow_selftest(T, 0.2, n / 64, 0x9e3779b97f4a7c15ull * (t + 1));
	printf("operands %ld mismatches %ld\n", (n / 64) * 64, bad);
	return bad != 0;
}
'''


@pytest.fixture(scope="module")
def powtest(tmp_path_factory):
    d = tmp_path_factory.mktemp("powtwin")
    (d / "powtest.cpp").write_text(SRC)
    exe = str(d / "powtest")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fopenmp", f"-I{ROOT}/porousfreezethaw_b200/csrc", "-o", exe,
                    str(d / "powtest.cpp"), "-ldl", "-lm"], check=True)
    return exe


def test_host_twin_matches_libm(powtest):
    """1e8 operands by default (about 5 s on 8 threads); PFT_POW_TEST_N=1000000000 for the 1e9 run recorded in
    profiles/r02_pow_twin.txt"""
    n = int(os.environ.get("PFT_POW_TEST_N", "100000000"))
    r = subprocess.run([powtest, str(n)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout
