"""The CUDA sources compiled for the CPU emulation with AddressSanitizer: global-memory accesses of the kernels (field
arrays with their ghost planes, completion words, control blocks, sync block, staging buffers) and the host side of
pft_b200.cu run under redzones.  (compute-sanitizer is closed on the GPU pool; this is the memory check there is.)"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "cuda_emu")


def test_device_loop_under_address_sanitizer():
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan is not installed")
    r = subprocess.run(["make", "-C", EMU, "libpft_b200_emu_asan.so"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "asan_driver.py"), os.path.join(EMU, "libpft_b200_emu_asan.so")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "ASAN-DRIVER-OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in r.stderr
