"""pytest configuration: the `gpu` marker, checker builds, shared fixtures."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_DIR = os.path.join(ROOT, "tests", "cuda_emu")
EMU_LIB = os.path.join(EMU_DIR, "libpft_b200_emu.so")
EMU_NCCL = os.path.join(EMU_DIR, "libnccl_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Build the CPU oracle (always) and the reference checker (only where /root/reference exists)."""
    from oracle import pyoracle
    pyoracle.build("oracle")
    if os.path.isdir("/root/reference"):
        pyoracle.build("ref")
    yield


@pytest.fixture(scope="session")
def emu_lib():
    """CPU emulation of the CUDA library (tests/cuda_emu): same sources, fibers instead of a GPU."""
    subprocess.run(["make", "-C", EMU_DIR], check=True, stdout=subprocess.DEVNULL)
    return EMU_LIB


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; GPU tests must run the nvcc-built sm_100a code, nothing else."""
    if os.environ.get("PFT_TEST_FORCE_EMU") == "1":
        # debugging aid for THIS container only (no GPU): dry-run the GPU test bodies on the emulation
        subprocess.run(["make", "-C", EMU_DIR], check=True, stdout=subprocess.DEVNULL)
        return EMU_LIB
    from porousfreezethaw_b200 import _lib
    _lib.load()          # raises loudly if the extension is missing
    return _lib.LIB_PATH
