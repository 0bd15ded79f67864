import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import ua3reo_loader
    mod = ua3reo_loader.load()
    # DEVELOPMENT AID: UA3REO_DEV_EMU=1 points the tests at the host-emulation build of the CUDA sources
    # (tools/emu) so that `-m gpu` test logic can be debugged in the GPU-less build container.
    emu = os.path.join(ROOT, "tools", "emu", "_build", "libua3reo_emu.so")
    if os.environ.get("UA3REO_DEV_EMU") == "1" and os.path.exists(emu):
        mod.LIB_PATH = emu
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


REFERENCE = "/root/reference"


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "FPGA"))
