import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_device_count() -> int:
    """Number of CUDA devices as the product library sees them (tb200_create is the authority: it fails with
    TB200_ERR_CUDA when there is none).  Probed once per session."""
    try:
        import ctypes as C

        from tardis_b200 import capi

        lib = capi.load()
        h = C.c_void_p()
        if lib.tb200_create(0, C.byref(h)) != 0:
            return 0
        lib.tb200_destroy(h)
        try:
            import torch

            return max(1, torch.cuda.device_count())
        except Exception:
            return 1
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (plain `pytest` on a CPU box stays green); with `-m gpu`
    on a box that has no device they still fail loudly -- the product has no CPU path to hide behind."""
    if not any("gpu" in item.keywords for item in items):
        return
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (tardis_b200 has no CPU path); run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle

    cpu_oracle.build()
    return cpu_oracle
