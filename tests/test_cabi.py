"""CPU-only checks of the drop-in boundary: the library builds/loads, exports every symbol that
include/tardis_b200.h declares, and the ctypes structs match the header field for field."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tardis_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tb200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tardis_b200 import capi

    lib = capi.load()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tardis_b200.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == names
    assert b"sm_100a" in lib.tb200_version()


def _header_struct_fields(name):
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct \{([^{}]*)\} " + name + ";", src).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"[A-Za-z_0-9]+", part)[-1])
    return fields


@pytest.mark.parametrize("cname, pyname", [("tb200_model", "Model"), ("tb200_config", "Config"), ("tb200_packets", "Packets"),
                                           ("tb200_counters", "Counters"), ("tb200_outputs", "Outputs"),
                                           ("tb200_packet_source", "PacketSource"), ("tb200_estimator_layout", "EstimatorLayout"),
                                           ("tb200_radfield_params", "RadfieldParams"), ("tb200_atomic_data", "AtomicData"),
                                           ("tb200_plasma_state", "PlasmaState"), ("tb200_source_function_params", "SourceFunctionParams"),
                                           ("tb200_formal_integral_params", "FormalIntegralParams")])
def test_ctypes_structs_match_header(cname, pyname):
    from tardis_b200 import capi

    assert [f[0] for f in getattr(capi, pyname)._fields_] == _header_struct_fields(cname)


def test_event_struct_size():
    from tardis_b200.engine import EVENT_DTYPE

    assert EVENT_DTYPE.itemsize == 14 * 8
    assert list(EVENT_DTYPE.names) == _header_struct_fields("tb200_event")


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tardis_b200.engine import Engine, EngineError

    with pytest.raises(EngineError, match="no CPU path|no CUDA device|CUDA"):
        Engine(0)
