"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports exactly the
symbols include/*.h declare; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT

import __graft_entry__ as G


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gps_[a-z0-9_]+)\s*\(", src)))


def test_build_and_exports():
    lib_path = G.build()
    lib = C.CDLL(lib_path)
    names = _declared("gpscore.h") + _declared("gpscore_debug.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libgpscore.so does not export " + n


def test_ctypes_table_matches_header():
    from gpscore_b200 import lib as L

    assert sorted(L.SIGNATURES) == _declared("gpscore.h")
    assert sorted(L.DEBUG_SIGNATURES) == _declared("gpscore_debug.h")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from gpscore_b200 import api, lib as L

    with pytest.raises(L.GpsError) as e:
        api.Context(0)
    assert e.value.code == L.GPS_ENODEVICE


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "scoring-rules-for-gaussian-process-regression-a-new-approach-to-inference_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn
