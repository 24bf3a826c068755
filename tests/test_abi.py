"""not-gpu: libvgs_b200.so loads, exports every symbol include/vgs_b200.h declares, and refuses to
run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "vgs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vgs_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported(built_lib):
    from vgs_svgs_segmentation_b200 import capi
    L = C.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/vgs_b200.h but not exported"
    assert sorted(capi.EXPORTED) == names


def test_no_cpu_fallback(built_lib):
    import torch
    from vgs_svgs_segmentation_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.VgsError) as e:
        capi.Handle()
    assert e.value.status == 5 and "no CPU path" in str(e.value)
    with pytest.raises(capi.VgsError) as e:      # the pooled lifecycle of the drop-in classes fails the same way
        capi.Handle(pooled=True)
    assert e.value.status == 5 and "no CPU path" in str(e.value)
    capi.load().vgs_pool_trim()                  # nothing parked: a no-op


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing in the product package may reference it"""
    pkg = os.path.join(ROOT, "vgs_svgs_segmentation_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "vgs_oracle" not in txt, f
