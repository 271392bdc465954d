"""The C-ABI libraries load and export every symbol their headers declare
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmr(?:gpu|c|_b200)_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product():
    path = os.path.join(ROOT, "tmr_b200", "lib", "libtmr_b200.so")
    if not os.path.exists(path):
        import __graft_entry__

        __graft_entry__.build()
    return ctypes.CDLL(path)


@pytest.mark.parametrize("header", ["tmrgpu.h", "tmr_capi.h", "tmr_b200_ext.h"])
def test_product_exports_every_declared_symbol(product, header):
    names = declared_functions(header)
    assert len(names) > (20 if header != "tmr_b200_ext.h" else 3)
    missing = [n for n in names if not hasattr(product, n)]
    assert not missing, missing


def test_product_is_the_cuda_build(product):
    product.tmrgpu_build_kind.restype = ctypes.c_char_p
    product.tmrc_backend.restype = ctypes.c_char_p
    assert product.tmrgpu_build_kind() == b"cuda-sm_100a"
    assert product.tmrc_backend() == b"b200-cuda"


def test_oracle_exports_the_same_class_binding(ref_lib):
    for n in declared_functions("tmr_capi.h"):
        assert hasattr(ref_lib, n), n
    assert ref_lib.tmrc_backend() == b"reference-cpu"


def test_python_binding_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import tmr_b200

    with pytest.raises(RuntimeError):
        tmr_b200.require_gpu()


def test_octant_record_layout(product):
    """24-byte TMROctant layout (reference src/TMROctant.h:49-53)."""
    from tmr_b200 import _capi

    assert _capi.OCT_DTYPE.itemsize == 24
    assert [_capi.OCT_DTYPE.fields[n][1] for n in
            ("block", "x", "y", "z", "tag", "level", "info")] == [0, 4, 8, 12, 16, 20, 22]
