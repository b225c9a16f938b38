"""CPU-side checks of the C-ABI boundary: the in-tree library loads and exports every symbol include/ofb_b200.h
declares, and the ctypes mirrors match the header's struct layouts. No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "ofb_b200.h")).read()


def test_library_exports_every_declared_symbol():
    import ofb_b200  # noqa: F401
    from ofb_b200 import _lib
    lib = _lib.lib()
    declared = set(re.findall(r"^int (ofb_[a-z0-9_]+)\(", _header(), flags=re.M))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ofb_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.ofb_version() >= 1


def _struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), _header(), flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.replace("\n", " ").split(";"):
        decl = re.sub(r"\b(const|void|float|int32_t|int64_t|uint8_t)\b|\*", " ", decl)
        out += [x.strip() for x in decl.split(",") if x.strip()]
    return out


def test_struct_mirrors_match_header():
    from ofb_b200 import _lib
    assert _struct_fields("ofb_gemm_args") == [f[0] for f in _lib.GemmArgs._fields_]
    assert _struct_fields("ofb_bimask_module") == [f[0] for f in _lib.BimaskModule._fields_]
    assert C.sizeof(_lib.BimaskModule) == 64


def test_missing_library_fails_loudly(monkeypatch):
    from ofb_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libofb_b200.so")
    with pytest.raises(_lib.OfbError):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch
    from ofb_b200 import _lib, ops
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(_lib.OfbError):
        ops.gemm(ops.EPI_STORE, a, a, M=128, N=128, K=64, out0=a)
