"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/bbdm_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

import conftest

HDR = os.path.join(conftest.REPO, "include", "bbdm_b200.h")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bbdm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from bbdm_b200 import build, cabi
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    decl = declared_symbols()
    assert len(decl) >= 25
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes binding covers the same set
    assert sorted(cabi.SYMBOLS) == decl, (set(decl) ^ set(cabi.SYMBOLS))
    l2 = cabi.load()
    assert l2.bbdm_abi_version() == cabi.ABI_VERSION
    assert isinstance(l2.bbdm_last_error(), bytes)


def test_struct_layouts_match_header_field_order():
    """ctypes.Structure field names/order of the three argument structs == the header's."""
    from bbdm_b200 import cabi
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)

    def fields(name):
        end = src.index("} " + name + ";")
        body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                m = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())
                if m:
                    out.append(m[0])
        return out

    assert fields("BbdmPSampleCoef") == [f[0] for f in cabi.PSampleCoef._fields_]
    assert fields("BbdmPrepArgs") == [f[0] for f in cabi.PrepArgs._fields_]
    assert fields("BbdmConvArgs") == [f[0] for f in cabi.ConvArgs._fields_]
    assert fields("BbdmWinoInputArgs") == [f[0] for f in cabi.WinoInputArgs._fields_]
    assert fields("BbdmWinoOutputArgs") == [f[0] for f in cabi.WinoOutputArgs._fields_]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from bbdm_b200 import cabi
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(cabi.BbdmError, match="no PyTorch/CPU fallback"):
        cabi.load()
