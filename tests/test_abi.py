"""The C-ABI shared library loads and exports every symbol include/gillb200.h declares (no compute calls)."""
import ctypes
import os
import re

from gill_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_reports_version():
    L = _lib.lib()
    assert L.gillb200_version() == 100
    assert L.gillb200_last_error() is not None


def test_every_declared_symbol_is_exported():
    names = _lib.exported_symbols()
    assert len(names) >= 20, names
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/gillb200.h but not exported: {missing}"


def test_struct_mirrors_match_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "gillb200.h")).read()
    for struct, cls in (("gillb200_gemm_args", _lib.GemmArgs), ("gillb200_attn_args", _lib.AttnArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                fields.append(re.findall(r"[A-Za-z_0-9]+", part)[-1])
        assert fields == [f[0] for f in cls._fields_], (struct, fields)


def test_bad_arguments_fail_loudly_without_a_gpu():
    # argument validation happens before any CUDA call
    g = _lib.GemmArgs()
    rc = _lib.lib().gillb200_gemm(ctypes.byref(g), None)
    assert rc < 0
    assert b"shape" in _lib.lib().gillb200_last_error() or b"null" in _lib.lib().gillb200_last_error()


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: NOTHING under gill_b200/ may reference it (the synthetic-weight builders and the
    smoke check live in harness/, outside the product package)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gill_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("ORACLE", ""), f"{fn} references the oracle"
