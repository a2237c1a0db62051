"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/tokzig_b200.h declares."""
import ctypes as C
import os
import re

import pytest

import tokzig_b200 as tz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tokzig_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tkz[hm]?_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    L = C.CDLL(tz.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/tokzig_b200.h but not exported"
    assert sorted(tz.EXPORTED_SYMBOLS) == decl


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", tz.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out, out


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly (TKZ_ERR_CUDA), never compute on the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(tz.TokzigError) as e:
        tz.Context(0)
    assert e.value.code == tz.ERR_CUDA
    t = tz.Tokenizer.from_json('{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "a": 1}}}', device=None)
    with pytest.raises(tz.TokzigError) as e:
        t.encode("a")
    assert e.value.code == tz.ERR_CUDA


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under tokenizer-zig_b200/ may import, link or open it."""
    pkg = os.path.join(ROOT, "tokenizer-zig_b200")
    for dp, _dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".zig")):
                s = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in s.lower(), f
