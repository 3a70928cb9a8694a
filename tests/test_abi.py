"""CPU tests of the drop-in boundary: the library builds, loads and exports every declared symbol;
codes agree between the oracle and the ABI; without a GPU the engine refuses to exist."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "methylasso_b200.h")


@pytest.fixture(scope="module")
def lib():
    from methylasso_b200 import build, _native
    build.build()
    return _native.load()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^ML_API[^;(]*?\b(ml_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from methylasso_b200 import _native
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_native.SIGNATURES) == names, "ctypes table out of sync with the header"
    assert lib.ml_abi_version() == 1


def test_error_codes_in_sync_with_oracle_header(lib):
    def enum(path, prefix):
        src = open(path).read()
        return {k: int(v) for k, v in re.findall(prefix + r"_(ERR_[A-Z0-9_]+|OK)\s*=\s*(\d+)", src)}
    a = enum(HEADER, "ML")
    b = enum(os.path.join(ROOT, "oracle", "methylasso_oracle.h"), "ORC")
    for k, v in b.items():
        assert a[k] == v, k
    assert lib.ml_strerror(6).decode() == "data is not sorted by dataset and position"


def test_params_struct_layout(lib):
    from methylasso_b200._native import MlParams
    from oracle.oracle import OrcParams
    assert C.sizeof(MlParams) == C.sizeof(OrcParams) == 72
    assert [f[0] for f in MlParams._fields_] == [f[0] for f in OrcParams._fields_]


def test_no_gpu_means_no_engine(lib):
    """The product path must fail loudly when there is no CUDA device (no CPU fallback)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    h = C.c_void_p()
    rc = lib.ml_engine_create(-1, C.byref(h))
    assert rc == 101 and not h.value
    assert b"no CPU path" in lib.ml_last_error()
    from methylasso_b200 import api
    with pytest.raises(api.MethyLassoError):
        api.Engine()


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "methylasso_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("test oracle", "").lower() or f in (), (dirpath, f)
