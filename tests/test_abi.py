"""The C-ABI library loads, exports every symbol include/emosaic_cuda.h declares, and fails loudly
(no CPU fallback) when no GPU is present.  CPU only — no compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "emosaic_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from emosaic_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/emosaic_cuda.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "python binding list and header disagree"


def test_abi_version_and_error_string():
    from emosaic_b200 import _lib
    lib = _lib.load()
    assert lib.emo_abi_version() == 2
    assert isinstance(lib.emo_last_error(), bytes)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import emosaic_b200
    with pytest.raises(emosaic_b200.EmosaicError) as e:
        emosaic_b200.Context(0)
    assert e.value.code == -6 and "no CPU path" in str(e.value)


def test_null_ctx_rejected():
    from emosaic_b200 import _lib
    lib = _lib.load()
    assert lib.emo_sync(None) == -1
    assert b"NULL" in lib.emo_last_error()
    assert lib.emo_match(None, None, 4, 4, None, None) == -1


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "emosaic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 on its own (what a Rust bindgen / cgo would parse)."""
    import shutil
    import subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not cc:
        pytest.skip("no C compiler")
    r = subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "emosaic_cuda.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_resize_null_and_argument_checks_need_no_gpu():
    """Argument validation of emo_resize happens before any CUDA call (NULL ctx is rejected first)."""
    from emosaic_b200 import _lib
    lib = _lib.load()
    assert lib.emo_resize(None, None, 1, 4, 4, 0, 0, 4, 4, 2, 2, None) == -1
    assert b"ctx is NULL" in lib.emo_last_error()


def test_stripe_bounds_matches_python_sharding():
    """emo_stripe_bounds (the C ABI's partition of block rows / tiles over GPUs) == sharding.stripe_bounds, and the parts tile the range."""
    import emosaic_b200 as emo
    from emosaic_b200 import sharding
    for units in (0, 1, 7, 37, 4096, 1_000_003):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                a, b = emo.stripe_bounds(units, world, r)
                assert (a, b) == sharding.stripe_bounds(units, world, r)
                assert a == prev and b >= a
                prev = b
            assert prev == units
    assert emo.stripe_bounds(10, 2, 5) == (0, 0)        # rank outside the world: empty range, no crash


def test_group_and_comm_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import emosaic_b200 as emo
    with pytest.raises(emo.EmosaicError) as e:
        emo.Group(2)
    assert e.value.code == -6
    from emosaic_b200 import _lib
    lib = _lib.load()
    assert lib.emo_group_mosaic(None, None, 4, 4, 3, 0, None, None, None) == -1
    assert lib.emo_comm_init_rank(None, None, 0, 1) == -1
    assert lib.emo_group_size(None) == 0 and lib.emo_group_ctx(None, 0) is None


def test_probe_library_is_separate():
    """Measurement code is not part of the product ABI: emo_probe_* live in tools/libemosaic_probe.so only."""
    from emosaic_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    assert not hasattr(lib, "emo_probe_int_pipe")
    from tools import probe
    if os.path.exists(probe.LIB_PATH):
        p = ctypes.CDLL(probe.LIB_PATH)
        assert hasattr(p, "emo_probe_int_pipe") and hasattr(p, "emo_probe_host_copy")
