"""CPU tests of the drop-in boundary: libb200feat.so loads without a GPU, exports every symbol
include/b2f.h declares, refuses to run without a device (no CPU fallback), and the host-side
mirror keeps the R wrapper's argument conventions."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "b2f.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(b2f_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from image_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/b2f.h but not exported: %s" % missing


def test_no_cpu_fallback_when_no_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from image_b200 import _lib, image_harris
    with pytest.raises(_lib.B2FError) as e:
        image_harris(np.zeros((64, 64)))
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_match_arg_and_enum_quirks():
    from image_b200 import harris as H
    assert H._match_arg(H.GAUSSIAN, H.GAUSSIAN, "gaussian") == 0          # default 'fast Gaussian' -> 0 = STD in C++
    assert H._match_arg("precise Gaussian", H.GAUSSIAN, "gaussian") == 1  # -> FAST (SII) in C++ (R/pkg.R:61,70)
    assert H._match_arg(H.PRECISION, H.PRECISION, "precision") == 0       # 'quadratic approximation' -> NO_INTERPOLATION
    assert H._match_arg("no subpixel", H.PRECISION, "precision") == 2     # -> QUARTIC in C++
    assert H._match_arg("Sobel", H.GRADIENT, "gradient") == 1             # partial matching like match.arg
    with pytest.raises(ValueError):
        H._match_arg("nonsense", H.MEASURE, "measure")


def test_product_never_imports_oracle():
    """The package must not reach into oracle/ (SPEC: oracle is test infrastructure only)."""
    pkg = os.path.join(ROOT, "image_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("SURVEY", ""), f
