"""The C-ABI library builds for sm_100a, loads, and exports every symbol that
include/mvs_ncc.h declares.  No compute calls: there is no GPU on the CPU tier."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    from mvs_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mvs_ncc.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mvs_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(built_lib, name)
    assert built_lib.mvs_abi_version() == _lib.ABI_VERSION


def test_library_has_sm100a_code(built_lib):
    import shutil
    import subprocess
    from mvs_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mvs_b200
    with pytest.raises(mvs_b200.MvsError, match="no CPU fallback"):
        mvs_b200.MvsContext(np.zeros((2, 16, 16, 3), np.uint8), np.tile(np.eye(3), (2, 1, 1)),
                            np.tile(np.eye(3), (2, 1, 1)), np.zeros((2, 3)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "simple-implementation-of-structure-from-motion-and-multi-view-stereo-by-python_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_mode_b_gathers_take_uniform_texture_handles(built_lib):
    """The Mode B kernels rely on the texture handle being provably warp-uniform: otherwise ptxas wraps
    every TLD4 in a per-lane "waterfall" loop (BRA.U.ANY), which serialises the gathers of a batch of
    views (measured 1.4x slower).  The proof is fragile (an IEEE fp32 division before the gathers breaks
    it), so the SASS is checked at build time."""
    import shutil
    import subprocess
    from mvs_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    kernels = {}
    name = None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            kernels[name] = [0, 0]
        elif name and "TLD4" in line:
            kernels[name][0] += 1
        elif name and "BRA.U.ANY" in line:
            kernels[name][1] += 1
    pmvs = {k: v for k, v in kernels.items() if "ncc_score_pmvs" in k and "Lb0" in k}
    assert pmvs, "no Mode B kernels found in the library"
    for k, (tld4, waterfall) in pmvs.items():
        assert tld4 > 0, k
        assert waterfall <= 2, (k, tld4, waterfall)      # only the reference view's own handle may need one
