"""The C-ABI library loads here (no GPU) and exports every symbol include/logmel_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from audio_classification_icbhi_b200 import _lib, build
    build.build(verbose=False)      # in-tree nvcc build (cross-compiles without a GPU)
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "logmel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from audio_classification_icbhi_b200 import _lib
    names = declared_functions()
    assert len(names) >= 13
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.EXPORTS, f"{name} has no ctypes signature in _lib.EXPORTS"
    assert sorted(_lib.EXPORTS) == names


def test_struct_layouts_match_the_header():
    from audio_classification_icbhi_b200 import _lib
    import numpy as np
    assert C.sizeof(_lib.LmAug) == 40
    assert np.dtype(_lib.AUG_DTYPE).itemsize == 40
    assert [n for n, _ in _lib.LmAug._fields_] == [n for n, _ in _lib.AUG_DTYPE]
    assert _lib.LmAug.seed.offset == 32 and _lib.LmAug.gain.offset == 8
    assert C.sizeof(_lib.LmInfo) == 40
    assert C.sizeof(_lib.LmConfig) == 48


def test_no_compute_entry_points_without_gpu(lib):
    """Error paths only: nothing here touches a device."""
    assert lib.lm_abi_version() == 1
    assert lib.lm_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.lm_strerror(-5)
    assert lib.lm_plan_create(None, 0, None) == -1
    assert lib.lm_plan_frames(None) == -1
    assert lib.lm_plan_destroy(None) == 0
    assert lib.lm_plan_launch_count(None) == 0
    import torch
    if not torch.cuda.is_available():
        from audio_classification_icbhi_b200 import LogMelPlan
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            LogMelPlan()


def test_library_contains_sm100a_blackwell_code():
    """The shipped .so holds sm_100a SASS with packed fp32x2 math, TMA bulk copies and TF32 MMA."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "audio_classification_icbhi_b200", "liblogmel_b200.so")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UBLKCP", "FFMA2", "FADD2", "HMMA.1688.F32.TF32", "SHFL.IDX"):
        assert mnemonic in sass, mnemonic
