"""The C-ABI shared library builds, loads, and exports exactly what include/comet_b200.h declares.
No compute call is made here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import importlib.util

    spec = importlib.util.spec_from_file_location("_comet_b200_build",
                                                  os.path.join(ROOT, "comet_pose_estimation_b200", "build.py"))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)  # by path: the package import needs the library this fixture builds
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "comet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(comet_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(built)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/comet_b200.h but not exported"


def test_binding_table_matches_header(built):
    from comet_pose_estimation_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib.comet_version() >= 100


def test_pyramid_geometry_helpers(built):
    from comet_pose_estimation_b200 import _lib

    lib = _lib.lib
    # coarse: 64 -> 32 -> 16 -> 8 -> 4 ; fine: 31 -> 15 -> 7 (floor)
    assert lib.comet_pyramid_elems(16, 128, 64, 64, 5) == 16 * 128 * (32 * 32 + 16 * 16 + 8 * 8 + 4 * 4)
    assert lib.comet_pyramid_elems(2, 32, 31, 31, 3) == 2 * 32 * (15 * 15 + 7 * 7)
    assert lib.comet_pyramid_offset(2, 32, 31, 31, 1) == 0
    assert lib.comet_pyramid_offset(2, 32, 31, 31, 2) == 2 * 32 * 15 * 15
    assert lib.comet_pyramid_elems(1, 1, 8, 8, 1) == 0


def test_invalid_arguments_are_reported_without_a_gpu(built):
    """Argument validation happens before any CUDA call and maps to AssertionError like the reference's asserts."""
    from comet_pose_estimation_b200 import _lib

    rc = _lib.lib.comet_pyramid_f32(None, None, 1, 4, 8, 8, 9, None)
    assert rc == _lib.ERR_INVALID and "num_levels" in _lib.last_error()
    with pytest.raises(AssertionError):
        _lib.check(rc)
    rc = _lib.lib.comet_corr_lookup_f32(None, None, None, 0, 0, 0, 0, None, 0, 0, 0, None, 0, 0, 0,
                                        1, 1, 1, 4, 8, 8, 2, 9, 0, 0, 0, None)
    assert rc == _lib.ERR_INVALID and "radius" in _lib.last_error()
    rc = _lib.lib.comet_sincos2d_f32(None, 10, 4, 4, None)
    assert rc == _lib.ERR_INVALID


def test_no_cpu_fallback(built):
    import torch

    import comet_pose_estimation_b200 as cb

    with pytest.raises(cb._lib.CometB200Error):
        cb.CorrBlock(torch.zeros(1, 1, 4, 8, 8))
    with pytest.raises(cb._lib.CometB200Error):
        cb.bilinear_sampler(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 1, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "comet_pose_estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{f} imports the oracle"


def test_options_are_an_explicit_abi_not_environment_variables(built):
    """A/B switches go through comet_set_option; the launch path reads no environment variable, and the production
    header exports no debug entry point (trace hooks exist only in -DCOMET_TC_TRACE builds)."""
    from comet_pose_estimation_b200 import _lib

    assert _lib.lib.comet_get_option(_lib.OPT_TENSOR_PATH) == 1 and _lib.lib.comet_get_option(_lib.OPT_TMA_LOOKUP) == 1
    assert _lib.lib.comet_get_option(_lib.OPT_GEMM_BK32) == 0          # measured slower: off unless asked for
    assert _lib.set_option(_lib.OPT_TMA_LOOKUP, False) is True
    assert _lib.lib.comet_get_option(_lib.OPT_TMA_LOOKUP) == 0
    _lib.set_option(_lib.OPT_TMA_LOOKUP, True)
    assert _lib.lib.comet_set_option(99, 1) == _lib.ERR_INVALID
    assert "comet_tc_debug_stamps" not in _declared()
    assert not hasattr(ctypes.CDLL(built), "comet_tc_debug_stamps")
    pkg = os.path.join(ROOT, "comet_pose_estimation_b200")
    for f in os.listdir(os.path.join(pkg, "csrc")):
        txt = open(os.path.join(pkg, "csrc", f)).read()
        txt = re.sub(r"#ifdef COMET_TC_TRACE.*?#endif", "", txt, flags=re.S)
        assert "getenv" not in txt, f"{f} reads the environment outside a trace build"
    for f in os.listdir(pkg):
        if f.endswith(".py") and f not in ("build.py", "launch.py"):   # launch.py reads torchrun's RANK / WORLD_SIZE
            assert "os.environ" not in open(os.path.join(pkg, f)).read(), f
