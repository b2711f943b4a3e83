"""The C-ABI boundary (CPU only, no compute calls): the library loads, exports every symbol that
include/sb/sb_api.h declares, struct sizes agree, and the product refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

from strelka_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sb", "sb_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_all_bound_in_python():
    assert _declared_symbols() == sorted(_abi.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} missing from {_abi.library_path()}"


def test_struct_sizes_match_header():
    # sizes implied by the header (checked against the ctypes/numpy mirrors)
    assert C.sizeof(_abi.sb_settings) == 16 * 4 + 16
    assert C.sizeof(_abi.sb_device_cfg) == 16
    assert C.sizeof(_abi.sb_counters) == 14 * 8 + 16 + 8 + 8 * 8 + 8 * 8 + 16 + 16
    # ... and the sizeof() values the compiled library itself reports (load_library refuses a mismatch)
    lib = _abi.load_library()
    assert lib.sb_abi_version() == _abi.SB_API_VERSION
    assert lib.sb_abi_struct_size(2) == C.sizeof(_abi.sb_counters) and lib.sb_abi_struct_size(3) == C.sizeof(_abi.sb_scene_view)
    assert lib.sb_abi_struct_size(4) == 96 and lib.sb_abi_struct_size(99) == 0
    assert _abi.VERTEX_DTYPE.itemsize == 32 and _abi.LIGHT_DTYPE.itemsize == 112
    assert _abi.INSTANCE_DTYPE.itemsize == 80 and _abi.MATERIAL_DTYPE.itemsize == 96


def test_settings_default_matches_reference_app_defaults():
    lib = _abi.load_library()
    s = _abi.sb_settings()
    lib.sb_settings_default(C.byref(s))
    # src/hdRunner/main.cpp:510-542
    assert (s.depth, s.spp, s.enable_acc, s.rect_light_sampling_method, s.tonemapper_type) == (4, 1, 1, 0, 0)
    assert (s.film_iso, s.cm2_factor, s.f_stop, s.shutter_speed) == (100.0, 1.0, 4.0, 100.0)
    assert abs(s.gamma - 2.4) < 1e-6 and s.sample_stride == 1


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from strelka_b200 import RenderFactory, RenderType, SbError

    r = RenderFactory.createRender(RenderType.eCompute)
    with pytest.raises(SbError, match="no usable CUDA device|CUDA"):
        r.init()


def test_factory_returns_none_for_foreign_backends():
    from strelka_b200 import RenderFactory, RenderType

    assert RenderFactory.createRender(RenderType.eOptiX) is None  # render.cpp:17-18,25: unsupported -> nullptr
    assert RenderFactory.createRender(RenderType.eMetal) is None


def test_product_does_not_reference_oracle():
    # the oracle is test infrastructure: nothing under strelka_b200/ may import, include or link it
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "strelka_b200")):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle/|import oracle|from oracle|liboracle|pyoracle", txt):
                    # comments that merely mention the oracle by name are fine; includes/imports are not
                    if re.search(r"#include\s+\"[^\"]*oracle|import oracle|from oracle|liboracle\.so|pyoracle", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
