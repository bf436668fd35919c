"""The C ABI library loads and exports every symbol include/sk_engine.h declares (no compute without a GPU)."""
import ctypes
import os
import re

from skirt9_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sk_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sk_(?:engine_\w+|abi_version|last_error))\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for n in abi.ABI_FUNCTIONS + abi.ENGINE_ONLY_FUNCTIONS:
        assert "sk_engine_" + n in names


def test_engine_library_exports_every_declared_symbol():
    from skirt9_b200 import build
    lib = ctypes.CDLL(build.build())
    for n in declared_symbols():
        assert hasattr(lib, n), n
    lib.sk_abi_version.restype = ctypes.c_int
    assert lib.sk_abi_version() == 6


def test_struct_sizes_match_the_c_header():
    """ctypes mirrors must have the C layout: compile a probe with gcc and compare sizeof."""
    import subprocess
    import tempfile
    src = '#include "sk_engine.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",' \
          "sizeof(sk_config_t),sizeof(sk_wavelength_grid_t),sizeof(sk_dustmix_t),sizeof(sk_source_t)," \
          "sizeof(sk_instrument_t),sizeof(sk_secondary_t),sizeof(sk_counters_t));return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"),
                               os.path.join(d, "p.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    mine = [ctypes.sizeof(t) for t in (abi.SkConfig, abi.SkWavelengthGrid, abi.SkDustMix, abi.SkSource,
                                       abi.SkInstrument, abi.SkSecondary, abi.SkCounters)]
    assert sizes == mine


def test_no_gpu_means_loud_failure_not_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(abi.SkError) as ei:
        abi.Engine(abi.SkConfig(0, 1, 0, 0.5, 1e4, 0, 0))
    assert ei.value.code == abi.SK_ERR_CUDA


def test_product_package_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "skirt9_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "liboracle" not in text and "sko_" not in text.replace("``sko_``", ""), f
