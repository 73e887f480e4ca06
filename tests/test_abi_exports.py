"""The C-ABI library loads on a machine without a GPU and exports every symbol include/b200slam.h declares
(no compute calls here); compute entry points refuse to run without a device instead of falling back."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200slam.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    # function declarations only: `b2s_status (*callback)(...)` typedefs and struct members are not symbols
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\((?!\s*\*)", text)))


def test_all_declared_symbols_exported(pkg):
    M = pkg.load("matcher")
    L = M.lib()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/b200slam.h but not exported by libb200slam.so"
    assert L.b2s_abi_version() == 1


def test_struct_sizes_match_header(pkg):
    abi = pkg.abi
    assert C.sizeof(abi.MatcherParams) == 11 * 8 + 8
    assert C.sizeof(abi.Laser) == 8 + 8 * 8
    assert C.sizeof(abi.GridInfo) == 40
    assert C.sizeof(abi.Search) == 6 * 8 + 8
    assert C.sizeof(abi.MatchResult) == 8 + 24 + 72 + 8
    assert C.sizeof(abi.OccGridInfo) == 16 + 16 + 8 + 8


def test_no_cpu_fallback_without_device(pkg):
    """On a box without a GPU every create call must fail with NO_DEVICE (never silently run on the CPU)."""
    M = pkg.load("matcher")
    if M.device_count() > 0:
        return
    abi, synth = pkg.abi, pkg.synth
    import pytest
    with pytest.raises(M.B2SError) as e:
        M.ScanMatcher(abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser()), 1)
    assert e.value.status == abi.B2S_ERR_NO_DEVICE
    for mod, ctor in (("occgrid", lambda m: m.OccupancyGrid(abi.laser_from(synth.Laser()), [[1.0] * 1081], [[0, 0, 0]], 0.05)),
                      ("hector", lambda m: m.HectorMap(64, 64, 0.05)), ("gmapping", lambda m: m.GMap())):
        with pytest.raises(M.B2SError) as e:
            ctor(pkg.load(mod))
        assert e.value.status == abi.B2S_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """The package (product) must not reference oracle/ in any way."""
    pkgdir = os.path.join(ROOT, "creating-2d-laser-slam-from-scratch_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"
