"""CPU: the C-ABI library builds, loads, and exports every symbol include/rcs.h declares; without a
device the compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from rmf_crowdsim_b200 import _build, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rcs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rcs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    syms = header_symbols()
    assert "rcs_step" in syms and "rcs_sim_create" in syms and len(syms) >= 30


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in header_symbols():
        assert hasattr(lib, name), f"librcs.so does not export {name}"
    assert set(_native.SIGNATURES) == set(header_symbols()), "ctypes table and header disagree"
    assert lib.rcs_abi_version() == 1


def test_sources_are_sm100a_only():
    assert "arch=compute_100a,code=sm_100a" in " ".join(_build.NVCC_FLAGS)
    assert "--fmad=false" in _build.NVCC_FLAGS


def test_product_does_not_touch_the_oracle():
    """Nothing under rmf_crowdsim_b200/ or include/ may reference oracle/."""
    bad = []
    for base in ("rmf_crowdsim_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if "oracle" in txt.lower() and f not in ("scenes.py",):
                        for line in txt.splitlines():
                            if re.search(r"(import|include|CDLL|dlopen).*oracle", line, re.I):
                                bad.append((f, line.strip()))
    assert not bad, bad


def test_no_device_fails_loudly(gpu_available):
    if gpu_available:
        pytest.skip("a device is present")
    lib = _native.load()
    desc = _native.SimDesc(10.0, 10.0, 1.0, 0.0, 0.0, 16, 0, 0)
    h = ctypes.c_void_p()
    rc = lib.rcs_sim_create(ctypes.byref(desc), ctypes.byref(h))
    assert rc == _native.RCS_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.rcs_last_error(None)
    import rmf_crowdsim_b200 as R

    with pytest.raises(R.CrowdsimError):
        R.LocationHash2D(10, 10, 1, (0, 0))


def test_rust_sys_crate_declares_only_header_symbols():
    """bindings/rust/rmf_crowdsim_gpu-sys is written by hand (no Rust toolchain here): at least every function it
    declares must exist in include/rcs.h, and the per-step entry points a maintainer needs must all be there."""
    import re

    src = open(os.path.join(ROOT, "bindings", "rust", "rmf_crowdsim_gpu-sys", "src", "lib.rs")).read()
    rust = set(re.findall(r"pub fn (rcs_[a-z0-9_]+)\(", src))
    hdr = set(header_symbols())
    assert rust and rust <= hdr, sorted(rust - hdr)
    need = {"rcs_sim_create", "rcs_sim_destroy", "rcs_last_error", "rcs_lp_none", "rcs_lp_zanlungo", "rcs_add_agents",
            "rcs_remove_agents", "rcs_step", "rcs_read_agents", "rcs_poll_events", "rcs_add_source_sink",
            "rcs_query_radius", "rcs_query_knn", "rcs_index_add_or_update", "rcs_index_remove"}
    assert need <= rust
