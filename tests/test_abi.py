"""CPU-only checks of the drop-in boundary: the library builds, loads without a GPU, exports every symbol
include/rekf.h declares, and fails loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rekf.h")).read()
    return sorted(set(re.findall(r"REKF_API\s+[\w\s\*]+?\b(rekf_\w+)\s*\(", text)))


def test_header_symbols_match_binding_table():
    from reflector_ekf_slam_b200 import engine
    assert _declared_symbols() == sorted(engine.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(engine_lib):
    for name in _declared_symbols():
        assert hasattr(engine_lib, name), f"librekf_b200.so does not export {name}"


def test_options_struct_layout_matches_header(engine_lib):
    from reflector_ekf_slam_b200._abi import RekfOptions
    o = RekfOptions()
    engine_lib.rekf_default_options(C.byref(o))
    assert o.odom_model == 0 and o.use_imu == 0
    assert abs(o.linear_velocity_cov - 0.0025) < 1e-18 and abs(o.angular_velocity_cov - 0.0064) < 1e-18
    assert abs(o.observation_cov - 0.0025) < 1e-18
    assert (o.max_landmarks, o.max_observations, o.max_map_landmarks) == (1024, 128, 1024)
    assert o.cov_update == 2 and o.map_loader == 0 and not o.stream and o.use_graphs == 0
    assert engine_lib.rekf_version().startswith(b"rekf-b200")


def test_no_gpu_fails_loudly(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from reflector_ekf_slam_b200.engine import RekfError, ReflectorEKFSLAM
    with pytest.raises(RekfError) as e:
        ReflectorEKFSLAM()
    assert e.value.code in (-6, -2)


def test_product_package_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under reflector_ekf_slam_b200/ may reference it."""
    pkg = os.path.join(ROOT, "reflector_ekf_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "rekf_oracle" not in text and "liboracle" not in text, f


def test_sass_contains_blackwell_tensor_and_tma_instructions(engine_lib):
    """UTCHMMA / UTCIMMA (tcgen05.mma kind::tf32 / kind::i8), UTMALDG / UTMAREDG.3D.ADD (TMA tensor loads; the covariance
    downdate leaves the SM as a TMA reduce-add, so there is no plain tensor store), UBLKCP
    (cp.async.bulk: the Cholesky triangle), LDTM (tcgen05.ld) and DMMA (fp64 tensor pipe) must be in the shipped binary."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump")
    if not exe:
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "reflector_ekf_slam_b200", "librekf_b200.so")
    sass = subprocess.run([exe, "-sass", lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTCIMMA", "UTMALDG.5D", "UTMAREDG.3D.ADD", "UBLKCP", "LDTM", "DMMA"):
        assert mnemonic in sass, mnemonic
    # the TRSM that runs beside the Cholesky (solve_ll.cuh): fp64 tensor pipe, cp.async staging, the bounded flag poll
    body = sass.split("k_solve_ll", 1)[1].split("Function :", 1)[0]
    for mnemonic in ("DMMA", "LDGSTS", "NANOSLEEP"):
        assert mnemonic in body, mnemonic
