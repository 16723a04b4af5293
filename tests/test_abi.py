"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/grid_b200.h
declares.  No compute calls here (no GPU)."""
import ctypes
import os
import re

import pytest

from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.build import build_robot_library, INCLUDE
from gridcodegenerator_b200.runtime import EXPORTS, load_library


def declared_symbols():
    text = open(os.path.join(INCLUDE, "grid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grid_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def iiwa_lib():
    so, _ = build_robot_library(load_named_robot("iiwa14"))
    return so


def test_header_and_python_binding_agree():
    assert set(declared_symbols()) == set(EXPORTS)


def test_library_exports_every_declared_symbol(iiwa_lib):
    lib = ctypes.CDLL(iiwa_lib)
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym


def test_identity_calls(iiwa_lib):
    robot = load_named_robot("iiwa14")
    lib = load_library(iiwa_lib)
    assert lib.grid_abi_version() == 1
    assert lib.grid_num_joints() == 7
    assert lib.grid_robot_name() == b"iiwa14"
    assert lib.grid_robot_hash().decode() == robot.param_hash()
    assert b"tps" in lib.grid_kernel_kind(b"fd_grad")
    assert lib.grid_kernel_kind(b"bogus") == b"none"
    assert lib.grid_traced_flops(b"fd_grad") > 0


def test_argument_validation_without_gpu(iiwa_lib):
    lib = load_library(iiwa_lib)
    # negative count and null pointers are rejected before any CUDA call
    assert lib.grid_forward_dynamics_gradient_device(None, None, 21, None, None, -1, 9.81, None) != 0
    assert b"num_timesteps" in lib.grid_last_error()
    assert lib.grid_forward_dynamics_gradient_device(None, None, 21, None, None, 4, 9.81, None) != 0
    assert lib.grid_inverse_dynamics_device(1 << 20, 1 << 20, 3, None, 4, 9.81, None) != 0   # stride < 2n
    assert b"stride" in lib.grid_last_error()
    assert lib.grid_forward_dynamics_device(None, None, 21, 0, 9.81, None) == 0              # empty batch is a no-op


def test_sass_is_sm100a_and_barrier_free(iiwa_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", iiwa_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
