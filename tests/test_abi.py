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
    assert lib.grid_abi_version() == 2
    assert lib.grid_num_joints() == 7
    assert lib.grid_robot_name() == b"iiwa14"
    assert lib.grid_robot_hash().decode() == robot.param_hash()
    assert b"tps" in lib.grid_kernel_kind(b"fd_grad")
    assert lib.grid_kernel_kind(b"bogus") == b"none"
    assert lib.grid_traced_flops(b"fd_grad") > 0
    # fused consumers ride on the thread-per-state family for iiwa14
    assert lib.grid_kernel_kind(b"fd_vjp") == b"tps" and lib.grid_kernel_kind(b"fd_lin") == b"tps"
    assert 0 < lib.grid_traced_flops(b"fd_vjp") < lib.grid_traced_flops(b"fd_grad")


def test_argument_validation_without_gpu(iiwa_lib):
    lib = load_library(iiwa_lib)
    # negative count and null pointers are rejected before any CUDA call
    assert lib.grid_forward_dynamics_gradient_device(None, None, 21, None, None, -1, 9.81, None) != 0
    assert b"num_timesteps" in lib.grid_last_error()
    assert lib.grid_forward_dynamics_gradient_device(None, None, 21, None, None, 4, 9.81, None) != 0
    assert lib.grid_inverse_dynamics_device(1 << 20, 1 << 20, 3, None, 4, 9.81, None) != 0   # stride < 2n
    assert b"stride" in lib.grid_last_error()
    assert lib.grid_forward_dynamics_device(None, None, 21, 0, 9.81, None) == 0              # empty batch is a no-op


def test_options_are_set_through_the_abi_not_getenv(iiwa_lib):
    """GRID_FORCE_KERNEL & co. are read once; afterwards only grid_set_option changes them, and bad
    values are rejected with a message (no launch-path getenv)."""
    lib = load_library(iiwa_lib)
    assert lib.grid_set_option(b"GRID_FORCE_KERNEL", b"cps") == 0
    assert lib.grid_set_option(b"GRID_FORCE_KERNEL", None) == 0
    assert lib.grid_set_option(b"GRID_FORCE_KERNEL", b"bogus") != 0
    assert b"GRID_FORCE_KERNEL" in lib.grid_last_error()
    assert lib.grid_set_option(b"GRID_PIPE_MODE", b"fused") == 0 and lib.grid_set_option(b"GRID_PIPE_MODE", b"") == 0
    assert lib.grid_set_option(b"GRID_PIPE_MODE", b"sideways") != 0
    assert lib.grid_set_option(b"GRID_PIPE_CHUNK", b"4096") == 0 and lib.grid_set_option(b"GRID_PIPE_CHUNK", None) == 0
    assert lib.grid_set_option(b"NOT_AN_OPTION", b"1") != 0
    import subprocess
    nm = subprocess.run(["nm", "-D", "--undefined-only", iiwa_lib], capture_output=True, text=True).stdout
    assert "getenv" in nm          # still read once at first use ...
    src = open(os.path.join(os.path.dirname(INCLUDE), "gridcodegenerator_b200", "csrc", "grid_pipe.cuh")).read()
    assert "getenv" not in src     # ... and never from the launchers


def test_consumer_entry_points_validate_without_gpu(iiwa_lib):
    lib = load_library(iiwa_lib)
    assert lib.grid_forward_dynamics_gradient_vjp_device(None, None, 21, None, 4, 0.01, 9.81, None) != 0
    assert lib.grid_forward_dynamics_gradient_vjp_device(1 << 20, 1 << 20, 21, None, 4, 0.01, 9.81, None) != 0
    assert b"d_lambda" in lib.grid_last_error()
    assert lib.grid_forward_dynamics_linearize_device(1 << 20, 1 << 20, 20, 4, 0.01, 9.81, None) != 0     # stride < 3n
    assert lib.grid_forward_dynamics_linearize_device(None, None, 21, 0, 0.01, 9.81, None) == 0
    assert lib.grid_graph_launch(None, None) != 0


def test_library_name_depends_on_the_kernel_plan():
    """a library built with one KernelPlan must never be returned for another."""
    from gridcodegenerator_b200.build import lib_path
    from gridcodegenerator_b200.codegen import KernelPlan, plan_signature
    robot = load_named_robot("mixed5")
    default, same, other = KernelPlan(robot), KernelPlan(robot, cps_max_states=2048), KernelPlan(robot, cps_max_states=64)
    assert plan_signature(None) == "" and plan_signature(default) == "" and plan_signature(same) == ""
    assert plan_signature(other) != ""
    assert lib_path(robot, plan=other) != lib_path(robot) == lib_path(robot, plan=default)
    assert lib_path(robot, plan=KernelPlan(robot, cps_max_states=64)) == lib_path(robot, plan=other)


def test_sass_is_sm100a_and_barrier_free(iiwa_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", iiwa_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_kernel_plan_families_per_robot():
    """Which kernel family serves what: chain kernels only for serial chains without single-thread programs
    (or when forced, for the test builds), fused consumers wherever the FD gradient has tps / pipe / lps kernels, a
    second set of column programs for small Atlas batches, and `only_algs` for experiment builds."""
    from gridcodegenerator_b200.codegen import KernelPlan
    from helpers import cached_plan
    chain = cached_plan("chain64")
    assert chain.lps == {"minv", "fd", "id_grad", "fd_grad"} and chain.kind["id"] == "tps"
    assert all(k.endswith("+lps") for a, k in chain.kind.items() if a != "id")
    assert chain.consumers == {"fd_vjp": "lps", "fd_lin": "lps"}
    iiwa = cached_plan("iiwa14")
    assert not iiwa.lps and iiwa.consumers == {"fd_vjp": "tps", "fd_lin": "tps"} and iiwa.pipe_small is None
    forced = KernelPlan(load_named_robot("pchain4"), lps_force=True)
    assert forced.lps and forced.lps_forced_only and "lps" in forced.consumers["fd_vjp"]
    assert not KernelPlan(load_named_robot("mixed5"), lps_force=True).lps          # a tree is not a chain
    atlas = KernelPlan(load_named_robot("atlas"), only_algs=("fd_grad",))
    assert atlas.kind["minv"] == "none" and "pipe" in atlas.kind["fd_grad"] and atlas.consumers["fd_vjp"] == "none"
    assert atlas.pipe_small is not None and len(atlas.pipe_small.tasks) > len(atlas.pipe["fd_grad"].tasks)
