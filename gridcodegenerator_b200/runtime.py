"""ctypes loader for a robot-specialised libgrid_<robot>.so and the Python-side engine.

``GridEngine`` is the host-side mirror of the reference's emitted host layer
(GRiDCodeGenerator.py:87-203 and the ``gen_*_host`` generators): the ``*_device`` methods
are the ``_compute_only`` functions (device pointers in, device pointers out, asynchronous
on the caller's stream - torch tensors are accepted as a convenience for their pointers),
the ``GridData`` methods are the mode-0 host functions (H2D, kernel, D2H through pinned
buffers).  There is no CPU fallback: if the library cannot be built or loaded, or no CUDA
device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from .build import build_robot_library
from .codegen import KernelPlan
from .robot import Robot

c_float_p = ctypes.POINTER(ctypes.c_float)

EXPORTS = {
    # name: (restype, argtypes)
    "grid_abi_version": (ctypes.c_int, []),
    "grid_num_joints": (ctypes.c_int, []),
    "grid_robot_name": (ctypes.c_char_p, []),
    "grid_robot_hash": (ctypes.c_char_p, []),
    "grid_last_error": (ctypes.c_char_p, []),
    "grid_kernel_kind": (ctypes.c_char_p, [ctypes.c_char_p]),
    "grid_traced_flops": (ctypes.c_longlong, [ctypes.c_char_p]),
    "grid_inverse_dynamics_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                    ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "grid_direct_minv_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_void_p]),
    "grid_forward_dynamics_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_float, ctypes.c_void_p]),
    "grid_inverse_dynamics_gradient_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                             ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                             ctypes.c_void_p]),
    "grid_forward_dynamics_gradient_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                             ctypes.c_float, ctypes.c_void_p]),
    "grid_data_create": (ctypes.c_void_p, [ctypes.c_int]),
    "grid_data_destroy": (None, [ctypes.c_void_p]),
    "grid_data_capacity": (ctypes.c_int, [ctypes.c_void_p]),
    "grid_data_ptr": (c_float_p, [ctypes.c_void_p, ctypes.c_char_p]),
    "grid_inverse_dynamics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int]),
    "grid_direct_minv": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "grid_forward_dynamics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float]),
    "grid_inverse_dynamics_gradient": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                                      ctypes.c_int]),
    "grid_forward_dynamics_gradient": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int]),
    "grid_measure_fp32_tflops": (ctypes.c_double, [ctypes.c_int]),
    "grid_launch_count": (ctypes.c_longlong, []),
    "grid_time_launches": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_int, c_float_p]),
}


class GridError(RuntimeError):
    pass


def load_library(path: str) -> ctypes.CDLL:
    if not os.path.exists(path):
        raise GridError("GRiD library %s does not exist (no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


def _ptr(x) -> Optional[int]:
    """Device pointer of a torch tensor / raw int address / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        if not x.is_cuda:
            raise GridError("device entry points need CUDA tensors (use GridData for host buffers)")
        if x.dtype.itemsize != 4 or not x.dtype.is_floating_point:
            raise GridError("tensors must be float32")
        if not x.is_contiguous():
            raise GridError("tensors must be contiguous")
        return x.data_ptr()
    raise GridError("unsupported buffer type %r" % type(x))


def _stream(stream) -> Optional[int]:
    if stream is None:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.cuda.current_stream().cuda_stream
        except ImportError:
            pass
        return None
    return getattr(stream, "cuda_stream", stream)


class GridData:
    """gridData<T> twin: pinned host buffers h_* exposed as numpy views, device buffers d_*."""

    FIELDS = {"q_qd_u": lambda n: 3 * n, "q_qd": lambda n: 2 * n, "q": lambda n: n, "c": lambda n: n,
              "Minv": lambda n: n * n, "qdd": lambda n: n, "dc_du": lambda n: 2 * n * n, "df_du": lambda n: 2 * n * n}

    def __init__(self, engine: "GridEngine", max_timesteps: int):
        self.engine, self.lib, self.n = engine, engine.lib, engine.n
        self.handle = self.lib.grid_data_create(int(max_timesteps))
        if not self.handle:
            raise GridError("grid_data_create failed: %s" % self.lib.grid_last_error().decode())
        self.capacity = int(max_timesteps)
        self.h = {}
        self.d = {}
        for f, words in self.FIELDS.items():
            hp = self.lib.grid_data_ptr(self.handle, ("h_" + f).encode())
            dp = self.lib.grid_data_ptr(self.handle, ("d_" + f).encode())
            self.h[f] = np.ctypeslib.as_array(hp, shape=(self.capacity, words(self.n)))
            self.d[f] = ctypes.cast(dp, ctypes.c_void_p).value

    def close(self):
        if self.handle:
            self.h = {}
            self.lib.grid_data_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise GridError("%s failed (%d): %s" % (what, rc, self.lib.grid_last_error().decode()))

    def inverse_dynamics(self, T, gravity=9.81, use_qdd=False, compressed=False):
        self._check(self.lib.grid_inverse_dynamics(self.handle, T, gravity, int(use_qdd), int(compressed)), "inverse_dynamics")
        return self.h["c"][:T]

    def direct_minv(self, T, compressed=False):
        self._check(self.lib.grid_direct_minv(self.handle, T, int(compressed)), "direct_minv")
        return self.h["Minv"][:T]

    def forward_dynamics(self, T, gravity=9.81):
        self._check(self.lib.grid_forward_dynamics(self.handle, T, gravity), "forward_dynamics")
        return self.h["qdd"][:T]

    def inverse_dynamics_gradient(self, T, gravity=9.81, use_qdd=False, compressed=False):
        self._check(self.lib.grid_inverse_dynamics_gradient(self.handle, T, gravity, int(use_qdd), int(compressed)),
                    "inverse_dynamics_gradient")
        return self.h["dc_du"][:T]

    def forward_dynamics_gradient(self, T, gravity=9.81, use_qdd_minv=False):
        self._check(self.lib.grid_forward_dynamics_gradient(self.handle, T, gravity, int(use_qdd_minv)),
                    "forward_dynamics_gradient")
        return self.h["df_du"][:T]


class GridEngine:
    """Robot-specialised dynamics engine: build (cached) + load + call."""

    def __init__(self, robot: Robot, plan: Optional[KernelPlan] = None, force_build: bool = False, tag: str = "",
                 lib_path: Optional[str] = None):
        self.robot = robot
        self.n = robot.n
        if lib_path is None:
            lib_path, self.build_info = build_robot_library(robot, plan, force=force_build, tag=tag)
        else:
            self.build_info = {}
        self.lib_path = lib_path
        self.lib = load_library(lib_path)
        if self.lib.grid_num_joints() != self.n or self.lib.grid_robot_hash().decode() != robot.param_hash():
            raise GridError("library %s was built for a different robot" % lib_path)

    # ---- info -----------------------------------------------------------------------
    def kernel_kind(self, alg: str) -> str:
        return self.lib.grid_kernel_kind(alg.encode()).decode()

    def traced_flops(self, alg: str) -> int:
        return int(self.lib.grid_traced_flops(alg.encode()))

    def launch_count(self) -> int:
        return int(self.lib.grid_launch_count())

    def measure_fp32_tflops(self, repeats: int = 5) -> float:
        v = float(self.lib.grid_measure_fp32_tflops(repeats))
        if v <= 0:
            raise GridError("fp32 microbenchmark failed: %s" % self.lib.grid_last_error().decode())
        return v

    def time_launches(self, alg: str, out, inp, num_timesteps=None, stride=None, gravity=9.81, reps=200):
        """Per-launch GPU durations (us) of `reps` back-to-back launches, event pairs recorded in C."""
        stride = int(inp.shape[-1]) if stride is None else stride
        T = int(inp.shape[0]) if num_timesteps is None else num_timesteps
        buf = (ctypes.c_float * reps)()
        self._check(self.lib.grid_time_launches(alg.encode(), _ptr(out), _ptr(inp), stride, T, gravity, reps, buf),
                    "grid_time_launches")
        return np.ctypeslib.as_array(buf).copy()

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise GridError("%s failed (%d): %s" % (what, rc, self.lib.grid_last_error().decode()))

    @staticmethod
    def _rows(t, stride):
        return int(t.shape[0]) if stride is None else int(t.numel() // stride)

    # ---- device entry points (tensors or raw pointers) --------------------------------
    def inverse_dynamics_device(self, c, q_qd, qdd=None, num_timesteps=None, stride=None, gravity=9.81, stream=None):
        stride = int(q_qd.shape[-1]) if stride is None else stride
        T = int(q_qd.shape[0]) if num_timesteps is None else num_timesteps
        self._check(self.lib.grid_inverse_dynamics_device(_ptr(c), _ptr(q_qd), stride, _ptr(qdd), T, gravity,
                                                          _stream(stream)), "inverse_dynamics_device")
        return c

    def direct_minv_device(self, Minv, q, num_timesteps=None, stride=None, stream=None):
        stride = int(q.shape[-1]) if stride is None else stride
        T = int(q.shape[0]) if num_timesteps is None else num_timesteps
        self._check(self.lib.grid_direct_minv_device(_ptr(Minv), _ptr(q), stride, T, _stream(stream)),
                    "direct_minv_device")
        return Minv

    def forward_dynamics_device(self, qdd, q_qd_u, num_timesteps=None, stride=None, gravity=9.81, stream=None):
        stride = int(q_qd_u.shape[-1]) if stride is None else stride
        T = int(q_qd_u.shape[0]) if num_timesteps is None else num_timesteps
        self._check(self.lib.grid_forward_dynamics_device(_ptr(qdd), _ptr(q_qd_u), stride, T, gravity,
                                                          _stream(stream)), "forward_dynamics_device")
        return qdd

    def inverse_dynamics_gradient_device(self, dc_du, q_qd, qdd=None, num_timesteps=None, stride=None, gravity=9.81,
                                         stream=None):
        stride = int(q_qd.shape[-1]) if stride is None else stride
        T = int(q_qd.shape[0]) if num_timesteps is None else num_timesteps
        self._check(self.lib.grid_inverse_dynamics_gradient_device(_ptr(dc_du), _ptr(q_qd), stride, _ptr(qdd), T,
                                                                   gravity, _stream(stream)),
                    "inverse_dynamics_gradient_device")
        return dc_du

    def forward_dynamics_gradient_device(self, df_du, q_qd_u, qdd=None, Minv=None, num_timesteps=None, stride=None,
                                         gravity=9.81, stream=None):
        stride = int(q_qd_u.shape[-1]) if stride is None else stride
        T = int(q_qd_u.shape[0]) if num_timesteps is None else num_timesteps
        self._check(self.lib.grid_forward_dynamics_gradient_device(_ptr(df_du), _ptr(q_qd_u), stride, _ptr(qdd),
                                                                   _ptr(Minv), T, gravity, _stream(stream)),
                    "forward_dynamics_gradient_device")
        return df_du

    # ---- host entry points ------------------------------------------------------------
    def make_data(self, max_timesteps: int) -> GridData:
        return GridData(self, max_timesteps)


_ENGINES = {}


def get_engine(robot_or_name, **kw) -> GridEngine:
    """Process-wide cache of engines keyed by robot hash."""
    from .urdf import load_named_robot
    robot = load_named_robot(robot_or_name) if isinstance(robot_or_name, str) else robot_or_name
    key = (robot.param_hash(), kw.get("tag", ""))
    if key not in _ENGINES:
        _ENGINES[key] = GridEngine(robot, **kw)
    return _ENGINES[key]
