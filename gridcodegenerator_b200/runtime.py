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
    "grid_forward_dynamics_gradient_vjp_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                                 ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                                 ctypes.c_float, ctypes.c_void_p]),
    "grid_forward_dynamics_linearize_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                              ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                              ctypes.c_void_p]),
    "grid_crba_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "grid_aba_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_void_p]),
    "grid_forward_dynamics_gradient_vjp": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float]),
    "grid_forward_dynamics_linearize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float]),
    "grid_graph_create": (ctypes.c_void_p, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                            ctypes.c_float]),
    "grid_graph_launch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "grid_graph_destroy": (None, [ctypes.c_void_p]),
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
    "grid_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p]),
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


class _Handle:
    """Owns one grid_data*.  Every numpy view of the pinned buffers references this object through
    its base chain, so the buffers are freed (cudaFreeHost) only after the GridData object has been
    closed or collected AND the last view is gone - a result kept past close() stays valid."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            if self.ptr:
                self.lib.grid_data_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _pinned_view(owner: _Handle, addr: int, rows: int, words: int) -> np.ndarray:
    buf = (ctypes.c_float * (rows * words)).from_address(addr)
    buf._grid_owner = owner                      # ctypes array -> handle: rides along in arr.base
    return np.ctypeslib.as_array(buf).reshape(rows, words)


class GridData:
    """gridData<T> twin: pinned host buffers h_* exposed as numpy views, device buffers d_*."""

    FIELDS = {"q_qd_u": lambda n: 3 * n, "q_qd": lambda n: 2 * n, "q": lambda n: n, "c": lambda n: n,
              "Minv": lambda n: n * n, "qdd": lambda n: n, "dc_du": lambda n: 2 * n * n, "df_du": lambda n: 2 * n * n}
    # buffers of the fused FD-gradient consumers: allocated by the library on first access
    CONSUMER_FIELDS = {"lambda": lambda n: 2 * n, "vjp": lambda n: 5 * n, "lin": lambda n: 2 * n + 3 * n * n}

    def __init__(self, engine: "GridEngine", max_timesteps: int):
        self.engine, self.lib, self.n = engine, engine.lib, engine.n
        ptr = self.lib.grid_data_create(int(max_timesteps))
        if not ptr:
            raise GridError("grid_data_create failed: %s" % self.lib.grid_last_error().decode())
        self._owner = _Handle(self.lib, ptr)
        self.capacity = int(max_timesteps)
        self.h = {}
        self.d = {}
        self._map(self.FIELDS)

    def _map(self, fields):
        ptr = self.handle
        for f, words in fields.items():
            hp = self.lib.grid_data_ptr(ptr, ("h_" + f).encode())
            dp = self.lib.grid_data_ptr(ptr, ("d_" + f).encode())
            if not hp or not dp:
                raise GridError("grid_data_ptr(%s) failed: %s" % (f, self.lib.grid_last_error().decode()))
            self.h[f] = _pinned_view(self._owner, ctypes.cast(hp, ctypes.c_void_p).value, self.capacity, words(self.n))
            self.d[f] = ctypes.cast(dp, ctypes.c_void_p).value

    def consumer_buffers(self):
        """Maps h_lambda / h_vjp / h_lin (allocated by the library on first use) into self.h / self.d."""
        if "lambda" not in self.h:
            self._map(self.CONSUMER_FIELDS)
        return self.h

    @property
    def handle(self):
        if self._owner is None:
            raise GridError("this GridData has been closed")
        return self._owner.ptr

    def close(self):
        """Drops this object's references; the pinned/device buffers are released as soon as no numpy
        view returned earlier is alive any more (never under a live view)."""
        self.h = {}
        self.d = {}
        self._owner = None

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise GridError("%s failed (%d): %s" % (what, rc, self.lib.grid_last_error().decode()))

    def inverse_dynamics(self, T, gravity=9.81, use_qdd=False, compressed=False):
        self.engine._sync_options()
        self._check(self.lib.grid_inverse_dynamics(self.handle, T, gravity, int(use_qdd), int(compressed)), "inverse_dynamics")
        return self.h["c"][:T]

    def direct_minv(self, T, compressed=False):
        self.engine._sync_options()
        self._check(self.lib.grid_direct_minv(self.handle, T, int(compressed)), "direct_minv")
        return self.h["Minv"][:T]

    def forward_dynamics(self, T, gravity=9.81):
        self.engine._sync_options()
        self._check(self.lib.grid_forward_dynamics(self.handle, T, gravity), "forward_dynamics")
        return self.h["qdd"][:T]

    def inverse_dynamics_gradient(self, T, gravity=9.81, use_qdd=False, compressed=False):
        self.engine._sync_options()
        self._check(self.lib.grid_inverse_dynamics_gradient(self.handle, T, gravity, int(use_qdd), int(compressed)),
                    "inverse_dynamics_gradient")
        return self.h["dc_du"][:T]

    def forward_dynamics_gradient(self, T, gravity=9.81, use_qdd_minv=False):
        self.engine._sync_options()
        self._check(self.lib.grid_forward_dynamics_gradient(self.handle, T, gravity, int(use_qdd_minv)),
                    "forward_dynamics_gradient")
        return self.h["df_du"][:T]

    def forward_dynamics_gradient_vjp(self, T, dt, gravity=9.81):
        """h_q_qd_u, h_lambda -> h_vjp = [x+ | A^T lam | B^T lam] (5n per state); see include/grid_b200.h."""
        self.consumer_buffers()
        self.engine._sync_options()
        self._check(self.lib.grid_forward_dynamics_gradient_vjp(self.handle, T, dt, gravity), "forward_dynamics_gradient_vjp")
        return self.h["vjp"][:T]

    def forward_dynamics_linearize(self, T, dt, gravity=9.81):
        """h_q_qd_u -> h_lin = [x+ | A21 | A22 | B2] (2n + 3n^2 per state)."""
        self.consumer_buffers()
        self.engine._sync_options()
        self._check(self.lib.grid_forward_dynamics_linearize(self.handle, T, dt, gravity), "forward_dynamics_linearize")
        return self.h["lin"][:T]


class GridGraph:
    """One captured fixed-shape launch (grid_graph_*); the tensors it was built on must outlive it."""

    def __init__(self, engine: "GridEngine", alg: str, out, inp, in1=None, in2=None, num_timesteps=None, stride=None,
                 dt: float = 0.0, gravity: float = 9.81):
        self.engine, self.lib = engine, engine.lib
        self.keep = (out, inp, in1, in2)
        stride = int(inp.shape[-1]) if stride is None else stride
        T = int(inp.shape[0]) if num_timesteps is None else num_timesteps
        engine._sync_options()
        self.ptr = self.lib.grid_graph_create(alg.encode(), _ptr(out), _ptr(inp), stride, _ptr(in1), _ptr(in2), T, dt,
                                              gravity)
        if not self.ptr:
            raise GridError("grid_graph_create failed: %s" % self.lib.grid_last_error().decode())

    def launch(self, stream=None):
        rc = self.lib.grid_graph_launch(self.ptr, _stream(stream))
        if rc != 0:
            raise GridError("grid_graph_launch failed (%d): %s" % (rc, self.lib.grid_last_error().decode()))

    def close(self):
        if self.ptr:
            self.lib.grid_graph_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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

    # ---- options: the library reads the environment once; later changes of the same variables in
    # THIS process (tests, sweeps) are pushed through grid_set_option before a call -----------------
    _OPTION_KEYS = ("GRID_FORCE_KERNEL", "GRID_PIPE_MODE", "GRID_PIPE_CHUNK", "GRID_PIPE_WARPS", "GRID_PIPE_STAGGER_NS",
                    "GRID_PIPE_ONLY_TASK", "GRID_PIPE_ORDER_CHUNK")

    def _sync_options(self):
        cur = tuple(os.environ.get(k) for k in self._OPTION_KEYS)
        if cur != getattr(self, "_pushed_options", None):
            for k, v in zip(self._OPTION_KEYS, cur):
                rc = self.lib.grid_set_option(k.encode(), None if v is None else v.encode())
                if rc != 0:
                    raise GridError("grid_set_option(%s=%r) failed: %s" % (k, v, self.lib.grid_last_error().decode()))
            self._pushed_options = cur

    def set_option(self, key: str, value: Optional[str]):
        """Explicit form: also updates os.environ so that the next call does not push the old value back."""
        if value is None:
            os.environ.pop(key, None)
        else:
            os.environ[key] = value
        self._sync_options()

    def time_launches(self, alg: str, out, inp, num_timesteps=None, stride=None, gravity=9.81, reps=200):
        """Per-launch GPU durations (us) of `reps` back-to-back launches, event pairs recorded in C.
        alg may end in "@graph": the launch is captured once and the timed launches replay the graph."""
        stride = int(inp.shape[-1]) if stride is None else stride
        T = int(inp.shape[0]) if num_timesteps is None else num_timesteps
        n = self.n
        words = {"id": n, "minv": n * n, "fd": n, "aba": n, "crba": n * n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n,
                 "fd_vjp": 5 * n, "fd_lin": 2 * n + 3 * n * n, "noop": 0}.get(alg.split("@")[0])
        if words is None:
            raise GridError("time_launches: unknown algorithm %r" % alg)
        if hasattr(out, "numel") and out.numel() < T * words:
            raise GridError("time_launches: the output buffer holds %d floats, %d states of %s need %d"
                            % (out.numel(), T, alg, T * words))
        buf = (ctypes.c_float * reps)()
        self._sync_options()
        self._check(self.lib.grid_time_launches(alg.encode(), _ptr(out), _ptr(inp), stride, T, gravity, reps, buf),
                    "grid_time_launches")
        return np.ctypeslib.as_array(buf).copy()

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise GridError("%s failed (%d): %s" % (what, rc, self.lib.grid_last_error().decode()))

    # ---- argument validation of the device entry points -------------------------------------------
    def _shape(self, what: str, inp, min_words: int, num_timesteps, stride, outs):
        """Resolves (T, stride) and checks every buffer: a mismatch must raise here, not become a silent
        out-of-bounds access on the GPU.  `outs`: (name, tensor or None, words per state)."""
        tensor = hasattr(inp, "data_ptr")
        if stride is None:
            if not tensor or inp.dim() != 2:
                raise GridError("%s: the state array must be 2-D [states, words] (or pass stride=)" % what)
            stride = int(inp.shape[1])
        if num_timesteps is None:
            if not tensor:
                raise GridError("%s: raw pointers need num_timesteps=" % what)
            num_timesteps = int(inp.shape[0]) if inp.dim() == 2 else int(inp.numel() // stride)
        T = int(num_timesteps)
        if T < 0 or stride < min_words:
            raise GridError("%s: stride %d is smaller than the %d words the algorithm reads per state" % (what, stride, min_words))
        dev = None
        if tensor:
            dev = inp.device
            if T and inp.numel() < (T - 1) * stride + min_words:
                raise GridError("%s: the state array holds fewer than %d states of stride %d" % (what, T, stride))
        for name, t, words in outs:
            if t is None or not hasattr(t, "data_ptr"):
                continue
            if t.numel() < T * words:
                raise GridError("%s: %s holds %d floats, %d states need %d" % (what, name, t.numel(), T, T * words))
            if dev is not None and t.device != dev:
                raise GridError("%s: %s is on %s, the states are on %s" % (what, name, t.device, dev))
        if dev is not None:
            import torch
            if dev.index is not None and dev.index != torch.cuda.current_device():
                raise GridError("%s: tensors are on cuda:%d but the current device is cuda:%d (use torch.cuda.device)"
                                % (what, dev.index, torch.cuda.current_device()))
        self._sync_options()
        return T, int(stride)

    # ---- device entry points (tensors or raw pointers) --------------------------------
    def inverse_dynamics_device(self, c, q_qd, qdd=None, num_timesteps=None, stride=None, gravity=9.81, stream=None):
        n = self.n
        T, stride = self._shape("inverse_dynamics_device", q_qd, 2 * n, num_timesteps, stride,
                                [("c", c, n), ("qdd", qdd, n)])
        self._check(self.lib.grid_inverse_dynamics_device(_ptr(c), _ptr(q_qd), stride, _ptr(qdd), T, gravity,
                                                          _stream(stream)), "inverse_dynamics_device")
        return c

    def direct_minv_device(self, Minv, q, num_timesteps=None, stride=None, stream=None):
        n = self.n
        T, stride = self._shape("direct_minv_device", q, n, num_timesteps, stride, [("Minv", Minv, n * n)])
        self._check(self.lib.grid_direct_minv_device(_ptr(Minv), _ptr(q), stride, T, _stream(stream)),
                    "direct_minv_device")
        return Minv

    def forward_dynamics_device(self, qdd, q_qd_u, num_timesteps=None, stride=None, gravity=9.81, stream=None):
        n = self.n
        T, stride = self._shape("forward_dynamics_device", q_qd_u, 3 * n, num_timesteps, stride, [("qdd", qdd, n)])
        self._check(self.lib.grid_forward_dynamics_device(_ptr(qdd), _ptr(q_qd_u), stride, T, gravity,
                                                          _stream(stream)), "forward_dynamics_device")
        return qdd

    # ---- further algorithms (include/grid_b200.h): mass matrix by CRBA, forward dynamics by ABA -------------
    def crba_device(self, M, q, num_timesteps=None, stride=None, stream=None):
        """M[n*n] per state: the joint-space mass matrix, column-major, both triangles."""
        n = self.n
        T, stride = self._shape("crba_device", q, n, num_timesteps, stride, [("M", M, n * n)])
        self._check(self.lib.grid_crba_device(_ptr(M), _ptr(q), stride, T, _stream(stream)), "crba_device")
        return M

    def aba_device(self, qdd, q_qd_u, num_timesteps=None, stride=None, gravity=9.81, stream=None):
        """qdd[n] per state by the articulated-body algorithm (same contract as forward_dynamics_device)."""
        n = self.n
        T, stride = self._shape("aba_device", q_qd_u, 3 * n, num_timesteps, stride, [("qdd", qdd, n)])
        self._check(self.lib.grid_aba_device(_ptr(qdd), _ptr(q_qd_u), stride, T, gravity, _stream(stream)), "aba_device")
        return qdd

    def inverse_dynamics_gradient_device(self, dc_du, q_qd, qdd=None, num_timesteps=None, stride=None, gravity=9.81,
                                         stream=None):
        n = self.n
        T, stride = self._shape("inverse_dynamics_gradient_device", q_qd, 2 * n, num_timesteps, stride,
                                [("dc_du", dc_du, 2 * n * n), ("qdd", qdd, n)])
        self._check(self.lib.grid_inverse_dynamics_gradient_device(_ptr(dc_du), _ptr(q_qd), stride, _ptr(qdd), T,
                                                                   gravity, _stream(stream)),
                    "inverse_dynamics_gradient_device")
        return dc_du

    def forward_dynamics_gradient_device(self, df_du, q_qd_u, qdd=None, Minv=None, num_timesteps=None, stride=None,
                                         gravity=9.81, stream=None):
        n = self.n
        pre = qdd is not None or Minv is not None
        T, stride = self._shape("forward_dynamics_gradient_device", q_qd_u, (2 if pre else 3) * n, num_timesteps, stride,
                                [("df_du", df_du, 2 * n * n), ("qdd", qdd, n), ("Minv", Minv, n * n)])
        self._check(self.lib.grid_forward_dynamics_gradient_device(_ptr(df_du), _ptr(q_qd_u), stride, _ptr(qdd),
                                                                   _ptr(Minv), T, gravity, _stream(stream)),
                    "forward_dynamics_gradient_device")
        return df_du

    # ---- consumers fused after the FD gradient (include/grid_b200.h) -------------------------------
    def forward_dynamics_gradient_vjp_device(self, out, q_qd_u, lam, dt, num_timesteps=None, stride=None, gravity=9.81,
                                             stream=None):
        """out[5n] = [x+ | A^T lam | B^T lam] per state, lam = [lam_q | lam_v] (2n)."""
        n = self.n
        T, stride = self._shape("forward_dynamics_gradient_vjp_device", q_qd_u, 3 * n, num_timesteps, stride,
                                [("out", out, 5 * n), ("lambda", lam, 2 * n)])
        self._check(self.lib.grid_forward_dynamics_gradient_vjp_device(_ptr(out), _ptr(q_qd_u), stride, _ptr(lam), T, dt,
                                                                       gravity, _stream(stream)),
                    "forward_dynamics_gradient_vjp_device")
        return out

    def forward_dynamics_linearize_device(self, out, q_qd_u, dt, num_timesteps=None, stride=None, gravity=9.81,
                                          stream=None):
        """out[2n + 3n^2] = [x+ | A21 | A22 | B2] per state."""
        n = self.n
        T, stride = self._shape("forward_dynamics_linearize_device", q_qd_u, 3 * n, num_timesteps, stride,
                                [("out", out, 2 * n + 3 * n * n)])
        self._check(self.lib.grid_forward_dynamics_linearize_device(_ptr(out), _ptr(q_qd_u), stride, T, dt, gravity,
                                                                    _stream(stream)),
                    "forward_dynamics_linearize_device")
        return out

    def make_graph(self, alg: str, out, inp, in1=None, in2=None, **kw) -> "GridGraph":
        return GridGraph(self, alg, out, inp, in1, in2, **kw)

    # ---- host entry points ------------------------------------------------------------
    def make_data(self, max_timesteps: int) -> GridData:
        return GridData(self, max_timesteps)


_ENGINES = {}


def get_engine(robot_or_name, **kw) -> GridEngine:
    """Process-wide cache of engines keyed by robot hash."""
    from .urdf import load_named_robot
    robot = load_named_robot(robot_or_name) if isinstance(robot_or_name, str) else robot_or_name
    from .codegen import plan_signature
    key = (robot.param_hash(), kw.get("tag", ""), plan_signature(kw.get("plan")), kw.get("lib_path"))
    if key not in _ENGINES:
        _ENGINES[key] = GridEngine(robot, **kw)
    return _ENGINES[key]
