"""CUDA emission: traced programs -> robot-specialised sm_100a translation unit.

The generated ``grid_<robot>.cu`` contains, per algorithm variant, one ``struct Alg*``
whose ``eval`` is the straight-line per-state program (thread-per-state, see
csrc/grid_tps.cuh), the launcher functions ``gridb200::gen::launch_*`` that
csrc/grid_abi.cuh exposes through the C ABI of include/grid_b200.h, and (for robots too
large for one thread per state) the constant tables of the warp-per-state kernels
(csrc/grid_wps.cuh).

Input words of ``eval`` (what each kernel variant reads per state):
  in0 = the [q | qd | u] row (only the first IN0 words are read, any stride),
  in1 = qdd (n words), in2 = Minv (n*n words).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

from .algorithms import TRACERS, algorithmic_flops
from .ir import Program
from .robot import Robot

CODEGEN_VERSION = "1"

# variant -> (struct name, IN0 words / n, IN1 words / n, IN2 words / n^2, output array, OUT words as f(n))
VARIANTS = {
    "id":               ("AlgId",           2, 0, 0, "c",     lambda n: n),
    "id_qdd":           ("AlgIdQdd",        2, 1, 0, "c",     lambda n: n),
    "minv":             ("AlgMinv",         1, 0, 0, "Minv",  lambda n: n * n),
    "fd":               ("AlgFd",           3, 0, 0, "qdd",   lambda n: n),
    "id_grad":          ("AlgIdGrad",       2, 0, 0, "dc_du", lambda n: 2 * n * n),
    "id_grad_qdd":      ("AlgIdGradQdd",    2, 1, 0, "dc_du", lambda n: 2 * n * n),
    "fd_grad":          ("AlgFdGrad",       3, 0, 0, "df_du", lambda n: 2 * n * n),
    "fd_grad_qdd_minv": ("AlgFdGradPre",    2, 1, 1, "df_du", lambda n: 2 * n * n),
}

_NAME_RE = re.compile(r"^(qdd|qd|q|u|Minv)(\d+)$")


def _flit(x: float) -> str:
    s = "%.9g" % x
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


def emit_eval(p: Program, n: int, in0_words: int, in1_words: int, out_name: str, out_words: int,
              indent: str = "        ", sync_every: int = 0) -> Tuple[List[str], Dict[str, int]]:
    """Prints the live part of ``p`` as straight-line CUDA.  Stores are emitted at the
    point in the trace where the algorithm produced them, so finished output columns do
    not occupy registers."""
    live = p.live_nodes()
    lines: List[str] = []
    # inputs first, then one __syncwarp(): the output tile aliases the input tile
    sincos_done = set()
    body: List[str] = []
    outs_by_pos: Dict[int, List[Tuple[int, object]]] = {}
    n_nodes = len(p.nodes)
    seen_out = set()
    for (name, idx, v) in p.outputs:
        if name != out_name:
            raise ValueError("unexpected output array %r" % name)
        if idx in seen_out:
            raise ValueError("output %s[%d] written twice" % (name, idx))
        seen_out.add(idx)
        pos = -1 if v.is_const else v.i
        outs_by_pos.setdefault(pos, []).append((idx, v))
    if seen_out != set(range(out_words)):
        raise ValueError("traced program does not cover every output word")

    def word_of(name: str) -> int:
        m = _NAME_RE.match(name)
        kind, idx = m.group(1), int(m.group(2))
        if kind == "q":
            return idx
        if kind == "qd":
            return n + idx
        if kind == "u":
            return 2 * n + idx
        if kind == "qdd":
            return in0_words + idx
        return in0_words + in1_words + idx        # Minv

    for i, k in enumerate(p.nodes):
        if live[i] and k[0] == "in" and k[1] != "gravity":
            lines.append("%sconst float t%d = s_in[%d];   // %s" % (indent, i, word_of(k[1]), k[1]))
    lines.append(indent + "__syncwarp();")
    for idx, v in outs_by_pos.get(-1, []):
        body.append("%ss_out[%d] = %s;" % (indent, idx, _flit(v.c)))
    emitted = 0
    for i, k in enumerate(p.nodes):
        if not live[i]:
            continue
        op = k[0]
        emitted += 1
        if sync_every and emitted % sync_every == 0:
            body.append(indent + "__syncthreads();")
        if op == "in":
            if k[1] == "gravity":
                body.append("%sconst float t%d = gravity;" % (indent, i))
        elif op in ("sin", "cos"):
            a = k[1]
            if a not in sincos_done:
                sincos_done.add(a)
                si, ci = p._cse.get(("sin", a)), p._cse.get(("cos", a))
                sname = "t%d" % si if si is not None and live[si] else "unused_s%d" % a
                cname = "t%d" % ci if ci is not None and live[ci] else "unused_c%d" % a
                body.append("%sfloat %s, %s; sincosf(t%d, &%s, &%s);" % (indent, sname, cname, a, sname, cname))
        elif op == "rcp":
            body.append("%sconst float t%d = 1.0f / t%d;" % (indent, i, k[1]))
        elif op == "mul":
            body.append("%sconst float t%d = t%d * t%d;" % (indent, i, k[1], k[2]))
        elif op == "mulc":
            body.append("%sconst float t%d = t%d * %s;" % (indent, i, k[1], _flit(k[2])))
        elif op == "add":
            body.append("%sconst float t%d = t%d %s t%d;" % (indent, i, k[1], "+" if k[3] > 0 else "-", k[2]))
        elif op == "addc":
            body.append("%sconst float t%d = t%d + %s;" % (indent, i, k[1], _flit(k[2])))
        for idx, v in outs_by_pos.get(i, []):
            body.append("%ss_out[%d] = %st%d;" % (indent, idx, "-" if v.s < 0 else "", v.i))
    return lines + body, p.op_counts()


def emit_alg_struct(robot: Robot, variant: str, p: Optional[Program] = None,
                    sync_every: int = 0) -> Tuple[str, Dict[str, int]]:
    n = robot.n
    sname, m0, m1, m2, out_name, out_fn = VARIANTS[variant]
    in0, in1, in2, out = m0 * n, m1 * n, m2 * n * n, out_fn(n)
    p = p if p is not None else TRACERS[variant](robot)
    body, cnt = emit_eval(p, n, in0, in1, out_name, out, sync_every=sync_every)
    txt = ["struct %s {" % sname,
           "    static constexpr int IN0 = %d, IN1 = %d, IN2 = %d, OUT = %d;" % (in0, in1, in2, out),
           "    static constexpr long long TRACED_FLOPS = %d;   // %d mul + %d add per state" % (
               cnt["flops"], cnt["mul"], cnt["add"]),
           "    static __device__ __forceinline__ void eval(const float *s_in, float *s_out, const float gravity) {"]
    txt += body
    txt += ["    }", "};", ""]
    return "\n".join(txt), cnt


class KernelPlan:
    """Which kernel family serves each algorithm of a robot, and its launch shape."""

    def __init__(self, robot: Robot, tps_max_flops: int = 60000, tps_warps: int = 1,
                 tps_min_blocks: Optional[Dict[str, int]] = None, tps_sync_every: int = 0):
        self.robot = robot
        self.tps_warps = tps_warps
        self.tps_sync_every = tps_sync_every if tps_warps > 1 else 0
        alg = algorithmic_flops(robot)
        self.kind: Dict[str, str] = {}
        for a in ("id", "minv", "fd", "id_grad", "fd_grad"):
            # the dense reference count over-states traced work by ~4x; gate on it
            self.kind[a] = "tps" if alg[a] <= tps_max_flops else "none"
        # resident single-warp CTAs per SM = register cap 65536/(32*min_blocks).  Measured on B200
        # (profiles/r1_sweep_tps.md): the gradient programs spill at 128/168 registers and run
        # 2.1x faster at 255 registers with no spills; the small programs fit 128.
        self.min_blocks = {"id": 16, "minv": 16, "fd": 16, "id_grad": 8, "fd_grad": 8}
        if tps_min_blocks:
            self.min_blocks.update(tps_min_blocks)


_LAUNCHERS = r'''
namespace GRID_NS { namespace gen {
%(launchers)s
const char *kernel_kind(const char *alg) {
    if (!alg) return "none";
%(kinds)s
    return "none";
}
long long traced_flops(const char *alg) {
    if (!alg) return 0;
%(flops)s
    return 0;
}
}}  // namespace GRID_NS::gen
'''


def generate_translation_unit(robot: Robot, plan: Optional[KernelPlan] = None,
                              ns_tag: str = "") -> Tuple[str, Dict[str, dict]]:
    plan = plan or KernelPlan(robot)
    ns_tag = re.sub(r"\W", "_", ns_tag)
    n = robot.n
    out: List[str] = []
    stats: Dict[str, dict] = {}
    out.append("// GENERATED by gridcodegenerator_b200.codegen v%s for robot '%s' (hash %s) - do not edit.\n"
               % (CODEGEN_VERSION, robot.name, robot.param_hash()))
    # a robot-unique namespace: several robot libraries are loaded into one process and C++
    # vague-linkage symbols (template statics) would otherwise be shared between them
    out.append("#define GRID_NS grid_%s_%s%s\n" % (re.sub(r"\W", "_", robot.name), robot.param_hash(), ns_tag))
    out.append('#include <cuda_runtime.h>\n#include "grid_tps.cuh"\n')
    out.append('#define GRID_ROBOT_NAME "%s"\n#define GRID_ROBOT_HASH "%s"\n#define GRID_N %d\n'
               % (robot.name, robot.param_hash(), n))
    out.append("namespace GRID_NS { namespace gen {\n")
    needed = {"id": ("id", "id_qdd"), "minv": ("minv",), "fd": ("fd",), "id_grad": ("id_grad", "id_grad_qdd"),
              "fd_grad": ("fd_grad", "fd_grad_qdd_minv")}
    for a, variants in needed.items():
        if plan.kind[a] != "tps":
            continue
        for v in variants:
            txt, cnt = emit_alg_struct(robot, v, sync_every=plan.tps_sync_every)
            out.append(txt)
            stats[v] = cnt
    out.append("}}  // namespace GRID_NS::gen\n")

    W = plan.tps_warps

    def tps(a, struct):
        return "tps_launch<%s, %d, %d>" % (struct, W, plan.min_blocks[a])

    L: List[str] = []
    unsupported = "    return cudaErrorNotSupported;"
    L.append("cudaError_t launch_id(float *d_c, const float *d_q_qd, int stride, const float *d_qdd, int N, float g,"
             " cudaStream_t s) {")
    if plan.kind["id"] == "tps":
        L.append("    if (d_qdd) return %s(d_c, d_q_qd, stride, d_qdd, nullptr, N, g, s);" % tps("id", "AlgIdQdd"))
        L.append("    return %s(d_c, d_q_qd, stride, nullptr, nullptr, N, g, s);" % tps("id", "AlgId"))
    else:
        L.append(unsupported)
    L.append("}")
    L.append("cudaError_t launch_minv(float *d_Minv, const float *d_q, int stride, int N, cudaStream_t s) {")
    L.append("    return %s(d_Minv, d_q, stride, nullptr, nullptr, N, 0.f, s);" % tps("minv", "AlgMinv")
             if plan.kind["minv"] == "tps" else unsupported)
    L.append("}")
    L.append("cudaError_t launch_fd(float *d_qdd, const float *d_q_qd_u, int stride, int N, float g, cudaStream_t s) {")
    L.append("    return %s(d_qdd, d_q_qd_u, stride, nullptr, nullptr, N, g, s);" % tps("fd", "AlgFd")
             if plan.kind["fd"] == "tps" else unsupported)
    L.append("}")
    L.append("cudaError_t launch_id_grad(float *d_dc_du, const float *d_q_qd, int stride, const float *d_qdd, int N,"
             " float g, cudaStream_t s) {")
    if plan.kind["id_grad"] == "tps":
        L.append("    if (d_qdd) return %s(d_dc_du, d_q_qd, stride, d_qdd, nullptr, N, g, s);"
                 % tps("id_grad", "AlgIdGradQdd"))
        L.append("    return %s(d_dc_du, d_q_qd, stride, nullptr, nullptr, N, g, s);" % tps("id_grad", "AlgIdGrad"))
    else:
        L.append(unsupported)
    L.append("}")
    L.append("cudaError_t launch_fd_grad(float *d_df_du, const float *d_in, int stride, const float *d_qdd,"
             " const float *d_Minv, int N, float g, cudaStream_t s) {")
    if plan.kind["fd_grad"] == "tps":
        L.append("    if (d_qdd) return %s(d_df_du, d_in, stride, d_qdd, d_Minv, N, g, s);"
                 % tps("fd_grad", "AlgFdGradPre"))
        L.append("    return %s(d_df_du, d_in, stride, nullptr, nullptr, N, g, s);" % tps("fd_grad", "AlgFdGrad"))
    else:
        L.append(unsupported)
    L.append("}")

    kinds = "\n".join('    if (!strcmp(alg, "%s")) return "%s";' % (a, k) for a, k in plan.kind.items())
    fl = "\n".join('    if (!strcmp(alg, "%s")) return %d;' % (a, stats[needed[a][0]]["flops"])
                   for a in plan.kind if plan.kind[a] == "tps")
    out.append("#include <cstring>\n")
    out.append(_LAUNCHERS % {"launchers": "\n".join(L), "kinds": kinds, "flops": fl})
    out.append('#include "grid_abi.cuh"\n')
    return "".join(out), stats
