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

CODEGEN_VERSION = "2"

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
    # the two halves of the mid-size-batch FD-gradient kernel (tps_half_kernel): d/dq block, d/dqd block
    "fd_grad_q":        ("AlgFdGradQ",      3, 0, 0, "df_du", lambda n: n * n),
    "fd_grad_qd":       ("AlgFdGradQd",     3, 0, 0, "df_du", lambda n: n * n),
    # consumers fused after the FD gradient (algorithms.trace_fd_consumer); in1 = lam (2n words)
    "fd_vjp":           ("AlgFdVjp",        3, 2, 0, "fd_vjp", lambda n: 5 * n),
    "fd_lin":           ("AlgFdLin",        3, 0, 0, "fd_lin", lambda n: 2 * n + 3 * n * n),
    # further algorithms: mass matrix by CRBA (both triangles), forward dynamics by ABA (no Minv)
    "crba":             ("AlgCrba",         1, 0, 0, "M",     lambda n: n * n),
    "aba":              ("AlgAba",          3, 0, 0, "qdd",   lambda n: n),
}

_NAME_RE = re.compile(r"^(qdd|qd|q|u|Minv|lam)(\d+)$")
_SCALARS = ("gravity", "dt")


def _flit(x: float) -> str:
    s = "%.9g" % x
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


def emit_eval(p: Program, n: int, in0_words: int, in1_words: int, out_name: str, out_words: int,
              indent: str = "        ", sync_every: int = 0, col_flush: bool = False) -> Tuple[List[str], Dict[str, int]]:
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
        if kind in ("qdd", "lam"):
            return in0_words + idx
        return in0_words + in1_words + idx        # Minv

    for i, k in enumerate(p.nodes):
        if live[i] and k[0] == "in" and k[1] not in _SCALARS:
            lines.append("%sconst float t%d = s_in[%d];   // %s" % (indent, i, word_of(k[1]), k[1]))
        elif live[i] and k[0] == "in":
            lines.append("%sconst float t%d = %s;" % (indent, i, k[1]))
    lines.append(indent + "__syncwarp();")
    parks = getattr(p, "parks", {})

    def store_stmt(idx, expr):
        if not col_flush:
            return "%ss_out[%d] = %s;" % (indent, idx, expr)
        sgrp, rem = divmod(idx, n * n)
        return "%ss_col[%d] = %s;" % (indent, sgrp * n + rem % n, expr)

    # column-pair groups: (idx % n*n) // n ; flushed after their last traced output
    const_by_group: Dict[int, list] = {}
    last_pos: Dict[int, int] = {}
    if col_flush:
        for (_, idx, v) in p.outputs:
            j = (idx % (n * n)) // n
            if v.is_const:
                const_by_group.setdefault(j, []).append((idx, v))
            else:
                last_pos[j] = max(last_pos.get(j, -1), v.i)
        flush_at: Dict[int, list] = {}
        for j, pos in last_pos.items():
            flush_at.setdefault(pos, []).append(j)
    else:
        for idx, v in outs_by_pos.get(-1, []):
            body.append(store_stmt(idx, _flit(v.c)))
    # packed-pair peephole: a pair product with a single use that is a pair addition becomes FFMA2
    uses = [0] * len(p.nodes)
    for i in range(len(p.nodes)):
        if live[i]:
            for o in p.operands(i):
                uses[o] += 1
    for (_, _, v) in p.outputs:
        if not v.is_const:
            uses[v.i] += 1
    fused_into: Dict[int, int] = {}          # product node -> the addition that absorbs it
    for i, k in enumerate(p.nodes):
        if live[i] and k[0] == "add2":
            for cand in (k[2], k[1]):
                kc = p.nodes[cand]
                if kc[0] in ("mul2", "mul2c") and uses[cand] == 1 and cand not in fused_into:
                    fused_into[cand] = i
                    break

    def pair_scalar(kc, negate):
        """Second operand of a pair product as a broadcast float2 expression."""
        if kc[0] == "mul2":
            return "make_float2(%st%d, %st%d)" % (("-" if negate else ""), kc[2], ("-" if negate else ""), kc[2])
        c = _flit(-kc[2] if negate else kc[2])
        return "make_float2(%s, %s)" % (c, c)

    def neg_pair(idx, negate):
        return "make_float2(-p%d.x, -p%d.y)" % (idx, idx) if negate else "p%d" % idx

    def half_expr(h):
        return _flit(h[1]) if h[0] == "c" else "%st%d" % ("-" if h[2] < 0 else "", h[1])

    emitted = 0
    for i, k in enumerate(p.nodes):
        if not live[i]:
            continue
        op = k[0]
        emitted += 1
        if sync_every and emitted % sync_every == 0:
            body.append(indent + "__syncthreads();")
        if op == "in":
            pass                                   # loaded above (scalars included)
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
        elif op == "pk":
            body.append("%sconst float2 p%d = make_float2(%s, %s);" % (indent, i, half_expr(k[1]), half_expr(k[2])))
        elif op == "half":
            body.append("%sconst float t%d = p%d.%s;" % (indent, i, k[1], "xy"[k[2]]))
        elif op in ("mul2", "mul2c"):
            if i not in fused_into:
                body.append("%sconst float2 p%d = __fmul2_rn(p%d, %s);" % (indent, i, k[1], pair_scalar(k, False)))
        elif op == "add2":
            lo, hi, rel = k[1], k[2], k[3]
            if fused_into.get(hi) == i:          # lo + rel * (P * s)
                kc = p.nodes[hi]
                body.append("%sconst float2 p%d = __ffma2_rn(p%d, %s, p%d);" % (indent, i, kc[1], pair_scalar(kc, rel < 0), lo))
            elif fused_into.get(lo) == i:        # (P * s) + rel * hi
                kc = p.nodes[lo]
                body.append("%sconst float2 p%d = __ffma2_rn(p%d, %s, %s);" % (indent, i, kc[1], pair_scalar(kc, False),
                                                                            neg_pair(hi, rel < 0)))
            else:
                body.append("%sconst float2 p%d = __fadd2_rn(p%d, %s);" % (indent, i, lo, neg_pair(hi, rel < 0)))
        elif op == "add2k":
            body.append("%sconst float2 p%d = __fadd2_rn(p%d, make_float2(%s, %s));" % (indent, i, k[1], _flit(k[2]), _flit(k[3])))
        elif op == "ld":
            body.append("%sconst float t%d = s_park[%d];" % (indent, i, 32 * k[1]))
        if i in parks:
            body.append("%ss_park[%d] = t%d;" % (indent, 32 * parks[i], i))
        for idx, v in outs_by_pos.get(i, []):
            body.append(store_stmt(idx, "%st%d" % ("-" if v.s < 0 else "", v.i)))
        if col_flush and i in flush_at:
            for j in sorted(flush_at[i]):
                for idx, v in const_by_group.get(j, []):
                    body.append(store_stmt(idx, _flit(v.c)))
                body.append("%s__syncwarp();" % indent)
                body.append("%sflush_colpair<%d>(g_tile, s_warp, %d, cnt, lane);" % (indent, n, j))
                body.append("%s__syncwarp();" % indent)
    return lines + body, p.op_counts()


def emit_col_struct(robot: Robot, name: str, alg: str, use_qdd: bool = False) -> Tuple[str, Dict[str, int]]:
    """Lane-uniform column program (algorithms.trace_column_program) as a struct for csrc/grid_cps.cuh:
    inputs straight from global memory (all lanes of a state read the same row: broadcast), masks
    from the lane's column index, each lane stores its own column."""
    from .algorithms import trace_column_program
    n = robot.n
    p = trace_column_program(robot, alg, use_qdd)
    live = p.live_nodes()
    in0 = 3 * n if alg == "fd_grad" else 2 * n
    in1 = n if use_qdd else 0
    L = ["struct %s {" % name,
         "    static constexpr int IN0 = %d, IN1 = %d, COLS = %d, ROWS = %d;" % (in0, max(in1, 1), 2 * n, n),
         "    static __device__ __forceinline__ void eval(const float *__restrict__ g_in, const float *__restrict__ g_in1,"
         " float *__restrict__ g_out, const int col, const bool wr, const float gravity) {"]
    ind = "        "
    outs: Dict[int, list] = {}
    for (_, idx, v) in p.outputs:
        outs.setdefault(-1 if v.is_const else v.i, []).append((idx, v))
    for idx, v in outs.get(-1, []):
        L.append("%sif (wr) g_out[%d] = %s;" % (ind, idx, _flit(v.c)))
    done = set()
    for i, k in enumerate(p.nodes):
        if not live[i]:
            continue
        op = k[0]
        if op == "in":
            nm = k[1]
            if nm == "gravity":
                L.append("%sconst float t%d = gravity;" % (ind, i))
            elif nm.startswith("mqd"):
                L.append("%sconst float t%d = (col == %d) ? 1.0f : 0.0f;" % (ind, i, n + int(nm[3:])))
            elif nm.startswith("mq"):
                L.append("%sconst float t%d = (col == %d) ? 1.0f : 0.0f;" % (ind, i, int(nm[2:])))
            elif nm.startswith("qdd"):
                L.append("%sconst float t%d = __ldg(g_in1 + %d);" % (ind, i, int(nm[3:])))
            else:
                m = _NAME_RE.match(nm)
                kind, idx = m.group(1), int(m.group(2))
                L.append("%sconst float t%d = __ldg(g_in + %d);" % (ind, i, {"q": 0, "qd": n, "u": 2 * n}[kind] + idx))
        elif op in ("sin", "cos"):
            a = k[1]
            if a not in done:
                done.add(a)
                si, ci = p._cse.get(("sin", a)), p._cse.get(("cos", a))
                sn = "t%d" % si if si is not None and live[si] else "us%d" % a
                cn = "t%d" % ci if ci is not None and live[ci] else "uc%d" % a
                L.append("%sfloat %s, %s; sincosf(t%d, &%s, &%s);" % (ind, sn, cn, a, sn, cn))
        elif op == "rcp":
            L.append("%sconst float t%d = 1.0f / t%d;" % (ind, i, k[1]))
        elif op == "mul":
            L.append("%sconst float t%d = t%d * t%d;" % (ind, i, k[1], k[2]))
        elif op == "mulc":
            L.append("%sconst float t%d = t%d * %s;" % (ind, i, k[1], _flit(k[2])))
        elif op == "add":
            L.append("%sconst float t%d = t%d %s t%d;" % (ind, i, k[1], "+" if k[3] > 0 else "-", k[2]))
        elif op == "addc":
            L.append("%sconst float t%d = t%d + %s;" % (ind, i, k[1], _flit(k[2])))
        for idx, v in outs.get(i, []):
            L.append("%sif (wr) g_out[%d] = %st%d;" % (ind, idx, "-" if v.s < 0 else "", v.i))
    L += ["    }", "};", ""]
    return "\n".join(L), p.op_counts()


def emit_alg_struct_looped(robot: Robot, sname: str, alg: str, use_qdd: bool = False,
                           pairs: bool = False) -> Tuple[str, Dict[str, int]]:
    """Thread-per-state program with the du-columns ROLLED into a loop: the column-independent
    part of algorithms.trace_column_program runs once, then one lane-uniform column body runs
    2n times with 0/1 masks derived from the loop counter.  The body (~1k instructions) stays in
    the instruction cache, which the fully unrolled program (8.5k instructions, 136 KB) does not
    (profiles/r1b: no_instruction = 58 % of stall cycles)."""
    from .algorithms import trace_column_program
    n = robot.n
    p = trace_column_program(robot, alg, use_qdd)
    live = p.live_nodes()
    dep = [False] * len(p.nodes)
    for i, k in enumerate(p.nodes):
        if k[0] == "in":
            dep[i] = k[1].startswith("mq")
        elif k[0] in ("sin", "cos", "rcp", "mulc", "addc"):
            dep[i] = dep[k[1]]
        elif k[0] in ("mul", "add"):
            dep[i] = dep[k[1]] or dep[k[2]]
    in0 = 3 * n if alg == "fd_grad" else 2 * n
    in1 = n if use_qdd else 0
    out_words = 2 * n * n
    ind = "        "
    pre: List[str] = []
    body: List[str] = []
    outs: Dict[int, list] = {}
    for (_, idx, v) in p.outputs:
        outs.setdefault(-1 if v.is_const else v.i, []).append((idx, v))
    done = set()
    for i, k in enumerate(p.nodes):
        if live[i] and k[0] == "in" and not dep[i] and k[1] != "gravity":
            m = _NAME_RE.match(k[1])
            kind, idx = m.group(1), int(m.group(2))
            word = {"q": idx, "qd": n + idx, "u": 2 * n + idx, "qdd": in0 + idx}[kind]
            pre.append("%sconst float t%d = s_in[%d];   // %s" % (ind, i, word, k[1]))
    pre.append(ind + "__syncwarp();")
    for i, k in enumerate(p.nodes):
        if not live[i]:
            continue
        op = k[0]
        tgt = body if dep[i] else pre
        pad = ind + ("    " if dep[i] else "")
        if op == "in":
            nm = k[1]
            if nm == "gravity":
                tgt.append("%sconst float t%d = gravity;" % (pad, i))
            elif nm.startswith("mqd"):
                tgt.append("%sconst float t%d = (col == %d) ? 1.0f : 0.0f;" % (pad, i, n + int(nm[3:])))
            elif nm.startswith("mq"):
                tgt.append("%sconst float t%d = (col == %d) ? 1.0f : 0.0f;" % (pad, i, int(nm[2:])))
        elif op in ("sin", "cos"):
            a = k[1]
            if a not in done:
                done.add(a)
                si, ci = p._cse.get(("sin", a)), p._cse.get(("cos", a))
                sn = "t%d" % si if si is not None and live[si] else "us%d" % a
                cn = "t%d" % ci if ci is not None and live[ci] else "uc%d" % a
                tgt.append("%sfloat %s, %s; sincosf(t%d, &%s, &%s);" % (pad, sn, cn, a, sn, cn))
        elif op == "rcp":
            tgt.append("%sconst float t%d = 1.0f / t%d;" % (pad, i, k[1]))
        elif op == "mul":
            tgt.append("%sconst float t%d = t%d * t%d;" % (pad, i, k[1], k[2]))
        elif op == "mulc":
            tgt.append("%sconst float t%d = t%d * %s;" % (pad, i, k[1], _flit(k[2])))
        elif op == "add":
            tgt.append("%sconst float t%d = t%d %s t%d;" % (pad, i, k[1], "+" if k[3] > 0 else "-", k[2]))
        elif op == "addc":
            tgt.append("%sconst float t%d = t%d + %s;" % (pad, i, k[1], _flit(k[2])))
        for idx, v in outs.get(i, []):
            body.append("%s    s_out[col * %d + %d] = %st%d;" % (ind, n, idx, "-" if v.s < 0 else "", v.i))
    if -1 in outs:
        raise ValueError("column program produced a constant output")
    cnt = p.op_counts()
    n_body = sum(1 for i, k in enumerate(p.nodes) if live[i] and dep[i] and k[0] in ("mul", "mulc", "add", "addc"))
    cnt["loop_body_flops"] = n_body
    cnt["flops"] = cnt["flops"] - n_body + 2 * n * n_body
    txt = ["struct %s {" % sname,
           "    static constexpr int IN0 = %d, IN1 = %d, IN2 = 0, OUT = %d;" % (in0, in1, out_words),
           "    static constexpr long long TRACED_FLOPS = %d;   // %d once + %d columns x %d" % (
               cnt["flops"], cnt["flops"] - 2 * n * n_body, 2 * n, n_body),
           "    static __device__ __forceinline__ void eval(const float *s_in, float *s_out, const float gravity,"
           " const float dt) {"]
    txt += pre
    txt.append(ind + "#pragma unroll 1")
    txt.append(ind + "for (int col = 0; col < %d; ++col) {" % (2 * n))
    txt += body
    txt.append(ind + "}")
    txt += ["    }", "};", ""]
    return "\n".join(txt), cnt


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
           "    static __device__ __forceinline__ void eval(const float *s_in, float *s_out, const float gravity,"
           " const float dt) {"]
    txt += body
    txt += ["    }", "};", ""]
    return "\n".join(txt), cnt


def _alloc_f_slots(robot: Robot):
    """Slot indices for the F columns of the Minv passes (csrc/grid_wps.cuh): a joint's F
    lives from the step its first child writes it (backward) / it is produced (forward)
    until its last reader is done, so only joints on the current root path hold a slot."""
    n = robot.n
    children = [[c for c in range(n) if robot.parent[c] == i] for i in range(n)]
    slot_b, live, nslot = [-1] * n, set(), 0

    def take():
        nonlocal nslot
        k = 0
        while k in live:
            k += 1
        live.add(k)
        nslot = max(nslot, k + 1)
        return k

    for i in range(n - 1, -1, -1):
        par = robot.parent[i]
        if par >= 0 and slot_b[par] < 0:
            slot_b[par] = take()
        if slot_b[i] >= 0:
            live.discard(slot_b[i])
    slot_f, live = [-1] * n, set()
    for i in range(n):
        par = robot.parent[i]
        if children[i]:
            slot_f[i] = take()
        if par >= 0 and children[par][-1] == i:
            live.discard(slot_f[par])
    return slot_b, slot_f, max(nslot, 1)


def wps_layout(robot: Robot) -> Dict[str, object]:
    """Generation-time tables and sizes of the wide (CTA-per-state) kernels."""
    n = robot.n
    level = [robot.get_bfs_level_by_id(i) for i in range(n)]
    nsub = [len(robot.get_subtree_by_id(i)) for i in range(n)]
    nchild = [sum(1 for c in range(n) if robot.parent[c] == i) for i in range(n)]
    slot_b, slot_f, nslot = _alloc_f_slots(robot)
    save = [-1] * n
    for i in range(n):
        if nchild[i] >= 2:
            save[i] = sum(1 for a in robot.get_ancestors_by_id(i) if nchild[a] >= 2)
    nsave = max([x + 1 for x in save] + [1])
    dfbase, acc = [], 0
    for i in range(n):
        dfbase.append(acc)
        acc += 6 * (level[i] + 1)
    df_words = acc
    minv_col_warps = (n + 31) // 32
    col_warps = (2 * n + 31) // 32
    nt = 32 * max(col_warps, minv_col_warps + 2)
    # mirror of struct L in csrc/grid_wps.cuh
    Er = (5 * n + 3) // 4 * 4
    IA = Er + 12 * n + 30 * n + 36 * n + n * n
    minv_end = IA + 36 * n + 13 * n + nslot * 6 * n
    grad_end = IA + 2 * df_words + nsave * 24 * n + 2 * n * n
    total = (max(minv_end, grad_end) + 3) // 4 * 4
    nlevels = max(level) + 1
    level_joints = sorted(range(n), key=lambda i: (level[i], i))
    level_start = [sum(1 for i in range(n) if level[i] < l) for l in range(nlevels + 1)]
    return dict(N=n, NT=nt, NSLOT=nslot, NSAVE=nsave, DF_WORDS=df_words, IA_LANE0=32 * minv_col_warps,
                NLEVELS=nlevels, level_joints=level_joints, level_start=level_start,
                RNEA_TID=nt - 32, COL_WARPS=col_warps, level=level, nsub=nsub, slot_b=slot_b, slot_f=slot_f,
                save=save, dfbase=dfbase, smem_bytes=4 * total)


def emit_wps_tables(robot: Robot, lay: Dict[str, object], include: bool = True) -> str:
    n = robot.n

    def ints(name, vals):
        return "__constant__ int %s[%d] = {%s};\n" % (name, len(vals), ", ".join(str(int(v)) for v in vals))

    def floats(name, vals):
        return "__constant__ float %s[%d] = {%s};\n" % (name, len(vals), ", ".join(_flit(float(v)) for v in vals))

    t = ["namespace GRID_NS { namespace gen {\n", "struct WT {\n"]
    for k in ("N", "NT", "NSLOT", "NSAVE", "DF_WORDS", "IA_LANE0", "RNEA_TID", "COL_WARPS", "NLEVELS"):
        t.append("    static constexpr int %s = %d;\n" % (k, lay[k]))
    t.append("    static constexpr bool TC_MATMUL = %s;\n" % ("true" if lay.get("TC_MATMUL") else "false"))
    t.append("};\n")
    t.append(ints("wt_parent", robot.parent))
    t.append(ints("wt_S", robot.S_ind))
    t.append(ints("wt_nsub", lay["nsub"]))
    t.append(ints("wt_level", lay["level"]))
    t.append(ints("wt_fslot_b", lay["slot_b"]))
    t.append(ints("wt_fslot_f", lay["slot_f"]))
    t.append(ints("wt_saveslot", lay["save"]))
    t.append(ints("wt_dfbase", lay["dfbase"]))
    t.append(ints("wt_level_start", lay["level_start"]))
    t.append(ints("wt_level_joints", lay["level_joints"]))
    t.append(floats("wt_E0", [x for i in range(n) for x in robot.E0[i].flatten()]))
    t.append(floats("wt_r0", [x for i in range(n) for x in robot.r0[i]]))
    t.append(floats("wt_I", [x for i in range(n) for x in robot.Imats[i].flatten()]))
    t.append(floats("wt_I_g", [x for i in range(n) for x in robot.Imats[i].flatten()]).replace("__constant__", "__device__ const"))
    t.append(floats("wt_damping", robot.damping))
    t.append("}}  // namespace GRID_NS::gen\n")
    if include:
        t.append('#include "grid_wps.cuh"\n')
    return "".join(t)


def emit_alg_struct_v2(robot: Robot, variant: str, park=()) -> Tuple[str, Dict[str, int]]:
    """Gradient program for csrc/grid_tps.cuh tps2_kernel: paired (float2) columns, per-column flush,
    optional parking of per-joint data in lane-private shared memory."""
    from .algorithms import trace_fd_grad_paired, trace_id_grad_paired
    n = robot.n
    sname, m0, m1, m2, out_name, out_fn = VARIANTS[variant]
    in0, in1, in2, out = m0 * n, m1 * n, m2 * n * n, out_fn(n)
    if variant.startswith("fd_grad"):
        p = trace_fd_grad_paired(robot, variant == "fd_grad_qdd_minv", park)
    else:
        p = trace_id_grad_paired(robot, variant == "id_grad_qdd", [x for x in park if x != "Minv"])
    body, cnt = emit_eval(p, n, in0, in1, out_name, out, col_flush=True)
    slots = len(getattr(p, "parks", {}))
    cnt["park_slots"] = slots
    txt = ["struct %s {" % sname,
           "    static constexpr int IN0 = %d, IN1 = %d, IN2 = %d, OUT = %d, NJ = %d, PARK_SLOTS = %d;" % (
               in0, in1, in2, out, n, slots),
           "    static constexpr long long TRACED_FLOPS = %d;   // %d mul + %d add per state, %d packed instructions" % (
               cnt["flops"], cnt["mul"], cnt["add"], cnt.get("packed", 0)),
           "    static __device__ __forceinline__ void eval(const float *s_in, float *s_col, float *s_park,"
           " float *__restrict__ g_tile, const int cnt, const int lane, const float *s_warp, const float gravity) {"]
    txt += body
    txt += ["    }", "};", ""]
    return "\n".join(txt), cnt


class KernelPlan:
    """Which kernel family serves each algorithm of a robot, and its launch shape."""

    def __init__(self, robot: Robot, tps_max_flops: int = 60000, tps_warps: int = 1,
                 tps_min_blocks: Optional[Dict[str, int]] = None, tps_sync_every: int = 0,
                 wps_max_states: int = 0, cps_max_states: int = 2048, tps_loop_columns: bool = False,
                 tps_pairs: bool = False, tps_v2_park=None, pipe_algs=None, pipe_min_states: int = 0,
                 pipe_opts: Optional[Dict[str, int]] = None, pipe_min_blocks: Tuple[int, int] = (1, 1),
                 pipe_warps: int = 8, pipe_sync_every: int = 256, pipe_scratch_lead: int = 160,
                 wps_tc_matmul: bool = False, only_algs=None, pipe_small_states: int = 24576,
                 pipe_small_group_flops: int = 4000, lps_min_states: int = 256, lps_force: bool = False,
                 tps_half: bool = False, pipe_x2: bool = False, extras_max_flops: int = 26000,
                 fd_via_aba: Optional[bool] = None):
        # every constructor argument except the robot: build.py hashes this into the library name, so
        # a library built with one plan is never returned for another
        self._args = {k: v for k, v in locals().items() if k not in ("self", "robot")}
        self.robot = robot
        # -Minv dc_du of the wide FD-gradient kernel on the tensor cores (3xTF32 mma.sync): measured in
        # profiles/r2_tc_*; off unless it wins end to end
        self.wps_tc_matmul = bool(wps_tc_matmul) and robot.n % 16 == 0
        # experiments: build only these algorithms (the others report "none"); None = all
        self.only_algs = None if only_algs is None else tuple(only_algs)
        self.tps_warps = tps_warps
        self.tps_sync_every = tps_sync_every if tps_warps > 1 else 0
        self.tps_loop_columns = tps_loop_columns
        self.tps_pairs = tps_pairs
        self.tps_v2_park = tps_v2_park          # None = variant 1; tuple of parked groups = variant 2
        alg = algorithmic_flops(robot)
        self.kind: Dict[str, str] = {}
        self.wps = wps_layout(robot)
        self.wps["TC_MATMUL"] = self.wps_tc_matmul
        self.wps_ok = self.wps["smem_bytes"] <= 227 * 1024
        # batches up to this many states go to the wide kernel when both exist (latency mode)
        self.wps_max_states = wps_max_states
        for a in ("id", "minv", "fd", "id_grad", "fd_grad"):
            # the dense reference count over-states traced work by ~4x; gate on it
            tps = alg[a] <= tps_max_flops
            wps = self.wps_ok and a != "id"
            if self.only_algs is not None and a not in self.only_algs:
                tps = wps = False
            self.kind[a] = "tps+wps" if tps and wps else "tps" if tps else "wps" if wps else "none"
            # latency kernels: gradient algorithms of robots whose 2n columns fit one warp
            if tps and a in ("id_grad", "fd_grad") and 2 * robot.n <= 32:
                self.kind[a] += "+cps"
        # phase-split kernels (pipeline.py): by default for every algorithm that has no
        # thread-per-state program; `pipe_algs` forces a set (experiments on small robots)
        from .pipeline import PipeVariant, components
        self.pipe_opts = dict(pipe_opts or {})
        self.pipe_min_states = pipe_min_states
        self.pipe_min_blocks = tuple(pipe_min_blocks)
        self.pipe_warps = pipe_warps
        self.pipe_sync_every = pipe_sync_every
        self.pipe_scratch_lead = pipe_scratch_lead
        # stage-1 (column) programs of the two-stage variants with two states per lane on the packed FP32
        # instructions (pipeline.emit_task x2)
        self.pipe_x2 = pipe_x2 if isinstance(pipe_x2, dict) else bool(pipe_x2)
        self.pipe: Dict[str, "PipeVariant"] = {}
        if pipe_algs is None:
            # a forest of several trees: one thread per (state, tree) beats one thread per state
            # (HyQ FD-gradient 2.25x, profiles/r1_matrix_pipe.jsonl); otherwise only where no
            # thread-per-state program exists
            forest = len(components(robot)) > 1
            want = [a for a in self.kind if (forest and a != "id") or "tps" not in self.kind[a]]
            if self.only_algs is not None:
                want = [a for a in want if a in self.only_algs]
        else:
            want = list(pipe_algs)
        variants = {"id": ("id", "id_qdd"), "minv": ("minv",), "fd": ("fd",), "id_grad": ("id_grad", "id_grad_qdd"),
                    "fd_grad": ("fd_grad", "fd_grad_qdd_minv")}
        for a in want:
            pvs = []
            for v in variants[a]:                 # stop at the first variant that does not fit (long chains: each
                pvs.append(PipeVariant(robot, v, **self.pipe_opts))      # rejected candidate costs a full trace)
                if not pvs[-1].feasible:
                    break
            if all(pv.feasible for pv in pvs):
                for pv in pvs:
                    self.pipe[pv.variant] = pv
                self.kind[a] = "pipe" if self.kind[a] == "none" else self.kind[a] + "+pipe"
        # Serial chains whose Minv / FD / gradients have no thread-per-state program (64-link chain): lane-per-state
        # kernels with rolled joint loops (csrc/grid_lps.cuh) for batches of lps_min_states and more; the
        # CTA-per-state wide kernels keep the small batches and the USE_QDD_MINV_FLAG overload
        self.lps_min_states = lps_min_states
        self.lps = set()
        if robot.is_serial_chain() and self.wps_ok:
            for a in ("minv", "fd", "id_grad", "fd_grad"):
                # lps_force: also for small chains that have thread-per-state programs (test builds: the chain
                # kernels run on iiwa14 and on a chain with prismatic joints when GRID_FORCE_KERNEL=lps)
                if "wps" in self.kind[a] and (lps_force or ("tps" not in self.kind[a] and "pipe" not in self.kind[a])):
                    self.lps.add(a)
                    self.kind[a] += "+lps"
        self.lps_forced_only = bool(lps_force)
        # Second set of FD-gradient column programs for small and mid-size batches (the per-GPU shard when 65 536
        # states are split over 4-8 GPUs): groups of ~4 k flops instead of 6.5 k give more (task, tiles) items to
        # spread over the SMs - Atlas 8 192 states 128 vs 152 us, 16 384 states 221 vs 242 us, but 779 vs 737 us at
        # 65 536 (profiles/r2_atlas_pipe_variants.jsonl).  Only with the default grouping options.
        self.pipe_small_states = pipe_small_states
        self.pipe_small = None
        if ("fd_grad" in self.pipe and self.pipe["fd_grad"].scratch_words > 0 and not self.pipe_opts
                and pipe_small_states > 0):
            pv = PipeVariant(robot, "fd_grad", group_flops=pipe_small_group_flops, struct_suffix="Small")
            if pv.feasible and len(pv.tasks) > len(self.pipe["fd_grad"].tasks):
                self.pipe_small = pv
        # EXPERIMENT (off; measured slower): mid-size batches of a thread-per-state FD gradient (between the latency
        # kernels and one tile per resident warp) with a warp per (tile, half) - the d/dq or the d/dqd block, each with
        # its own copy of the column-independent part: critical path 0.61x (6 490 / 5 906 vs 10 606 traced flops), twice
        # the warps in flight.  iiwa14, 4 096 ... 18 944 states: 22.9-24.9 us vs 20.8 us for the whole program per
        # thread (profiles/r2_exp_tps_half_split.jsonl).  The ~20 us of a single pass is not the dependent chain of one
        # warp but the cold stream of the program's code through every SM (139 KB at ~3.5 B/cycle); two half programs
        # are 1.17x the code.
        self.tps_half = (bool(tps_half) and "tps" in self.kind["fd_grad"] and "pipe" not in self.kind["fd_grad"]
                         and plan_default(tps_warps, tps_v2_park, tps_loop_columns, tps_pairs))
        # consumers fused after the FD gradient ride on the family that serves fd_grad at large batches
        self.consumers: Dict[str, str] = {}
        for c in ("fd_vjp", "fd_lin"):
            fams = []
            if self.only_algs is not None and c not in self.only_algs:
                self.consumers[c] = "none"
                continue
            if "tps" in self.kind["fd_grad"]:
                fams.append("tps")
            if "pipe" in self.kind["fd_grad"]:
                pv = PipeVariant(robot, c, **self.pipe_opts)
                if pv.feasible:
                    self.pipe[c] = pv
                    fams.append("pipe")
            if "fd_grad" in self.lps:
                fams.append("lps")
            self.consumers[c] = "+".join(fams) if fams else "none"
        # Further algorithms, thread-per-state programs: the mass matrix by the composite-rigid-body
        # algorithm and forward dynamics by the articulated-body algorithm (O(n), no Minv).  Gate: traced flops
        # (a program beyond ~26 k operations is a megabyte of straight-line code: the 64-link chain's CRBA is out,
        # its ABA is in).
        from .algorithms import TRACERS as _T
        self.extras: Dict[str, str] = {}
        self.extra_programs: Dict[str, Program] = {}
        for x in ("crba", "aba"):
            if self.only_algs is not None and x not in self.only_algs:
                self.extras[x] = "none"
                continue
            prog = _T[x](robot)
            fl = prog.op_counts()["flops"]
            self.extras[x] = "tps" if fl <= extras_max_flops else "none"
            if self.extras[x] == "tps":
                self.extra_programs[x] = prog
        # forward_dynamics served by the ABA program: the same qdd through fewer operations (iiwa14 1 691 vs 2 331
        # traced flops, Atlas 8 758 in ONE thread vs 15 784 over the phase-split Minv + RNEA, 64-link chain 22 879 vs the
        # rolled chain kernels).  Measured (profiles/r2_aba_crba_timings.jsonl, 65 536 states): iiwa14 10.8 vs 12.7 us,
        # HyQ 12.8 vs 23.6 us, Atlas 59.9 vs 79.3 us, chain 422 vs 805 us; at small batches everything sits on the
        # launch floor, so robots that have a thread-per-state FD switch only from 32 768 states on, the others from 256.
        if fd_via_aba is None:
            fd_via_aba = True
        self.fd_via_aba = bool(fd_via_aba) and self.extras.get("aba") == "tps"
        self.fd_aba_min_states = 32768 if "tps" in self.kind["fd"] else 256
        # largest batch the phase-split kernels take when a thread-per-state program exists as well (see
        # generate_translation_unit.body): Minv / FD up to 32 768 states, gradients while the output fits ~96 MB of L2
        self.pipe_max_states = {"minv": 32768, "fd": 32768, "id_grad": (96 << 20) // (8 * robot.n * robot.n),
                                "fd_grad": (96 << 20) // (8 * robot.n * robot.n)}
        self.cps_lanes = 16 if 2 * robot.n <= 16 else 32
        # where phase-split kernels exist they are at least as fast as the latency kernels at every batch
        # size once their CTA size follows the batch (HyQ FD gradient N = 128: 8.6 vs 10.6 us, N = 512: 9.5
        # vs 10.6 us, profiles/r1_matrix_final_all_robots.jsonl): no latency kernels for those algorithms
        for a in ("id_grad", "fd_grad"):
            if "pipe" in self.kind[a] and "+cps" in self.kind[a]:
                self.kind[a] = self.kind[a].replace("+cps", "")
        self.cps_max_states = cps_max_states
        # resident single-warp CTAs per SM = register cap 65536/(32*min_blocks).  Measured on B200
        # (profiles/r1_sweep_tps.md): the gradient programs spill at 128/168 registers and run
        # 2.1x faster at 255 registers with no spills; the small programs fit 128.
        self.min_blocks = {"id": 16, "minv": 16, "fd": 16, "id_grad": 8, "fd_grad": 8}
        if tps_min_blocks:
            self.min_blocks.update(tps_min_blocks)


def plan_default(tps_warps, tps_v2_park, tps_loop_columns, tps_pairs) -> bool:
    """The half-split kernel is only generated beside the default thread-per-state programs."""
    return tps_warps == 1 and tps_v2_park is None and not tps_loop_columns and not tps_pairs


def plan_signature(plan: Optional["KernelPlan"]) -> str:
    """'' for the default plan, else a short hash of the plan's constructor arguments."""
    if plan is None:
        return ""
    import hashlib
    default = KernelPlan.__init__.__defaults__
    names = KernelPlan.__init__.__code__.co_varnames[2:2 + len(default)]
    args = {k: plan._args[k] for k in names}
    if all(args[k] == d for k, d in zip(names, default)):
        return ""
    return hashlib.sha256(repr(sorted(args.items(), key=lambda kv: kv[0])).encode()).hexdigest()[:8]


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
    if plan.pipe:
        out.append('#include "grid_pipe.cuh"\n')
    out.append('#define GRID_ROBOT_NAME "%s"\n#define GRID_ROBOT_HASH "%s"\n#define GRID_N %d\n'
               % (robot.name, robot.param_hash(), n))
    out.append("namespace GRID_NS { namespace gen {\n")
    needed = {"id": ("id", "id_qdd"), "minv": ("minv",), "fd": ("fd",), "id_grad": ("id_grad", "id_grad_qdd"),
              "fd_grad": ("fd_grad", "fd_grad_qdd_minv")}
    for a, variants in needed.items():
        if "tps" not in plan.kind[a]:
            continue
        for v in variants:
            if plan.tps_loop_columns and v in ("fd_grad", "id_grad", "id_grad_qdd"):
                txt, cnt = emit_alg_struct_looped(robot, VARIANTS[v][0], "fd_grad" if v == "fd_grad" else "id_grad",
                                                  use_qdd=(v == "id_grad_qdd"))
                out.append(txt)
                stats[v] = cnt
                continue
            if plan.tps_v2_park is not None and v in ("fd_grad", "fd_grad_qdd_minv", "id_grad", "id_grad_qdd"):
                txt, cnt = emit_alg_struct_v2(robot, v, plan.tps_v2_park)
                out.append(txt)
                stats[v] = cnt
                continue
            prog = None
            if plan.tps_pairs:
                from .algorithms import PAIRED_TRACERS
                if v in PAIRED_TRACERS:
                    prog = PAIRED_TRACERS[v](robot)
            txt, cnt = emit_alg_struct(robot, v, p=prog, sync_every=plan.tps_sync_every)
            out.append(txt)
            stats[v] = cnt
    for c, fam in plan.consumers.items():
        if "tps" in fam:
            txt, cnt = emit_alg_struct(robot, c, sync_every=plan.tps_sync_every)
            out.append(txt)
            stats[c] = cnt
    for x, fam in plan.extras.items():
        if fam == "tps":
            txt, cnt = emit_alg_struct(robot, x, p=plan.extra_programs[x], sync_every=plan.tps_sync_every)
            out.append(txt)
            stats[x] = cnt
    if plan.tps_half:
        for v in ("fd_grad_q", "fd_grad_qd"):
            txt, cnt = emit_alg_struct(robot, v)
            out.append(txt)
            stats[v] = cnt
    has_cps = any("cps" in k for k in plan.kind.values())
    if has_cps:
        for nm, alg, uq in (("ColIdGrad", "id_grad", False), ("ColIdGradQdd", "id_grad", True),
                            ("ColFdGrad", "fd_grad", False)):
            txt, cnt = emit_col_struct(robot, nm, alg, uq)
            out.append(txt)
            stats["cps_" + nm] = cnt
    if plan.pipe:
        from .pipeline import emit_pipe_struct
        for v, pv in list(plan.pipe.items()) + ([("fd_grad_small", plan.pipe_small)] if plan.pipe_small else []):
            txt, summ = emit_pipe_struct(pv, plan.pipe_min_blocks, plan.pipe_warps, plan.pipe_sync_every,
                                         plan.pipe_scratch_lead, x2=plan.pipe_x2)
            out.append(txt)
            stats["pipe_" + v] = summ
    out.append("}}  // namespace GRID_NS::gen\n")
    if has_cps:
        out.append('#include "grid_cps.cuh"\n')

    W = plan.tps_warps
    has_tps = lambda a: "tps" in plan.kind[a]
    has_wps = lambda a: "wps" in plan.kind[a]
    if any(has_wps(a) for a in plan.kind):
        out.append(emit_wps_tables(robot, plan.wps))
    if plan.lps:
        out.append('#include "grid_lps.cuh"\n')

    def tps(a, struct):
        if plan.tps_v2_park is not None and a in ("id_grad", "fd_grad"):
            return "tps2_launch<%s, %d, %d>" % (struct, W, plan.min_blocks[a])
        return "tps_launch<%s, %d, %d>" % (struct, W, plan.min_blocks[a])

    def body(a, tps_call, wps_call, cps_call=(), pipe_call=(), lps_call=()):
        """tps_call / wps_call / cps_call: list of (condition or None, expression)."""
        lines = []
        if a in plan.lps:
            lines.append("    if (use_lps(N)) {")
            lines += ["        %sreturn %s;" % ("if (%s) " % c if c else "", e) for c, e in lps_call]
            lines.append("    }")
        if "cps" in plan.kind[a]:
            lines.append("    if (use_cps(N)) {")
            lines += ["        %sreturn %s;" % ("if (%s) " % c if c else "", e) for c, e in cps_call]
            lines.append("    }")
        if "pipe" in plan.kind[a]:
            others = plan.kind[a] != "pipe"
            # where a thread-per-state program exists too, the phase-split kernels keep the batches they win
            # (profiles/r2_hyq_kernel_families.jsonl: HyQ Minv 8.4 vs 10.5 us at 128 states but 24 vs 17 us at 65 536;
            # gradients 44 vs 54 us at 65 536 but 305 vs 177 us at 262 144, where the output no longer fits the L2 and
            # the 48-byte column fragments of four leg programs reach DRAM as partial sectors)
            cap = plan.pipe_max_states.get(a, 0) if has_tps(a) else 0
            lines.append("    if (%s) {" % (("use_pipe(N, %d)" % cap) if others else "true"))
            lines += ["        %sreturn %s;" % ("if (%s) " % c if c else "", e) for c, e in pipe_call]
            lines.append("    }")
        if has_tps(a) and has_wps(a):
            lines.append("    if (use_wide(N)) {")
            lines += ["        %sreturn %s;" % ("if (%s) " % c if c else "", e) for c, e in wps_call]
            lines.append("    }")
            lines += ["    %sreturn %s;" % ("if (%s) " % c if c else "", e) for c, e in tps_call]
        elif has_tps(a):
            lines += ["    %sreturn %s;" % ("if (%s) " % c if c else "", e) for c, e in tps_call]
        elif has_wps(a):
            lines += ["    %sreturn %s;" % ("if (%s) " % c if c else "", e) for c, e in wps_call]
        else:
            lines.append("    return cudaErrorNotSupported;")
        return lines

    L: List[str] = []
    L.append("// Kernel-family choice.  The override comes from options() (environment read once at first use,\n"
             "// grid_set_option() afterwards): no getenv on the launch path.\n"
             "// batches up to WPS_MAX_STATES use the wide kernels when both families exist\n"
             "static bool use_wide(int N) {\n"
             "    const int f = options().force_kernel;\n"
             "    if (f == kWps) return true;\n"
             "    if (f == kTps) return false;\n"
             "    return N <= %d;\n}" % plan.wps_max_states)
    L.append("// small batches go to the lane-per-column latency kernels\n"
             "static bool use_cps(int N) {\n"
             "    const int f = options().force_kernel;\n"
             "    if (f != kAuto) return f == kCps;\n"
             "    return N <= %d;\n}" % plan.cps_max_states)
    L.append("// large batches of robots with phase-split kernels (grid_pipe.cuh)\n"
             "static bool use_pipe(int N, int max_states = 0) {\n"
             "    const int f = options().force_kernel;\n"
             "    if (f != kAuto) return f == kPipe;\n"
             "    return N >= %d && (max_states <= 0 || N <= max_states);\n}" % plan.pipe_min_states)
    L.append("// serial chains without thread-per-state programs: lane-per-state kernels with rolled joint loops\n"
             "static bool use_lps(int N) {\n"
             "    const int f = options().force_kernel;\n"
             "    if (f != kAuto) return f == kLps;\n"
             "    return %s;\n}" % ("false" if plan.lps_forced_only else "N >= %d" % plan.lps_min_states))
    G = plan.cps_lanes
    PL = lambda struct, out, inp, in1: "pipe::pipe_launch<gen::%s>(%s, %s, stride, %s, N, g, s)" % (struct, out, inp, in1)
    L.append("cudaError_t launch_id(float *d_c, const float *d_q_qd, int stride, const float *d_qdd, int N, float g,"
             " cudaStream_t s) {")
    L += body("id", [("d_qdd", "%s(d_c, d_q_qd, stride, d_qdd, nullptr, N, g, s)" % tps("id", "AlgIdQdd")),
                     (None, "%s(d_c, d_q_qd, stride, nullptr, nullptr, N, g, s)" % tps("id", "AlgId"))], [],
              pipe_call=[("d_qdd", PL("PipeIdQdd", "d_c", "d_q_qd", "d_qdd")),
                         (None, PL("PipeId", "d_c", "d_q_qd", "nullptr"))])
    L.append("}")
    L.append("cudaError_t launch_minv(float *d_Minv, const float *d_q, int stride, int N, cudaStream_t s) {")
    L += body("minv", [(None, "%s(d_Minv, d_q, stride, nullptr, nullptr, N, 0.f, s)" % tps("minv", "AlgMinv"))],
              [(None, "wps::wps_launch<0, false>(d_Minv, d_q, stride, nullptr, nullptr, N, 0.f, s)")],
              pipe_call=[(None, "pipe::pipe_launch<gen::PipeMinv>(d_Minv, d_q, stride, nullptr, N, 0.f, s)")],
              lps_call=[(None, "lps::lps_launch<0, false>(d_Minv, d_q, stride, nullptr, N, 0.f, s)")])
    L.append("}")
    xmb = lambda x: 16 if stats[x]["flops"] < 4000 else 8          # register cap of the extra programs: 128 / 255
    L.append("cudaError_t launch_aba(float *d_qdd, const float *d_q_qd_u, int stride, int N, float g, cudaStream_t s) {")
    L.append("    return %s;" % ("tps_launch<AlgAba, %d, %d>(d_qdd, d_q_qd_u, stride, nullptr, nullptr, N, g, s)" % (W, xmb("aba"))
                                if plan.extras["aba"] == "tps" else "cudaErrorNotSupported"))
    L.append("}")
    L.append("cudaError_t launch_crba(float *d_M, const float *d_q, int stride, int N, cudaStream_t s) {")
    L.append("    return %s;" % ("tps_launch<AlgCrba, %d, %d>(d_M, d_q, stride, nullptr, nullptr, N, 0.f, s)" % (W, xmb("crba"))
                                if plan.extras["crba"] == "tps" else "cudaErrorNotSupported"))
    L.append("}")
    L.append("cudaError_t launch_fd(float *d_qdd, const float *d_q_qd_u, int stride, int N, float g, cudaStream_t s) {")
    if plan.fd_via_aba:
        L.append("    if (options().force_kernel == kAuto && N >= %d) return launch_aba(d_qdd, d_q_qd_u, stride, N, g, s);"
                 % plan.fd_aba_min_states)
    L += body("fd", [(None, "%s(d_qdd, d_q_qd_u, stride, nullptr, nullptr, N, g, s)" % tps("fd", "AlgFd"))],
              [(None, "wps::wps_launch<1, false>(d_qdd, d_q_qd_u, stride, nullptr, nullptr, N, g, s)")],
              pipe_call=[(None, PL("PipeFd", "d_qdd", "d_q_qd_u", "nullptr"))],
              lps_call=[(None, "lps::lps_launch<1, false>(d_qdd, d_q_qd_u, stride, nullptr, N, g, s)")])
    L.append("}")
    L.append("cudaError_t launch_id_grad(float *d_dc_du, const float *d_q_qd, int stride, const float *d_qdd, int N,"
             " float g, cudaStream_t s) {")
    L += body("id_grad",
              [("d_qdd", "%s(d_dc_du, d_q_qd, stride, d_qdd, nullptr, N, g, s)" % tps("id_grad", "AlgIdGradQdd")),
               (None, "%s(d_dc_du, d_q_qd, stride, nullptr, nullptr, N, g, s)" % tps("id_grad", "AlgIdGrad"))],
              [("d_qdd", "wps::wps_launch<2, true>(d_dc_du, d_q_qd, stride, d_qdd, nullptr, N, g, s)"),
               (None, "wps::wps_launch<2, false>(d_dc_du, d_q_qd, stride, nullptr, nullptr, N, g, s)")],
              [("d_qdd", "cps_launch<gen::ColIdGradQdd, %d>(d_dc_du, d_q_qd, stride, d_qdd, N, g, s)" % G),
               (None, "cps_launch<gen::ColIdGrad, %d>(d_dc_du, d_q_qd, stride, nullptr, N, g, s)" % G)],
              pipe_call=[("d_qdd", PL("PipeIdGradQdd", "d_dc_du", "d_q_qd", "d_qdd")),
                         (None, PL("PipeIdGrad", "d_dc_du", "d_q_qd", "nullptr"))],
              lps_call=[("d_qdd", "lps::lps_launch<2, true>(d_dc_du, d_q_qd, stride, d_qdd, N, g, s)"),
                        (None, "lps::lps_launch<2, false>(d_dc_du, d_q_qd, stride, nullptr, N, g, s)")])
    L.append("}")
    L.append("cudaError_t launch_fd_grad(float *d_df_du, const float *d_in, int stride, const float *d_qdd,"
             " const float *d_Minv, int N, float g, cudaStream_t s) {")
    half = ([("!d_qdd && options().force_kernel == kAuto && tps_half_fits<AlgFdGradQ, AlgFdGradQd, %d>(N)" % plan.min_blocks["fd_grad"],
              "tps_half_launch<AlgFdGradQ, AlgFdGradQd, %d>(d_df_du, d_in, stride, N, g, s)" % plan.min_blocks["fd_grad"])]
            if plan.tps_half else [])
    L += body("fd_grad",
              [("d_qdd", "%s(d_df_du, d_in, stride, d_qdd, d_Minv, N, g, s)" % tps("fd_grad", "AlgFdGradPre"))] + half +
              [(None, "%s(d_df_du, d_in, stride, nullptr, nullptr, N, g, s)" % tps("fd_grad", "AlgFdGrad"))],
              [("d_qdd", "wps::wps_launch<3, true>(d_df_du, d_in, stride, d_qdd, d_Minv, N, g, s)"),
               (None, "wps::wps_launch<3, false>(d_df_du, d_in, stride, nullptr, nullptr, N, g, s)")],
              [("!d_qdd", "cps_launch<gen::ColFdGrad, %d>(d_df_du, d_in, stride, nullptr, N, g, s)" % G)],
              pipe_call=([("!d_qdd && N <= %d" % plan.pipe_small_states, PL("PipeFdGradSmall", "d_df_du", "d_in", "nullptr"))]
                         if plan.pipe_small else []) + [
                  ("!d_qdd", PL("PipeFdGrad", "d_df_du", "d_in", "nullptr")),
                  ("d_qdd", "pipe::pipe_launch<gen::PipeFdGradPre>(d_df_du, d_in, stride, d_qdd, N, g, s, 0.f, d_Minv)")],
              lps_call=[("d_qdd", "lps::lps_launch<3, true>(d_df_du, d_in, stride, d_qdd, N, g, s, 0.f, d_Minv)"),
                        (None, "lps::lps_launch<3, false>(d_df_du, d_in, stride, nullptr, N, g, s)")])
    L.append("}")

    for c, struct, pstruct, lam in (("fd_vjp", "AlgFdVjp", "PipeFdVjp", "d_lam"), ("fd_lin", "AlgFdLin", "PipeFdLin", "nullptr")):
        fam = plan.consumers[c]
        sig = "float *d_out, const float *d_in, int stride, %sint N, float g, float dt, cudaStream_t s" % (
            "const float *d_lam, " if c == "fd_vjp" else "")
        L.append("cudaError_t launch_%s(%s) {" % (c, sig))
        pipe_call = "pipe::pipe_launch<gen::%s>(d_out, d_in, stride, %s, N, g, s, dt)" % (pstruct, lam)
        tps_call = "tps_launch<%s, %d, %d>(d_out, d_in, stride, %s, nullptr, N, g, s, dt)" % (
            struct, W, plan.min_blocks["fd_grad"], lam)
        if "lps" in fam:
            lps_call = "lps::lps_launch<%d, false>(d_out, d_in, stride, %s, N, g, s, dt)" % (4 if c == "fd_vjp" else 5, lam)
            if fam == "lps":
                L.append("    return %s;" % lps_call)
            else:                                    # test builds of small chains: only when forced
                L.append("    if (options().force_kernel == kLps) return %s;" % lps_call)
            fam = fam.replace("+lps", "")
        if fam == "lps":
            pass
        elif fam == "tps+pipe":
            L.append("    if (use_pipe(N)) return %s;" % pipe_call)
            L.append("    return %s;" % tps_call)
        elif fam == "pipe":
            L.append("    return %s;" % pipe_call)
        elif fam == "tps":
            L.append("    return %s;" % tps_call)
        else:
            L.append("    return cudaErrorNotSupported;")
        L.append("}")

    kinds = "\n".join('    if (!strcmp(alg, "%s")) return "%s";' % (a, k)
                      for a, k in list(plan.kind.items()) + list(plan.consumers.items()) + list(plan.extras.items()))
    if plan.fd_via_aba:                               # reported by grid_kernel_kind("fd@large")
        kinds += '\n    if (!strcmp(alg, "fd@large")) return "tps(aba)";'
    fl = "\n".join('    if (!strcmp(alg, "%s")) return %d;' % (
        a, stats[needed[a][0]]["flops"] if "tps" in plan.kind[a] else plan.pipe[needed[a][0]].flops)
        for a in plan.kind if "tps" in plan.kind[a] or "pipe" in plan.kind[a])
    fl += "\n" + "\n".join('    if (!strcmp(alg, "%s")) return %d;' % (
        c, stats[c]["flops"] if "tps" in fam else plan.pipe[c].flops)
        for c, fam in plan.consumers.items() if "tps" in fam or "pipe" in fam)
    fl += "\n" + "\n".join('    if (!strcmp(alg, "%s")) return %d;' % (x, stats[x]["flops"])
                            for x, fam in plan.extras.items() if fam == "tps")
    out.append("#include <cstring>\n#include <cstdlib>\n")
    out.append(_LAUNCHERS % {"launchers": "\n".join(L), "kinds": kinds, "flops": fl})
    out.append('#include "grid_abi.cuh"\n')
    return "".join(out), stats
