"""Drop-in facade: the reference's Python API surface on top of the B200 generator.

``GRiDCodeGenerator(robot).gen_all_code()`` writes ``<FILE_NAMESPACE>.cuh`` into the current
directory exactly like the reference (GRiDCodeGenerator.py:241-310), and every public
generator keeps its name and meaning (README.md:33-57): ``gen_<alg>_inner_temp_mem_size``,
``gen_<alg>_inner_function_call``, ``gen_<alg>_inner``, ``gen_<alg>_device``,
``gen_<alg>_kernel``, ``gen_<alg>_host``, ``gen_<alg>``; the text-emitter helpers
(``gen_add_code_line`` ...); and the numpy ``test_*`` methods.  What is emitted is different:

  * ``*_inner`` / ``*_device``  keep the reference signatures and memory contract (inputs and
    outputs in shared memory, every thread of the block calls) but run the robot-specialised
    straight-line program traced by algorithms.py; ``s_XImats`` / ``s_temp`` are accepted and
    ignored (no X/I tables, no scratch: ``gen_*_temp_mem_size`` returns 0);
  * ``*_kernel`` keep the reference signatures (``d_out, d_in, stride, [d_qdd, d_Minv,]
    d_robotModel, gravity, NUM_TIMESTEPS``) and work for any launch shape, but map one thread
    to one state and stage tiles through shared memory (csrc/grid_tps.cuh);
  * host functions keep ``<alg><T,FLAGS>(hd_data, d_robotModel, gravity, num_timesteps,
    block_dimms, thread_dimms, streams)`` (+ ``_single_timing``, ``_compute_only``); the launch
    shape is chosen by the library, ``block_dimms`` / ``thread_dimms`` are accepted and ignored.

``use_thread_group=True`` is unfinished in the reference (``cgrps::thread_group tgrp = TBD;``,
SURVEY.md Appendix A): the kwarg is accepted and ignored.  T must be ``float``.
The in-process fast path does not go through this text at all: see runtime.GridEngine.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np

from . import algorithms as A
from .codegen import _flit, emit_alg_struct, emit_wps_tables, KernelPlan, VARIANTS
from .ir import Program
from .robot import Robot

_PKG = os.path.dirname(os.path.abspath(__file__))


def _emit_pointer_eval(p: Program, in_expr, out_expr, indent="    ") -> List[str]:
    """Straight-line CUDA for the live part of ``p`` with inputs/outputs addressed through
    caller-supplied expressions: in_expr(name) -> C expression, out_expr(array, idx) -> lvalue."""
    live = p.live_nodes()
    lines: List[str] = []
    outs: Dict[int, list] = {}
    for (name, idx, v) in p.outputs:
        outs.setdefault(-1 if v.is_const else v.i, []).append((name, idx, v))
    for name, idx, v in outs.get(-1, []):
        lines.append("%s%s = %s;" % (indent, out_expr(name, idx), _flit(v.c)))
    done = set()
    for i, k in enumerate(p.nodes):
        if not live[i]:
            continue
        op = k[0]
        if op == "in":
            lines.append("%sconst float t%d = %s;" % (indent, i, in_expr(k[1])))
        elif op in ("sin", "cos"):
            a = k[1]
            if a not in done:
                done.add(a)
                si, ci = p._cse.get(("sin", a)), p._cse.get(("cos", a))
                sn = "t%d" % si if si is not None and live[si] else "us%d" % a
                cn = "t%d" % ci if ci is not None and live[ci] else "uc%d" % a
                lines.append("%sfloat %s, %s; sincosf(t%d, &%s, &%s);" % (indent, sn, cn, a, sn, cn))
        elif op == "rcp":
            lines.append("%sconst float t%d = 1.0f / t%d;" % (indent, i, k[1]))
        elif op == "mul":
            lines.append("%sconst float t%d = t%d * t%d;" % (indent, i, k[1], k[2]))
        elif op == "mulc":
            lines.append("%sconst float t%d = t%d * %s;" % (indent, i, k[1], _flit(k[2])))
        elif op == "add":
            lines.append("%sconst float t%d = t%d %s t%d;" % (indent, i, k[1], "+" if k[3] > 0 else "-", k[2]))
        elif op == "addc":
            lines.append("%sconst float t%d = t%d + %s;" % (indent, i, k[1], _flit(k[2])))
        for name, idx, v in outs.get(i, []):
            lines.append("%s%s = %st%d;" % (indent, out_expr(name, idx), "-" if v.s < 0 else "", v.i))
    return lines


def _split_name(name: str):
    i = len(name)
    while i > 0 and name[i - 1].isdigit():
        i -= 1
    return name[:i], (int(name[i:]) if i < len(name) else 0)


class GRiDCodeGenerator:
    ALGS = ("inverse_dynamics", "direct_minv", "forward_dynamics", "inverse_dynamics_gradient",
            "forward_dynamics_gradient")

    def __init__(self, robotObj: Robot, DEBUG_MODE=False, NEED_PRINT_MAT=False, USE_DYNAMIC_SHARED_MEM=True,
                 FILE_NAMESPACE="grid"):
        self.robot = robotObj
        self.code_str = ""
        self.indent_level = 0
        self.DEBUG_MODE = DEBUG_MODE
        self.gen_print_mat = DEBUG_MODE or NEED_PRINT_MAT
        self.use_dynamic_shared_mem_flag = USE_DYNAMIC_SHARED_MEM or (self.robot.get_num_pos() > 12)
        self.file_namespace = FILE_NAMESPACE
        self._plan = KernelPlan(self.robot)
        self._impl_ns = "%s_b200_impl" % FILE_NAMESPACE
        # kernel family behind each emitted kernel: straight-line thread-per-state where the traced
        # program is small enough, else the wide CTA-per-state kernels
        self._family = {a: ("tps" if "tps" in k else "wps" if "wps" in k else None) for a, k in self._plan.kind.items()}
        self._unsupported = {a for a, f in self._family.items() if f != "tps"}

    # ------------------------------------------------------------------ text emitter helpers
    def gen_add_code_line(self, new_code_line, add_indent_after=False):
        self.code_str += self.indent_level * "    " + new_code_line + "\n"
        if add_indent_after:
            self.indent_level += 1

    def gen_add_code_lines(self, new_code_lines, add_indent_after=False):
        for line in new_code_lines:
            self.gen_add_code_line(line)
        if add_indent_after:
            self.indent_level += 1

    def gen_add_end_control_flow(self):
        self.indent_level -= 1
        self.gen_add_code_line("}")

    def gen_add_end_function(self):
        self.indent_level -= 1
        self.gen_add_code_line("}\n")

    def gen_add_func_doc(self, func_desc, notes=[], params=[], return_val=None):
        doc = ["/**", " * " + func_desc, " *"]
        if notes:
            doc += [" * Notes:"] + [" *   " + x for x in notes] + [" *"]
        doc += [" * @param " + x for x in params]
        if return_val is not None:
            doc.append(" * @return " + return_val)
        self.gen_add_code_lines(doc + [" */"])

    def gen_add_serial_ops(self, use_thread_group=False):
        self.gen_add_code_line("if(threadIdx.x == 0 && threadIdx.y == 0){", True)

    def gen_add_parallel_loop(self, var_name, max_val, use_thread_group=False, block_level=False):
        if block_level:
            code = "for(int %s = blockIdx.x + blockIdx.y*gridDim.x; %s < %s; %s += gridDim.x*gridDim.y){" % (
                var_name, var_name, max_val, var_name)
        else:
            code = "for(int %s = threadIdx.x + threadIdx.y*blockDim.x; %s < %s; %s += blockDim.x*blockDim.y){" % (
                var_name, var_name, max_val, var_name)
        self.gen_add_code_line(code, True)

    def gen_add_sync(self, use_thread_group=False):
        self.gen_add_code_line("__syncthreads();")

    def gen_static_array_ind_2d(self, col, row, col_stride=6):
        return col_stride * col + row

    def gen_static_array_ind_3d(self, ind, col, row, ind_stride=36, col_stride=6):
        return ind_stride * ind + col_stride * col + row

    def gen_var_in_list(self, var_name, option_list):
        return "(" + " || ".join("(%s == %s)" % (var_name, o) for o in option_list) + ")"

    def gen_var_not_in_list(self, var_name, option_list):
        return "(" + " && ".join("(%s != %s)" % (var_name, o) for o in option_list) + ")"

    def gen_kernel_load_inputs(self, name, stride, amount, use_thread_group=False, name2=None, stride2=1, amount2=1,
                               name3=None, stride3=1, amount3=1):
        self.gen_add_code_line("// load to shared mem")
        for nm, st, am in ((name, stride, amount), (name2, stride2, amount2), (name3, stride3, amount3)):
            if nm is None:
                continue
            self.gen_add_code_line("const T *d_%s_k = &d_%s[k*%s];" % (nm, nm, st))
            self.gen_add_parallel_loop("ind", str(am), use_thread_group)
            self.gen_add_code_line("s_%s[ind] = d_%s_k[ind];" % (nm, nm))
            self.gen_add_end_control_flow()
        self.gen_add_sync(use_thread_group)

    def gen_kernel_load_inputs_single_timing(self, name, amount, use_thread_group=False, name2=None, amount2=1,
                                             name3=None, amount3=1):
        self.gen_add_code_line("// load to shared mem")
        for nm, am in ((name, amount), (name2, amount2), (name3, amount3)):
            if nm is None:
                continue
            self.gen_add_parallel_loop("ind", str(am), use_thread_group)
            self.gen_add_code_line("s_%s[ind] = d_%s[ind];" % (nm, nm))
            self.gen_add_end_control_flow()
        self.gen_add_sync(use_thread_group)

    def gen_kernel_save_result_single_timing(self, store_to_name, amount, use_thread_group=False, load_from_name=None):
        src = load_from_name or ("s_" + store_to_name)
        self.gen_add_code_line("// save down to global")
        self.gen_add_parallel_loop("ind", str(amount), use_thread_group)
        self.gen_add_code_line("d_%s[ind] = %s[ind];" % (store_to_name, src))
        self.gen_add_end_control_flow()
        self.gen_add_sync(use_thread_group)

    def gen_add_multi_threaded_select(self, loop_counter, comparator, counts, select_tuples, USE_NON_BRANCH_ALWAYS=False):
        """if / else-if / else selection of several variables by ranges of a loop counter
        (helpers/_code_generation_helpers.py:81-130): non-branching ternaries for one variable,
        an if-chain otherwise."""
        n = len(counts)
        if len(select_tuples) == 1 or USE_NON_BRANCH_ALWAYS:
            for (typ, name, values) in select_tuples:
                expr = str(values[-1])
                for idx in range(n - 2, -1, -1):
                    expr = "(%s %s %s) ? %s : (%s)" % (loop_counter, comparator, counts[idx], values[idx], expr)
                self.gen_add_code_line("%s %s = %s;" % (typ, name, expr))
            return
        self.gen_add_code_line(" ".join("%s %s;" % (typ, name) for (typ, name, _) in select_tuples))
        for idx in range(n):
            head = ("if (%s %s %s){" % (loop_counter, comparator, counts[idx]) if idx == 0 else
                    "else if (%s %s %s){" % (loop_counter, comparator, counts[idx]) if idx < n - 1 else "else {")
            self.gen_add_code_line(head + " ".join("%s = %s;" % (name, values[idx]) for (_, name, values) in select_tuples) + "}")

    def gen_kernel_save_result(self, store_to_name, stride, amount, use_thread_group=False, load_from_name=None):
        src = load_from_name or ("s_" + store_to_name)
        self.gen_add_code_line("// save down to global")
        self.gen_add_code_line("T *d_%s_k = &d_%s[k*%s];" % (store_to_name, store_to_name, stride))
        self.gen_add_parallel_loop("ind", str(amount), use_thread_group)
        self.gen_add_code_line("d_%s_k[ind] = %s[ind];" % (store_to_name, src))
        self.gen_add_end_control_flow()
        self.gen_add_sync(use_thread_group)

    # ------------------------------------------------------------------ topology helpers (python values)
    def gen_topology_helpers_size(self):
        """Ints in the emitted topology table (fixed layout, see _topology_table).  The traced
        programs do not read it: topology is resolved at trace time."""
        return 6 * self.robot.get_num_pos() + 1

    def gen_topology_sparsity_helpers_python(self, INIT_MODE=False):
        """Same python values as helpers/_topology_helpers.py:193-215."""
        n = self.robot.get_num_pos()
        num_anc = [len(self.robot.get_ancestors_by_id(j)) for j in range(n)]
        num_sub = [len(self.robot.get_subtree_by_id(j)) for j in range(n)]
        run_anc = [sum(num_anc[:j]) for j in range(n + 1)]
        run_sub = [sum(num_sub[:j]) for j in range(n)]
        dva_cols_per_partial = self.robot.get_total_ancestor_count() + n
        dva_cols_per_jid = [num_anc[j] + 1 for j in range(n)]
        running_sum_dva = [run_anc[j] + j for j in range(n)]
        df_cols_per_partial = self.robot.get_total_ancestor_count() + self.robot.get_total_subtree_count()
        df_cols_per_jid = [num_anc[j] + num_sub[j] for j in range(n)]
        running_sum_df = [run_anc[j] + run_sub[j] for j in range(n)]
        df_col_that_is_jid = num_anc
        if INIT_MODE:
            return [str(x) for x in num_anc], [str(x) for x in num_sub], [str(x) for x in run_anc], \
                   [str(x) for x in run_sub]
        return dva_cols_per_partial, dva_cols_per_jid, running_sum_dva, df_cols_per_partial, df_cols_per_jid, \
            running_sum_df, df_col_that_is_jid

    # ------------------------------------------------------------------ file-level pieces
    def gen_add_includes(self, use_thread_group=False):
        self.gen_add_code_lines(["", "#include <assert.h>", "#include <stdio.h>", "#include <stdlib.h>",
                                 "#include <string.h>", "#include <time.h>", "#include <type_traits>",
                                 "#include <cuda_runtime.h>",
                                 "// single kernel timing helper code",
                                 "#define time_delta_us_timespec(start,end) (1e6*static_cast<double>(end.tv_sec - "
                                 "start.tv_sec)+1e-3*static_cast<double>(end.tv_nsec - start.tv_nsec))", ""])

    def gen_add_gpu_err(self):
        self.gen_add_func_doc("Check for runtime errors using the CUDA API", [], [], None)
        self.gen_add_code_lines([
            "#ifndef gpuErrchk",
            "__host__ inline void gpuAssert(cudaError_t code, const char *file, const int line, bool abort=true){",
            "    if (code != cudaSuccess){",
            "        fprintf(stderr,\"GPUassert: %s %s %d\\n\", cudaGetErrorString(code), file, line);",
            "        if (abort){cudaDeviceReset(); exit(code);}",
            "    }",
            "}",
            "#define gpuErrchk(err) {gpuAssert(err, __FILE__, __LINE__);}",
            "#endif", ""])
        if self.gen_print_mat:
            self.gen_add_code_lines([
                "template <typename T, int M, int N>",
                "__host__ __device__ void printMat(const T *A, int lda){",
                "    for(int i=0; i<M; i++){ for(int j=0; j<N; j++){printf(\"%.4f \",A[i + lda*j]);} printf(\"\\n\"); }",
                "}", ""])

    def _shared_counts(self) -> Dict[str, int]:
        """Dynamic shared memory (floats) the emitted kernels need for SUGGESTED_THREADS threads."""
        n = self.robot.get_num_pos()
        warps = max(1, self._suggested_threads() // 32)
        sizes = {"ID": (3 * n, n), "MINV": (n, n * n), "FD": (3 * n, n), "ID_DU": (3 * n, 2 * n * n),
                 "FD_DU": (3 * n + n * n, 2 * n * n)}
        # upper bound of csrc/grid_tps.cuh TpsShape::WARP_WORDS over the variants of each code
        out = {k: warps * ((32 * max(i | 1, o | 1) + 3) // 4 * 4) for k, (i, o) in sizes.items()}
        for code, alg in (("MINV", "minv"), ("FD", "fd"), ("ID_DU", "id_grad"), ("FD_DU", "fd_grad")):
            if self._family[alg] == "wps":
                out[code] = self._plan.wps["smem_bytes"] // 4
        return out

    def _suggested_threads(self) -> int:
        """Block size of the emitted kernels: the wide kernels need exactly WT::NT threads."""
        return self._plan.wps["NT"] if "wps" in self._family.values() else 128

    def gen_add_constants_helpers(self):
        n = self.robot.get_num_pos()
        c = self._shared_counts()
        self.gen_add_code_lines([
            "const int NUM_JOINTS = %d;" % n,
            "const int ID_DYNAMIC_SHARED_MEM_COUNT = %d;" % c["ID"],
            "const int MINV_DYNAMIC_SHARED_MEM_COUNT = %d;" % c["MINV"],
            "const int FD_DYNAMIC_SHARED_MEM_COUNT = %d;" % c["FD"],
            "const int ID_DU_DYNAMIC_SHARED_MEM_COUNT = %d;" % c["ID_DU"],
            "const int FD_DU_DYNAMIC_SHARED_MEM_COUNT = %d;" % c["FD_DU"],
            "const int ID_DU_MAX_SHARED_MEM_COUNT = %d;" % c["ID_DU"],
            "const int FD_DU_MAX_SHARED_MEM_COUNT = %d;" % c["FD_DU"],
            "const int SUGGESTED_THREADS = %d;" % self._suggested_threads(),
            "// Define custom structs",
            "template <typename T>", "struct robotModel {", "    T *d_XImats;", "    int *d_topology_helpers;", "};",
            "template <typename T>", "struct gridData {",
            "    // GPU INPUTS", "    T *d_q_qd_u;", "    T *d_q_qd;", "    T *d_q;",
            "    // CPU INPUTS", "    T *h_q_qd_u;", "    T *h_q_qd;", "    T *h_q;",
            "    // GPU OUTPUTS", "    T *d_c;", "    T *d_Minv;", "    T *d_qdd;", "    T *d_dc_du;", "    T *d_df_du;",
            "    // CPU OUTPUTS", "    T *h_c;", "    T *h_Minv;", "    T *h_qdd;", "    T *h_dc_du;", "    T *h_df_du;",
            "};", ""])

    def gen_spatial_algebra_helpers(self):
        """dot_prod, mx0..mx5 (+_peq/_scaled/_peq_scaled), mxX*, fx, fx_zeroed, fx_times_v(_peq) with the
        reference's names and signatures (helpers/_spatial_algebra_helpers.py:35-256), for user code
        written against them.  The traced programs do not call them: their cross products are folded
        at trace time.  Entries are generated from the definition (v x) = [[w x, 0], [l x, w x]],
        fx = -(v x)^T."""
        L = self.gen_add_code_line
        L("// spatial algebra helpers kept for API compatibility (the traced programs fold these away)")
        L("template <typename T, int N, int S1, int S2, typename P1, typename P2>")
        L("__device__ T dot_prod(const P1 *vec1, const P2 *vec2) {")
        L("    T result = 0;")
        L("    for (int i = 0; i < N; i++) { result += vec1[i*S1] * vec2[i*S2]; }")
        L("    return result;")
        L("}")

        def skew_entry(r, c, base):
            """(a x)[r][c] as (sign, component index) or None, for the 3-vector stored at offset base."""
            if r == c:
                return None
            k = 3 - r - c
            sign = 1 if (c - r) % 3 == 2 else -1      # [[0,-a2,a1],[a2,0,-a0],[-a1,a0,0]]
            return sign, base + k

        def crm_entry(r, c):
            """motion cross matrix (v x)[r][c]"""
            if r < 3 and c < 3:
                return skew_entry(r, c, 0)
            if r >= 3 and c >= 3:
                return skew_entry(r - 3, c - 3, 0)
            if r >= 3 and c < 3:
                return skew_entry(r - 3, c, 3)
            return None

        def term(e, vec):
            return "static_cast<T>(0)" if e is None else "%s%s[%d]" % ("-" if e[0] < 0 else "", vec, e[1])

        for k in range(6):
            col = [crm_entry(r, k) for r in range(6)]
            for peq, scaled in ((False, False), (True, False), (False, True), (True, True)):
                name = "mx%d%s%s" % (k, "_peq" if peq else "", "_scaled" if scaled else "")
                L("template <typename T>")
                L("__device__ void %s(T *s_vecX, const T *s_vec%s) {" % (name, ", const T alpha" if scaled else ""))
                for r in range(6):
                    if col[r] is None and peq:
                        continue
                    L("    s_vecX[%d] %s %s%s;" % (r, "+=" if peq else "=", term(col[r], "s_vec"),
                                                  "*alpha" if scaled and col[r] is not None else ""))
                L("}")
        for peq, scaled in ((False, False), (True, False), (False, True), (True, True)):
            suffix = ("_peq" if peq else "") + ("_scaled" if scaled else "")
            L("template <typename T>")
            L("__device__ void mxX%s(T *s_vecX, const T *s_vec, %sconst int S_ind) {" % (
                suffix, "const T alpha, " if scaled else ""))
            L("    switch(S_ind){")
            for k in range(6):
                L("        case %d: mx%d%s<T>(s_vecX, s_vec%s); break;" % (k, k, suffix, ", alpha" if scaled else ""))
            L("    }")
            L("}")
        fx_entry = lambda r, c: (lambda e: None if e is None else (-e[0], e[1]))(crm_entry(c, r))   # fx = -(v x)^T
        for zeroed in (False, True):
            L("template <typename T>")
            L("__device__ void fx%s(T *s_matX, const T *s_vecX) {" % ("_zeroed" if zeroed else ""))
            for c in range(6):
                for r in range(6):
                    e = fx_entry(r, c)
                    if e is None and zeroed:
                        continue
                    L("    s_matX[6*%d + %d] = %s;" % (c, r, term(e, "s_vecX")))
            L("}")
        for peq in (False, True):
            L("template <typename T>")
            L("__device__ void fx_times_v%s(T *s_result, const T *s_fxVec, const T *s_timesVec) {" % ("_peq" if peq else ""))
            for r in range(6):
                terms = []
                for c in range(6):
                    e = fx_entry(r, c)
                    if e is not None:
                        terms.append("%s s_fxVec[%d] * s_timesVec[%d]" % ("-" if e[0] < 0 else "+", e[1], c))
                L("    s_result[%d] %s %s;" % (r, "+=" if peq else "=", " ".join(terms)))
            L("}")
        L("")

    def gen_mx_func_call_for_cpp(self, inds=None, PEQ_FLAG=False, SCALE_FLAG=False, updated_var_names=None):
        """Emits the call to the statically selected mx<k> helper when all joints in `inds` share an axis,
        else to the run-time mxX variant (helpers/_spatial_algebra_helpers.py:1-33)."""
        v = dict(S_ind_name="S_ind", s_dst_name="s_dst", s_src_name="s_src", s_scale_name="s_scale")
        v.update(updated_var_names or {})
        n = self.robot.get_num_pos()
        inds = list(range(n)) if inds is None else inds
        same = self.robot.are_Ss_identical(inds)
        k = str(self.robot.S_ind[inds[0]]) if same else "X"
        name = "mx%s%s%s<T>" % (k, "_peq" if PEQ_FLAG else "", "_scaled" if SCALE_FLAG else "")
        args = [v["s_dst_name"], v["s_src_name"]] + ([v["s_scale_name"]] if SCALE_FLAG else []) + \
               ([] if same else [v["S_ind_name"]])
        self.gen_add_code_line("%s(%s);" % (name, ", ".join(args)))

    def _topology_table(self):
        """[parent(n) | S_ind(n) | num_ancestors(n) | num_subtree(n) | running_sum_anc(n+1) | running_sum_sub(n)]"""
        n = self.robot.get_num_pos()
        num_anc = [len(self.robot.get_ancestors_by_id(j)) for j in range(n)]
        num_sub = [len(self.robot.get_subtree_by_id(j)) for j in range(n)]
        run_anc = [sum(num_anc[:j]) for j in range(n + 1)]
        run_sub = [sum(num_sub[:j]) for j in range(n)]
        return list(self.robot.get_parent_id_array()) + list(self.robot.S_ind) + num_anc + num_sub + run_anc + run_sub

    def gen_topology_helpers_pointers_for_cpp(self, inds=None, updated_var_names=None, NO_GRAD_FLAG=False):
        """C++ index expressions for parent / S axis / compressed-column offsets of joint `jid`
        (helpers/_topology_helpers.py:260-332): literals for a single joint, closed forms for serial
        chains, otherwise look-ups in the emitted `topology_helpers` table (a namespace-scope device
        array here; the same layout is also reachable through robotModel::d_topology_helpers)."""
        v = dict(jid_name="jid", s_topology_helpers_name="topology_helpers")
        v.update(updated_var_names or {})
        n = self.robot.get_num_pos()
        inds = list(range(n)) if inds is None else inds
        jid, tab = v["jid_name"], v["s_topology_helpers_name"]
        _, _, run_dva, _, _, run_df, df_col = self.gen_topology_sparsity_helpers_python()
        same_S = self.robot.are_Ss_identical(inds)
        if len(inds) == 1:
            i, par = inds[0], self.robot.get_parent_id(inds[0])
            dva_p1 = run_dva[i + 1] if i + 1 < n else run_dva[i] + len(self.robot.get_ancestors_by_id(i)) + 1
            out = [str(par), str(self.robot.S_ind[i]), str(run_dva[i]), str(run_df[i]),
                   str(run_dva[par]) if par >= 0 else "-1", str(run_df[par]) if par >= 0 else "-1", str(dva_p1),
                   str(df_col[i])]
        elif self.robot.is_serial_chain():
            out = ["(%s-1)" % jid, str(self.robot.S_ind[inds[0]]) if same_S else "%s[%d + %s]" % (tab, n, jid),
                   "%s*(%s+1)/2" % (jid, jid), "%d*%s" % (n, jid), "%s*(%s-1)/2" % (jid, jid), "%d*(%s-1)" % (n, jid),
                   "(%s+1)*(%s+2)/2" % (jid, jid), jid]
        else:
            par = "%s[%s]" % (tab, jid)
            ra, rs = 4 * n, 5 * n + 1
            out = [par, str(self.robot.S_ind[inds[0]]) if same_S else "%s[%d + %s]" % (tab, n, jid),
                   "(%s[%d + %s] + %s)" % (tab, ra, jid, jid),
                   "(%s[%d + %s] + %s[%d + %s])" % (tab, ra, jid, tab, rs, jid),
                   "(%s[%d + %s] + %s)" % (tab, ra, par, par),
                   "(%s[%d + %s] + %s[%d + %s])" % (tab, ra, par, tab, rs, par),
                   "(%s[%d + %s + 1] + %s + 1)" % (tab, ra, jid, jid),
                   "%s[%d + %s]" % (tab, 2 * n, jid)]
        return tuple(out[:2]) if NO_GRAD_FLAG else tuple(out)

    def gen_insert_helpers_function_call(self, updated_var_names=None):
        """Argument splice for the helper pointers every *_inner takes.  There is no topology
        table here, so only s_XImats (accepted, unused) is threaded through."""
        names = dict(s_XImats_name="s_XImats")
        names.update(updated_var_names or {})
        return names["s_XImats_name"] + ", "

    def gen_insert_helpers_func_def_params(self, func_def, func_params, param_insert_position=-1,
                                           updated_var_names=None):
        names = dict(s_XImats_name="s_XImats")
        names.update(updated_var_names or {})
        func_def += "T *" + names["s_XImats_name"] + ", "
        func_params.insert(param_insert_position, "s_XImats is accepted for API compatibility and unused")
        return func_def, func_params

    def gen_init_topology_helpers(self):
        tab = self._topology_table()
        self.gen_add_code_lines([
            "// topology table for user code (the traced programs resolve topology at generation time):",
            "// [parent(n) | S_ind(n) | num_ancestors(n) | num_subtree(n) | running_sum_anc(n+1) | running_sum_sub(n)]",
            "__device__ const int topology_helpers[%d] = {%s};" % (len(tab), ", ".join(str(x) for x in tab)),
            "const int h_topology_helpers[%d] = {%s};" % (len(tab), ", ".join(str(x) for x in tab)),
            "template <typename T>", "__host__", "int *init_topology_helpers(){",
            "    int *d_topology_helpers; gpuErrchk(cudaMalloc((void**)&d_topology_helpers,%d*sizeof(int)));" % len(tab),
            "    gpuErrchk(cudaMemcpy(d_topology_helpers,h_topology_helpers,%d*sizeof(int),cudaMemcpyHostToDevice));" % len(tab),
            "    return d_topology_helpers;", "}", ""])

    # ---- XI table: [X_0 .. X_{n-1} | (I_base) | I_0 .. I_{n-1}], 36 floats each, column-major
    # (element (row, col) of matrix k at 36 k + 6 col + row: reference helpers/_topology_helpers.py:19-47).
    # The traced programs never read it (their X_tree and inertias are immediates), but the reference's
    # users do: a third-party kernel written against grid.cuh loads s_XImats with
    # load_update_XImats_helpers() and applies X_i(q) / I_i itself.
    def _ximats_entries(self):
        """Per joint: {(row, col): (a, b, c)} with X[row, col](q) = a sin q + b cos q + c (revolute) or
        a q + c (prismatic, b = 0), found numerically from the robot's own X(q) functions."""
        out = []
        for i in range(self.robot.get_num_pos()):
            X = self.robot.get_Xmat_Func_by_id(i)
            ent = {}
            if self.robot.S_ind[i] < 3:
                X0, X1, X2 = X(0.0), X(0.5 * np.pi), X(np.pi)
                c = 0.5 * (X0 + X2)
                a, b = X1 - c, 0.5 * (X0 - X2)
            else:
                X0, X1 = X(0.0), X(1.0)
                c, a, b = X0, X1 - X0, np.zeros((6, 6))
            for col in range(6):
                for row in range(6):
                    t = tuple(0.0 if abs(x) < 1e-14 else float("%.13g" % x) for x in (a[row, col], b[row, col], c[row, col]))
                    ent[(row, col)] = t
            out.append(ent)
        return out

    def gen_init_XImats(self, include_base_inertia=False):
        n = self.robot.get_num_pos()
        size = 72 * n + (36 if include_base_inertia else 0)
        self.gen_add_func_doc("Initializes the Xmats and Imats in GPU memory",
                              ["Memory order is X[0...N], %sI[0...N]; entries of X that depend on q are 0 until "
                               "load_update_XImats_helpers fills them" % ("Ibase, " if include_base_inertia else "")],
                              [], "A pointer to the XI memory in the GPU")
        lines = ["template <typename T>", "__host__", "T* init_XImats() {",
                 "    T *h_XImats = (T *)malloc(%d*sizeof(T));" % size]
        for i, ent in enumerate(self._ximats_entries()):
            lines.append("    // X[%d]" % i)
            for col in range(6):
                for row in range(6):
                    a, b, c = ent[(row, col)]
                    lines.append("    h_XImats[%d] = static_cast<T>(%s);" % (36 * i + 6 * col + row,
                                                                          repr(c) if a == 0.0 and b == 0.0 else "0"))
        Imats = self.robot.get_Imats_ordered_by_id()
        if not include_base_inertia:
            Imats = Imats[1:]
        for k, I in enumerate(Imats):
            lines.append("    // %s" % ("Base Inertia" if include_base_inertia and k == 0 else
                                        "I[%d]" % (k - int(include_base_inertia))))
            for col in range(6):
                for row in range(6):
                    lines.append("    h_XImats[%d] = static_cast<T>(%s);" % (36 * (n + k) + 6 * col + row,
                                                                          repr(float(I[row, col]))))
        lines += ["    T *d_XImats; gpuErrchk(cudaMalloc((void**)&d_XImats,%d*sizeof(T)));" % size,
                  "    gpuErrchk(cudaMemcpy(d_XImats,h_XImats,%d*sizeof(T),cudaMemcpyHostToDevice));" % size,
                  "    free(h_XImats);", "    return d_XImats;", "}", ""]
        self.gen_add_code_lines(lines)

    def gen_init_robotModel(self):
        self.gen_add_func_doc("Initializes the robotModel helpers in GPU memory (XI table + topology table)", [], [],
                              "A pointer to the robotModel struct on the GPU")
        self.gen_add_code_lines([
            "template <typename T>", "__host__", "robotModel<T>* init_robotModel() {",
            "    robotModel<T> h_robotModel; h_robotModel.d_XImats = init_XImats<T>();",
            "    h_robotModel.d_topology_helpers = init_topology_helpers<T>();",
            "    robotModel<T> *d_robotModel; gpuErrchk(cudaMalloc((void**)&d_robotModel,sizeof(robotModel<T>)));",
            "    gpuErrchk(cudaMemcpy(d_robotModel,&h_robotModel,sizeof(robotModel<T>),cudaMemcpyHostToDevice));",
            "    return d_robotModel;", "}", ""])

    def gen_init_gridData(self):
        body = ["    gridData<T> *hd_data = (gridData<T> *)malloc(sizeof(gridData<T>));",
                "    const size_t n = NUM_JOINTS, Tn = NUM_TIMESTEPS;"]
        for nm, words in (("q_qd_u", "3*n"), ("q_qd", "2*n"), ("q", "n"), ("c", "n"), ("Minv", "n*n"), ("qdd", "n"),
                          ("dc_du", "2*n*n"), ("df_du", "2*n*n")):
            body.append("    gpuErrchk(cudaMalloc((void**)&hd_data->d_%s, %s*Tn*sizeof(T)));" % (nm, words))
            body.append("    gpuErrchk(cudaMallocHost((void**)&hd_data->h_%s, %s*Tn*sizeof(T)));  // pinned" % (nm, words))
        body.append("    return hd_data;")
        for tmpl, sig in (("template <typename T, int NUM_TIMESTEPS>", "gridData<T> *init_gridData(){"),
                          ("template <typename T>", "gridData<T> *init_gridData(int NUM_TIMESTEPS){")):
            self.gen_add_func_doc("Allocates device and (pinned) host memory for all computations", [], [],
                                  "A pointer to the gridData struct of pointers")
            self.gen_add_code_lines([tmpl, "__host__", sig] + body + ["}", ""])

    def gen_load_update_XImats_helpers_temp_mem_size(self):
        return 2 * self.robot.get_num_pos()          # sin | cos, as in the reference (_topology_helpers.py:56-58)

    def _needs_topology_param(self):
        n = self.robot.get_num_pos()
        return not self.robot.is_serial_chain() or not self.robot.are_Ss_identical(list(range(n)))

    def gen_load_update_XImats_helpers_function_call(self, use_thread_group=False, updated_var_names=None):
        v = dict(s_XImats_name="s_XImats", s_q_name="s_q", d_robotModel_name="d_robotModel", s_temp_name="s_temp",
                 s_topology_helpers_name="s_topology_helpers")
        v.update(updated_var_names or {})
        args = [v["s_XImats_name"], v["s_q_name"]]
        if self._needs_topology_param():
            args.append(v["s_topology_helpers_name"])
        self.gen_add_code_line("load_update_XImats_helpers<T>(%s);" % ", ".join(
            args + [v["d_robotModel_name"], v["s_temp_name"]]))

    def gen_XImats_helpers_temp_shared_memory_code(self, temp_mem_size=None):
        n = self.robot.get_num_pos()
        words = self.gen_load_update_XImats_helpers_temp_mem_size() if temp_mem_size is None else temp_mem_size
        if not self.use_dynamic_shared_mem_flag:
            self.gen_add_code_line("__shared__ T s_XImats[%d];" % (72 * n))
            if self._needs_topology_param():
                self.gen_add_code_line("__shared__ int s_topology_helpers[%d];" % self.gen_topology_helpers_size())
            self.gen_add_code_line("__shared__ T s_temp[%d];" % max(1, words))
        else:
            self.gen_add_code_line("extern __shared__ T s_XITemp[]; T *s_XImats = s_XITemp; ")
            if self._needs_topology_param():
                self.gen_add_code_line("int *s_topology_helpers = (int *)&s_XImats[%d];" % (72 * n))
                self.gen_add_code_line("T *s_temp = (T *)&s_topology_helpers[%d];" % self.gen_topology_helpers_size())
            else:
                self.gen_add_code_line("T *s_temp = &s_XImats[%d];" % (72 * n))

    def gen_load_update_XImats_helpers(self, use_thread_group=False):
        """X_i(q) for every joint into s_XImats (and the I block and topology table copied beside it): what
        reference helpers/_topology_helpers.py:90-182 emits, with every q-dependent entry written by the
        thread that owns the joint instead of thread 0."""
        n = self.robot.get_num_pos()
        topo = self._needs_topology_param()
        params = ["s_XImats is the (shared) memory destination location for the XImats (72*NUM_JOINTS)",
                  "s_q is the (shared) memory location of the current configuration",
                  "d_robotModel is the pointer to the initialized model specific helpers",
                  "s_temp is temporary (shared) memory for sin and cos of size: %d" % (2 * n)]
        if topo:
            params.insert(2, "s_topology_helpers is the (shared) memory destination location for the topology_helpers")
        self.gen_add_func_doc("Updates the Xmats in (shared) GPU memory acording to the configuration", [], params, None)
        if topo:
            self.gen_add_code_line("#define GRID_XIMATS_TAKES_TOPOLOGY_HELPERS 1   // extra int *s_topology_helpers argument")
        lines = ["template <typename T>", "__device__",
                 "void load_update_XImats_helpers(T *s_XImats, const T *s_q, %sconst robotModel<T> *d_robotModel, "
                 "T *s_temp) {" % ("int *s_topology_helpers, " if topo else ""),
                 "    const int tid_ = threadIdx.x + threadIdx.y*blockDim.x, nthr_ = blockDim.x*blockDim.y;",
                 "    for(int ind = tid_; ind < %d; ind += nthr_){ s_XImats[ind] = d_robotModel->d_XImats[ind]; }" % (72 * n)]
        if topo:
            lines.append("    for(int ind = tid_; ind < %d; ind += nthr_){ s_topology_helpers[ind] = "
                         "d_robotModel->d_topology_helpers[ind]; }" % self.gen_topology_helpers_size())
        lines += ["    for(int k = tid_; k < %d; k += nthr_){ s_temp[k] = static_cast<T>(sin(s_q[k])); "
                  "s_temp[k+%d] = static_cast<T>(cos(s_q[k])); }" % (n, n),
                  "    __syncthreads();",
                  "    for(int k = tid_; k < %d; k += nthr_){" % n,
                  "        const T sn_ = s_temp[k], cs_ = s_temp[k+%d], th_ = s_q[k]; T *X_ = &s_XImats[36*k];" % n,
                  "        (void)sn_; (void)cs_; (void)th_;",
                  "        switch(k){"]
        for i, ent in enumerate(self._ximats_entries()):
            lines.append("        case %d:" % i)
            rev = self.robot.S_ind[i] < 3
            for col in range(6):
                for row in range(6):
                    a, b, c = ent[(row, col)]
                    if a == 0.0 and b == 0.0:
                        continue
                    terms = []
                    if a != 0.0:
                        terms.append("static_cast<T>(%r)*%s" % (a, "sn_" if rev else "th_"))
                    if b != 0.0:
                        terms.append("static_cast<T>(%r)*cs_" % b)
                    if c != 0.0:
                        terms.append("static_cast<T>(%r)" % c)
                    lines.append("            X_[%d] = %s;" % (6 * col + row, " + ".join(terms)))
            lines.append("            break;")
        lines += ["        }", "    }", "    __syncthreads();", "}", ""]
        self.gen_add_code_lines(lines)

    def gen_init_close_grid(self):
        if None in self._family.values():
            self.gen_add_code_line("// some kernels do not exist for this robot: init_grid/close_grid omitted")
            return
        self.gen_add_func_doc("Sets shared mem needed for gradient kernels and initializes streams for host functions",
                              [], [], "A pointer to the array of streams")
        self.gen_add_code_lines([
            "template <typename T>", "__host__", "cudaStream_t *init_grid(){",
            "    // kernels may need more than the default 48 KB of dynamic shared memory",
            "    typedef void (*k5_t)(T *, const T *, const int, const robotModel<T> *, const int);",
            "    typedef void (*k6_t)(T *, const T *, const int, const robotModel<T> *, const T, const int);",
            "    typedef void (*k7_t)(T *, const T *, const int, const T *, const robotModel<T> *, const T, const int);",
            "    typedef void (*k8_t)(T *, const T *, const int, const T *, const T *, const robotModel<T> *, const T, const int);",
            "    const int minv_bytes = MINV_DYNAMIC_SHARED_MEM_COUNT*sizeof(T), fd_bytes = FD_DYNAMIC_SHARED_MEM_COUNT*sizeof(T);",
            "    const int id_du_bytes = ID_DU_MAX_SHARED_MEM_COUNT*sizeof(T), fd_du_bytes = FD_DU_MAX_SHARED_MEM_COUNT*sizeof(T);",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k5_t>(&direct_minv_kernel<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,minv_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k5_t>(&direct_minv_kernel_single_timing<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,minv_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k6_t>(&forward_dynamics_kernel<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,fd_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k6_t>(&forward_dynamics_kernel_single_timing<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,fd_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k6_t>(&inverse_dynamics_gradient_kernel<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,id_du_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k7_t>(&inverse_dynamics_gradient_kernel<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,id_du_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k6_t>(&inverse_dynamics_gradient_kernel_single_timing<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,id_du_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k7_t>(&inverse_dynamics_gradient_kernel_single_timing<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,id_du_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k6_t>(&forward_dynamics_gradient_kernel<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,fd_du_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k8_t>(&forward_dynamics_gradient_kernel<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,fd_du_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k6_t>(&forward_dynamics_gradient_kernel_single_timing<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,fd_du_bytes));",
            "    gpuErrchk(cudaFuncSetAttribute(static_cast<k8_t>(&forward_dynamics_gradient_kernel_single_timing<T>),cudaFuncAttributeMaxDynamicSharedMemorySize,fd_du_bytes));",
            "    cudaStream_t *streams = (cudaStream_t *)malloc(3*sizeof(cudaStream_t));",
            "    int minPriority, maxPriority; gpuErrchk(cudaDeviceGetStreamPriorityRange(&minPriority, &maxPriority));",
            "    for(int i=0; i<3; i++){ gpuErrchk(cudaStreamCreateWithPriority(&(streams[i]),cudaStreamNonBlocking,maxPriority)); }",
            "    return streams;", "}", ""])
        self.gen_add_func_doc("Frees the memory used by grid", [],
                              ["streams allocated by init_grid", "robotModel allocated by init_robotModel",
                               "data allocated by init_gridData"], None)
        frees = ["    robotModel<T> h_robotModel; gpuErrchk(cudaMemcpy(&h_robotModel,d_robotModel,sizeof(robotModel<T>),"
                 "cudaMemcpyDeviceToHost));",
                 "    gpuErrchk(cudaFree(h_robotModel.d_XImats)); gpuErrchk(cudaFree(h_robotModel.d_topology_helpers));",
                 "    gpuErrchk(cudaFree(d_robotModel));"]
        for nm in ("q_qd_u", "q_qd", "q", "c", "Minv", "qdd", "dc_du", "df_du"):
            frees.append("    gpuErrchk(cudaFree(hd_data->d_%s)); gpuErrchk(cudaFreeHost(hd_data->h_%s));" % (nm, nm))
        self.gen_add_code_lines(["template <typename T>", "__host__",
                                 "void close_grid(cudaStream_t *streams, robotModel<T> *d_robotModel, gridData<T> *hd_data){"]
                                + frees + ["    for(int i=0; i<3; i++){gpuErrchk(cudaStreamDestroy(streams[i]));} free(streams);",
                                           "    free(hd_data);", "}", ""])

    # ------------------------------------------------------------------ implementation namespace
    def _gen_impl_namespace(self):
        """Traced programs + thread-per-state shell, in <namespace>_b200_impl."""
        text = open(os.path.join(_PKG, "csrc", "grid_tps.cuh")).read().replace("#pragma once", "")
        self.gen_add_code_line("#define GRID_NS %s" % self._impl_ns)
        self.code_str += text
        if self._plan.pipe:     # phase-split kernels: the host functions of the large robots launch them
            self.code_str += open(os.path.join(_PKG, "csrc", "grid_pipe.cuh")).read().replace(
                "#pragma once", "").replace('#include "grid_tps.cuh"', "")
        self.gen_add_code_line("namespace GRID_NS { namespace gen {")
        if self._plan.pipe:
            from .pipeline import emit_pipe_struct
            for v, pv in self._plan.pipe.items():
                txt, _ = emit_pipe_struct(pv, self._plan.pipe_min_blocks, self._plan.pipe_warps,
                                          self._plan.pipe_sync_every, self._plan.pipe_scratch_lead)
                self.code_str += txt
        for alg, variants in (("id", ("id", "id_qdd")), ("minv", ("minv",)), ("fd", ("fd",)),
                              ("id_grad", ("id_grad", "id_grad_qdd")), ("fd_grad", ("fd_grad", "fd_grad_qdd_minv"))):
            if self._family[alg] != "tps":
                continue
            for v in variants:
                txt, _ = emit_alg_struct(self.robot, v)
                self.code_str += txt
        self.gen_add_code_line("}}  // namespace GRID_NS::gen")
        if "wps" in self._family.values():
            self.code_str += emit_wps_tables(self.robot, self._plan.wps, include=False)
            self.code_str += open(os.path.join(_PKG, "csrc", "grid_wps.cuh")).read().replace("#pragma once", "")
        self.gen_add_code_line("#undef GRID_NS")
        self.gen_add_code_line("")

    # ------------------------------------------------------------------ per-algorithm generators
    def _serial_device_fn(self, doc, signature, prog: Program, in_expr, out_expr):
        """Emits the traced program as a plain (non-template) float function plus a thin template
        wrapper with the reference signature: a 10k-statement template body makes the CUDA
        front-end re-instantiate the whole AST and takes minutes to compile."""
        self._impl_counter = getattr(self, "_impl_counter", 0) + 1
        name = signature.split("(")[0].split()[-1]
        params = signature[signature.index("(") + 1:signature.rindex(")")]
        impl = "%s_traced_%d" % (name, self._impl_counter)
        fparams = params.replace("const T *", "const float *").replace("T *", "float *").replace(
            "const T ", "const float ").replace("const robotModel<T> *d_robotModel", "const void *d_robotModel")
        argnames = ", ".join(x.strip().split()[-1].lstrip("*") for x in params.split(","))
        self.code_str += "static __device__ void %s(%s) {\n" % (impl, fparams)
        self.code_str += "\n".join(_emit_pointer_eval(prog, in_expr, out_expr, "    ")) + "\n}\n"
        self.gen_add_func_doc(doc, ["thread 0 of the block evaluates the traced program; all threads must call"],
                              [], None)
        self.gen_add_code_lines(["template <typename T>", "__device__", signature])
        self.code_str += "    static_assert(std::is_same<T,float>::value, \"T must be float\");\n"
        self.code_str += "    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) %s(%s);\n" % (impl, argnames)
        self.code_str += "    __syncthreads();\n}\n\n"

    @staticmethod
    def _smem_in(mapping):
        def f(name):
            if name == "gravity":
                return "gravity"
            base, idx = _split_name(name)
            return "%s[%d]" % (mapping[base], idx)
        return f

    def _too_large(self, alg_key, fn):
        if self._family[alg_key] != "tps":
            self.gen_add_code_line("// %s: the traced single-thread program is too large for this robot; use the "
                                   "%s kernel / host function (wide CTA-per-state kernels) instead" % (fn, alg_key))
            return True
        return False

    def _wide(self, alg_key):
        """True when the _inner/_device functions of this algorithm are served by the wide CTA-per-state body
        (csrc/grid_wps.cuh wps_inner): robots whose single-thread program is too large (Atlas, 64-link chain)."""
        return self._family[alg_key] == "wps"

    def _wide_temp_words(self):
        return self._plan.wps["smem_bytes"] // 4

    def _wide_fn(self, doc, signature, call, device=False):
        """Reference-signature wrapper around a wide body.  `call` is the wps function call with `S_WORK` standing
        for the scratch block: the caller's s_temp (_inner) or this block's dynamic shared memory (_device)."""
        notes = ["all threads of the block must call; blockDim.x must be SUGGESTED_THREADS (%d)" % self._plan.wps["NT"],
                 ("uses %d floats of dynamic shared memory (launch the calling kernel with that much)"
                  if device else "s_temp: %d floats, 16-byte aligned") % self._wide_temp_words()]
        self.gen_add_func_doc(doc, notes, [], None)
        self.gen_add_code_lines(["template <typename T>", "__device__", signature])
        self.code_str += "    static_assert(std::is_same<T,float>::value, \"T must be float\");\n"
        if device:
            self.code_str += "    extern __shared__ float4 s_wide_dyn4_[]; float *s_work_ = reinterpret_cast<float *>(s_wide_dyn4_);\n"
        else:
            self.code_str += "    float *s_work_ = s_temp;\n"
        self.code_str += "    %s::wps::%s;\n}\n\n" % (self._impl_ns, call.replace("S_WORK", "s_work_"))

    # -- inverse dynamics
    def gen_inverse_dynamics_inner_temp_mem_size(self):
        return 0

    def gen_inverse_dynamics_device_temp_mem_size(self, compute_c=False):
        return 0

    def gen_inverse_dynamics_inner_function_call(self, use_thread_group=False, compute_c=False, use_qdd_input=False,
                                                 updated_var_names=None):
        v = dict(s_c_name="s_c", s_vaf_name="s_vaf", s_q_name="s_q", s_qd_name="s_qd", s_qdd_name="s_qdd",
                 s_temp_name="s_temp", gravity_name="gravity")
        v.update(updated_var_names or {})
        call = "inverse_dynamics_inner%s<T>(" % ("" if compute_c else "_vaf")
        args = ([v["s_c_name"]] if compute_c else []) + [v["s_vaf_name"], v["s_q_name"], v["s_qd_name"]]
        if use_qdd_input:
            args.append(v["s_qdd_name"])
        self.gen_add_code_line(call + ", ".join(args + ["s_XImats", v["s_temp_name"], v["gravity_name"]]) + ");")

    def gen_inverse_dynamics_inner(self, use_thread_group=False, compute_c=False, use_qdd_input=False):
        if self._too_large("id", "inverse_dynamics_inner"):
            return
        p = A.trace_id_full(self.robot, use_qdd_input)
        if not compute_c:
            p.outputs = [o for o in p.outputs if o[0] == "vaf"]
        sig = "void inverse_dynamics_inner%s(%sT *s_vaf, const T *s_q, const T *s_qd, %sT *s_XImats, T *s_temp, " \
              "const T gravity) {" % ("" if compute_c else "_vaf", "T *s_c, " if compute_c else "",
                                     "const T *s_qdd, " if use_qdd_input else "")
        self._serial_device_fn("Compute the RNEA (Recursive Newton-Euler Algorithm)", sig, p,
                               self._smem_in({"q": "s_q", "qd": "s_qd", "qdd": "s_qdd"}),
                               lambda a, i: "s_%s[%d]" % (a, i))

    def gen_inverse_dynamics_device(self, use_thread_group=False, compute_c=False, use_qdd_input=False):
        if self._too_large("id", "inverse_dynamics_device"):
            return
        p = A.trace_id_full(self.robot, use_qdd_input)
        p.outputs = [o for o in p.outputs if o[0] == ("c" if compute_c else "vaf")]
        sig = "void inverse_dynamics%s_device(T *s_%s, const T *s_q, const T *s_qd, %sconst robotModel<T> *d_robotModel, " \
              "const T gravity) {" % ("" if compute_c else "_vaf", "c" if compute_c else "vaf",
                                     "const T *s_qdd, " if use_qdd_input else "")
        self._serial_device_fn("Compute the RNEA (Recursive Newton-Euler Algorithm)", sig, p,
                               self._smem_in({"q": "s_q", "qd": "s_qd", "qdd": "s_qdd"}),
                               lambda a, i: "s_%s[%d]" % (a, i))

    def _kernel(self, name, out, in_name, stride_name, extra_params, struct_plain, struct_extra, extra_cond,
                single_call_timing, alg_key, doc):
        """Reference-signature __global__ wrapper around the thread-per-state tile loop."""
        if self._family[alg_key] is None:
            self.gen_add_code_line("// %s: no kernel for this robot (shared memory of the wide kernels exceeds 227 KB)" % name)
            return
        ns = self._impl_ns
        fn = name + ("_single_timing" if single_call_timing else "")
        params = "T *%s, const T *%s, const int %s, %sconst robotModel<T> *d_robotModel, const T gravity, " \
                 "const int NUM_TIMESTEPS" % (out, in_name, stride_name, extra_params)
        if name == "direct_minv_kernel":
            params = params.replace("const T gravity, ", "")
        wide = self._family[alg_key] == "wps"
        self.gen_add_func_doc(doc, ["one CTA per state, lanes = columns: launch with exactly SUGGESTED_THREADS threads"
                                    if wide else "one thread per state; any 1-D/2-D launch shape",
                                    "dynamic shared memory = <CODE>_DYNAMIC_SHARED_MEM_COUNT*sizeof(T)"], [], None)
        self.gen_add_code_lines(["template <typename T>", "__global__", "__launch_bounds__(SUGGESTED_THREADS)",
                                 "void %s(%s) {" % (fn, params)])
        if wide:
            code = {"minv": 0, "fd": 1, "id_grad": 2, "fd_grad": 3}[alg_key]
            in1 = "d_qdd" if "d_qdd" in extra_params else "nullptr"
            in2 = "d_Minv" if "d_Minv" in extra_params else "nullptr"
            g = "0.f" if name == "direct_minv_kernel" else "gravity"
            call = "%s::wps::wps_body<%d, %s>(%s, %s, %s, %s, %s, %s, %s);" % (
                ns, code, "true" if extra_cond else "false", out, in_name, stride_name, in1, in2,
                "1" if single_call_timing else "NUM_TIMESTEPS", g)
            body = ["    static_assert(std::is_same<T,float>::value, \"T must be float\");",
                    "    if (blockDim.x != SUGGESTED_THREADS) { if (threadIdx.x == 0 && blockIdx.x == 0) "
                    "printf(\"%s needs SUGGESTED_THREADS threads per block\\n\"); return; }" % fn]
            if single_call_timing:
                body.append("    for (int rep = 0; rep < NUM_TIMESTEPS; rep++)")
            body.append("    " + call)
            self.gen_add_code_lines(body + ["}", ""])
            return
        struct = struct_extra if extra_cond else struct_plain
        in1 = "d_qdd" if "d_qdd" in extra_params else "nullptr"
        in2 = "d_Minv" if "d_Minv" in extra_params else "nullptr"
        g = "0.f" if name == "direct_minv_kernel" else "gravity"
        count = "1" if single_call_timing else "NUM_TIMESTEPS"
        body = ["    static_assert(std::is_same<T,float>::value, \"T must be float\");"]
        if single_call_timing:
            body.append("    for (int rep = 0; rep < NUM_TIMESTEPS; rep++)")
        body.append("    %s::tps_body<%s::gen::%s>(%s, %s, %s, %s, %s, %s, %s, 0.f);" % (
            ns, ns, struct, out, in_name, stride_name, in1, in2, count, g))
        self.gen_add_code_lines(body + ["}", ""])

    def gen_inverse_dynamics_kernel(self, use_thread_group=False, use_qdd_input=False, single_call_timing=False):
        self._kernel("inverse_dynamics_kernel", "d_c", "d_q_qd", "stride_q_qd",
                     "const T *d_qdd, " if use_qdd_input else "", "AlgId", "AlgIdQdd", use_qdd_input,
                     single_call_timing, "id", "Compute the RNEA (Recursive Newton-Euler Algorithm)")

    def _host(self, fn, mode, tmpl, flags_doc, alg_key, copies_in, kernel_calls, out_name, out_words, code, takes_g=True):
        single, compute_only = mode == 1, mode == 2
        if self._family[alg_key] is None:
            return
        wide = self._family[alg_key] == "wps"
        name = fn + ("_single_timing" if single else "_compute_only" if compute_only else "")
        sig = "void %s(gridData<T> *hd_data, const robotModel<T> *d_robotModel, %sconst int num_timesteps, " \
              "const dim3 block_dimms, const dim3 thread_dimms%s) {" % (
                  name, "const T gravity, " if takes_g else "", "" if compute_only else ", cudaStream_t *streams")
        self.gen_add_func_doc("%s host wrapper (%s)" % (fn, ["H2D + kernel + D2H", "single-call timing",
                                                             "compute only"][mode]),
                              ["block_dimms / thread_dimms are accepted for compatibility; the launch shape is "
                               "chosen by the library", flags_doc], [], None)
        T = "1" if single else "num_timesteps"
        lines = [tmpl, "__host__", sig, "    const int T_ = %s; (void)block_dimms; (void)thread_dimms;" % T,
                 ("    const int blocks_ = T_ < 148*16 ? T_ : 148*16;   // one CTA per state, grid-stride beyond" if wide
                  else "    const int blocks_ = (T_ + SUGGESTED_THREADS - 1) / SUGGESTED_THREADS;"),
                 "    const size_t smem_ = %s_DYNAMIC_SHARED_MEM_COUNT*sizeof(T);" % code]
        if not compute_only:
            lines += ["    " + c for c in copies_in] + ["    gpuErrchk(cudaDeviceSynchronize());"]
        if single:
            lines.append("    struct timespec start, end; clock_gettime(CLOCK_MONOTONIC,&start);")
        if not single and "pipe" in self._plan.kind.get(alg_key, ""):
            kernel_calls = [self._pipe_host_call(k, alg_key) for k in kernel_calls]
        lines += ["    " + k.replace("KERNEL<T>", ("%s_kernel%s<T>" % (fn, "_single_timing" if single else "")))
                  .replace("<<<>>>", "<<<blocks_,SUGGESTED_THREADS,smem_>>>")
                  .replace("NT_", "num_timesteps") for k in kernel_calls]
        lines.append("    gpuErrchk(cudaGetLastError()); gpuErrchk(cudaDeviceSynchronize());")
        if single:
            lines.append("    clock_gettime(CLOCK_MONOTONIC,&end);")
        if not compute_only:
            lines.append("    gpuErrchk(cudaMemcpy(hd_data->h_%s,hd_data->d_%s,%s*T_*sizeof(T),cudaMemcpyDeviceToHost));"
                         % (out_name, out_name, out_words))
        if single:
            lines.append("    printf(\"Single Call %s %%fus\\n\",time_delta_us_timespec(start,end)/"
                         "static_cast<double>(num_timesteps));" % code)
        self.gen_add_code_lines(lines + ["}", ""])

    def _pipe_host_call(self, line: str, alg_key: str) -> str:
        """Rewrites `KERNEL<T><<<>>>(out, in, stride, [d_qdd,] d_robotModel, [gravity,] NT_)` into a launch of
        the phase-split kernels (csrc/grid_pipe.cuh) on the default stream, where a variant exists; the
        reference-signature `_kernel` (wide kernel) stays available for callers that launch it themselves."""
        import re
        names = {"minv": ("PipeMinv", None), "fd": ("PipeFd", None), "id": ("PipeId", "PipeIdQdd"),
                 "id_grad": ("PipeIdGrad", "PipeIdGradQdd"), "fd_grad": ("PipeFdGrad", None)}[alg_key]

        def sub(m):
            args = [a.strip() for a in m.group(1).split(",")]
            extras = args[3:args.index("d_robotModel")]
            g = "gravity" if "gravity" in args else "0.f"
            if len(extras) == 2 and alg_key == "fd_grad" and "fd_grad_qdd_minv" in self._plan.pipe:
                # USE_QDD_MINV_FLAG: qdd through the tile, the caller's Minv read straight from global memory
                return "gpuErrchk(%s::pipe::pipe_launch<%s::gen::PipeFdGradPre>(%s, %s, %s, %s, num_timesteps, %s, 0, 0.f, %s));" % (
                    self._impl_ns, self._impl_ns, args[0], args[1], args[2], extras[0], g, extras[1])
            if len(extras) > 1 or (extras and names[1] is None):
                return m.group(0)                      # no phase-split variant: wide kernel
            struct = names[1] if extras else names[0]
            return "gpuErrchk(%s::pipe::pipe_launch<%s::gen::%s>(%s, %s, %s, %s, num_timesteps, %s, 0));" % (
                self._impl_ns, self._impl_ns, struct, args[0], args[1], args[2], extras[0] if extras else "nullptr", g)
        return re.sub(r"KERNEL<T><<<>>>\(([^;]*)\);", sub, line)

    _H2D = "gpuErrchk(cudaMemcpyAsync(hd_data->d_%s,hd_data->h_%s,%s*T_*sizeof(T),cudaMemcpyHostToDevice,streams[%d]));"

    def gen_inverse_dynamics_host(self, mode=0):
        copies = ["if (USE_COMPRESSED_MEM) {" + self._H2D % ("q_qd", "q_qd", "2*NUM_JOINTS", 0) + "}",
                  "else {" + self._H2D % ("q_qd_u", "q_qd_u", "3*NUM_JOINTS", 0) + "}",
                  "if (USE_QDD_FLAG) {" + self._H2D % ("qdd", "qdd", "NUM_JOINTS", 1) + "}"]
        calls = ["const T *in_ = USE_COMPRESSED_MEM ? hd_data->d_q_qd : hd_data->d_q_qd_u; "
                 "const int stride_ = USE_COMPRESSED_MEM ? 2*NUM_JOINTS : 3*NUM_JOINTS;",
                 "if (USE_QDD_FLAG) {KERNEL<T><<<>>>(hd_data->d_c,in_,stride_,hd_data->d_qdd,d_robotModel,gravity,NT_);}",
                 "else {KERNEL<T><<<>>>(hd_data->d_c,in_,stride_,d_robotModel,gravity,NT_);}"]
        self._host("inverse_dynamics", mode, "template <typename T, bool USE_QDD_FLAG = false, bool USE_COMPRESSED_MEM = false>",
                   "USE_QDD_FLAG reads d_qdd; USE_COMPRESSED_MEM reads d_q_qd (stride 2n)", "id", copies, calls,
                   "c", "NUM_JOINTS", "ID")

    def gen_inverse_dynamics(self, use_thread_group=False):
        self.gen_inverse_dynamics_inner(use_thread_group, True, True)
        self.gen_inverse_dynamics_inner(use_thread_group, True, False)
        self.gen_inverse_dynamics_inner(use_thread_group, False, True)
        self.gen_inverse_dynamics_inner(use_thread_group, False, False)
        self.gen_inverse_dynamics_device(use_thread_group, True, True)
        self.gen_inverse_dynamics_device(use_thread_group, True, False)
        self.gen_inverse_dynamics_device(use_thread_group, False, True)
        self.gen_inverse_dynamics_device(use_thread_group, False, False)
        for timing in (False, True):
            self.gen_inverse_dynamics_kernel(use_thread_group, True, timing)
            self.gen_inverse_dynamics_kernel(use_thread_group, False, timing)
        for mode in (0, 1, 2):
            self.gen_inverse_dynamics_host(mode)

    # -- direct minv
    def gen_direct_minv_inner_temp_mem_size(self):
        return self._wide_temp_words() if self._wide("minv") else 0

    def gen_direct_minv_inner_function_call(self, use_thread_group=False, updated_var_names=None):
        v = dict(s_Minv_name="s_Minv", s_q_name="s_q", s_temp_name="s_temp")
        v.update(updated_var_names or {})
        self.gen_add_code_line("direct_minv_inner<T>(%s, %s, s_XImats, %s);" % (v["s_Minv_name"], v["s_q_name"],
                                                                                 v["s_temp_name"]))

    def gen_direct_minv_inner(self, use_thread_group=False):
        if self._wide("minv"):
            self._wide_fn("Compute the inverse of the mass matrix (upper triangle, column-major)",
                          "void direct_minv_inner(T *s_Minv, const T *s_q, T *s_XImats, T *s_temp) {",
                          "wps_inner<0, false>(s_Minv, s_q, nullptr, nullptr, nullptr, S_WORK, 0.f)")
            return
        if self._too_large("minv", "direct_minv_inner"):
            return
        self._serial_device_fn("Compute the inverse of the mass matrix (upper triangle, column-major)",
                               "void direct_minv_inner(T *s_Minv, const T *s_q, T *s_XImats, T *s_temp) {",
                               A.trace_minv(self.robot), self._smem_in({"q": "s_q"}), lambda a, i: "s_Minv[%d]" % i)

    def gen_direct_minv_device(self, use_thread_group=False):
        if self._wide("minv"):
            self._wide_fn("Compute the inverse of the mass matrix (upper triangle, column-major)",
                          "void direct_minv_device(T *s_Minv, const T *s_q, const robotModel<T> *d_robotModel){",
                          "wps_inner<0, false>(s_Minv, s_q, nullptr, nullptr, nullptr, S_WORK, 0.f)", device=True)
            return
        if self._too_large("minv", "direct_minv_device"):
            return
        self._serial_device_fn("Compute the inverse of the mass matrix (upper triangle, column-major)",
                               "void direct_minv_device(T *s_Minv, const T *s_q, const robotModel<T> *d_robotModel){",
                               A.trace_minv(self.robot), self._smem_in({"q": "s_q"}), lambda a, i: "s_Minv[%d]" % i)

    def gen_direct_minv_kernel(self, use_thread_group=False, single_call_timing=False):
        self._kernel("direct_minv_kernel", "d_Minv", "d_q", "stride_q", "", "AlgMinv", "AlgMinv", False,
                     single_call_timing, "minv", "Compute the inverse of the mass matrix")

    def gen_direct_minv_host(self, mode=0):
        copies = ["if (USE_COMPRESSED_MEM) {" + self._H2D % ("q", "q", "NUM_JOINTS", 0) + "}",
                  "else {" + self._H2D % ("q_qd_u", "q_qd_u", "3*NUM_JOINTS", 0) + "}"]
        calls = ["if (USE_COMPRESSED_MEM) {KERNEL<T><<<>>>(hd_data->d_Minv,hd_data->d_q,NUM_JOINTS,d_robotModel,NT_);}",
                 "else {KERNEL<T><<<>>>(hd_data->d_Minv,hd_data->d_q_qd_u,3*NUM_JOINTS,d_robotModel,NT_);}"]
        self._host("direct_minv", mode, "template <typename T, bool USE_COMPRESSED_MEM = false>",
                   "USE_COMPRESSED_MEM reads d_q (stride n)", "minv", copies, calls, "Minv", "NUM_JOINTS*NUM_JOINTS",
                   "MINV", takes_g=False)

    def gen_direct_minv(self, use_thread_group=False):
        self.gen_direct_minv_inner(use_thread_group)
        self.gen_direct_minv_device(use_thread_group)
        self.gen_direct_minv_kernel(use_thread_group, True)
        self.gen_direct_minv_kernel(use_thread_group, False)
        for mode in (0, 1, 2):
            self.gen_direct_minv_host(mode)

    # -- forward dynamics
    def gen_forward_dynamics_inner_temp_mem_size(self):
        return self._wide_temp_words() if self._wide("fd") else 0

    def gen_forward_dynamics_finish_function_call(self, updated_var_names=None):
        v = dict(s_qdd_name="s_qdd", s_u_name="s_u", s_c_name="s_c", s_Minv_name="s_Minv")
        v.update(updated_var_names or {})
        self.gen_add_code_line("forward_dynamics_finish<T>(%s, %s, %s, %s);" % (v["s_qdd_name"], v["s_u_name"],
                                                                               v["s_c_name"], v["s_Minv_name"]))

    def gen_forward_dynamics_finish(self):
        self._serial_device_fn("Finish the forward dynamics computation qdd = Minv*(u-c)",
                               "void forward_dynamics_finish(T *s_qdd, const T *s_u, const T *s_c, const T *s_Minv) {",
                               A.trace_fd_finish(self.robot), self._smem_in({"u": "s_u", "c": "s_c", "Minv": "s_Minv"}),
                               lambda a, i: "s_qdd[%d]" % i)

    def gen_forward_dynamics_inner_function_call(self, use_thread_group=False, updated_var_names=None):
        v = dict(s_qdd_name="s_qdd", s_q_name="s_q", s_qd_name="s_qd", s_u_name="s_u", s_temp_name="s_temp",
                 gravity_name="gravity")
        v.update(updated_var_names or {})
        self.gen_add_code_line("forward_dynamics_inner<T>(%s, %s, %s, %s, s_XImats, %s, %s);" % (
            v["s_qdd_name"], v["s_q_name"], v["s_qd_name"], v["s_u_name"], v["s_temp_name"], v["gravity_name"]))

    def gen_forward_dynamics_inner(self, use_thread_group=False):
        if self._wide("fd"):
            self._wide_fn("Computes forward dynamics",
                          "void forward_dynamics_inner(T *s_qdd, const T *s_q, const T *s_qd, const T *s_u, "
                          "T *s_XImats, T *s_temp, const T gravity) {",
                          "wps_inner<1, false>(s_qdd, s_q, s_qd, s_u, nullptr, S_WORK, gravity)")
            return
        if self._too_large("fd", "forward_dynamics_inner"):
            return
        self._serial_device_fn("Computes forward dynamics",
                               "void forward_dynamics_inner(T *s_qdd, const T *s_q, const T *s_qd, const T *s_u, "
                               "T *s_XImats, T *s_temp, const T gravity) {", A.trace_fd(self.robot),
                               self._smem_in({"q": "s_q", "qd": "s_qd", "u": "s_u"}), lambda a, i: "s_qdd[%d]" % i)

    def gen_forward_dynamics_device(self, use_thread_group=False):
        if self._wide("fd"):
            self._wide_fn("Computes forward dynamics",
                          "void forward_dynamics_device(T *s_qdd, const T *s_q, const T *s_qd, const T *s_u, "
                          "const robotModel<T> *d_robotModel, const T gravity) {",
                          "wps_inner<1, false>(s_qdd, s_q, s_qd, s_u, nullptr, S_WORK, gravity)", device=True)
            return
        if self._too_large("fd", "forward_dynamics_device"):
            return
        self._serial_device_fn("Computes forward dynamics",
                               "void forward_dynamics_device(T *s_qdd, const T *s_q, const T *s_qd, const T *s_u, "
                               "const robotModel<T> *d_robotModel, const T gravity) {", A.trace_fd(self.robot),
                               self._smem_in({"q": "s_q", "qd": "s_qd", "u": "s_u"}), lambda a, i: "s_qdd[%d]" % i)

    def gen_forward_dynamics_kernel(self, use_thread_group=False, single_call_timing=False):
        self._kernel("forward_dynamics_kernel", "d_qdd", "d_q_qd_u", "stride_q_qd_u", "", "AlgFd", "AlgFd", False,
                     single_call_timing, "fd", "Computes forward dynamics")

    def gen_forward_dynamics_host(self, mode=0):
        self._host("forward_dynamics", mode, "template <typename T>", "", "fd",
                   [self._H2D % ("q_qd_u", "q_qd_u", "3*NUM_JOINTS", 0)],
                   ["KERNEL<T><<<>>>(hd_data->d_qdd,hd_data->d_q_qd_u,3*NUM_JOINTS,d_robotModel,gravity,NT_);"],
                   "qdd", "NUM_JOINTS", "FD")

    def gen_forward_dynamics(self, use_thread_group=False):
        self.gen_forward_dynamics_finish()
        self.gen_forward_dynamics_inner(use_thread_group)
        self.gen_forward_dynamics_device(use_thread_group)
        self.gen_forward_dynamics_kernel(use_thread_group, True)
        self.gen_forward_dynamics_kernel(use_thread_group, False)
        for mode in (0, 1, 2):
            self.gen_forward_dynamics_host(mode)

    # -- inverse dynamics gradient
    def gen_inverse_dynamics_gradient_inner_temp_mem_size(self):
        return self._wide_temp_words() if self._wide("id_grad") else 0

    def gen_inverse_dynamics_gradient_kernel_max_temp_mem_size(self):
        return self._wide_temp_words() if self._wide("id_grad") else 0

    def gen_inverse_dynamics_gradient_inner_function_call(self, use_thread_group=False, updated_var_names=None):
        v = dict(s_dc_du_name="s_dc_du", s_q_name="s_q", s_qd_name="s_qd", s_vaf_name="s_vaf", s_temp_name="s_temp",
                 gravity_name="gravity")
        v.update(updated_var_names or {})
        self.gen_add_code_line("inverse_dynamics_gradient_inner<T>(%s, %s, %s, %s, s_XImats, %s, %s);" % (
            v["s_dc_du_name"], v["s_q_name"], v["s_qd_name"], v["s_vaf_name"], v["s_temp_name"], v["gravity_name"]))

    def gen_inverse_dynamics_gradient_inner(self, use_thread_group=False):
        if self._wide("id_grad"):
            self._wide_fn("Computes the gradient of inverse dynamics from v, a, f",
                          "void inverse_dynamics_gradient_inner(T *s_dc_du, const T *s_q, const T *s_qd, "
                          "const T *s_vaf, T *s_XImats, T *s_temp, const T gravity) {",
                          "wps_grad_inner_vaf(s_dc_du, s_q, s_qd, s_vaf, S_WORK)")
            return
        if self._too_large("id_grad", "inverse_dynamics_gradient_inner"):
            return
        self._serial_device_fn("Computes the gradient of inverse dynamics from v, a, f",
                               "void inverse_dynamics_gradient_inner(T *s_dc_du, const T *s_q, const T *s_qd, "
                               "const T *s_vaf, T *s_XImats, T *s_temp, const T gravity) {",
                               A.trace_id_grad_from_vaf(self.robot),
                               self._smem_in({"q": "s_q", "qd": "s_qd", "vaf": "s_vaf"}),
                               lambda a, i: "s_dc_du[%d]" % i)

    def gen_inverse_dynamics_gradient_device(self, use_thread_group=False, use_qdd_input=False):
        if self._wide("id_grad"):
            self._wide_fn("Computes the gradient of inverse dynamics",
                          "void inverse_dynamics_gradient_device(T *s_dc_du, const T *s_q, const T *s_qd, %s"
                          "const robotModel<T> *d_robotModel, const T gravity) {" % (
                              "const T *s_qdd, " if use_qdd_input else ""),
                          "wps_inner<2, %s>(s_dc_du, s_q, s_qd, %s, nullptr, S_WORK, gravity)" % (
                              ("true", "s_qdd") if use_qdd_input else ("false", "nullptr")), device=True)
            return
        if self._too_large("id_grad", "inverse_dynamics_gradient_device"):
            return
        self._serial_device_fn("Computes the gradient of inverse dynamics",
                               "void inverse_dynamics_gradient_device(T *s_dc_du, const T *s_q, const T *s_qd, %s"
                               "const robotModel<T> *d_robotModel, const T gravity) {" % (
                                   "const T *s_qdd, " if use_qdd_input else ""),
                               A.trace_id_grad(self.robot, use_qdd_input),
                               self._smem_in({"q": "s_q", "qd": "s_qd", "qdd": "s_qdd"}),
                               lambda a, i: "s_dc_du[%d]" % i)

    def gen_inverse_dynamics_gradient_kernel(self, use_thread_group=False, use_qdd_input=False, single_call_timing=False):
        self._kernel("inverse_dynamics_gradient_kernel", "d_dc_du", "d_q_qd", "stride_q_qd",
                     "const T *d_qdd, " if use_qdd_input else "", "AlgIdGrad", "AlgIdGradQdd", use_qdd_input,
                     single_call_timing, "id_grad", "Computes the gradient of inverse dynamics")

    def gen_inverse_dynamics_gradient_host(self, mode=0):
        copies = ["if (USE_COMPRESSED_MEM) {" + self._H2D % ("q_qd", "q_qd", "2*NUM_JOINTS", 0) + "}",
                  "else {" + self._H2D % ("q_qd_u", "q_qd_u", "3*NUM_JOINTS", 0) + "}",
                  "if (USE_QDD_FLAG) {" + self._H2D % ("qdd", "qdd", "NUM_JOINTS", 1) + "}"]
        calls = ["const T *in_ = USE_COMPRESSED_MEM ? hd_data->d_q_qd : hd_data->d_q_qd_u; "
                 "const int stride_ = USE_COMPRESSED_MEM ? 2*NUM_JOINTS : 3*NUM_JOINTS;",
                 "if (USE_QDD_FLAG) {KERNEL<T><<<>>>(hd_data->d_dc_du,in_,stride_,hd_data->d_qdd,d_robotModel,gravity,NT_);}",
                 "else {KERNEL<T><<<>>>(hd_data->d_dc_du,in_,stride_,d_robotModel,gravity,NT_);}"]
        self._host("inverse_dynamics_gradient", mode,
                   "template <typename T, bool USE_QDD_FLAG = false, bool USE_COMPRESSED_MEM = false>",
                   "USE_QDD_FLAG reads d_qdd; USE_COMPRESSED_MEM reads d_q_qd (stride 2n)", "id_grad", copies, calls,
                   "dc_du", "2*NUM_JOINTS*NUM_JOINTS", "ID_DU")

    def gen_inverse_dynamics_gradient(self, use_thread_group=False):
        self.gen_inverse_dynamics_gradient_inner(use_thread_group)
        self.gen_inverse_dynamics_gradient_device(use_thread_group, True)
        self.gen_inverse_dynamics_gradient_device(use_thread_group, False)
        for timing in (False, True):
            self.gen_inverse_dynamics_gradient_kernel(use_thread_group, True, timing)
            self.gen_inverse_dynamics_gradient_kernel(use_thread_group, False, timing)
        for mode in (0, 1, 2):
            self.gen_inverse_dynamics_gradient_host(mode)

    # -- forward dynamics gradient
    def gen_forward_dynamics_gradient_inner_temp_mem_size(self):
        return self._wide_temp_words() if self._wide("fd_grad") else 0

    def gen_forward_dynamics_gradient_kernel_max_temp_mem_size(self):
        return self._wide_temp_words() if self._wide("fd_grad") else 0

    def gen_forward_dynamics_gradient_inner_python(self, use_thread_group=False, use_qdd_input=False):
        """The reference composes the FD gradient in Python out of the other inners
        (algorithms/_forward_dynamics_gradient.py:7-57).  Here the composition happened at trace
        time; this emits the call into the single traced program."""
        self.gen_add_code_line("// df_du = -Minv*dc_du at qdd = FD(q,qd,u): one traced program (see *_device below)")

    def gen_forward_dynamics_gradient_device(self, use_thread_group=False, use_qdd_input=False):
        if self._wide("fd_grad"):
            self._wide_fn("Computes the gradient of forward dynamics",
                          "void forward_dynamics_gradient_device(T *s_df_du, const T *s_q, const T *s_qd, %s"
                          "const robotModel<T> *d_robotModel, const T gravity) {" % (
                              "const T *s_qdd, const T *s_Minv, " if use_qdd_input else "const T *s_u, "),
                          "wps_inner<3, %s>(s_df_du, s_q, s_qd, %s, S_WORK, gravity)" % (
                              ("true", "s_qdd, s_Minv") if use_qdd_input else ("false", "s_u, nullptr")), device=True)
            return
        if self._too_large("fd_grad", "forward_dynamics_gradient_device"):
            return
        sig = "void forward_dynamics_gradient_device(T *s_df_du, const T *s_q, const T *s_qd, %s" \
              "const robotModel<T> *d_robotModel, const T gravity) {" % (
                  "const T *s_qdd, const T *s_Minv, " if use_qdd_input else "const T *s_u, ")
        self._serial_device_fn("Computes the gradient of forward dynamics", sig,
                               A.trace_fd_grad(self.robot, use_qdd_input),
                               self._smem_in({"q": "s_q", "qd": "s_qd", "u": "s_u", "qdd": "s_qdd", "Minv": "s_Minv"}),
                               lambda a, i: "s_df_du[%d]" % i)

    def gen_forward_dynamics_gradient_kernel(self, use_thread_group=False, use_qdd_input=False, single_call_timing=False):
        self._kernel("forward_dynamics_gradient_kernel", "d_df_du", "d_q_qd" if use_qdd_input else "d_q_qd_u",
                     "stride_q_qd", "const T *d_qdd, const T *d_Minv, " if use_qdd_input else "", "AlgFdGrad",
                     "AlgFdGradPre", use_qdd_input, single_call_timing, "fd_grad",
                     "Computes the gradient of forward dynamics")

    def gen_forward_dynamics_gradient_host(self, mode=0):
        copies = [self._H2D % ("q_qd_u", "q_qd_u", "3*NUM_JOINTS", 0),
                  "if (USE_QDD_MINV_FLAG) {" + self._H2D % ("qdd", "qdd", "NUM_JOINTS", 1) + " " +
                  self._H2D % ("Minv", "Minv", "NUM_JOINTS*NUM_JOINTS", 2) + "}"]
        calls = ["if (USE_QDD_MINV_FLAG) {KERNEL<T><<<>>>(hd_data->d_df_du,hd_data->d_q_qd_u,3*NUM_JOINTS,hd_data->d_qdd,"
                 "hd_data->d_Minv,d_robotModel,gravity,NT_);}",
                 "else {KERNEL<T><<<>>>(hd_data->d_df_du,hd_data->d_q_qd_u,3*NUM_JOINTS,d_robotModel,gravity,NT_);}"]
        self._host("forward_dynamics_gradient", mode, "template <typename T, bool USE_QDD_MINV_FLAG = false>",
                   "USE_QDD_MINV_FLAG reads d_qdd and d_Minv", "fd_grad", copies, calls, "df_du",
                   "2*NUM_JOINTS*NUM_JOINTS", "FD_DU")

    def gen_forward_dynamics_gradient(self, use_thread_group=False):
        self.gen_forward_dynamics_gradient_device(use_thread_group, False)
        self.gen_forward_dynamics_gradient_device(use_thread_group, True)
        for timing in (False, True):
            self.gen_forward_dynamics_gradient_kernel(use_thread_group, True, timing)
            self.gen_forward_dynamics_gradient_kernel(use_thread_group, False, timing)
        for mode in (0, 1, 2):
            self.gen_forward_dynamics_gradient_host(mode)

    # ------------------------------------------------------------------ the whole file
    def gen_all_code(self, use_thread_group=False, include_base_inertia=False):
        self.code_str, self.indent_level = "", 0
        notes = ["Drop-in for the header GRiDCodeGenerator emits (same namespace, structs, constants and the",
                 "ALGORITHM_inner/_device/_kernel/host contract), generated by gridcodegenerator_b200 for sm_100a:",
                 "robot-specialised straight-line programs, one thread per state.  Compile with",
                 "nvcc -gencode arch=compute_100a,code=sm_100a.  Suggested (required) type T is float.",
                 "Kernels need dynamic shared memory <CODE>_DYNAMIC_SHARED_MEM_COUNT*sizeof(T), CODE in",
                 "[ID, MINV, FD, ID_DU, FD_DU], and at most SUGGESTED_THREADS threads per block use it.",
                 "Kernel families in this file: " + ", ".join("%s=%s" % kv for kv in self._family.items())]
        self.gen_add_func_doc("This instance of %s.cuh is optimized for the urdf: %s" % (self.file_namespace,
                                                                                         self.robot.name), notes)
        self.gen_add_includes(use_thread_group)
        self.gen_add_gpu_err()
        self._gen_impl_namespace()
        self.gen_add_func_doc("All functions are kept in this namespace")
        self.gen_add_code_line("namespace " + self.file_namespace + " {", True)
        self.gen_add_constants_helpers()
        self.gen_spatial_algebra_helpers()
        self.gen_init_topology_helpers()
        self.gen_init_XImats(include_base_inertia)
        self.gen_init_robotModel()
        self.gen_init_gridData()
        self.gen_load_update_XImats_helpers(use_thread_group)
        self.gen_inverse_dynamics(use_thread_group)
        self.gen_direct_minv(use_thread_group)
        self.gen_forward_dynamics(use_thread_group)
        self.gen_inverse_dynamics_gradient(use_thread_group)
        self.gen_forward_dynamics_gradient(use_thread_group)
        self.gen_init_close_grid()
        self.gen_add_end_control_flow()
        with open(self.file_namespace + ".cuh", "w") as f:
            f.write(self.code_str)

    # ------------------------------------------------------------------ numpy test functions
    # The reference binds its numpy implementation here (_test.py).  These evaluate the SAME
    # traced programs the CUDA emitter prints, in float64 on the host, so a user can compare
    # GPU output against them exactly as with the reference.
    def _eval(self, prog: Program, **arrays):
        ins = {}
        for k, v in arrays.items():
            if np.isscalar(v):
                ins[k] = np.array([float(v)])
            else:
                for i, x in enumerate(np.asarray(v, dtype=np.float64).flatten()):
                    ins["%s%d" % (k, i)] = np.array([x])
        return {k: v[0] for k, v in prog.evaluate(ins, np.float64).items()}

    def test_rnea(self, q, qd, qdd=None, GRAVITY=-9.81):
        n = self.robot.get_num_pos()
        kw = dict(q=q, qd=qd, gravity=-GRAVITY)
        if qdd is not None:
            kw["qdd"] = qdd
        out = self._eval(A.trace_id_full(self.robot, qdd is not None), **kw)
        vaf = out["vaf"]
        v, a, f = (vaf[k * 6 * n:(k + 1) * 6 * n].reshape(n, 6).T for k in range(3))
        return out["c"], v, a, f

    # pass-level functions (reference _test.py:5-107, 117-202, 229-488): the forward/backward halves
    # with the reference's argument lists and return tuples.  Halves that START an algorithm evaluate
    # the traced program up to that point; halves that take intermediate arrays from the caller
    # (test_rnea_bpass, test_minv_fpass) are a function of those arrays and run in numpy on them.
    def test_rnea_fpass(self, q, qd, qdd=None, GRAVITY=-9.81):
        n = self.robot.get_num_pos()
        kw = dict(q=q, qd=qd, gravity=-GRAVITY)
        if qdd is not None:
            kw["qdd"] = qdd
        vaf = self._eval(A.trace_rnea_fpass(self.robot, qdd is not None), **kw)["vaf"]
        v, a, f = (vaf[k * 6 * n:(k + 1) * 6 * n].reshape(n, 6).T.copy() for k in range(3))
        return v, a, f

    def test_rnea_bpass(self, q, qd, f):
        """c_i = S_i^T f_i after f_parent += X_i^T f_i, children before parents; updates f in place
        like the reference (_test.py:78-107)."""
        n = self.robot.get_num_pos()
        c = np.zeros(n)
        for i in range(n - 1, -1, -1):                     # ids are a DFS pre-order: child > parent
            c[i] = f[int(np.argmax(self.robot.get_S_by_id(i))), i] + self.robot.get_damping_by_id(i) * qd[i]
            par = self.robot.get_parent_id(i)
            if par >= 0:
                f[:, par] += self.robot.get_Xmat_Func_by_id(i)(q[i]).T @ f[:, i]
        return c, f

    def test_minv_bpass(self, q):
        n = self.robot.get_num_pos()
        out = self._eval(A.trace_minv_bpass(self.robot), q=q)
        return (out["Minv"].reshape(n, n).copy(), out["F"].reshape(n, 6, n).copy(), out["U"].reshape(n, 6).copy(),
                out["Dinv"].copy())

    def test_minv_fpass(self, q, Minv, F, U, Dinv):
        """Forward pass on caller-supplied (Minv, F, U, Dinv), joints in id order (_test.py:186-202)."""
        n = self.robot.get_num_pos()
        for i in range(n):
            par = self.robot.get_parent_id(i)
            k = int(np.argmax(self.robot.get_S_by_id(i)))
            if par >= 0:
                X = self.robot.get_Xmat_Func_by_id(i)(q[i])
                XF = X @ F[par][:, i:]
                Minv[i, i:] -= Dinv[i] * (U[i] @ XF)
                F[i][:, i:] = XF
            else:
                F[i][:, i:] = 0.0
            F[i][k, i:] += Minv[i, i:]
        return Minv

    def test_minv(self, q, output_dense=True):
        n = self.robot.get_num_pos()
        M = self._eval(A.trace_minv(self.robot), q=q)["Minv"].reshape(n, n).T
        return self.test_densify_Minv(M) if output_dense else M

    def test_rnea_grad_inner(self, q, qd, v, a, f, GRAVITY=-9.81):
        """(dc_dq, dc_dqd, dv_dq, dv_dqd, da_dq, da_dqd, df_fp_dq, df_fp_dqd, df_dq, df_dqd) from the RNEA
        results, 6 x n x n arrays indexed [row, column, joint] (_test.py:229-488)."""
        n = self.robot.get_num_pos()
        vaf = np.concatenate([np.asarray(x, dtype=np.float64).T.flatten() for x in (v, a, f)])
        out = self._eval(A.trace_id_grad_inner(self.robot), q=q, qd=qd, vaf=vaf)
        dc = out["dc_du"].reshape(2 * n, n).T
        return (dc[:, :n].copy(), dc[:, n:].copy()) + tuple(out[k].reshape(6, n, n).copy() for k in A.GRAD_INNER_ARRAYS)

    def test_densify_Minv(self, Minv):
        return np.triu(Minv) + np.triu(Minv, 1).T

    def test_rnea_grad(self, q, qd, qdd=None, GRAVITY=-9.81):
        n = self.robot.get_num_pos()
        kw = dict(q=q, qd=qd, gravity=-GRAVITY)
        if qdd is not None:
            kw["qdd"] = qdd
        return self._eval(A.trace_id_grad(self.robot, qdd is not None), **kw)["dc_du"].reshape(2 * n, n).T

    def test_fd_grad(self, q, qd, u, GRAVITY=-9.81):
        n = self.robot.get_num_pos()
        return self._eval(A.trace_fd_grad(self.robot), q=q, qd=qd, u=u, gravity=-GRAVITY)["df_du"].reshape(2 * n, n).T

    # spatial-algebra primitives (_test.py:522-681)
    def mxS(self, S, vec, alpha=1.0):
        vec = np.asarray(vec, dtype=np.float64).flatten()
        k = int(np.argmax(np.asarray(S) == 1)) if np.any(np.asarray(S) == 1) else -1
        return np.zeros(6) if k < 0 else self.mx(vec)[:, k] * alpha

    def mx0(self, vec, alpha=1.0): return self.mxS(np.eye(6)[0], vec, alpha)
    def mx1(self, vec, alpha=1.0): return self.mxS(np.eye(6)[1], vec, alpha)
    def mx2(self, vec, alpha=1.0): return self.mxS(np.eye(6)[2], vec, alpha)
    def mx3(self, vec, alpha=1.0): return self.mxS(np.eye(6)[3], vec, alpha)
    def mx4(self, vec, alpha=1.0): return self.mxS(np.eye(6)[4], vec, alpha)
    def mx5(self, vec, alpha=1.0): return self.mxS(np.eye(6)[5], vec, alpha)

    def mx(self, vec):
        return -self.fx(vec).T

    def fx(self, vec):
        w, l = np.asarray(vec, dtype=np.float64)[:3], np.asarray(vec, dtype=np.float64)[3:]
        sk = lambda a: np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=np.float64)
        out = np.zeros((6, 6))
        out[:3, :3] = out[3:, 3:] = sk(w)
        out[:3, 3:] = sk(l)
        return out

    def fxS(self, S, vec, alpha=1.0):
        return -self.mxS(S, vec, alpha)

    def fxv(self, fxVec, timesVec):
        return self.fx(fxVec) @ np.asarray(timesVec, dtype=np.float64)

    def mxv(self, fxVec, timesVec):
        return self.mx(fxVec) @ np.asarray(timesVec, dtype=np.float64)
