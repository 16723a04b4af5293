/*
 * grid_b200.h - C ABI of a robot-specialised GRiD dynamics library (libgrid_<robot>.so).
 *
 * One shared library is generated and compiled per robot (topology, X_tree and inertias
 * are baked in at compile time).  The entry points are what a binding of the
 * reference's emitted header would bind - the reference emits C++ templates in
 * `namespace grid` (GRiDCodeGenerator.py:241-310); each function below cites the
 * emitted host function it replaces.  All matrices are column-major, all state arrays
 * are state-major exactly as in the reference's gridData (GRiDCodeGenerator.py:94-114).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; grid_last_error() then
 *     holds a message (the reference calls exit() instead, GRiDCodeGenerator.py:211-218);
 *   - `gravity` is the positive magnitude (9.81f), applied as a_base = X[:,5]*gravity
 *     (algorithms/_inverse_dynamics.py:123);
 *   - `*_device` entry points take DEVICE pointers, are asynchronous on `stream`
 *     (a cudaStream_t passed as void*, NULL = default stream) and never touch the host:
 *     they are the reference's `_compute_only` host functions (mode 2);
 *   - `grid_*` entry points taking a grid_data* are the reference's mode-0 host functions:
 *     H2D from the handle's pinned h_* inputs, kernel, D2H into the pinned h_* outputs,
 *     then synchronise;
 *   - there is no CPU fallback: without a CUDA device every compute call fails;
 *   - several devices in one process are supported: every launcher keeps its cached state (shared-
 *     memory opt-in, occupancy caps) per device and uses the CURRENT device (cudaSetDevice); a
 *     grid_data handle belongs to the device that was current when it was created and its host
 *     functions fail with a message when called under another device.
 */
#ifndef GRID_B200_H
#define GRID_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GRID_B200_ABI_VERSION 2

/* ---- identity --------------------------------------------------------------------- */
int grid_abi_version(void);
int grid_num_joints(void);              /* const int NUM_JOINTS  (GRiDCodeGenerator.py:75)   */
const char *grid_robot_name(void);
const char *grid_robot_hash(void);      /* hash of every robot parameter compiled in         */
const char *grid_last_error(void);      /* message of the last failure on this thread        */
/* "+"-joined kernel families available: "tps" (thread per state), "wps" (CTA per state),
 * "cps" (lane per column, latency), or "none"; for
 * alg in {"id","minv","fd","id_grad","fd_grad"} */
const char *grid_kernel_kind(const char *alg);
/* live traced FP32 ops per state of the straight-line kernels (0 when not tps) */
long long grid_traced_flops(const char *alg);

/* ---- device-pointer entry points (reference *_compute_only, mode 2) ---------------- */

/* inverse_dynamics<T,USE_QDD_FLAG,USE_COMPRESSED_MEM>  algorithms/_inverse_dynamics.py:423-495
 * d_q_qd: [q(n) | qd(n) | ...] per state with `stride` floats between states (3n for the
 * d_q_qd_u buffer, 2n for the compressed d_q_qd buffer); d_qdd: n per state or NULL;
 * d_c: n per state. */
int grid_inverse_dynamics_device(float *d_c, const float *d_q_qd, int stride, const float *d_qdd,
                                 int num_timesteps, float gravity, void *stream);

/* direct_minv<T,USE_COMPRESSED_MEM>  algorithms/_direct_minv.py:456-517
 * d_Minv: n*n per state, column-major, upper triangle filled, strict lower triangle 0. */
int grid_direct_minv_device(float *d_Minv, const float *d_q, int stride, int num_timesteps, void *stream);

/* forward_dynamics<T>  algorithms/_forward_dynamics.py:196-252
 * d_q_qd_u: [q | qd | u] per state; d_qdd: n per state. */
int grid_forward_dynamics_device(float *d_qdd, const float *d_q_qd_u, int stride, int num_timesteps,
                                 float gravity, void *stream);

/* inverse_dynamics_gradient<T,USE_QDD_FLAG,USE_COMPRESSED_MEM>
 * algorithms/_inverse_dynamics_gradient.py:762-834
 * d_dc_du: 2*n*n per state = column-major n x 2n [dc/dq | dc/dqd]. */
int grid_inverse_dynamics_gradient_device(float *d_dc_du, const float *d_q_qd, int stride, const float *d_qdd,
                                          int num_timesteps, float gravity, void *stream);

/* forward_dynamics_gradient<T,USE_QDD_MINV_FLAG>  algorithms/_forward_dynamics_gradient.py:179-242
 * d_qdd and d_Minv both NULL: inputs are (q, qd, u) and everything is computed in one kernel;
 * both non-NULL: inputs are (q, qd) + qdd (n) + Minv (n*n, upper triangle read symmetrically).
 * d_df_du: 2*n*n per state = column-major n x 2n [dqdd/dq | dqdd/dqd]. */
int grid_forward_dynamics_gradient_device(float *d_df_du, const float *d_q_qd_u, int stride, const float *d_qdd,
                                          const float *d_Minv, int num_timesteps, float gravity, void *stream);

/* ---- consumers fused after the FD gradient (beyond the reference) -------------
 * The reference's forward_dynamics_gradient stops at writing df_du (2n^2 floats per state) to
 * global memory and copying it to the host (algorithms/_forward_dynamics_gradient.py:159-161,
 * 235-238).  These two entry points trace the consumer INTO the gradient kernel, so df_du never
 * leaves the registers.  Explicit Euler with step dt on x = [q; qd]:
 *     x+ = x + dt [qd; qdd(q,qd,u)],   A = dx+/dx = [[I, dt I], [dt dqdd/dq, I + dt dqdd/dqd]],
 *     B = dx+/du = [[0], [dt Minv]].
 *
 * grid_forward_dynamics_gradient_vjp_device: costate step of a shooting / DDP backward pass.
 *   d_lambda: [lam_q (n) | lam_v (n)] per state;  d_out: 5n per state =
 *   [x+ (2n) | A^T lam (2n) | B^T lam (n)].
 * grid_forward_dynamics_linearize_device: the non-constant blocks of A and B.
 *   d_out: 2n + 3n^2 per state = [x+ (2n) | A21 = dt dqdd/dq | A22 = I + dt dqdd/dqd | B2 = dt Minv],
 *   the three n x n blocks column-major, B2 full symmetric.
 * Available when the robot's FD gradient is served by the thread-per-state or phase-split kernels
 * (grid_kernel_kind("fd_vjp") / ("fd_lin") != "none"); otherwise they fail with a message. */
int grid_forward_dynamics_gradient_vjp_device(float *d_out, const float *d_q_qd_u, int stride, const float *d_lambda,
                                              int num_timesteps, float dt, float gravity, void *stream);
int grid_forward_dynamics_linearize_device(float *d_out, const float *d_q_qd_u, int stride, int num_timesteps, float dt,
                                           float gravity, void *stream);

/* ---- further algorithms (what the GRiD family added after this reference) ----
 * No counterpart in /root/reference; pinned to it through its own algorithms (tests/golden/*_mass.npz: the columns
 * of M from the reference's RNEA; M Minv = I; ABA = Minv (u - c)).
 * grid_crba_device: joint-space mass matrix M(q) by the composite-rigid-body algorithm.
 *   d_q: q per state (stride words apart); d_M: n*n per state, column-major, BOTH triangles.
 * grid_aba_device: qdd = FD(q, qd, u) by the articulated-body algorithm (O(n), no Minv); same inputs, output and
 *   conventions (gravity, joint damping) as grid_forward_dynamics_device.  Where it is the faster program
 *   grid_forward_dynamics_device itself runs it for large batches (grid_kernel_kind("fd@large") == "tps(aba)").
 * Available when grid_kernel_kind("crba") / ("aba") != "none"; otherwise they fail with a message. */
int grid_crba_device(float *d_M, const float *d_q, int stride, int num_timesteps, void *stream);
int grid_aba_device(float *d_qdd, const float *d_q_qd_u, int stride, int num_timesteps, float gravity, void *stream);

/* ---- gridData-style handle (reference init_gridData / init_grid / close_grid) ------- */
typedef struct grid_data grid_data;

/* init_gridData<T>(NUM_TIMESTEPS) + init_grid<T>()  GRiDCodeGenerator.py:116-189.
 * Host buffers are pinned (the reference uses malloc, GRiDCodeGenerator.py:123-137). */
grid_data *grid_data_create(int max_timesteps);
/* close_grid<T>  GRiDCodeGenerator.py:191-203 */
void grid_data_destroy(grid_data *hd);
int grid_data_capacity(const grid_data *hd);

/* field access; names are the reference's gridData members:
 * "h_q_qd_u","h_q_qd","h_q","h_c","h_Minv","h_qdd","h_dc_du","h_df_du" and the d_* twins, plus the
 * consumer buffers "h_lambda" (2n), "h_vjp" (5n), "h_lin" (2n + 3n^2) and their d_* twins, which are
 * allocated on first access.  Returns NULL for an unknown name. */
float *grid_data_ptr(grid_data *hd, const char *field);

/* mode-0 host functions: copy h_* -> d_*, run, copy result d_* -> h_*, synchronise.
 * use_qdd != 0 reads h_qdd as an input (USE_QDD_FLAG); use_qdd_minv != 0 reads h_qdd and
 * h_Minv (USE_QDD_MINV_FLAG).  Inputs are read from h_q_qd_u (stride 3n) unless
 * compressed != 0 (USE_COMPRESSED_MEM: h_q_qd stride 2n / h_q stride n). */
int grid_inverse_dynamics(grid_data *hd, int num_timesteps, float gravity, int use_qdd, int compressed);
int grid_direct_minv(grid_data *hd, int num_timesteps, int compressed);
int grid_forward_dynamics(grid_data *hd, int num_timesteps, float gravity);
int grid_inverse_dynamics_gradient(grid_data *hd, int num_timesteps, float gravity, int use_qdd, int compressed);
int grid_forward_dynamics_gradient(grid_data *hd, int num_timesteps, float gravity, int use_qdd_minv);
/* host forms of the fused consumers: h_q_qd_u (+ h_lambda) in, h_vjp / h_lin out */
int grid_forward_dynamics_gradient_vjp(grid_data *hd, int num_timesteps, float dt, float gravity);
int grid_forward_dynamics_linearize(grid_data *hd, int num_timesteps, float dt, float gravity);

/* ---- repeated fixed-shape calls: CUDA graph ------------------------------------------------
 * grid_graph_create captures ONE launch of `alg` ("id","minv","fd","id_grad","fd_grad","fd_vjp",
 * "fd_lin") on the given device buffers (d_in1 = qdd or lambda, d_in2 = Minv, NULL when unused);
 * grid_graph_launch replays it on `stream`.  For phase-split kernels the graph holds the scratch
 * allocation, the ticket memset and both kernels.  The buffers must stay valid while the graph lives. */
typedef struct grid_graph grid_graph;
grid_graph *grid_graph_create(const char *alg, float *d_out, const float *d_in, int stride, const float *d_in1,
                              const float *d_in2, int num_timesteps, float dt, float gravity);
int grid_graph_launch(grid_graph *g, void *stream);
void grid_graph_destroy(grid_graph *g);

/* ---- options ------------------------------------------------------------------------------
 * The environment variables GRID_FORCE_KERNEL (tps|wps|cps|pipe), GRID_PIPE_MODE (staged|fused),
 * GRID_PIPE_CHUNK (states), GRID_PIPE_WARPS (CTA width of the phase-split kernels), GRID_PIPE_STAGGER_NS
 * (experiment: start offset between CTAs that share an SM), GRID_PIPE_ORDER_CHUNK (experiment: stage-1 items in
 * chunk-major order, states per chunk) and GRID_PIPE_ONLY_TASK (profiling: 100 * stage + task runs that one task
 * program only - the results are then incomplete) are read ONCE, at the first launch; afterwards they change only through
 * this call (value NULL or "" restores the default).  Nothing on the launch path calls getenv. */
int grid_set_option(const char *key, const char *value);

/* ---- measurement helpers (not part of the reference API) ------------------------------ */
/* Runs an FFMA-only microbenchmark on the current device and returns the measured FP32
 * (non-tensor) throughput in TFLOP/s (the roofline denominator); <0 on error. */
double grid_measure_fp32_tflops(int repeats);
/* Times `reps` back-to-back launches of one algorithm ("id","minv","fd","id_grad","fd_grad"; with the
 * suffix "@graph" the launch is captured once into a CUDA graph and the timed launches replay it; inputs
 * as for the matching *_device call with d_qdd = d_Minv = NULL) with one CUDA event pair per launch,
 * recorded from C so that no interpreter time sits between the events; h_us receives the `reps`
 * per-launch durations in microseconds.  This is how the N = 128 latency is measured.  alg = "noop" times an
 * empty kernel the same way: the floor of the method (event pair + launch), reported beside the latency. */
int grid_time_launches(const char *alg, float *d_out, const float *d_in, int stride, int num_timesteps,
                       float gravity, int reps, float *h_us);
/* Number of kernels this library has launched since load (the bench's gpu_launches). */
long long grid_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GRID_B200_H */
