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
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef GRID_B200_H
#define GRID_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GRID_B200_ABI_VERSION 1

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

/* ---- gridData-style handle (reference init_gridData / init_grid / close_grid) ------- */
typedef struct grid_data grid_data;

/* init_gridData<T>(NUM_TIMESTEPS) + init_grid<T>()  GRiDCodeGenerator.py:116-189.
 * Host buffers are pinned (the reference uses malloc, GRiDCodeGenerator.py:123-137). */
grid_data *grid_data_create(int max_timesteps);
/* close_grid<T>  GRiDCodeGenerator.py:191-203 */
void grid_data_destroy(grid_data *hd);
int grid_data_capacity(const grid_data *hd);

/* field access; names are the reference's gridData members:
 * "h_q_qd_u","h_q_qd","h_q","h_c","h_Minv","h_qdd","h_dc_du","h_df_du" and the d_* twins.
 * Returns NULL for an unknown name. */
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

/* ---- measurement helpers (not part of the reference API) ------------------------------ */
/* Runs an FFMA-only microbenchmark on the current device and returns the measured FP32
 * (non-tensor) throughput in TFLOP/s (the roofline denominator); <0 on error. */
double grid_measure_fp32_tflops(int repeats);
/* Times `reps` back-to-back launches of one algorithm ("id","minv","fd","id_grad","fd_grad", inputs
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
