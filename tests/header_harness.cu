// Exercises an emitted <namespace>.cuh the way a GRiD-wrapper user would: init_*, fill h_*,
// call the host functions, read h_* back; plus one kernel that calls the _inner/_device
// functions on state 0.  argv: <input.bin> <output.bin>
// input : int32 N, float q_qd_u[N*3n], float qdd[N*n]
// output: c, c_qdd, Minv, fd_qdd, dc_du, dc_du_qdd, df_du, df_du_pre, df_du_compute_only  (N states each)
//         then device-function results for state 0: df_du_dev, dc_du_inner, Minv_inner, qdd_finish, c_inner,
//         mxX(v1, k) for k = 0..5 (column 2 tripled), fx_times_v(v1, f1), fx(v1) * f1 via dot_prod
//         wide robots (HARNESS_WIDE_DEVICE_FNS) instead: df_du_dev, dc_du_inner(vaf at qdd_in), dc_du_dev(qdd_in),
//         Minv_inner, qdd_inner, df_du_dev(USE_QDD_MINV: FD's qdd, Minv)
//         last (every robot): s_XImats (72 n) after load_update_XImats_helpers on state 0
#include "grid.cuh"
#include <vector>
using namespace grid;

// what a third-party kernel written against grid.cuh does first: X_i(q), I_i into shared memory
template <typename T>
__global__ void ximats_kernel(T *out, const T *q_qd_u, const robotModel<T> *d_robotModel) {
    constexpr int n = NUM_JOINTS;
    extern __shared__ T s_dyn[];
    T *s_XImats = s_dyn, *s_q = s_XImats + 72 * n, *s_temp = s_q + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_q[i] = q_qd_u[i];
    __syncthreads();
#ifdef GRID_XIMATS_TAKES_TOPOLOGY_HELPERS
    __shared__ int s_topology_helpers[8 * n];
    load_update_XImats_helpers<T>(s_XImats, s_q, s_topology_helpers, d_robotModel, s_temp);
#else
    load_update_XImats_helpers<T>(s_XImats, s_q, d_robotModel, s_temp);
#endif
    for (int i = threadIdx.x; i < 72 * n; i += blockDim.x) out[i] = s_XImats[i];
}

#ifdef HARNESS_DEVICE_FNS
template <typename T>
__global__ void device_fn_kernel(T *out, const T *q_qd_u, const T *qdd_in, const T *Minv_in,
                                 const robotModel<T> *d_robotModel, T gravity) {
    constexpr int n = NUM_JOINTS;
    __shared__ T s_q[n], s_qd[n], s_u[n], s_qdd[n], s_c[n], s_vaf[18 * n], s_Minv[n * n], s_out[2 * n * n], s_fin[n];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_q[i] = q_qd_u[i]; s_qd[i] = q_qd_u[n + i]; s_u[i] = q_qd_u[2 * n + i]; s_qdd[i] = qdd_in[i];
    }
    __syncthreads();
    T *o = out;
    forward_dynamics_gradient_device<T>(s_out, s_q, s_qd, s_u, d_robotModel, gravity);
    for (int i = threadIdx.x; i < 2 * n * n; i += blockDim.x) o[i] = s_out[i];
    o += 2 * n * n;
    __syncthreads();
    inverse_dynamics_inner<T>(s_c, s_vaf, s_q, s_qd, s_qdd, (T *)nullptr, (T *)nullptr, gravity);
    inverse_dynamics_gradient_inner<T>(s_out, s_q, s_qd, s_vaf, (T *)nullptr, (T *)nullptr, gravity);
    for (int i = threadIdx.x; i < 2 * n * n; i += blockDim.x) o[i] = s_out[i];
    o += 2 * n * n;
    __syncthreads();
    direct_minv_inner<T>(s_Minv, s_q, (T *)nullptr, (T *)nullptr);
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) o[i] = s_Minv[i];
    o += n * n;
    __syncthreads();
    // c at qdd = 0, then qdd = Minv (u - c)
    inverse_dynamics_device<T>(s_c, s_q, s_qd, d_robotModel, gravity);
    forward_dynamics_finish<T>(s_fin, s_u, s_c, s_Minv);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { o[i] = s_fin[i]; o[n + i] = s_c[i]; }
    o += 2 * n;
    __syncthreads();
    // spatial algebra helpers: (v x) e_k for k = 0..5 via mxX, v x* f via fx_times_v and via fx + dot_prod
    if (threadIdx.x == 0) {
        const T *v = s_vaf + 6, *f = s_vaf + 12 * n + 6;          // v and f of joint 1
        for (int k = 0; k < 6; k++) mxX<T>(o + 6 * k, v, k);
        mx2_peq_scaled<T>(o + 12, v, static_cast<T>(2));           // column 2 becomes 3 * mx2(v)
        fx_times_v<T>(o + 36, v, f);
        T M[36];
        fx<T>(M, v);
        for (int r = 0; r < 6; r++) o[42 + r] = dot_prod<T, 6, 6, 1>(M + r, f);
    }
}

#endif

#ifdef HARNESS_WIDE_DEVICE_FNS
// Robots whose single-thread programs are too large (Atlas, 64-link chain): the _inner/_device functions are the
// wide CTA-per-state bodies behind the reference signatures.  _inner takes its scratch from the caller (s_temp,
// <alg>_inner_temp_mem_size() floats = FD_DU_DYNAMIC_SHARED_MEM_COUNT here), _device uses the block's dynamic
// shared memory; they run one after the other, so one dynamic block serves both.
template <typename T>
__global__ void __launch_bounds__(SUGGESTED_THREADS)
wide_device_fn_kernel(T *out, const T *q_qd_u, const T *qdd_in, const robotModel<T> *d_robotModel, T gravity) {
    constexpr int n = NUM_JOINTS;
    extern __shared__ float4 s_dyn4[];
    T *s_temp = reinterpret_cast<T *>(s_dyn4);
    __shared__ T s_q[n], s_qd[n], s_u[n], s_qdd[n], s_c[n], s_vaf[18 * n], s_Minv[n * n], s_out[2 * n * n], s_fin[n];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_q[i] = q_qd_u[i]; s_qd[i] = q_qd_u[n + i]; s_u[i] = q_qd_u[2 * n + i]; s_qdd[i] = qdd_in[i];
    }
    __syncthreads();
    T *o = out;
    forward_dynamics_gradient_device<T>(s_out, s_q, s_qd, s_u, d_robotModel, gravity);          // df_du
    for (int i = threadIdx.x; i < 2 * n * n; i += blockDim.x) o[i] = s_out[i];
    o += 2 * n * n;
    __syncthreads();
    inverse_dynamics_inner<T>(s_c, s_vaf, s_q, s_qd, s_qdd, (T *)nullptr, (T *)nullptr, gravity); // v, a, f at qdd_in
    inverse_dynamics_gradient_inner<T>(s_out, s_q, s_qd, s_vaf, (T *)nullptr, s_temp, gravity);  // dc_du from vaf
    for (int i = threadIdx.x; i < 2 * n * n; i += blockDim.x) o[i] = s_out[i];
    o += 2 * n * n;
    __syncthreads();
    inverse_dynamics_gradient_device<T>(s_out, s_q, s_qd, s_qdd, d_robotModel, gravity);         // dc_du (qdd given)
    for (int i = threadIdx.x; i < 2 * n * n; i += blockDim.x) o[i] = s_out[i];
    o += 2 * n * n;
    __syncthreads();
    direct_minv_inner<T>(s_Minv, s_q, (T *)nullptr, s_temp);
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) o[i] = s_Minv[i];
    o += n * n;
    __syncthreads();
    forward_dynamics_inner<T>(s_fin, s_q, s_qd, s_u, (T *)nullptr, s_temp, gravity);             // qdd
    for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = s_fin[i];
    o += n;
    __syncthreads();
    // USE_QDD_MINV_FLAG overload fed with FD's own qdd and Minv: must reproduce df_du
    forward_dynamics_gradient_device<T>(s_out, s_q, s_qd, s_fin, s_Minv, d_robotModel, gravity);
    for (int i = threadIdx.x; i < 2 * n * n; i += blockDim.x) o[i] = s_out[i];
}
#endif

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    const int n = NUM_JOINTS;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("input"); return 2; }
    int N = 0;
    if (fread(&N, sizeof(int), 1, f) != 1) return 2;
    std::vector<float> in(size_t(N) * 3 * n), qdd(size_t(N) * n);
    if (fread(in.data(), sizeof(float), in.size(), f) != in.size()) return 2;
    if (fread(qdd.data(), sizeof(float), qdd.size(), f) != qdd.size()) return 2;
    fclose(f);
    const float gravity = 9.81f;
    robotModel<float> *d_robotModel = init_robotModel<float>();
    cudaStream_t *streams = init_grid<float>();
    gridData<float> *hd = init_gridData<float>(N);
    dim3 blocks(N), threads(SUGGESTED_THREADS);
    memcpy(hd->h_q_qd_u, in.data(), in.size() * sizeof(float));
    FILE *o = fopen(argv[2], "wb");
    auto dump = [&](const float *p, size_t words) { fwrite(p, sizeof(float), words, o); };

    inverse_dynamics<float>(hd, d_robotModel, gravity, N, blocks, threads, streams);
    dump(hd->h_c, size_t(N) * n);
    memcpy(hd->h_qdd, qdd.data(), qdd.size() * sizeof(float));
    inverse_dynamics<float, true>(hd, d_robotModel, gravity, N, blocks, threads, streams);
    dump(hd->h_c, size_t(N) * n);
    direct_minv<float>(hd, d_robotModel, N, blocks, threads, streams);
    dump(hd->h_Minv, size_t(N) * n * n);
    std::vector<float> Minv(hd->h_Minv, hd->h_Minv + size_t(N) * n * n);
    forward_dynamics<float>(hd, d_robotModel, gravity, N, blocks, threads, streams);
    dump(hd->h_qdd, size_t(N) * n);
    std::vector<float> fd_qdd(hd->h_qdd, hd->h_qdd + size_t(N) * n);
    inverse_dynamics_gradient<float>(hd, d_robotModel, gravity, N, blocks, threads, streams);
    dump(hd->h_dc_du, size_t(N) * 2 * n * n);
    memcpy(hd->h_qdd, qdd.data(), qdd.size() * sizeof(float));
    inverse_dynamics_gradient<float, true>(hd, d_robotModel, gravity, N, blocks, threads, streams);
    dump(hd->h_dc_du, size_t(N) * 2 * n * n);
    forward_dynamics_gradient<float>(hd, d_robotModel, gravity, N, blocks, threads, streams);
    dump(hd->h_df_du, size_t(N) * 2 * n * n);
    // USE_QDD_MINV_FLAG: feed FD's own qdd and Minv back in
    memcpy(hd->h_qdd, fd_qdd.data(), fd_qdd.size() * sizeof(float));
    memcpy(hd->h_Minv, Minv.data(), Minv.size() * sizeof(float));
    forward_dynamics_gradient<float, true>(hd, d_robotModel, gravity, N, blocks, threads, streams);
    dump(hd->h_df_du, size_t(N) * 2 * n * n);
    // compute-only: inputs are already on the device from the previous call
    gpuErrchk(cudaMemset(hd->d_df_du, 0, size_t(N) * 2 * n * n * sizeof(float)));
    forward_dynamics_gradient_compute_only<float>(hd, d_robotModel, gravity, N, blocks, threads);
    gpuErrchk(cudaMemcpy(hd->h_df_du, hd->d_df_du, size_t(N) * 2 * n * n * sizeof(float), cudaMemcpyDeviceToHost));
    dump(hd->h_df_du, size_t(N) * 2 * n * n);

#ifdef HARNESS_DEVICE_FNS
    const size_t dev_words = 2 * n * n * 2 + n * n + 2 * n + 48;
    float *d_dev;
    gpuErrchk(cudaMalloc(&d_dev, dev_words * sizeof(float)));
    device_fn_kernel<float><<<1, 64>>>(d_dev, hd->d_q_qd_u, hd->d_qdd, hd->d_Minv, d_robotModel, gravity);
    gpuErrchk(cudaDeviceSynchronize());
    std::vector<float> dev(dev_words);
    gpuErrchk(cudaMemcpy(dev.data(), d_dev, dev_words * sizeof(float), cudaMemcpyDeviceToHost));
    dump(dev.data(), dev_words);
    cudaFree(d_dev);
#endif
#ifdef HARNESS_WIDE_DEVICE_FNS
    {
        const size_t dev_words = 2 * n * n * 4 + n * n + n;
        const size_t smem = FD_DU_DYNAMIC_SHARED_MEM_COUNT * sizeof(float);
        float *d_dev, *d_qdd_in;
        gpuErrchk(cudaMalloc(&d_dev, dev_words * sizeof(float)));
        gpuErrchk(cudaMalloc(&d_qdd_in, n * sizeof(float)));
        gpuErrchk(cudaMemcpy(d_qdd_in, qdd.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        gpuErrchk(cudaFuncSetAttribute(wide_device_fn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wide_device_fn_kernel<float><<<1, SUGGESTED_THREADS, smem>>>(d_dev, hd->d_q_qd_u, d_qdd_in, d_robotModel, gravity);
        gpuErrchk(cudaGetLastError());
        gpuErrchk(cudaDeviceSynchronize());
        std::vector<float> dev(dev_words);
        gpuErrchk(cudaMemcpy(dev.data(), d_dev, dev_words * sizeof(float), cudaMemcpyDeviceToHost));
        dump(dev.data(), dev_words);
        cudaFree(d_dev);
        cudaFree(d_qdd_in);
    }
#endif
    {
        float *d_xi;
        gpuErrchk(cudaMalloc(&d_xi, 72 * n * sizeof(float)));
        ximats_kernel<float><<<1, 96, (72 * n + 3 * n) * sizeof(float)>>>(d_xi, hd->d_q_qd_u, d_robotModel);
        gpuErrchk(cudaDeviceSynchronize());
        std::vector<float> xi(72 * n);
        gpuErrchk(cudaMemcpy(xi.data(), d_xi, xi.size() * sizeof(float), cudaMemcpyDeviceToHost));
        dump(xi.data(), xi.size());
        cudaFree(d_xi);
    }
    fclose(o);
    close_grid<float>(streams, d_robotModel, hd);
    printf("ok N=%d n=%d\n", N, n);
    return 0;
}
