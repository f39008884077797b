// msim_reduce.cu -- what follows the simulation kernel on the path: the per-ring / per-variable experiment statistics
// over a shard of replications (the [K,3] / [18,3] float64 vectors the ranks allreduce), and the device-side test hooks
// of the C ABI (Philox known answers, variate moments).
//
// Reference anchors: FactorySimulation.*_metrics (src/manusim/factory_sim.py:748-767: sum, sum / (run_until - warmup),
// or mean = sum / count per ring) and ExperimentMetrics.calculate_experiment_stats (src/manusim/metrics.py:809-816,
// 859-919: NaN runs dropped, per-variable = mean over keys first).
#include <cuda_runtime.h>
#include <math_constants.h>
#include "msim_device.cuh"
#include "philox.cuh"

namespace msim {

#define FULLMASK 0xFFFFFFFFu
#ifndef MSIM_LAUNCH
#define MSIM_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// ---------------------------------------------------------------------------------------------------------
// per-ring experiment statistics: x = per-replication metric value; out[ring] += {sum x, sum x^2, n}
__global__ void reduce_kernel(const double* __restrict__ acc_sum, const uint32_t* __restrict__ acc_cnt,
                              const int32_t* __restrict__ status, long long n_reps, int K, int P, int R, double span,
                              double* __restrict__ out) {
    const int ring = blockIdx.x;
    // metric class: 0 = sum, 1 = sum / span, 2 = mean  (metrics.py:43-67, factory_sim.py:758-763)
    int cls;
    if (ring < 3 * P) cls = 0;
    else if (ring < 11 * P) cls = 2;
    else if (ring < 11 * P + 3 * R) cls = 1;
    else cls = 2;
    double s1 = 0.0, s2 = 0.0, nn = 0.0;
    for (long long i = threadIdx.x; i < n_reps; i += blockDim.x) {
        if (status[i] != 0) continue;
        double s = acc_sum[i * K + ring];
        uint32_t c = acc_cnt[i * K + ring];
        double x;
        if (cls == 0) x = s;
        else if (cls == 1) x = s / span;
        else { if (c == 0) continue; x = s / (double)c; }
        s1 += x; s2 += x * x; nn += 1.0;
    }
    __shared__ double sh[3][32];
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_xor_sync(FULLMASK, s1, off);
        s2 += __shfl_xor_sync(FULLMASK, s2, off);
        nn += __shfl_xor_sync(FULLMASK, nn, off);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; sh[2][w] = nn; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        s1 = l < nw ? sh[0][l] : 0.0; s2 = l < nw ? sh[1][l] : 0.0; nn = l < nw ? sh[2][l] : 0.0;
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(FULLMASK, s1, off);
            s2 += __shfl_xor_sync(FULLMASK, s2, off);
            nn += __shfl_xor_sync(FULLMASK, nn, off);
        }
        if (l == 0) { out[ring * 3 + 0] += s1; out[ring * 3 + 1] += s2; out[ring * 3 + 2] += nn; }
    }
}

// per-VARIABLE experiment statistics (ExperimentMetrics' default, metrics.py:809-816): per replication the per-key
// metric values are first averaged over the keys (NaN-skipping), then reduced to {sum, sum^2, n} over replications
__global__ void reduce_variables_kernel(const double* __restrict__ acc_sum, const uint32_t* __restrict__ acc_cnt,
                                        const int32_t* __restrict__ status, long long n_reps, int K, int P, int R,
                                        double span, double* __restrict__ out) {
    const int var = blockIdx.x;                      // 0..10 product metrics, 11..14 resource metrics, 15..17 general
    int start, count, cls;
    if (var < 11) { start = var * P; count = P; cls = var < 3 ? 0 : 2; }
    else if (var < 15) { start = 11 * P + (var - 11) * R; count = R; cls = var < 14 ? 1 : 2; }
    else { start = 11 * P + 4 * R + (var - 15); count = 1; cls = 2; }
    double s1 = 0.0, s2 = 0.0, nn = 0.0;
    for (long long i = threadIdx.x; i < n_reps; i += blockDim.x) {
        if (status[i] != 0) continue;
        double tot = 0.0;
        int valid = 0;
        for (int k = 0; k < count; ++k) {
            double s = acc_sum[i * K + start + k];
            uint32_t c = acc_cnt[i * K + start + k];
            if (cls == 0) { tot += s; valid++; }
            else if (cls == 1) { tot += s / span; valid++; }
            else if (c) { tot += s / (double)c; valid++; }
        }
        if (!valid) continue;
        double x = tot / (double)valid;
        s1 += x; s2 += x * x; nn += 1.0;
    }
    __shared__ double sh[3][32];
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_xor_sync(FULLMASK, s1, off);
        s2 += __shfl_xor_sync(FULLMASK, s2, off);
        nn += __shfl_xor_sync(FULLMASK, nn, off);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; sh[2][w] = nn; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        s1 = l < nw ? sh[0][l] : 0.0; s2 = l < nw ? sh[1][l] : 0.0; nn = l < nw ? sh[2][l] : 0.0;
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(FULLMASK, s1, off);
            s2 += __shfl_xor_sync(FULLMASK, s2, off);
            nn += __shfl_xor_sync(FULLMASK, nn, off);
        }
        if (l == 0) { out[var * 3 + 0] += s1; out[var * 3 + 1] += s2; out[var * 3 + 2] += nn; }
    }
}

// per-replication metric values (the numbers of one run's metrics_*.csv, factory_sim.py:748-789): out[rep][ring] = sum,
// sum / span or sum / count (NaN when the ring is empty); every ring of a failed replication is NaN.  Elementwise,
// HBM-bound: reads 12 B and writes 8 B per (replication, ring).
__global__ void metric_values_kernel(const double* __restrict__ acc_sum, const uint32_t* __restrict__ acc_cnt,
                                     const int32_t* __restrict__ status, long long n_reps, int K, int P, int R, double span,
                                     double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // consecutive threads, consecutive elements
    if (i >= n_reps * K) return;
    const long long rep = i / K;
    const int ring = (int)(i - rep * K);
    double x = CUDART_NAN;
    if (status[rep] == 0) {
        const double s = acc_sum[i];
        const uint32_t c = acc_cnt[i];
        if (ring < 3 * P) x = s;
        else if (ring >= 11 * P && ring < 11 * P + 3 * R) x = s / span;
        else if (c) x = s / (double)c;
    }
    out[i] = x;
}

__global__ void test_philox_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = o[k];
}

__global__ void test_variates_kernel(DDist d, uint64_t seed, int stream_id, int batch_count, double* out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    PhiloxStream s{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)stream_id, (uint32_t)i};
    XU r = raw_xu(s.k0, s.k1, s.stream, s.draw, 0u);
    float v;
    if (d.kind == MSIM_DIST_CONSTANT) {
        out[i] = batch_count < 0 ? (double)(d.d0 > 0.f ? d.d0 : 0.f) : (double)(batch_count > 0 ? batch_count : 0) * d.raw0;
        return;
    }
    if (batch_count < 0) {
        v = sample_from_xu(s, d.kind, d.d0, d.d1, d.mt_d, d.mt_c, r);
        v = v > 0.0f ? v : 0.0f;
    } else if (batch_count == 0) {
        v = 0.0f;
    } else if (batch_count == 1) {
        v = sample_from_xu(s, d.kind, d.d0, d.d1, d.mt_d, d.mt_c, r);
        if (d.kind == MSIM_DIST_NORMAL && !(v > 0.0f)) v = 0.0f;
    } else {
        v = batch_many(s, d.kind, d.d0, d.d1, batch_count, r);
    }
    out[i] = (double)v;
}

int launch_reduce(const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status, int64_t n_reps, int K, int P,
                  int R, double span, double* out, void* stream) {
    MSIM_LAUNCH(reduce_kernel, K, 256, 0, (cudaStream_t)stream, acc_sum, acc_cnt, status, n_reps, K, P, R, span, out);
    return cudaGetLastError() == cudaSuccess ? 0 : MSIM_E_CUDA;
}

int launch_reduce_variables(const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status, int64_t n_reps, int K,
                            int P, int R, double span, double* out, void* stream) {
    MSIM_LAUNCH(reduce_variables_kernel, 18, 256, 0, (cudaStream_t)stream, acc_sum, acc_cnt, status, n_reps, K, P, R, span, out);
    return cudaGetLastError() == cudaSuccess ? 0 : MSIM_E_CUDA;
}

int launch_metric_values(const double* acc_sum, const uint32_t* acc_cnt, const int32_t* status, int64_t n_reps, int K, int P,
                         int R, double span, double* out, void* stream) {
    const long long blocks = ((long long)n_reps * K + 255) / 256;
    MSIM_LAUNCH(metric_values_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, acc_sum, acc_cnt, status, n_reps, K, P, R,
                span, out);
    return cudaGetLastError() == cudaSuccess ? 0 : MSIM_E_CUDA;
}

int launch_test_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int64_t n) {
    MSIM_LAUNCH(test_philox_kernel, (unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)0, ctr, key, out, n);
    return cudaGetLastError() == cudaSuccess ? 0 : MSIM_E_CUDA;
}

int launch_test_variates(DDist d, uint64_t seed, int stream_id, int batch_count, double* out, int64_t n) {
    MSIM_LAUNCH(test_variates_kernel, (unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)0, d, seed, stream_id, batch_count, out, n);
    return cudaGetLastError() == cudaSuccess ? 0 : MSIM_E_CUDA;
}

} // namespace msim
