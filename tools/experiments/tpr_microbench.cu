// tpr_microbench.cu -- how fast COULD a thread-per-replication engine run on a B200?  
//
// The production kernel gives one WARP to a replication (state in shared memory, lane 0 runs the handlers: 6-7 lanes
// active per instruction, instruction-fetch bound).  The alternative gives one THREAD to a replication: 32 replications
// per warp, every fetched instruction serves up to 32 lanes.  Then the state cannot live in shared memory (5.2-7.9 KB per
// replication, 227 KB per SM = 36-40 replications = 36-40 threads per SM), it has to live in global memory, and what
// bounds the engine is the memory system: an event of the factory model touches ~20 words of its replication's state,
// half of them dependent on the previous one (list links, order -> product -> route step -> resource).
//
// This micro-benchmark measures that bound, generously: every lane owns a slab of SLAB bytes (word-interleaved across the
// warp's 32 lanes, or contiguous per lane), and per "event" does K accesses inside its slab -- `dep` of them a dependent
// chain (address from the previous value), the rest independent, half of them read-modify-write -- plus a few ALU
// instructions; all 32 lanes converged, no handler divergence, no calendar, no log streaming, no variates.  Events/s of
// this loop is an UPPER bound for a thread-per-replication engine with that state footprint.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tpr_microbench tools/experiments/tpr_microbench.cu
//   ./tpr_microbench            # sweeps slab size x resident warps per SM x layout, prints JSON lines
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <bool INTERLEAVED>
__global__ void tpr_kernel(uint32_t* __restrict__ state, int slab_words, int events, int K, int dep, unsigned long long* out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    // word w of lane l: interleaved -> base + (w * 32 + l); contiguous -> base + l * slab_words + w
    uint32_t* base = state + warp * 32ll * slab_words;
    uint32_t x = (uint32_t)(warp * 32 + lane) * 2654435761u + 12345u;
    uint32_t acc = 0;
    for (int e = 0; e < events; ++e) {
        uint32_t idx = x % (uint32_t)slab_words;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            uint32_t* p = INTERLEAVED ? base + ((long long)idx * 32 + lane) : base + ((long long)lane * slab_words + idx);
            uint32_t v = *p;
            if (k & 1) *p = v + 1u;                       // half of the accesses update the state
            acc += v;
            x = x * 1664525u + 1013904223u;
            // a dependent access takes its address from the loaded value, an independent one from the generator
            idx = (k < dep ? (v ^ x) : x) % (uint32_t)slab_words;
        }
        acc = acc * 31u + (x >> 7);                       // a little ALU work per event
    }
    if (acc == 0xDEADBEEFu) out[0] = acc;                 // keep the loop alive
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int K = 20, dep = 10, events = argc > 1 ? atoi(argv[1]) : 2000;
    unsigned long long* d_out;
    cudaMalloc(&d_out, 8);
    const int slabs[] = {1536, 3072, 5632};               // bytes per replication: a compacted hot state ... the present slab
    const int warps_per_sm[] = {4, 8, 16, 32, 64};
    for (int layout = 0; layout < 2; ++layout)
        for (int s : slabs)
            for (int w : warps_per_sm) {
                const int slab_words = s / 4;
                const long long n_warps = (long long)sms * w;
                const size_t bytes = (size_t)n_warps * 32 * s;
                uint32_t* d_state;
                if (cudaMalloc(&d_state, bytes) != cudaSuccess) { printf("{\"error\": \"alloc %zu\"}\n", bytes); continue; }
                cudaMemset(d_state, 1, bytes);
                const int block = 128;                    // 4 warps per CTA
                const int grid = (int)(n_warps / 4);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                for (int rep = 0; rep < 2; ++rep) {       // first pass warms up
                    cudaEventRecord(e0);
                    if (layout == 0) tpr_kernel<true><<<grid, block>>>(d_state, slab_words, events, K, dep, d_out);
                    else tpr_kernel<false><<<grid, block>>>(d_state, slab_words, events, K, dep, d_out);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                const double ev = (double)n_warps * 32 * events;
                printf("{\"layout\": \"%s\", \"slab_bytes\": %d, \"warps_per_sm\": %d, \"replications\": %lld, \"state_MB\": %.1f, "
                       "\"accesses_per_event\": %d, \"dependent\": %d, \"ms\": %.2f, \"events_per_s\": %.4e, \"GB_per_s_sectors\": %.1f}\n",
                       layout == 0 ? "interleaved" : "contiguous", s, w, n_warps * 32, bytes / 1e6, K, dep, ms, ev / (ms * 1e-3),
                       ev * K * 32.0 * 1.5 / (ms * 1e-3) / 1e9);
                fflush(stdout);
                cudaFree(d_state);
            }
    return 0;
}
