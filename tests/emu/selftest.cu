// selftest.cu -- TEST INFRASTRUCTURE: semantics of the SIMT shim itself (tests/emu/include/cuda_runtime.h).
// Built by tests/test_emu_kernel.py into an executable; prints one JSON line.
#include <cuda_runtime.h>

__global__ void warp_prims(unsigned* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* o = out + (blockIdx.x * (blockDim.x >> 5) + warp) * 8;
    unsigned v = (unsigned)(lane * 7 + warp);
    unsigned sum = v;
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
    unsigned mn = __reduce_min_sync(0xFFFFFFFFu, 100u - (unsigned)lane);
    unsigned add = __reduce_add_sync(0xFFFFFFFFu, 1u);
    unsigned bal = __ballot_sync(0xFFFFFFFFu, lane % 3 == 0);
    unsigned grp = __match_any_sync(0xFFFFFFFFu, (unsigned)(lane & 3));
    unsigned from5 = __shfl_sync(0xFFFFFFFFu, v, 5);
    double d = __shfl_sync(0xFFFFFFFFu, (double)lane + 0.5, 31);
    if (lane == 9) { o[0] = sum; o[1] = mn; o[2] = add; o[3] = bal; o[4] = grp; o[5] = from5; o[6] = (unsigned)(d * 2.0); }
    // lanes diverge between collectives and reconverge at the next one
    if (lane & 1) { for (volatile int i = 0; i < lane; ++i) {} }
    __syncwarp();
    if (lane == 0) o[7] = 0xC0FFEEu;
}

__global__ void block_sum(const int* in, int* out) {
    __shared__ int sh[8];
    int v = in[blockIdx.x * blockDim.x + threadIdx.x];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
        out[blockIdx.x] = t;
    }
}

// lane 3 waits at the block barrier while the other 31 lanes wait on a warp collective: neither can complete
// (a hang on the GPU, an error here)
__global__ void divergent(int* out) {
    const int lane = threadIdx.x & 31;
    if (lane == 3) __syncthreads();
    out[lane] = __shfl_sync(0xFFFFFFFFu, lane, 0);
}

// lanes meet at DIFFERENT warp primitives (undefined on the GPU): reported as an error
__global__ void mismatched(int* out) {
    const int lane = threadIdx.x & 31;
    if (lane & 1) out[lane] = __shfl_sync(0xFFFFFFFFu, lane, 0);
    else out[lane] = (int)__ballot_sync(0xFFFFFFFFu, 1);
}

int main() {
    int fails = 0;
    {
        std::vector<unsigned> out(3 * 2 * 8, 0u);
        unsigned* o = out.data();
        MSIM_LAUNCH(warp_prims, 3, 64, 0, nullptr, o);
        if (cudaGetLastError() != cudaSuccess) fails++;
        for (int b = 0; b < 3; ++b)
            for (int w = 0; w < 2; ++w) {
                const unsigned* r = o + (b * 2 + w) * 8;
                unsigned sum = 0;
                for (int l = 0; l < 32; ++l) sum += (unsigned)(l * 7 + w);
                if (r[0] != sum || r[1] != 69u || r[2] != 32u || r[3] != 0x49249249u || r[4] != 0x22222222u ||
                    r[5] != (unsigned)(35 + w) || r[6] != 63u || r[7] != 0xC0FFEEu)
                    fails++;
            }
    }
    {
        std::vector<int> in(5 * 256), out(5, 0);
        long long want[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5 * 256; ++i) { in[i] = (i * 37) % 101 - 50; want[i / 256] += in[i]; }
        const int* pi = in.data();
        int* po = out.data();
        MSIM_LAUNCH(block_sum, 5, 256, 0, nullptr, pi, po);
        if (cudaGetLastError() != cudaSuccess) fails++;
        for (int b = 0; b < 5; ++b) if (out[b] != (int)want[b]) fails++;
    }
    int detected = 0;
    {
        std::vector<int> out(32, -1);
        int* po = out.data();
        MSIM_LAUNCH(divergent, 1, 32, 0, nullptr, po);
        detected = cudaGetLastError() != cudaSuccess;
    }
    int mismatch = 0;
    {
        std::vector<int> out(32, -1);
        int* po = out.data();
        MSIM_LAUNCH(mismatched, 1, 32, 0, nullptr, po);
        mismatch = cudaGetLastError() != cudaSuccess;
    }
    printf("{\"fails\": %d, \"divergence_detected\": %d, \"mismatch_detected\": %d}\n", fails, detected, mismatch);
    return fails == 0 && detected && mismatch ? 0 : 1;
}
