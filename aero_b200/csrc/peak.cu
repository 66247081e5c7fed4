// In-run measurement of the roofline denominator that binds this path: the issue rate of the ALU pipe
// (LOP3 / SHF / PRMT / IADD3), which is what the BLAKE2s row hash and the carry chains of the NTT
// butterflies saturate (the full instruction table is tools/int_peak.cu).
#include "kernels.cuh"

namespace aero {

constexpr int PEAK_ILP = 8, PEAK_UNROLL = 32;
__global__ void __launch_bounds__(256) alu_peak_kernel(uint32_t *out, uint32_t seed, int iters) {
    uint32_t x[PEAK_ILP], y[PEAK_ILP];
#pragma unroll
    for (int i = 0; i < PEAK_ILP; i++) {
        x[i] = seed + threadIdx.x * 77 + i;
        y[i] = seed * 3 + i + blockIdx.x + threadIdx.x * 0x9e3779b1u;
    }
    const uint32_t c = seed | 1;
    // 256 LOP3 per loop trip: the trip's own counter update and compare (ALU-pipe instructions too) stay
    // below 1 % of the stream, so the rate counted is the pipe's, not the loop's
    for (int it = 0; it < iters; it += PEAK_UNROLL) {
#pragma unroll
        for (int u = 0; u < PEAK_UNROLL; u++) {
#pragma unroll
            for (int i = 0; i < PEAK_ILP; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < PEAK_ILP; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// lane-operations per second of dependent-chain-free LOP3 streams, 8 warps per scheduler, ~1 ms per launch
// (best of three, so that a launch that ran into a clock ramp does not lower the denominator)
double measure_alu_peak(int num_sms, uint32_t *scratch /* num_sms * 4 * 256 words */, cudaStream_t s) {
    const int blocks = num_sms * 4, iters = 16384;
    static_assert(16384 % PEAK_UNROLL == 0, "whole loop trips");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    AERO_COUNT_LAUNCH(4);
    alu_peak_kernel<<<blocks, 256, 0, s>>>(scratch, 12345u, iters);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, s);
        alu_peak_kernel<<<blocks, 256, 0, s>>>(scratch, 12345u + rep, iters);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)blocks * 256 * iters * PEAK_ILP / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace aero
