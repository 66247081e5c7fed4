// INT-pipe microbenchmark for sm_100a: issue rates of the integer instructions the prover's kernels
// are built from (LOP3 / SHF / PRMT / IADD3 on the ALU pipe; IMAD / IMAD.HI / IMAD.WIDE on the FMA
// pipe) and of ALU+FMA mixes.  The result (lane-ops per SM-cycle) is the denominator of the INT
// roofline in DESIGN.md.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 16384

enum Op { LOP3 = 0, SHF, PRMT, IADD3, IMAD, IMADHI, IMADWIDE, MIX_LOP_IMAD, MIX_SHF_IMAD, MIX_LOP_IMADHI, MIX_LOP_IMADWIDE,
          MIX_2ALU_1FMA, NOPS };
static const char *names[NOPS] = {"lop3", "shf.r.wrap", "prmt", "iadd3", "imad.lo", "imad.hi", "imad.wide",
                                  "1 lop3 : 1 imad", "1 shf : 1 imad", "1 lop3 : 1 imad.hi", "1 lop3 : 1 imad.wide",
                                  "2 lop3 : 1 imad"};
static const int ops_per_iter[NOPS] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3};

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t x[ILP], y[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed + threadIdx.x * 77 + i; y[i] = seed * 3 + i + blockIdx.x + threadIdx.x * 0x9e3779b1u; w[i] = x[i]; }
    const uint32_t c = seed | 1;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
            if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[i]) : "r"(y[i]), "r"(c));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c), "r"(y[i]));
            if (OP == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c), "r"(y[i]));
            if (OP == IMADWIDE) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(c));
            if (OP == MIX_LOP_IMAD) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(c), "r"(c));
            }
            if (OP == MIX_SHF_IMAD) {
                asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(c), "r"(c));
            }
            if (OP == MIX_LOP_IMADHI) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(c), "r"(c));
            }
            if (OP == MIX_LOP_IMADWIDE) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
                asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(c));
            }
            if (OP == MIX_2ALU_1FMA) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(c), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= x[i] ^ y[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, uint32_t *d_out, long long *d_cyc, int warps_per_smsp) {
    const int threads = 256;
    const int blocks_per_sm = warps_per_smsp * 4 * 32 / threads;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(d_out, 12345u, d_cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d_out, 12345u, d_cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[4096];
    cudaMemcpy(cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)cyc[i];
    avg /= blocks;
    const double lane_ops = (double)blocks * threads * ITERS * ILP * ops_per_iter[OP];
    printf("%-22s warps/SMSP=%d  %8.3f ms  %7.1f Glane-ops/s  %6.1f lane-ops/SM-cycle (in-kernel clock64)\n", names[OP],
           warps_per_smsp, ms, lane_ops / ms / 1e6, lane_ops / sms / avg);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    uint32_t *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, 4096 * 256 * 4);
    cudaMalloc(&d_cyc, 4096 * 8);
    const int sms = p.multiProcessorCount;
    for (int w : {4, 8}) {
        run<LOP3>(sms, d_out, d_cyc, w);
        run<SHF>(sms, d_out, d_cyc, w);
        run<PRMT>(sms, d_out, d_cyc, w);
        run<IADD3>(sms, d_out, d_cyc, w);
        run<IMAD>(sms, d_out, d_cyc, w);
        run<IMADHI>(sms, d_out, d_cyc, w);
        run<IMADWIDE>(sms, d_out, d_cyc, w);
        run<MIX_LOP_IMAD>(sms, d_out, d_cyc, w);
        run<MIX_SHF_IMAD>(sms, d_out, d_cyc, w);
        run<MIX_LOP_IMADHI>(sms, d_out, d_cyc, w);
        run<MIX_LOP_IMADWIDE>(sms, d_out, d_cyc, w);
        run<MIX_2ALU_1FMA>(sms, d_out, d_cyc, w);
    }
    return 0;
}
