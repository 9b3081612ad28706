// tools/ubench_dpx.cu -- issue-rate microbenchmark of the instructions the DP kernels are built from (sm_100a).
// Prints warp-instructions per cycle per SM for independent-chain streams of each op (8 chains per thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dpx tools/ubench_dpx.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

#define ITER 4096
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a[8], acc0 = 0, acc1 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 17 + i * 1237;
    const uint32_t c1 = seed | 0x00030003u, c2 = seed ^ 0x7fff7fffu;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = __vmaxs2(a[i], c2 + i);                               // VIMNMX.S16x2
            if (OP == 1) a[i] = __viaddmax_s16x2(a[i], c1, c2);                        // VIADDMNMX.S16x2
            if (OP == 2) a[i] = __vadd2(a[i], c1);                                     // VIADD.16x2
            if (OP == 3) a[i] = a[i] - c1;                                             // plain 32-bit sub
            if (OP == 4) a[i] = __vimax3_s16x2(a[i], c1, c2 + i);                      // VIMNMX3
            if (OP == 5) {                                                            // VIMNMX with predicates + 2 predicated adds
                uint32_t v;
                asm("{\n\t.reg .pred plo, phi;\n\t.reg .s16 a0, a1, b0, b1;\n\t"
                    "max.s16x2 %0, %3, %4;\n\tmov.b32 {a0, a1}, %0;\n\tmov.b32 {b0, b1}, %3;\n\t"
                    "setp.eq.s16 plo, a0, b0;\n\tsetp.eq.s16 phi, a1, b1;\n\t"
                    "@plo add.u32 %1, %1, %5;\n\t@phi add.u32 %2, %2, %6;\n\t}"
                    : "=r"(v), "+r"(acc0), "+r"(acc1) : "r"(a[i]), "r"(c2 + i), "r"(1u << i), "r"(256u << i));
                a[i] = v + 1;
            }
            if (OP == 6) { uint32_t d; asm("lop3.b32 %0, %1, %2, 0, 0xC3;" : "=r"(d) : "r"(a[i]), "r"(c2)); a[i] = d; }
            if (OP == 7) a[i] = a[i] * 3 + c1;                                         // IMAD
            if (OP == 8) {                                                            // HMNMX2 (fp16x2 max on int patterns)
                __half2 x = *reinterpret_cast<__half2 *>(&a[i]); uint32_t cc = (c2 + it) & 0x3fff3fffu;
                __half2 y = *reinterpret_cast<__half2 *>(&cc);
                x = __hmax2(x, y); a[i] = *reinterpret_cast<uint32_t *>(&x) + 1;
            }
            if (OP == 9) {                                                            // VIADDMNMX + HMNMX2 interleaved
                if (i & 1) { a[i] = __viaddmax_s16x2(a[i], c1, c2); }
                else { __half2 x = *reinterpret_cast<__half2 *>(&a[i]); uint32_t cc = (c2 + it) & 0x3fff3fffu;
                       __half2 y = *reinterpret_cast<__half2 *>(&cc); x = __hmax2(x, y); a[i] = *reinterpret_cast<uint32_t *>(&x); }
            }
            if (OP == 10) {                                                           // VIADDMNMX + 32-bit sub interleaved
                if (i & 1) a[i] = __viaddmax_s16x2(a[i], c1, c2); else a[i] = a[i] - (c1 + it);
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = acc0 + acc1;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int OP> void run(const char *name, int per_iter_instr) {
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&cyc, 8);
    k<OP><<<148 * 2, 512>>>(out, 12345u, cyc);   // 32 warps / SM
    cudaDeviceSynchronize();
    k<OP><<<148 * 2, 512>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double warp_instr = 32.0 * ITER * 8 * per_iter_instr;   // per SM
    printf("%-34s %.3f warp-instr/clk/SM  (%.3f per SMSP)\n", name, warp_instr / h, warp_instr / h / 4);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("VIMNMX.S16x2", 1);
    run<1>("VIADDMNMX.S16x2", 1);
    run<2>("VIADD.16x2", 1);
    run<3>("32-bit sub (VIADD/IADD3)", 1);
    run<4>("VIMNMX3.S16x2", 1);
    run<5>("VIMNMX.P + 2 pred adds + add (4)", 4);
    run<6>("LOP3", 1);
    run<7>("IMAD", 1);
    run<8>("HMNMX2 + add (2)", 2);
    run<9>("VIADDMNMX / HMNMX2 alternating", 1);
    run<10>("VIADDMNMX / 32-bit sub alternating", 1);
    return 0;
}
