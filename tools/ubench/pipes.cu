// Micro-benchmark: per-SMSP issue rate of the integer instructions the prefilter is built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 16

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t a0, uint32_t b0) {
    uint32_t r[8];
    uint64_t w[8];
    for (int i = 0; i < 8; ++i) { r[i] = a0 + threadIdx.x * (i + 1); w[i] = r[i]; }
    uint32_t b = b0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(a0));
                if (OP == 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                if (OP == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[i]), "r"(b));
                if (OP == 4) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                if (OP == 5) asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
                if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                if (OP == 7) asm volatile("bfind.u32 %0, %0;" : "+r"(r[i]));
                if (OP == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
                if (OP == 9) {  // 1:1 mix LOP3 + IMAD
                    if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(a0));
                    else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                }
                if (OP == 10) {  // 1:1 mix LOP3 + dp4a
                    if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(a0));
                    else asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                }
                if (OP == 11) {  // 1:1 mix LOP3 + mad.wide
                    if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(a0));
                    else asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[i]), "r"(b));
                }
                if (OP == 12) {  // 1:2 mix LOP3 + 2x (IMAD, dp4a)
                    if ((i & 3) == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(a0));
                    else if (i & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                    else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a0));
                }
                if (OP == 13) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(b));
                if (OP == 14) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            }
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += r[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, uint32_t* d, int sms, double ghz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<sms, 1024>>>(d, 1, 3);
    cudaEventRecord(e0);
    k<OP><<<sms, 1024>>>(d, 1, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst_per_smsp = (double)ITERS * UNROLL * 8 * (1024 / 32) / 4;
    const double cycles = ms * 1e-3 * ghz * 1e9;
    printf("%-28s %8.3f ms  %6.3f warp-inst/cycle/SMSP (at %.3f GHz)\n", name, ms, warp_inst_per_smsp / cycles, ghz);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    uint32_t* d;
    cudaMalloc(&d, p.multiProcessorCount * 1024 * 4);
    run<0>("LOP3", d, p.multiProcessorCount, ghz);
    run<1>("SHF.R (reg shift)", d, p.multiProcessorCount, ghz);
    run<13>("SHF.L (imm shift)", d, p.multiProcessorCount, ghz);
    run<2>("IMAD", d, p.multiProcessorCount, ghz);
    run<3>("IMAD.WIDE", d, p.multiProcessorCount, ghz);
    run<14>("IMAD.HI", d, p.multiProcessorCount, ghz);
    run<4>("IDP4A", d, p.multiProcessorCount, ghz);
    run<5>("POPC", d, p.multiProcessorCount, ghz);
    run<6>("PRMT", d, p.multiProcessorCount, ghz);
    run<7>("FLO (bfind)", d, p.multiProcessorCount, ghz);
    run<8>("IADD", d, p.multiProcessorCount, ghz);
    run<9>("mix LOP3+IMAD 1:1", d, p.multiProcessorCount, ghz);
    run<10>("mix LOP3+IDP4A 1:1", d, p.multiProcessorCount, ghz);
    run<11>("mix LOP3+IMAD.WIDE 1:1", d, p.multiProcessorCount, ghz);
    run<12>("mix LOP3+IDP+IMAD 1:1:2", d, p.multiProcessorCount, ghz);
    return 0;
}
