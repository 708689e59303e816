// pipes.cu -- issue-rate microbenchmark of the integer instructions the pileup kernel leans on (sm_100a).
// Each test runs N independent dependency chains per thread (ILP 8), 8 warps per SMSP-quarter x 148 SMs, and reports
// warp-instructions per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int OP> __global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed, uint32_t one, long long *cyc) {
    __shared__ uint32_t tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = i * 2654435761u;
    __syncthreads();
    uint32_t a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = seed + threadIdx.x * 8 + j;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t x = a[j];
            if (OP == 0) x = (x & 0x7f7f7f7fu) ^ (x >> 3);                                 // LOP3 (+SHF)
            if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(seed), "r"(one));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(seed));          // IMAD
            if (OP == 3) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(0x00000400u), "r"(seed)); // IDP.4A
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(x) : "r"(seed));
            if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 8;" : "+r"(x) : "r"(seed));
            if (OP == 6) x = __popc(x) + seed;
            if (OP == 7) x = __vabsdiffu4(x, seed);
            if (OP == 8) x = tab[x & 1023u];                                                // LDS.32 random (+LOP)
            if (OP == 9) x = tab[(x & 3u) + (threadIdx.x & 31u) * 4u];                          // LDS.32 conflict-free
            if (OP == 10) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(seed));             // IADD
            if (OP == 11) x = __ffs(x) + seed;
            if (OP == 12) x = __vsadu4(x, seed) ;
            if (OP == 13) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(x) : "r"(seed), "r"(one));
            a[j] = x;
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, int inst_per) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, 1024>>>(out, 12345u, 1u, cyc);
    k<OP><<<148, 1024>>>(out, 12345u, 1u, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i]; avg /= 148;
    double winst = 32.0 * ITER * 8 * inst_per;     // warp-instructions per SM
    printf("%-28s %.2f warp-inst/clk/SM  (%.2f per SMSP)  [%d inst per step incl. helpers]\n", name, winst / avg, winst / avg / 4, inst_per);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("LOP3+SHF", 2); run<1>("LOP3", 1); run<2>("IMAD", 1); run<3>("IDP.4A", 1); run<4>("PRMT", 1); run<5>("SHF", 1);
    run<6>("POPC+IADD", 2); run<7>("VABSDIFF4", 1); run<8>("LDS random + LOP", 2); run<9>("LDS conflict-free + LOP..", 2);
    run<10>("IADD", 1); run<11>("FFS(BREV+FLO)+IADD", 3); run<12>("VABSDIFF4.ACC", 1); run<13>("VABSDIFF4 add", 1);
    return 0;
}
