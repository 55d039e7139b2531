// icache.cu -- how fast can an SM sub-partition be fed with straight-line code?
// For body sizes S (instructions) and W warps per sub-partition, runs a loop whose body is S
// independent integer instructions (alternating FMA-pipe IMAD and ALU-pipe LOP3/IADD3) and reports
// issued instructions per cycle per sub-partition.  Build: nvcc -arch=sm_100a -O3 -o icache icache.cu
#include <cstdio>
#include <cuda_runtime.h>

#define I2(a, b) \
    asm volatile("mad.lo.u32 %0, %0, %2, %3;\n\txor.b32 %1, %1, %0;" : "+r"(a), "+r"(b) : "r"(c0), "r"(c1));
#define G8 I2(x0, y0) I2(x1, y1) I2(x2, y2) I2(x3, y3) I2(x4, y4) I2(x5, y5) I2(x6, y6) I2(x7, y7)   // 16 instr
#define G64 G8 G8 G8 G8
#define G256 G64 G64 G64 G64
#define G1K G256 G256 G256 G256
#define G4K G1K G1K G1K G1K
#define G16K G4K G4K G4K G4K

template <int S> struct Body;
#define DEF(S, CODE)                                                                                      \
    template <> struct Body<S> {                                                                          \
        static __device__ __forceinline__ void run(unsigned &x0, unsigned &x1, unsigned &x2, unsigned &x3, \
                                                   unsigned &x4, unsigned &x5, unsigned &x6, unsigned &x7, \
                                                   unsigned &y0, unsigned &y1, unsigned &y2, unsigned &y3, \
                                                   unsigned &y4, unsigned &y5, unsigned &y6, unsigned &y7, \
                                                   unsigned c0, unsigned c1) { CODE }                     \
    };
DEF(64, G64)
DEF(128, G64 G64)
DEF(256, G256)
DEF(512, G256 G256)
DEF(1024, G1K)
DEF(2048, G1K G1K)
DEF(4096, G4K)
DEF(8192, G4K G4K)
DEF(16384, G16K)
DEF(32768, G16K G16K)

template <int S, bool SYNC>
__global__ void kern(unsigned *out, int iters, unsigned c0, unsigned c1, long long *cycles)
{
    unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned y0 = 1, y1 = 2, y2 = 3, y3 = 4, y4 = 5, y5 = 6, y6 = 7, y7 = 8;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        Body<S>::run(x0, x1, x2, x3, x4, x5, x6, x7, y0, y1, y2, y3, y4, y5, y6, y7, c0, c1);
        if (SYNC) __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int S, bool SYNC> void run(int warps_per_smsp, unsigned *out, long long *dcyc)
{
    int threads = warps_per_smsp * 4 * 32;
    long long total = 1 << 22;                // instructions per warp
    int iters = (int)(total / S);
    kern<S, SYNC><<<148, threads>>>(out, iters, 3, 5, dcyc);
    cudaDeviceSynchronize();
    kern<S, SYNC><<<148, threads>>>(out, iters, 3, 5, dcyc);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    double ipc = (double)iters * S * warps_per_smsp / (double)cyc;
    printf("S=%6d  warps/SMSP=%d  sync=%d  IPC/SMSP=%.3f\n", S, warps_per_smsp, (int)SYNC, ipc);
}

int main()
{
    unsigned *out;
    long long *dcyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&dcyc, 8);
    for (int w : {1, 2, 4, 8}) {
        run<64, false>(w, out, dcyc);
        run<256, false>(w, out, dcyc);
        run<512, false>(w, out, dcyc);
        run<1024, false>(w, out, dcyc);
        run<2048, false>(w, out, dcyc);
        run<4096, false>(w, out, dcyc);
        run<8192, false>(w, out, dcyc);
        run<16384, false>(w, out, dcyc);
        run<32768, false>(w, out, dcyc);
        run<32768, true>(w, out, dcyc);
        run<1024, true>(w, out, dcyc);
        run<256, true>(w, out, dcyc);
    }
    return 0;
}
