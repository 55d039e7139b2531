// alu.cu -- issue rate of the integer instruction forms the field arithmetic is made of, 4 warps per
// sub-partition, loop body of 1024 instructions (fits the instruction cache).
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(X) X X X X X X X X
#define REP64(X) REP8(REP8(X))
// (1) plain 32-bit adds, no carries
#define B_ADD  asm volatile("add.u32 %0, %0, %4; add.u32 %1, %1, %4; add.u32 %2, %2, %4; add.u32 %3, %3, %4;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(k));
// (2) 64-bit add as carry chain (IADD3 + IADD3.X)
#define B_CC   asm volatile("add.cc.u32 %0, %0, %4; addc.u32 %1, %1, %4; add.cc.u32 %2, %2, %4; addc.u32 %3, %3, %4;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(k));
// (3) modular subtraction pattern: sub.cc, subc.cc, subc(mask), sub.cc, subc  (5 instr) on two values -> use 4+? keep 5
#define B_SUB  asm volatile("sub.cc.u32 %0, %0, %4; subc.cc.u32 %1, %1, %4; subc.u32 %2, 0, 0; sub.cc.u32 %0, %0, %2; subc.u32 %1, %1, 0;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(k));
// (4) mix: IMAD.WIDE + adds
#define B_MIX  asm volatile("mad.lo.u32 %0, %0, %4, %1; add.cc.u32 %2, %2, %4; addc.u32 %3, %3, %4; mad.lo.u32 %1, %1, %4, %0;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(k));
// (5) IMAD.WIDE with a 64-bit addend (the fix-up instruction of ff_sub), written as the mad.lo.cc / madc.hi pair
//     that ptxas fuses; four register pairs, each multiplies a word of its neighbour (nothing loop-invariant)
#define WPAIR(L, H, M) asm volatile("mad.lo.cc.u32 %0, %2, %2, %0; madc.hi.u32 %1, %2, %2, %1;" : "+r"(L), "+r"(H) : "r"(M));
#define B_WIDE WPAIR(a, b, d) WPAIR(c, d, b2) WPAIR(a2, b2, d2) WPAIR(c2, d2, b)
// (6) 1:1 mix of IMAD.WIDE and carry adds: both pipes busy
#define B_WMIX WPAIR(a, b, d2) WPAIR(a2, b2, b) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(c), "+r"(d) : "r"(k)); \
               asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(c2), "+r"(d2) : "r"(k));
// (7) the new modular subtraction: sub.cc, subc.cc, subc, IMAD.WIDE (m * m + d), hi -= m
#define B_SUBW asm volatile("sub.cc.u32 %0, %0, %4; subc.cc.u32 %1, %1, %4; subc.u32 %2, 0, 0; mad.lo.cc.u32 %0, %2, %2, %0; madc.hi.u32 %1, %2, %2, %1; sub.u32 %1, %1, %2;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(k)); \
               asm volatile("sub.cc.u32 %0, %0, %4; subc.cc.u32 %1, %1, %4; subc.u32 %2, 0, 0; mad.lo.cc.u32 %0, %2, %2, %0; madc.hi.u32 %1, %2, %2, %1; sub.u32 %1, %1, %2;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2) : "r"(k));
template <int MODE> __global__ void kern(unsigned *out, int iters, unsigned k, long long *cycles)
{
    unsigned a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    unsigned a2 = a + 4, b2 = a + 5, c2 = a + 6, d2 = a + 7;
    unsigned long long w0 = a, w1 = b, w2 = c, w3 = d;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (MODE == 1) { REP64(B_ADD) }
        if (MODE == 2) { REP64(B_CC) }
        if (MODE == 3) { REP64(B_SUB) }
        if (MODE == 4) { REP64(B_MIX) }
        if (MODE == 5) { REP64(B_WIDE) }
        if (MODE == 6) { REP64(B_WMIX) }
        if (MODE == 7) { REP64(B_SUBW) }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ a2 ^ b2 ^ c2 ^ d2 ^ (unsigned)(w0 ^ w1 ^ w2 ^ w3) ^ (unsigned)((w0 ^ w1 ^ w2 ^ w3) >> 32);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int MODE> void run(const char *name, int per_body, unsigned *out, long long *dcyc)
{
    for (int w : {1, 2, 4, 8}) {
        int iters = 4096;
        kern<MODE><<<148, w * 128>>>(out, iters, 3, dcyc);
        cudaDeviceSynchronize();
        kern<MODE><<<148, w * 128>>>(out, iters, 3, dcyc);
        cudaDeviceSynchronize();
        long long cyc;
        cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps/SMSP=%d  IPC/SMSP=%.3f\n", name, w, (double)iters * per_body * w / (double)cyc);
    }
}
int main()
{
    unsigned *out; long long *dcyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&dcyc, 8);
    run<1>("add.u32 (IADD3)", 64 * 4, out, dcyc);
    run<2>("add.cc/addc (IADD3+IADD3.X)", 64 * 4, out, dcyc);
    run<3>("mod-sub pattern (5 instr)", 64 * 5, out, dcyc);
    run<4>("2 IMAD + IADD3 + IADD3.X", 64 * 4, out, dcyc);
    run<5>("IMAD.WIDE 64-bit addend", 64 * 4, out, dcyc);
    run<6>("2 IMAD.WIDE + 2 (IADD3 + IADD3.X)", 64 * 6, out, dcyc);
    run<7>("mod-sub, IMAD.WIDE fix (2x5)", 64 * 10, out, dcyc);
    return 0;
}
