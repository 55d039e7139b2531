// pipes.cu -- issue cost (SMSP cycles per warp-instruction) of every integer instruction form the field
// arithmetic compiles to, and of their mixes.  Each mode is a loop body of 64 repetitions of a small
// block on 8 registers; what ptxas made of it is read with `cuobjdump -sass` (tools/microbench/README in
// profiles/): the table in DESIGN.md combines both.  4 and 8 warps per sub-partition (throughput, not latency).
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(X) X X X X X X X X
#define REP64(X) REP8(REP8(X))
typedef unsigned u32;
typedef unsigned long long u64;

// 1: LOP3 only (no FMA-pipe twin exists)
#define B1 asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2; xor.b32 %2, %2, %3; xor.b32 %3, %3, %0;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); \
           asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2; xor.b32 %2, %2, %3; xor.b32 %3, %3, %0;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));
// 2: 4-limb carry chain: IADD3(P) + 2 IADD3.X(P) + IADD3.X
#define B2 asm volatile("add.cc.u32 %0, %0, %1; addc.cc.u32 %1, %1, %2; addc.cc.u32 %2, %2, %3; addc.u32 %3, %3, %0;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); \
           asm volatile("add.cc.u32 %0, %0, %1; addc.cc.u32 %1, %1, %2; addc.cc.u32 %2, %2, %3; addc.u32 %3, %3, %0;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));
// 3: IMAD (32-bit multiply-add, nothing foldable)
#define B3 asm volatile("mad.lo.u32 %0, %0, %1, %2; mad.lo.u32 %1, %1, %2, %3; mad.lo.u32 %2, %2, %3, %0; mad.lo.u32 %3, %3, %0, %1;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); \
           asm volatile("mad.lo.u32 %0, %0, %1, %2; mad.lo.u32 %1, %1, %2, %3; mad.lo.u32 %2, %2, %3, %0; mad.lo.u32 %3, %3, %0, %1;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));
// 4: IMAD.WIDE without addend: w_i = lo(w_j) * hi(w_j)
#define WMUL(D, S) asm volatile("{ .reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, l, h; }" : "=l"(D) : "l"(S));
#define B4 WMUL(w0, w1) WMUL(w1, w2) WMUL(w2, w3) WMUL(w3, w0) WMUL(w4, w5) WMUL(w5, w6) WMUL(w6, w7) WMUL(w7, w4)
// 5: IMAD.WIDE with a 64-bit addend (mad.lo.cc / madc.hi pair, fused by ptxas)
#define WPAIR(L, H, M) asm volatile("mad.lo.cc.u32 %0, %2, %2, %0; madc.hi.u32 %1, %2, %2, %1;" : "+r"(L), "+r"(H) : "r"(M));
#define B5 WPAIR(a, b, d) WPAIR(c, d, b2) WPAIR(a2, b2, d2) WPAIR(c2, d2, b)
// 6: IMAD.HI
#define B6 asm volatile("mad.hi.u32 %0, %0, %1, %2; mad.hi.u32 %1, %1, %2, %3; mad.hi.u32 %2, %2, %3, %0; mad.hi.u32 %3, %3, %0, %1;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); \
           asm volatile("mad.hi.u32 %0, %0, %1, %2; mad.hi.u32 %1, %1, %2, %3; mad.hi.u32 %2, %2, %3, %0; mad.hi.u32 %3, %3, %0, %1;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));
// 7: LOP3 + IMAD 1:1 (both pipes)
#define B7 asm volatile("xor.b32 %0, %0, %1; mad.lo.u32 %2, %2, %3, %0; xor.b32 %1, %1, %2; mad.lo.u32 %3, %3, %0, %1;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); \
           asm volatile("xor.b32 %0, %0, %1; mad.lo.u32 %2, %2, %3, %0; xor.b32 %1, %1, %2; mad.lo.u32 %3, %3, %0, %1;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));
// 8: 2 LOP3 + 1 IMAD.WIDE (no addend)
#define B8 WMUL(w0, w1) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2;" : "+r"(a), "+r"(b), "+r"(c)); \
           WMUL(w1, w0) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2;" : "+r"(c), "+r"(d), "+r"(a)); \
           WMUL(w2, w3) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2;" : "+r"(a2), "+r"(b2), "+r"(c2)); \
           WMUL(w3, w2) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2;" : "+r"(c2), "+r"(d2), "+r"(a2));
// 9: carry pair only: IADD3(P) + IADD3.X
#define B9 asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a), "+r"(b) : "r"(c), "r"(d)); \
           asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c), "+r"(d) : "r"(a2), "r"(b2)); \
           asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a2), "+r"(b2) : "r"(c2), "r"(d2)); \
           asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c2), "+r"(d2) : "r"(a), "r"(b));
// 10: SHF (funnel shift) only
#define B10 asm volatile("shf.l.wrap.b32 %0, %0, %1, %2; shf.l.wrap.b32 %1, %1, %2, %3; shf.l.wrap.b32 %2, %2, %3, %0; shf.l.wrap.b32 %3, %3, %0, %1;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); \
            asm volatile("shf.l.wrap.b32 %0, %0, %1, %2; shf.l.wrap.b32 %1, %1, %2, %3; shf.l.wrap.b32 %2, %2, %3, %0; shf.l.wrap.b32 %3, %3, %0, %1;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));
// 11: compare + select (ISETP + SEL)
#define B11 asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %2, %3, %0, p; setp.lt.u32 p, %1, %2; selp.u32 %3, %0, %1, p; }" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)); \
            asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %2, %3, %0, p; setp.lt.u32 p, %1, %2; selp.u32 %3, %0, %1, p; }" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));
// 12: the old 5-instruction modular subtraction, two independent copies
#define SUBOLD(A, B, C, D) asm volatile("sub.cc.u32 %0, %0, %2; subc.cc.u32 %1, %1, %3; subc.u32 %2, 0, 0; sub.cc.u32 %0, %0, %2; subc.u32 %1, %1, 0;" : "+r"(A), "+r"(B), "+r"(C), "+r"(D));
#define B12 SUBOLD(a, b, c, d) SUBOLD(a2, b2, c2, d2)
// 13: 3-limb carry chain (IADD3(P), IADD3.X(P), IADD3.X) -- is a carry-out on IADD3.X extra?
#define B13 asm volatile("add.cc.u32 %0, %0, %1; addc.cc.u32 %1, %1, %2; addc.u32 %2, %2, %3;" : "+r"(a), "+r"(b), "+r"(c) : "r"(d)); \
            asm volatile("add.cc.u32 %0, %0, %1; addc.cc.u32 %1, %1, %2; addc.u32 %2, %2, %3;" : "+r"(a2), "+r"(b2), "+r"(c2) : "r"(d2)); \
            asm volatile("add.cc.u32 %0, %0, %1; addc.cc.u32 %1, %1, %2; addc.u32 %2, %2, %3;" : "+r"(d), "+r"(a), "+r"(b) : "r"(c)); \
            asm volatile("add.cc.u32 %0, %0, %1; addc.cc.u32 %1, %1, %2; addc.u32 %2, %2, %3;" : "+r"(d2), "+r"(a2), "+r"(b2) : "r"(c2));
// 14: IMAD.HI with carry-out + consumer (mad.hi.cc / addc): the carry rides the FMA pipe?
#define B14 asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1; addc.u32 %2, %2, 0;" : "+r"(a), "+r"(b), "+r"(c) : "r"(d)); \
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1; addc.u32 %2, %2, 0;" : "+r"(a2), "+r"(b2), "+r"(c2) : "r"(d2)); \
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1; addc.u32 %2, %2, 0;" : "+r"(d), "+r"(c), "+r"(b) : "r"(a)); \
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1; addc.u32 %2, %2, 0;" : "+r"(d2), "+r"(c2), "+r"(b2) : "r"(a2));
// 15: LOP3 + IMAD.WIDE(addend) 2:1
#define B15 WPAIR(a, b, d) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2;" : "+r"(c), "+r"(d), "+r"(a)); \
            WPAIR(a2, b2, d2) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2;" : "+r"(c2), "+r"(d2), "+r"(a2));
// 16: 4 LOP3 + 1 IMAD.WIDE(addend)
#define B16 WPAIR(a, b, d) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2; xor.b32 %2, %2, %3; xor.b32 %3, %3, %0;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2)); \
            WPAIR(c, d, b) asm volatile("xor.b32 %0, %0, %1; xor.b32 %1, %1, %2; xor.b32 %2, %2, %3; xor.b32 %3, %3, %0;" : "+r"(a2), "+r"(b2), "+r"(c2), "+r"(d2));

template <int MODE> __global__ void __launch_bounds__(1024) kern(u32 *out, int iters, long long *cycles)
{
    u32 a = threadIdx.x + out[0], b = a + 1, c = a + 2, d = a + 3;
    u32 a2 = a + 4, b2 = a + 5, c2 = a + 6, d2 = a + 7;
    u64 w0 = a, w1 = b | 1ull << 33, w2 = c | 1ull << 34, w3 = d | 1ull << 35, w4 = a2 | 1ull << 36, w5 = b2 | 1ull << 37, w6 = c2 | 1ull << 38, w7 = d2 | 1ull << 39;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (MODE == 1) { REP64(B1) }
        if (MODE == 2) { REP64(B2) }
        if (MODE == 3) { REP64(B3) }
        if (MODE == 4) { REP64(B4) }
        if (MODE == 5) { REP64(B5) }
        if (MODE == 6) { REP64(B6) }
        if (MODE == 7) { REP64(B7) }
        if (MODE == 8) { REP64(B8) }
        if (MODE == 9) { REP64(B9) }
        if (MODE == 10) { REP64(B10) }
        if (MODE == 11) { REP64(B11) }
        if (MODE == 12) { REP64(B12) }
        if (MODE == 13) { REP64(B13) }
        if (MODE == 14) { REP64(B14) }
        if (MODE == 15) { REP64(B15) }
        if (MODE == 16) { REP64(B16) }
    }
    long long t1 = clock64();
    u64 w = w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7;
    out[1 + blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ a2 ^ b2 ^ c2 ^ d2 ^ (u32)w ^ (u32)(w >> 32);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int MODE> void run(const char *name, u32 *out, long long *dcyc)
{
    for (int w : {4, 8}) {
        int iters = 512;
        kern<MODE><<<148, w * 128>>>(out, iters, dcyc);
        cudaDeviceSynchronize();
        kern<MODE><<<148, w * 128>>>(out, iters, dcyc);
        cudaDeviceSynchronize();
        long long cyc;
        cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        // cycles per repetition of the block, per warp: SMSP cycles / (iters * 64 reps * warps on the SMSP)
        printf("mode %2d %-44s warps/SMSP=%d  cycles per block per warp = %.2f\n", MODE, name, w, (double)cyc / (iters * 64.0 * w));
    }
}
int main()
{
    u32 *out; long long *dcyc;
    cudaMalloc(&out, (148 * 1024 + 1) * 4); cudaMemset(out, 0, (148 * 1024 + 1) * 4); cudaMalloc(&dcyc, 8);
    run<1>("8 LOP3", out, dcyc);
    run<2>("2 x (IADD3.P + 2 IADD3.X.P + IADD3.X)", out, dcyc);
    run<3>("8 IMAD", out, dcyc);
    run<4>("8 IMAD.WIDE (no addend)", out, dcyc);
    run<5>("4 IMAD.WIDE (64-bit addend)", out, dcyc);
    run<6>("8 IMAD.HI", out, dcyc);
    run<7>("4 LOP3 + 4 IMAD", out, dcyc);
    run<8>("8 LOP3 + 4 IMAD.WIDE (no addend)", out, dcyc);
    run<9>("4 x (IADD3.P + IADD3.X)", out, dcyc);
    run<10>("8 SHF", out, dcyc);
    run<11>("4 ISETP + 4 SEL", out, dcyc);
    run<12>("2 x old mod-sub (5 instr)", out, dcyc);
    run<13>("4 x (IADD3.P + IADD3.X.P + IADD3.X)", out, dcyc);
    run<14>("4 x (mad.lo.cc + madc.hi.cc + addc)", out, dcyc);
    run<15>("2 x (IMAD.WIDE addend + 2 LOP3)", out, dcyc);
    run<16>("2 x (IMAD.WIDE addend + 4 LOP3)", out, dcyc);
    return 0;
}
