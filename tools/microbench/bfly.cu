// bfly.cu -- cycles per modular butterfly (a+b, a-b mod p) for candidate implementations, 16 independent
// elements per thread, 4 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nufhe_b200/csrc/ff.cuh"
using namespace nb;

__constant__ unsigned c_one = 1, c_m1 = 0xffffffffu, c_eps32 = 0xffffffffu;

// V0: the 5-instruction borrow chain used until r1 v2.3 (four of the five tied to the ALU pipe)
struct V0 {
    static __device__ __forceinline__ u64 sub(u64 a, u64 b)
    {
        u32 l, h, m;
        asm("sub.cc.u32 %0, %3, %5;\n\t"
            "subc.cc.u32 %1, %4, %6;\n\t"
            "subc.u32 %2, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %2;\n\t"
            "subc.u32 %1, %1, 0;"
            : "=&r"(l), "=&r"(h), "=&r"(m)
            : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
        return pack(l, h);
    }
    static __device__ __forceinline__ u64 add(u64 a, u64 b) { return sub(a, FF_P - b); }
};
// V1: current ff.cuh (borrow fix as one IMAD.WIDE)
struct V1 { static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return ff_sub(a, b); }
            static __device__ __forceinline__ u64 add(u64 a, u64 b) { return ff_add(a, b); } };
// V2: plain C, compare-based
struct V2 { static __device__ __forceinline__ u64 sub(u64 a, u64 b) { u64 d = a - b; return a < b ? d + FF_P : d; }
            static __device__ __forceinline__ u64 add(u64 a, u64 b) { u64 nb_ = FF_P - b; u64 d = a - nb_; return a < nb_ ? d + FF_P : d; } };
// V3: IMAD.WIDE subtraction (no carry flags), compare-based fix with predicated IMAD.WIDE
struct V3 {
    static __device__ __forceinline__ u64 sub(u64 a, u64 b)
    {
        u64 t;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(lo32(b)), "r"(c_m1), "l"(a));   // a + b_lo*2^32 - b_lo
        u32 hi = hi32(t) - lo32(b) - hi32(b);
        u64 d = pack(lo32(t), hi);
        if (a < b) {                                   // + p
            asm("mad.wide.u32 %0, %1, %1, %0;" : "+l"(d) : "r"(c_one));
            d = pack(lo32(d), hi32(d) - 1u);
        }
        return d;
    }
    static __device__ __forceinline__ u64 add(u64 a, u64 b) { return sub(a, FF_P - b); }
};
// V4: 3-limb lazy: no modular fix at all (upper bound on what laziness could buy): 64-bit add / sub only
struct V4 { static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return a - b; }
            static __device__ __forceinline__ u64 add(u64 a, u64 b) { return a + b; } };

// V5: borrow fix as an ADD chain (lo + beta, hi + mask + carry) so that ptxas may use IMAD.X (FMA pipe) for the
// carry-consuming halves
struct V5 {
    static __device__ __forceinline__ u64 sub(u64 a, u64 b)
    {
        u32 l, h, m, be;
        asm("sub.cc.u32 %0, %4, %6;\n\t"
            "subc.cc.u32 %1, %5, %7;\n\t"
            "subc.u32 %2, 0, 0;\n\t"          // m = -borrow
            "neg.s32 %3, %2;\n\t"             // beta = borrow
            "add.cc.u32 %0, %0, %3;\n\t"      // + p = (m : beta)
            "addc.u32 %1, %1, %2;"
            : "=&r"(l), "=&r"(h), "=&r"(m), "=&r"(be)
            : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
        return pack(l, h);
    }
    static __device__ __forceinline__ u64 add(u64 a, u64 b)
    {
        // a + b - p, then + p if that went negative:  a - (p - b) with p - b formed in the same chain
        u32 n0, n1;
        asm("sub.cc.u32 %0, 1, %2;\n\t"
            "subc.u32 %1, 0xffffffff, %3;"
            : "=&r"(n0), "=&r"(n1) : "r"(lo32(b)), "r"(hi32(b)));
        return sub(a, pack(n0, n1));
    }
};
// V6: like V5 but the add is a true add chain: s = a + b (carry c); t = s - p = s + eps (carry c2); pick
struct V6 {
    static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return V5::sub(a, b); }
    static __device__ __forceinline__ u64 add(u64 a, u64 b)
    {
        u32 s0, s1, t0, t1, k;
        asm("add.cc.u32 %0, %5, %7;\n\t"
            "addc.cc.u32 %1, %6, %8;\n\t"
            "addc.u32 %4, 0, 0;\n\t"
            "add.cc.u32 %2, %0, 0xffffffff;\n\t"
            "addc.cc.u32 %3, %1, 0;\n\t"
            "addc.u32 %4, %4, 0;"
            : "=&r"(s0), "=&r"(s1), "=&r"(t0), "=&r"(t1), "=&r"(k)
            : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
        return k ? pack(t0, t1) : pack(s0, s1);
    }
};

// V7: add through IMAD.WIDE carry chains: s = a + b0 (wide mad, carry out), hi += b1 (carry out), then + eps with
// carry out; select.  sub as in V1.
struct V7 {
    static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return ff_sub(a, b); }
    static __device__ __forceinline__ u64 add(u64 a, u64 b)
    {
        u32 s0, s1, t0, t1, k;
        const u32 one = c_one;
        asm("mad.lo.cc.u32 %0, %7, %9, %5;\n\t"       // s = a + b0 * one      (IMAD.WIDE with carry out)
            "madc.hi.cc.u32 %1, %7, %9, %6;\n\t"
            "addc.u32 %4, 0, 0;\n\t"
            "add.cc.u32 %1, %1, %8;\n\t"              // hi += b1
            "addc.u32 %4, %4, 0;\n\t"
            "mad.lo.cc.u32 %2, %10, %9, %0;\n\t"      // t = s + eps * one
            "madc.hi.cc.u32 %3, %10, %9, %1;\n\t"
            "addc.u32 %4, %4, 0;"
            : "=&r"(s0), "=&r"(s1), "=&r"(t0), "=&r"(t1), "=&r"(k)
            : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)), "r"(one), "r"(c_eps32));
        return k ? pack(t0, t1) : pack(s0, s1);
    }
};

template <class V> __global__ void __launch_bounds__(512) kern(u64 *io, int iters, long long *cycles)
{
    u64 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = io[(blockIdx.x * blockDim.x + threadIdx.x) * 16 + i] % FF_P;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int half = 8 >> s;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int i0 = (q / half) * 2 * half + (q % half), i1 = i0 + half;
                u64 a = v[i0], b = v[i1];
                v[i0] = V::add(a, b);
                v[i1] = V::sub(a, b);
            }
        }
    }
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) io[(blockIdx.x * blockDim.x + threadIdx.x) * 16 + i] = v[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <class V> void run(const char *name, u64 *io, long long *dcyc)
{
    int iters = 2000;
    kern<V><<<148, 512>>>(io, iters, dcyc);
    cudaDeviceSynchronize();
    kern<V><<<148, 512>>>(io, iters, dcyc);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    // 32 butterflies per iteration per thread; 4 warps per SMSP
    printf("%-40s %.2f cycles per butterfly per warp-slot (SMSP cycles / (32*iters*4 warps))\n", name,
           (double)cyc / (32.0 * iters * 4));
}
int main()
{
    u64 *io; long long *dcyc;
    cudaMalloc(&io, 148 * 512 * 16 * 8); cudaMemset(io, 0x5a, 148 * 512 * 16 * 8); cudaMalloc(&dcyc, 8);
    run<V0>("V0 5-instruction borrow chain (old)", io, dcyc);
    run<V1>("V1 IMAD.WIDE borrow fix (current)", io, dcyc);
    run<V2>("V2 plain C compare/select", io, dcyc);
    run<V3>("V3 IMAD.WIDE sub + predicated fix", io, dcyc);
    run<V4>("V4 no reduction (lower bound)", io, dcyc);
    run<V5>("V5 add-chain fix (IMAD.X friendly)", io, dcyc);
    run<V6>("V6 V5 sub + add via carry select", io, dcyc);
    run<V7>("V7 add via IMAD.WIDE carry chains", io, dcyc);
    return 0;
}
