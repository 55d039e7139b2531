// ff.cuh -- arithmetic in Z_p, p = 2^64 - 2^32 + 1 (the field of nufhe's NTT transform,
// reference: nufhe/transform/arithmetic.mako, nufhe/transform/ntt_cpu.py:23).
//
// All functions take and return CANONICAL representatives (< p) unless a name says otherwise.
// The same source is compiled by nvcc for sm_100a and by g++ for the host-side lane emulation
// (tests/ + csrc/host_emul.cpp), so the index/twiddle logic can be checked without a GPU.
//
// Identities used throughout (phi = 2^32):  phi^2 = phi - 1,  phi^3 = -1  (2^64 = 2^32 - 1, 2^96 = -1).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#define NB_HDC __host__ __device__ constexpr
#define NB_D __device__ __forceinline__
#else
#define NB_HD inline
#define NB_HDC constexpr
#define NB_D inline
#endif

namespace nb {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int32_t i32;

constexpr u64 FF_P = 0xffffffff00000001ULL;
constexpr u64 FF_EPS = 0xffffffffULL;   // 2^64 mod p

#if defined(__CUDACC__)
// Multipliers read from the constant bank so that ptxas keeps them as IMAD.WIDE operands (FMA pipe)
// instead of strength-reducing to SHF/IADD3 on the ALU pipe, which is the binding pipe of every kernel
// here (profiles/r1_v2_analysis.md).
__constant__ u32 nb_c_pow2[32] = {
    1u << 0, 1u << 1, 1u << 2, 1u << 3, 1u << 4, 1u << 5, 1u << 6, 1u << 7, 1u << 8, 1u << 9, 1u << 10,
    1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20,
    1u << 21, 1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30,
    1u << 31};
__constant__ u32 nb_c_eps = 0xffffffffu;
#endif

NB_HD u32 lo32(u64 x) { return (u32)x; }
NB_HD u32 hi32(u64 x) { return (u32)(x >> 32); }
NB_HD u64 pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// x in [0, 2^64) -> canonical (arithmetic.mako:164-194 `mod`)
NB_HD u64 ff_canon(u64 x) { return x >= FF_P ? x - FF_P : x; }

// a - b mod p; a may be any 64-bit value, b canonical; result canonical iff a canonical.
// (arithmetic.mako:122-161 `sub`): wrap-around of 2^64 is undone by subtracting 2^32-1.
#if defined(__CUDA_ARCH__)
#ifndef NB_SUB_CHAIN
#define NB_SUB_CHAIN 0      // experiment knob: bit 0 = ff_sub, bit 1 = ff_add use the 5-instruction borrow chain
#endif
// d = a - b mod 2^64; on a borrow add p.  CHAIN = false: + (m : beta) with beta = borrow, m = -borrow, as an add chain
// (6 instructions, 3 of them carry producers tied to the ALU pipe, the rest IMAD.X / IMAD.MOV).  CHAIN = true: a
// second borrow chain, - (0 : m) (5 instructions, 4 on the ALU pipe).
template <bool CHAIN> NB_D u64 ff_sub_dev(u64 a, u64 b)
{
    if constexpr (CHAIN) {
        u32 l, h, m;
        asm("sub.cc.u32 %0, %3, %5;\n\t"
            "subc.cc.u32 %1, %4, %6;\n\t"
            "subc.u32 %2, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %2;\n\t"
            "subc.u32 %1, %1, 0;"
            : "=&r"(l), "=&r"(h), "=&r"(m)
            : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
        (void)m;
        return pack(l, h);
    } else {
        u32 l, h, m, be;
        asm("sub.cc.u32 %0, %4, %6;\n\t"
            "subc.cc.u32 %1, %5, %7;\n\t"
            "subc.u32 %2, 0, 0;\n\t"
            "sub.u32 %3, 0, %2;\n\t"
            "add.cc.u32 %0, %0, %3;\n\t"
            "addc.u32 %1, %1, %2;"
            : "=&r"(l), "=&r"(h), "=&r"(m), "=&r"(be)
            : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
        (void)be;
        return pack(l, h);
    }
}
#endif

NB_HD u64 ff_sub(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    return ff_sub_dev<(NB_SUB_CHAIN & 1) != 0>(a, b);
#else
    u64 d = a - b;
    return a < b ? d - FF_EPS : d;
#endif
}

NB_HD u64 ff_neg(u64 a) { return a ? FF_P - a : 0; }

// a + b mod p, both canonical (arithmetic.mako:78-119 `add`), computed as a - (p - b).
NB_HD u64 ff_add(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    return ff_sub_dev<(NB_SUB_CHAIN & 2) != 0>(a, FF_P - b);
#else
    return ff_sub(a, FF_P - b);
#endif
}

// a + b for a, b in [0, p] as "some 64-bit value of the right residue": the 64-bit sum, minus p (= plus eps mod 2^64)
// when it wrapped -- no second wrap is possible (a + b - 2^64 + eps <= p).  A sum that did NOT wrap is left alone; it
// lies above p only if it falls into (p, 2^64), probability 2^-33 for uniform operands, and then its high limb is
// 2^32 - 1.  The butterfly networks that use this form watch the high limbs and redo their task with ff_add when one
// shows up (ntt_lane.cuh, br_phases.cuh); one ALU instruction per addition saved against ff_add = ff_sub(a, p - b).
NB_HD u64 ff_add_nc(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    u32 l, h, g, ng;
    asm("add.cc.u32 %0, %4, %6;\n\t"
        "addc.cc.u32 %1, %5, %7;\n\t"
        "addc.u32 %2, 0, 0;\n\t"           // g = carry out of the 64-bit sum
        "sub.u32 %3, 0, %2;\n\t"           // -g
        "sub.cc.u32 %0, %0, %2;\n\t"       // + g * eps = - g + g * 2^32: low limb, borrow iff it was 0
        "subc.u32 %1, %1, %3;"             // high limb + g - borrow
        : "=&r"(l), "=&r"(h), "=&r"(g), "=&r"(ng)
        : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
    (void)ng;
    return pack(l, h);
#else
    return ff_add(a, b);
#endif
}

// v * (2^32 - 1) for a 32-bit v: always canonical ((2^32-1)^2 < p).
NB_HD u64 ff_eps_mul(u32 v)
{
#if defined(__CUDA_ARCH__)
    u64 t;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(v), "r"(nb_c_eps));
    return t;
#else
    return ((u64)v << 32) - v;
#endif
}

// 128-bit (hi:lo) -> canonical  (arithmetic.mako:197-333 `mul`, reduction part :200-207)
NB_HD u64 ff_reduce128(u64 lo, u64 hi)
{
    u64 t = ff_sub(lo, (u64)hi32(hi));             // - hi_hi * 2^96
    // t is loose here (lo may be >= p); ff_eps_mul() is canonical, so one wrap fix is enough
    u64 u = ff_eps_mul(lo32(hi));                  // + hi_lo * 2^64
    u64 r = t + u;
    if (r < u) r += FF_EPS;
    return ff_canon(r);
}

#if defined(__CUDA_ARCH__)
// 64 x 64 -> 128 bit product as four 32-bit limbs.  Written as mad.lo.cc / madc.hi.cc chains: ptxas turns
// them into IMAD.WIDE / IMAD.HI with carry predicates, i.e. the carries ride on the FMA pipe and the ALU
// pipe (the binding one, profiles/r1_final_summary.txt) sees a single SEL.
NB_D void mul128(u64 a, u64 b, u32 &r0, u32 &r1, u32 &r2, u32 &r3)
{
    asm("{\n\t.reg .u64 t;\n\t"
        "mul.wide.u32 t, %4, %6;\n\t"          // one IMAD.WIDE; a mul.lo / mul.hi pair costs IMAD + IMAD.HI (2 + 5 cycles)
        "mov.b64 {%0, %1}, t;\n\t}\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.u32 %2, %4, %7, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, 0, 0;\n\t"
        "mad.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.u32 %3, %5, %7, %3;"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
}
// acc (5 limbs, acc4 small) += a * b
NB_D void mac128(u64 a, u64 b, u32 &c0, u32 &c1, u32 &c2, u32 &c3, u32 &c4)
{
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %8, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %8, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %6, %7, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4)
        : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
}
// v + k * eps for k in {0, 1}, as v - k + k * 2^32: one borrow-producing subtraction, the other two instructions
// can go to either pipe.  (An IMAD.WIDE with a 64-bit addend would be a single instruction, but it issues at a
// quarter of the IMAD rate and stalls the ALU pipe next to it: tools/microbench/pipes.cu)
// Used for "subtract p when v > p" (v + eps wraps to v - p) and for carry folds.
NB_D u64 ff_add_keps(u32 v0, u32 v1, u32 k)
{
    asm("sub.cc.u32 %0, %0, %2;\n\t"
        "subc.u32 %1, %1, 0;"
        : "+r"(v0), "+r"(v1) : "r"(k));
    return pack(v0, v1 + k);
}
// any 64-bit v -> [0, p]: v - p if v > p (v = p is left alone: "almost canonical")
NB_D u64 ff_canon_dev(u32 v0, u32 v1)
{
    u32 f;                                    // f = carry out of v + (2^32 - 2) = [v1 == 2^32 - 1 and v0 >= 2]
    asm("add.cc.u32 %0, %1, 0xfffffffe;\n\t"
        "addc.cc.u32 %0, %2, 0;\n\t"
        "addc.u32 %0, 0, 0;"
        : "=&r"(f) : "r"(v0), "r"(v1));
    return ff_add_keps(v0, v1, f);
}
// l + m phi + h0 phi^2 + h1 phi^3 -> [0, p].  With phi^2 = phi - 1 and phi^3 = -1 the value is
// l + (m + h0) phi - (h0 + h1); the carry c of mu = m + h0 is c phi^2 = c phi - c, and mu + c cannot overflow
// (c = 1 implies mu <= 2^32 - 2), so it is the 64-bit difference  ((mu + c) : l) - (h0 + h1 + c)  >= -2^33 - 1,
// one borrow fix (as in ff_sub) and one conditional subtraction of p.
// _nc ("not canonicalised"): the 64-bit value after the borrow fix -- the right residue, but anywhere in [0, 2^64).
// It exceeds p with probability 2^-32; callers that skip the conditional subtraction watch the high limbs instead
// (hi == 2^32 - 1 is necessary for v >= p - 1) and canonicalise on that rare path (br_phases.cuh).
NB_D u64 ff_reduce_limbs_nc(u32 l, u32 m, u32 h0, u32 h1)
{
    u32 r0, r1, d0, d1, k;
    asm("add.cc.u32 %1, %6, %7;\n\t"         // mu = m + h0
        "addc.u32 %1, %1, 0;\n\t"            // + c (the flag is left untouched)
        "addc.cc.u32 %2, %7, %8;\n\t"        // d = h0 + h1 + c
        "addc.u32 %3, 0, 0;\n\t"
        "sub.cc.u32 %0, %5, %2;\n\t"         // (r1 : r0) = (mu + c : l) - d
        "subc.cc.u32 %1, %1, %3;\n\t"
        "subc.u32 %4, 0, 0;\n\t"             // k = -borrow
        "sub.u32 %2, 0, %4;\n\t"             // borrow
        "add.cc.u32 %0, %0, %2;\n\t"         // + borrow * p = (k : borrow)
        "addc.u32 %1, %1, %4;"
        : "=&r"(r0), "=&r"(r1), "=&r"(d0), "=&r"(d1), "=&r"(k)
        : "r"(l), "r"(m), "r"(h0), "r"(h1));
    return pack(r0, r1);
}
NB_D u64 ff_reduce_limbs(u32 l, u32 m, u32 h0, u32 h1)
{
    const u64 v = ff_reduce_limbs_nc(l, m, h0, h1);
    return ff_canon_dev(lo32(v), hi32(v));
}
#endif

// [0, p] form of any 64-bit value: the rare-path fix of the _nc results
NB_HD u64 ff_canon_almost(u64 v)
{
#if defined(__CUDA_ARCH__)
    return ff_canon_dev(lo32(v), hi32(v));
#else
    return ff_canon(v);
#endif
}

NB_HD u64 ff_mul(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    u32 l, m, h0, h1;
    mul128(a, b, l, m, h0, h1);
    return ff_reduce_limbs(l, m, h0, h1);
#else
    unsigned __int128 pr = (unsigned __int128)a * b;
    return ff_reduce128((u64)pr, (u64)(pr >> 64));
#endif
}

// a * b as the right residue in [0, 2^64) on the device (see ff_reduce_limbs_nc); canonical on the host
NB_HD u64 ff_mul_nc(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    u32 l, m, h0, h1;
    mul128(a, b, l, m, h0, h1);
    return ff_reduce_limbs_nc(l, m, h0, h1);
#else
    return ff_mul(a, b);
#endif
}

// a0*b0 + a1*b1 + a2*b2 + a3*b3 mod p: the four 128-bit products are summed first (130 bits) and
// folded once; 2^128 = phi^4 = -phi.
NB_HD u64 ff_dot4(const u64 *a, const u64 *b)
{
#if defined(__CUDA_ARCH__)
    u32 c0, c1, c2, c3, c4 = 0;
    mul128(a[0], b[0], c0, c1, c2, c3);
#pragma unroll
    for (int k = 1; k < 4; k++) mac128(a[k], b[k], c0, c1, c2, c3, c4);
    return ff_sub(ff_reduce_limbs(c0, c1, c2, c3), (u64)c4 << 32);
#else
    unsigned __int128 acc = 0;
    u32 c = 0;
    for (int k = 0; k < 4; k++) {
        unsigned __int128 pr = (unsigned __int128)a[k] * b[k];
        unsigned __int128 nx = acc + pr;
        c += nx < acc;
        acc = nx;
    }
    u64 lo = (u64)acc, hi = (u64)(acc >> 64);
    return ff_sub(ff_reduce128(lo, hi), (u64)c << 32);
#endif
}

// the same dot product minus a canonical c, as a residue in [0, 2^64) on the device: ff_sub keeps any 64-bit
// minuend correct mod p (it only adds p back on a borrow), so the two subtractions need no canonical input
NB_HD u64 ff_dot4_sub_nc(const u64 *a, const u64 *b, u64 c)
{
#if defined(__CUDA_ARCH__)
    u32 c0, c1, c2, c3, c4 = 0;
    mul128(a[0], b[0], c0, c1, c2, c3);
#pragma unroll
    for (int k = 1; k < 4; k++) mac128(a[k], b[k], c0, c1, c2, c3, c4);
    return ff_sub(ff_sub(ff_reduce_limbs_nc(c0, c1, c2, c3), (u64)c4 << 32), c);
#else
    return ff_sub(ff_dot4(a, b), c);
#endif
}

// two-term version for the pair shape (br_phases.cuh: every CTA of a pair multiplies its own two digit polynomials):
// a0*b0 + a1*b1 - c as a residue in [0, 2^64) on the device, c canonical (0 allowed)
NB_HD u64 ff_dot2_sub_nc(const u64 *a, const u64 *b, u64 c)
{
#if defined(__CUDA_ARCH__)
    u32 c0, c1, c2, c3, c4 = 0;
    mul128(a[0], b[0], c0, c1, c2, c3);
    mac128(a[1], b[1], c0, c1, c2, c3, c4);
    return ff_sub(ff_sub(ff_reduce_limbs_nc(c0, c1, c2, c3), (u64)c4 << 32), c);
#else
    return ff_sub(ff_add(ff_mul(a[0], b[0]), ff_mul(a[1], b[1])), c);
#endif
}

// u * 2^(6*J1) for a small unsigned u < 2^10 (gadget digit + 512): the twist of the inner 16-point
// transform without a general shift.  Result canonical.
template <int J1> NB_HD u64 ff_twist_small(u32 u)
{
    constexpr int s = 6 * J1;
    if constexpr (s <= 54) {
        return (u64)u << s;                                        // <= 1023 * 2^54 < p
    } else if constexpr (s == 60) {
        return ((u64)(u & 15) << 60) + ff_eps_mul(u >> 4);         // (u>>4)*2^64 + (u&15)*2^60, no wrap
    } else if constexpr (s < 90) {
        return ff_eps_mul(u << (s - 64));                          // u*2^(s-64) < 2^32, times 2^64
    } else {
        // s == 90: u*2^26 = w0 + w1*2^32; times 2^64: w0*eps + w1*2^96 = w0*eps - w1
        return ff_sub(ff_eps_mul((u & 63) << 26), (u64)(u >> 6));
    }
}

// x * 2^(6 J1) for a Torus32 coefficient x (any int32): the twist of the stand-alone forward transform with
// i32_conversion (ntt.mako:395-399 followed by the first shift of the transform).  |x| <= 2^31 has two limbs after
// the bit shift, so the three limb combinations collapse to at most one modular subtraction.  Result in [0, p].
template <int J1> NB_HD u64 ff_twist_i32(i32 x)
{
    constexpr int s = 6 * J1, q = s / 32, r = s % 32;
    static_assert(s < 96, "twist exponent");
    const bool neg = x < 0;
    const u32 m = neg ? 0u - (u32)x : (u32)x;
    u64 v;
    if constexpr (q == 0) {
        v = (u64)m << s;                                              // < 2^62
    } else {
        const u32 y0 = m << r, y1 = r ? m >> (32 - r) : 0u;           // |x| * 2^r = y0 + y1 phi
        if constexpr (q == 1) {                                       // (y0 + y1) phi - y1, carry of y0 + y1 folded as eps
            const u64 t = (u64)y0 + y1;
            v = ff_sub((t << 32) + (t >> 32) * FF_EPS, (u64)y1);
        } else {                                                      // y0 phi^2 + y1 phi^3 = y0 eps - y1
            v = ff_sub(ff_eps_mul(y0), (u64)y1);
        }
    }
#if defined(__CUDA_ARCH__)
    return neg ? FF_P - v : v;                                        // p may stand for 0 on the device
#else
    return neg ? ff_neg(v) : v;
#endif
}

// a * b * 2^-64 mod p, the reference's Montgomery product (arithmetic.mako:355-419 `mul_prepared`)
constexpr u64 FF_RINV = 0xfffffffe00000001ULL;     // 2^-64 mod p (polynomial_transform_ntt.py:66)
NB_HD u64 ff_mul_prepared(u64 a, u64 b) { return ff_mul(ff_mul(a, b), FF_RINV); }
// a * 2^64 mod p (arithmetic.mako:336-352 `prepare_for_mul`)
NB_HD u64 ff_prepare_for_mul(u64 a) { return ff_mul(a, FF_EPS); }

// int32 -> field (ntt.mako:395-399) and field -> int32 (ntt.mako:402-408; ntt_cpu.py:74-80)
NB_HD u64 ff_from_i32(i32 x) { return x >= 0 ? (u64)(u32)x : FF_P - (u64)(u32)(-(int64_t)x); }
NB_HD i32 ff_to_i32(u64 v) { return (i32)(lo32(v) - (u32)(v > FF_P / 2)); }

// ---- multiplication by 2^S (arithmetic.mako:465-1045 `lsh`, extended to any S mod 192) -------
//
// x << r (0 <= r < 32) as three 32-bit limbs y0 + y1*phi + y2*phi^2, then times phi^q with
// phi^3 = -1; the three limbs land on {1, phi, phi^2} with signs and phi^2 is folded with phi - 1.
// POS/NEG are canonical partial sums; the result is POS - NEG.
struct Limbs3 { u32 y0, y1, y2; };

NB_HD Limbs3 ff_bitshift(u64 x, int r)     // r in [0, 32)
{
    Limbs3 y;
    u32 x0 = lo32(x), x1 = hi32(x);
#if defined(__CUDA_ARCH__)
    {
        // two IMAD.WIDE (FMA pipe) and one OR: the carry-in bits of the middle limb never overlap
        const u32 k = nb_c_pow2[r];
        u64 t, u;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x0), "r"(k));
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(u) : "r"(x1), "r"(k));
        y.y0 = lo32(t); y.y1 = lo32(u) | hi32(t); y.y2 = hi32(u);
        return y;
    }
#endif
    if (r == 0) { y.y0 = x0; y.y1 = x1; y.y2 = 0; }
    else {
        y.y0 = x0 << r;
        y.y1 = (x1 << r) | (x0 >> (32 - r));
        y.y2 = x1 >> (32 - r);
    }
    return y;
}

// value = sign * (y0 + y1*phi + y2*phi^2) * phi^q, q in {0,1,2}, negate in {false,true}
NB_HD u64 ff_limbs_combine(Limbs3 y, int q, bool negate)
{
    u64 pos, neg;
    if (q == 0) {            // (y0 - y2) + (y1 + y2) phi
        pos = ff_add(ff_canon(pack(y.y0, y.y1)), ff_eps_mul(y.y2));
        neg = 0;
    } else if (q == 1) {     // (-y2 - y1) + (y0 + y1) phi
        pos = ff_add((u64)y.y0 << 32, ff_eps_mul(y.y1));
        neg = y.y2;
    } else {                 // (-y0 - y1) + (y0 - y2) phi
        pos = ff_eps_mul(y.y0);
        neg = ff_canon(pack(y.y1, y.y2));
    }
    return negate ? ff_sub(neg, pos) : ff_sub(pos, neg);
}

#if defined(__CUDA_ARCH__)
// ---- device-only tight forms of the six (q, sign) limb combinations ------------------------------
// Inputs: y = x << r as limbs (y2 < 2^31).  Outputs are in [0, p] ("almost canonical": p itself may
// stand for 0; ff_sub / ff_add / ff_mul / the shifts are closed over that range and ff_to_i32(p) = 0).
NB_D void mulwide(u32 a, u32 b, u32 &lo, u32 &hi)
{
    u64 t;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b));
    lo = lo32(t); hi = hi32(t);
}
// pattern a: (y0 - y2) + (y1 + y2) phi = pack(y0, y1) + y2 * eps.  The product and the 64-bit accumulate are one
// mad.lo.cc / madc.hi.cc chain (carry on the FMA pipe).  y2 * eps < 2^63, so the sum is < 2^64 + 2^63: a carry
// folds as + eps (no second carry, result < p); without a carry the sum may exceed p.  The two cases exclude each
// other, so they share one IMAD.WIDE: + (carry | sum > p) * eps.
NB_D u64 ff_comb_a(u32 y0, u32 y1, u32 y2)
{
    u32 r0, r1, k, f;
    asm("mad.lo.cc.u32 %0, %6, %7, %4;\n\t"
        "madc.hi.cc.u32 %1, %6, %7, %5;\n\t"
        "addc.u32 %2, 0, 0;\n\t"
        "add.cc.u32 %3, %0, 0xfffffffe;\n\t"
        "addc.cc.u32 %3, %1, 0;\n\t"
        "addc.u32 %2, %2, 0;"
        : "=&r"(r0), "=&r"(r1), "=&r"(k), "=&r"(f)
        : "r"(y0), "r"(y1), "r"(y2), "r"(nb_c_eps));
    (void)f;
    return ff_add_keps(r0, r1, k);
}
// pattern b: (-y1 - y2) + (y0 + y1) phi.  The carry c of s = y0 + y1 is c phi^2 = c phi - c and s + c cannot
// overflow, so the value is the 64-bit difference ((s + c) : 0) - (y1 + y2 + c) in [-2^33, p - 1]: one borrow
// fix (as in ff_sub) and the result is canonical without a compare.
NB_D u64 ff_comb_b(u32 y0, u32 y1, u32 y2)
{
    u32 r0, r1, d0, d1, k;
    asm("add.cc.u32 %1, %5, %6;\n\t"         // s = y0 + y1
        "addc.u32 %1, %1, 0;\n\t"            // + c (flag untouched)
        "addc.cc.u32 %2, %6, %7;\n\t"        // d = y1 + y2 + c
        "addc.u32 %3, 0, 0;\n\t"
        "sub.cc.u32 %0, 0, %2;\n\t"          // (r1 : r0) = ((s + c) : 0) - d
        "subc.cc.u32 %1, %1, %3;\n\t"
        "subc.u32 %4, 0, 0;\n\t"             // k = -borrow
        "sub.u32 %2, 0, %4;\n\t"             // borrow
        "add.cc.u32 %0, %0, %2;\n\t"         // + borrow * p = (k : borrow)
        "addc.u32 %1, %1, %4;"
        : "=&r"(r0), "=&r"(r1), "=&r"(d0), "=&r"(d1), "=&r"(k)
        : "r"(y0), "r"(y1), "r"(y2));
    return pack(r0, r1);
}
// pattern c: (-y0 - y1) + (y0 - y2) phi = y0 * eps - pack(y1, y2);  pack(y1, y2) < 2^63
NB_D u64 ff_comb_c(u32 y0, u32 y1, u32 y2)
{
    u32 u0, u1;
    mulwide(y0, nb_c_eps, u0, u1);
    return ff_sub(pack(u0, u1), pack(y1, y2));
}
NB_D u64 ff_comb_c_neg(u32 y0, u32 y1, u32 y2)
{
    u32 u0, u1;
    mulwide(y0, nb_c_eps, u0, u1);
    return ff_sub(pack(y1, y2), pack(u0, u1));
}
template <int S> NB_D u64 ff_shl_dev(u64 x)
{
    constexpr int s = S % 192, s96 = s % 96, r = s96 % 32, q = s96 / 32;
    constexpr bool negate = s >= 96;
    u32 y0, y1, y2;
    if (r == 0) { y0 = lo32(x); y1 = hi32(x); y2 = 0; }
    else {
        y0 = lo32(x) << r;
        y1 = (hi32(x) << r) | (lo32(x) >> (32 - r));
        y2 = hi32(x) >> (32 - r);
    }
    if (q == 0) { u64 v = ff_comb_a(y0, y1, y2); return negate ? FF_P - v : v; }
    if (q == 1) { u64 v = ff_comb_b(y0, y1, y2); return negate ? FF_P - v : v; }
    return negate ? ff_comb_c_neg(y0, y1, y2) : ff_comb_c(y0, y1, y2);
}
#endif

// x * 2^S mod p, S a compile-time constant (any non-negative integer; reduced mod 192).
// Host: canonical in, canonical out.  Device: [0, p] in, [0, p] out (see above).
template <int S> NB_HD u64 ff_shl(u64 x)
{
    constexpr int s = S % 192;
    constexpr int s96 = s % 96;
    constexpr bool negate = s >= 96;
    if (s == 0) return x;
#if defined(__CUDA_ARCH__)
    if (s == 96) return FF_P - x;
    return ff_shl_dev<S>(x);
#else
    if (s == 96) return ff_neg(x);
    return ff_limbs_combine(ff_bitshift(x, s96 % 32), s96 / 32, negate);
#endif
}

// x * 2^s mod p for a run-time s in [0, 192)
NB_HD u64 ff_shl_var(u64 x, int s)
{
    bool negate = s >= 96;
    int s96 = negate ? s - 96 : s;
    return ff_limbs_combine(ff_bitshift(x, s96 & 31), s96 >> 5, negate);
}

}  // namespace nb
