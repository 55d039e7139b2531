// ntt_lane.cuh -- register-level building blocks of the 1024-point negacyclic NTT over Z_p.
//
// What the transform computes (reference semantics: nufhe/transform/ntt.py:30-44, natural order):
//     forward:  X[k] = sum_j x[j] psi^(j(2k+1)),        psi = 0xa70dc47e4cbdf43f^(2^32/2048)
//     inverse:  x[j] = N^-1 sum_k X[k] psi^(-j(2k+1))
// How (our own factorisation, not the reference's 8*2*8*8 one): with j = 64 j1 + j2, k = k1 + 16 k2,
//     X[k1 + 16 k2] = sum_j2 w64^(j2 k2) * psi^(j2 (2 k1 + 1)) * sum_j1 x[64 j1 + j2] w32^(j1 (2 k1 + 1))
// and psi^64 = w32 = 2^6, psi^32 = w64 = 2^3, so the inner 16-point and the outer 64-point transforms
// need only multiplications by powers of two; ONE general multiplication per element (the middle
// twiddle, which also absorbs the negacyclic twist) remains.  This header has the fully unrolled
// in-register radix-2 networks (compile-time shift amounts); br_phases.cuh arranges them into passes.
// The decimation-in-frequency forward leaves each register-level transform bit-reversed and the
// decimation-in-time inverse undoes it, so no reordering is ever needed between them.
#pragma once
#include "ff.cuh"

namespace nb {

constexpr int NTT_N = 1024;

template <int I> struct IC { static constexpr int value = I; };
template <int B, int E, typename F> struct StaticFor {
    NB_HD static void run(F &f) { f(IC<B>()); StaticFor<B + 1, E, F>::run(f); }
};
template <int E, typename F> struct StaticFor<E, E, F> { NB_HD static void run(F &) {} };
template <int B, int E, typename F> NB_HD void static_for(F f) { StaticFor<B, E, F>::run(f); }

NB_HDC int brev(int x, int bits) { int r = 0; for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i); return r; }

// One decimation-in-frequency layer set: size 2^LOGN at v[BASE ..], root 2^ROOTLOG.
template <int LOGN, int ROOTLOG, int BASE, int STRIDE = 1> NB_HD void dif_inlane(u64 *v)
{
    constexpr int n = 1 << LOGN;
    static_for<0, LOGN>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int half = n >> (s + 1);
        static_for<0, n / 2>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            constexpr int blk = q / half, k = q % half;
            constexpr int i0 = BASE + (blk * 2 * half + k) * STRIDE, i1 = i0 + half * STRIDE;
            u64 a = v[i0], b = v[i1];
            v[i0] = ff_add(a, b);
            v[i1] = ff_shl<(ROOTLOG * k * (1 << s)) % 192>(ff_sub(a, b));
        });
    });
}

// The same network with additions in the ff_add_nc form; hmax collects the high limbs of the sums (see ff_add_nc)
NB_HD u32 nb_umax(u32 a, u32 b) { return a > b ? a : b; }
template <int LOGN, int ROOTLOG, int BASE, int STRIDE = 1> NB_HD void dif_inlane_nc(u64 *v, u32 &hmax)
{
    constexpr int n = 1 << LOGN;
    static_for<0, LOGN>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int half = n >> (s + 1);
        static_for<0, n / 2>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            constexpr int blk = q / half, k = q % half;
            constexpr int i0 = BASE + (blk * 2 * half + k) * STRIDE, i1 = i0 + half * STRIDE;
            u64 a = v[i0], b = v[i1];
            v[i0] = ff_add_nc(a, b);
            hmax = nb_umax(hmax, hi32(v[i0]));
            v[i1] = ff_shl<(ROOTLOG * k * (1 << s)) % 192>(ff_sub(a, b));
        });
    });
}

// The exact inverse network (up to the factor 2^LOGN): decimation in time with the inverse root.
template <int LOGN, int ROOTLOG, int BASE, int STRIDE = 1> NB_HD void dit_inlane(u64 *v)
{
    constexpr int n = 1 << LOGN;
    static_for<0, LOGN>([&](auto S) {
        constexpr int s = LOGN - 1 - decltype(S)::value;
        constexpr int half = n >> (s + 1);
        static_for<0, n / 2>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            constexpr int blk = q / half, k = q % half;
            constexpr int i0 = BASE + (blk * 2 * half + k) * STRIDE, i1 = i0 + half * STRIDE;
            constexpr int e = (ROOTLOG * k * (1 << s)) % 192;
            // inverse twiddle 2^-e = -2^(96 - e) for 0 < e < 96: shift by the positive amount and swap the two
            // outputs instead of negating the product
            if constexpr (e > 0 && e < 96) {
                u64 a = v[i0], t = ff_shl<96 - e>(v[i1]);
                v[i0] = ff_sub(a, t);
                v[i1] = ff_add(a, t);
            } else {
                u64 a = v[i0], t = ff_shl<(192 - e) % 192>(v[i1]);
                v[i0] = ff_add(a, t);
                v[i1] = ff_sub(a, t);
            }
        });
    });
}

template <int LOGN, int ROOTLOG, int BASE, int STRIDE = 1> NB_HD void dit_inlane_nc(u64 *v, u32 &hmax)
{
    constexpr int n = 1 << LOGN;
    static_for<0, LOGN>([&](auto S) {
        constexpr int s = LOGN - 1 - decltype(S)::value;
        constexpr int half = n >> (s + 1);
        static_for<0, n / 2>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            constexpr int blk = q / half, k = q % half;
            constexpr int i0 = BASE + (blk * 2 * half + k) * STRIDE, i1 = i0 + half * STRIDE;
            constexpr int e = (ROOTLOG * k * (1 << s)) % 192;
            if constexpr (e > 0 && e < 96) {
                u64 a = v[i0], t = ff_shl<96 - e>(v[i1]);
                v[i0] = ff_sub(a, t);
                v[i1] = ff_add_nc(a, t);
                hmax = nb_umax(hmax, hi32(v[i1]));
            } else {
                u64 a = v[i0], t = ff_shl<(192 - e) % 192>(v[i1]);
                v[i0] = ff_add_nc(a, t);
                hmax = nb_umax(hmax, hi32(v[i0]));
                v[i1] = ff_sub(a, t);
            }
        });
    });
}

}  // namespace nb
