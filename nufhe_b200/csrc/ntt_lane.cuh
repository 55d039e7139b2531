// ntt_lane.cuh -- the per-lane stages of the warp-wide 1024-point negacyclic NTT over Z_p.
//
// What it computes (reference semantics: nufhe/transform/ntt.py:30-44, natural order):
//     forward:  X[k] = sum_j x[j] psi^(j(2k+1)),        psi = 0xa70dc47e4cbdf43f^(2^32/2048)
//     inverse:  x[j] = N^-1 sum_k X[k] psi^(-j(2k+1))
// How (our own factorisation, not the reference's 8*2*8*8 one): with j = 64 j1 + j2, k = k1 + 16 k2,
//     X[k1 + 16 k2] = sum_j2 w64^(j2 k2) * psi^(j2 (2 k1 + 1)) * sum_j1 x[64 j1 + j2] w32^(j1 (2 k1 + 1))
// and psi^64 = w32 = 2^6, psi^32 = w64 = 2^3, so the inner 16-point and the outer 64-point transforms
// need only multiplications by powers of two; ONE general multiplication per element (the middle
// twiddle, which also absorbs the negacyclic twist) remains.  A warp holds one polynomial, 32
// elements per lane:
//     stage A  (lane l holds x[64 j1 + 32 h + l] in slot 16 h + j1): two 16-point transforms in registers
//     stage T  multiply by the table psi^(j2 (2 k1 + 1))
//     stage C0 radix-2 step of the 64-point transform between slots i and 16 + i (lane-dependent shift 2^(3 l))
//     transpose (slot s, lane l) <-> (slot l, lane s)   -- the only cross-lane exchange
//     stage C1 32-point transform in registers.
// The forward output of lane L, slot t is X[ntt_out_index(L, t)]; the inverse consumes that layout,
// so pointwise products never need a reordering.  The decimation-in-frequency forward leaves each
// register-level transform bit-reversed and the decimation-in-time inverse undoes it.
#pragma once
#include "ff.cuh"

namespace nb {

constexpr int NTT_N = 1024;
constexpr int SLOTS = 32;

template <int I> struct IC { static constexpr int value = I; };
template <int B, int E, typename F> struct StaticFor {
    NB_HD static void run(F &f) { f(IC<B>()); StaticFor<B + 1, E, F>::run(f); }
};
template <int E, typename F> struct StaticFor<E, E, F> { NB_HD static void run(F &) {} };
template <int B, int E, typename F> NB_HD void static_for(F f) { StaticFor<B, E, F>::run(f); }

NB_HDC int brev(int x, int bits) { int r = 0; for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i); return r; }

// natural index of the element a lane holds before the forward / after the inverse transform
NB_HD int ntt_in_index(int lane, int slot) { return 64 * (slot & 15) + 32 * (slot >> 4) + lane; }
// natural index of the transformed element held by (lane, slot) after the forward transform
NB_HD int ntt_out_index(int lane, int slot)
{
    int k1 = brev(lane & 15, 4), g = lane >> 4, m = brev(slot, 5);
    return k1 + 16 * (2 * m + g);
}
// exponent e such that the stage-T twiddle of (lane, slot) is psi^e (forward) / psi^-e (inverse)
NB_HD int ntt_twiddle_exponent(int lane, int slot)
{
    int k1 = brev(slot & 15, 4), j2 = lane + 32 * (slot >> 4);
    return (j2 * (2 * k1 + 1)) % 2048;
}

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
        NB_LOCKSTEP();
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
            u64 a = v[i0], t = ff_shl<(192 - e) % 192>(v[i1]);
            v[i0] = ff_add(a, t);
            v[i1] = ff_sub(a, t);
        });
        NB_LOCKSTEP();
    });
}

// ---- forward ------------------------------------------------------------------------------------

// stage A + T + C0.  `twd` points at this lane's column of the forward table: twd[slot * 32].
NB_HD void ntt_fwd_pre(u64 *v, const u64 *twd, int lane)
{
    static_for<0, 2>([&](auto H) {
        constexpr int h = decltype(H)::value;
        static_for<1, 16>([&](auto J) {
            constexpr int j1 = decltype(J)::value;
            v[16 * h + j1] = ff_shl<6 * j1>(v[16 * h + j1]);     // w32^j1 twist of the inner transform
        });
        dif_inlane<4, 12, 16 * h>(v);
    });
    static_for<0, 32>([&](auto T) {
        constexpr int t = decltype(T)::value;
        v[t] = ff_mul(v[t], twd[t * 32]);
        if (t % 16 == 15) NB_LOCKSTEP();
    });
    const int sh = 3 * lane;
    static_for<0, 16>([&](auto I) {
        constexpr int i = decltype(I)::value;
        u64 a = v[i], b = v[16 + i];
        v[i] = ff_add(a, b);
        v[16 + i] = ff_shl_var(ff_sub(a, b), sh);
    });
    NB_LOCKSTEP();
}
// stage C1 (after the transpose)
NB_HD void ntt_fwd_post(u64 *v) { dif_inlane<5, 6, 0>(v); }

// ---- inverse ------------------------------------------------------------------------------------

NB_HD void ntt_inv_pre(u64 *v) { dit_inlane<5, 6, 0>(v); }
// after the transpose back: C0', T' (table holds psi^-e / 1024), A'
NB_HD void ntt_inv_post(u64 *v, const u64 *twd_inv, int lane)
{
    const int sh = (192 - 3 * lane) % 192;
    static_for<0, 16>([&](auto I) {
        constexpr int i = decltype(I)::value;
        u64 a = v[i], t = ff_shl_var(v[16 + i], sh);
        v[i] = ff_add(a, t);
        v[16 + i] = ff_sub(a, t);
    });
    NB_LOCKSTEP();
    static_for<0, 32>([&](auto T) {
        constexpr int t = decltype(T)::value;
        v[t] = ff_mul(v[t], twd_inv[t * 32]);
        if (t % 16 == 15) NB_LOCKSTEP();
    });
    static_for<0, 2>([&](auto H) {
        constexpr int h = decltype(H)::value;
        dit_inlane<4, 12, 16 * h>(v);
        static_for<1, 16>([&](auto J) {
            constexpr int j1 = decltype(J)::value;
            v[16 * h + j1] = ff_shl<(192 - 6 * j1) % 192>(v[16 * h + j1]);
        });
    });
}

}  // namespace nb
