// kernels.cuh -- sm_100a kernels of the TFHE gate-bootstrapping hot path.
//
// Work decomposition: ONE WARP PER POLYNOMIAL.  A ciphertext's accumulator (k+1 = 2 polynomials of
// 1024 Torus32 coefficients) is owned by a pair of warps; warp `mi` owns ACC[mi] in shared memory,
// transforms the two gadget digits of its polynomial, multiplies them with its half of the
// bootstrap-key row, swaps one partial product with its partner through shared memory, and runs the
// inverse transform of output polynomial `mi`.  The only block-level synchronisation on the path is
// a 64-thread named barrier around that swap.
//
// Reference counterparts: nufhe/blind_rotate.mako:18-226 (fused bootstrap), tgsw_gpu.py:110-169
// (external product), transform/computation.mako:18-143 (stand-alone transform), lwe_gpu.mako:59-118
// (key switch).  Nothing here is derived from those kernels' structure (128 threads/transform,
// 8*2*8*8 radix plan, 15 block barriers per step); see ntt_lane.cuh for the transform we use.
#pragma once
#include <cuda_runtime.h>
#include "ntt_lane.cuh"

namespace nb {

constexpr int LWE_N_MAX = 1024;          // upper bound on the LWE dimension n handled by the gate kernels
constexpr int TR_STRIDE = 33;            // padded row of the transpose scratch (u64 units)
constexpr int TR_WORDS = 32 * TR_STRIDE; // u64 per warp scratch

// ---- warp transpose through shared memory: (slot s, lane l) <-> (slot l, lane s) ----------------
NB_D void warp_transpose(u64 *v, u64 *scratch, int lane)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int s = 0; s < 32; s++) scratch[s * TR_STRIDE + lane] = v[s];
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 32; s++) v[s] = scratch[lane * TR_STRIDE + s];
    __syncwarp();
#endif
}

NB_D void warp_ntt_forward(u64 *v, u64 *scratch, const u64 *twd_fwd, int lane)
{
    ntt_fwd_pre(v, twd_fwd + lane, lane);
    warp_transpose(v, scratch, lane);
    ntt_fwd_post(v);
}

NB_D void warp_ntt_inverse(u64 *v, u64 *scratch, const u64 *twd_inv, int lane)
{
    ntt_inv_pre(v);
    warp_transpose(v, scratch, lane);
    ntt_inv_post(v, twd_inv + lane, lane);
}

// ---- stand-alone batched transforms (reference: transform/computation.mako `standalone_transform`) ---
// natural order in and out, one warp per polynomial, grid-stride over the batch.
template <bool IN_I32>
__global__ void __launch_bounds__(128) ntt_forward_kernel(const void *__restrict__ in, u64 *__restrict__ out,
                                                            const u64 *__restrict__ twd_fwd, size_t batch)
{
    __shared__ u64 scratch_all[4 * TR_WORDS];
    __shared__ u64 twd[NTT_N];
    for (int i = threadIdx.x; i < NTT_N; i += blockDim.x) twd[i] = twd_fwd[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *scratch = scratch_all + warp * TR_WORDS;
    // every warp of the CTA runs the same number of iterations (NB_LOCKSTEP barriers inside)
    for (size_t p0 = (size_t)blockIdx.x * 4; p0 < batch; p0 += (size_t)gridDim.x * 4) {
        const size_t p = p0 + warp;
        const bool live = p < batch;
        u64 v[32];
#pragma unroll
        for (int s = 0; s < 32; s++) {
            size_t idx = (live ? p : p0) * NTT_N + ntt_in_index(lane, s);
            if (IN_I32) v[s] = ff_from_i32(((const i32 *)in)[idx]);
            else v[s] = ff_canon(((const u64 *)in)[idx]);
        }
        warp_ntt_forward(v, scratch, twd, lane);
        if (live) {
#pragma unroll
            for (int s = 0; s < 32; s++) out[p * NTT_N + ntt_out_index(lane, s)] = v[s];
        }
    }
}

template <bool OUT_I32>
__global__ void __launch_bounds__(128) ntt_inverse_kernel(const u64 *__restrict__ in, void *__restrict__ out,
                                                            const u64 *__restrict__ twd_inv, size_t batch)
{
    __shared__ u64 scratch_all[4 * TR_WORDS];
    __shared__ u64 twd[NTT_N];
    for (int i = threadIdx.x; i < NTT_N; i += blockDim.x) twd[i] = twd_inv[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *scratch = scratch_all + warp * TR_WORDS;
    for (size_t p0 = (size_t)blockIdx.x * 4; p0 < batch; p0 += (size_t)gridDim.x * 4) {
        const size_t p = p0 + warp;
        const bool live = p < batch;
        u64 v[32];
#pragma unroll
        for (int s = 0; s < 32; s++) v[s] = ff_canon(in[(live ? p : p0) * NTT_N + ntt_out_index(lane, s)]);
        warp_ntt_inverse(v, scratch, twd, lane);
        if (live) {
#pragma unroll
            for (int s = 0; s < 32; s++) {
                size_t idx = p * NTT_N + ntt_in_index(lane, s);
                if (OUT_I32) ((i32 *)out)[idx] = ff_to_i32(v[s]);
                else ((u64 *)out)[idx] = v[s];
            }
        }
    }
}

// ---- element-wise field ops (unit tests of the arithmetic; key generation) ----------------------
enum FfOp { FF_OP_ADD = 0, FF_OP_SUB = 1, FF_OP_MUL = 2, FF_OP_MUL_PREPARED = 3, FF_OP_PREPARE = 4, FF_OP_LSH = 5 };

__global__ void ff_elementwise_kernel(int op, const u64 *__restrict__ a, const u64 *__restrict__ b,
                                      u64 *__restrict__ out, size_t n, size_t b_period)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u64 x = ff_canon(a[i]);
        u64 y = b ? b[b_period ? i % b_period : i] : 0;
        u64 r;
        switch (op) {
        case FF_OP_ADD: r = ff_add(x, ff_canon(y)); break;
        case FF_OP_SUB: r = ff_sub(x, ff_canon(y)); break;
        case FF_OP_MUL: r = ff_mul(x, ff_canon(y)); break;
        case FF_OP_MUL_PREPARED: r = ff_mul_prepared(x, ff_canon(y)); break;
        case FF_OP_PREPARE: r = ff_prepare_for_mul(x); break;
        default: r = ff_shl_var(x, (int)(y % 192)); break;
        }
        out[i] = r;
    }
}

// ---- bootstrap-key layout ------------------------------------------------------------------------
// reference row (nufhe/blind_rotate.py:112): [mi][j][mo][k], values NTT(bk) * 2^64 (Montgomery form).
// internal row: [mi][slot][j][lane][mo], plain values, so that in the MAC lane `lane` of warp `mi`
// reads, for transform slot `slot` and digit j, one 16-byte pair (mo = 0, 1).
__global__ void bk_prepare_kernel(const u64 *__restrict__ bk_ref, u64 *__restrict__ bk_int, size_t rows)
{
    const size_t total = rows * 8 * NTT_N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t row = i / (8 * NTT_N);
        int r = (int)(i % (8 * NTT_N));
        int mo = r & 1, lane = (r >> 1) & 31, j = (r >> 6) & 1, slot = (r >> 7) & 31, mi = r >> 12;
        int k = ntt_out_index(lane, slot);
        u64 x = bk_ref[row * 8 * NTT_N + (((mi * 2 + j) * 2 + mo) * NTT_N + k)];
        bk_int[i] = ff_mul(ff_canon(x), FF_RINV);
    }
}

// ---- the external-product step shared by the fused bootstrap and the stand-alone kernel ---------
struct WarpState {
    int lane;        // 0..31
    int mi;          // which accumulator polynomial this warp owns (0: mask, 1: body)
    int bar_id;      // named barrier shared with the partner warp
    i32 *acc;        // this warp's polynomial, 1024 Torus32 in shared memory, natural order
    u64 *scratch;    // transpose scratch, TR_WORDS
    u64 *xchg_out;   // partial products this warp hands to its partner, [slot][lane]
    u64 *xchg_in;    // the partner's buffer
    const u64 *twd_fwd, *twd_inv;   // shared-memory tables [slot][lane]
};

NB_D void pair_barrier(int id)
{
#if defined(__CUDA_ARCH__)
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
#endif
}

// gadget decomposition of one coefficient (tgsw_gpu.py:31-54; blind_rotate.mako:41-43,116-124)
NB_D void decompose(i32 c, i32 &d0, i32 &d1)
{
    const u32 offset = 0x80000000u + (1u << 21);
    i32 t = (i32)((u32)c + offset);
    d0 = ((t >> 22) & 1023) - 512;
    d1 = ((t >> 12) & 1023) - 512;
}

// ROTATE: acc += bk_row (x) ((X^a - 1) acc)      (mux_rotate, nufhe/bootstrap.py:96-109)
// else  : acc  = bk_row (x) acc                  (tgsw_transformed_external_mul, tgsw.py:165-172)
// bk_row: internal layout (bk_prepare_kernel), this kernel reads it through the read-only path.
template <bool ROTATE>
NB_D void external_product_step(const WarpState &w, const u64 *__restrict__ bk_row, int a)
{
    const int lane = w.lane;
    u64 v0[32], v1[32];
    // 1. (X^a - 1) * ACC[mi], decomposed into two digit polynomials, in forward-transform slot order
    {
        const int ar = a & (NTT_N - 1);
        const bool flip = (a >> 10) & 1;
#pragma unroll
        for (int s = 0; s < 32; s++) {
            const int idx = ntt_in_index(lane, s);
            i32 c;
            if (ROTATE) {
                i32 src = w.acc[(idx - ar) & (NTT_N - 1)];
                bool neg = (idx < ar) != flip;
                c = (i32)((neg ? 0u - (u32)src : (u32)src) - (u32)w.acc[idx]);
            } else {
                c = w.acc[idx];
            }
            i32 d0, d1;
            decompose(c, d0, d1);
            v0[s] = ff_from_i32(d0);
            v1[s] = ff_from_i32(d1);
        }
    }
    NB_LOCKSTEP();
    // 2. forward transforms of the two digits
    warp_ntt_forward(v0, w.scratch, w.twd_fwd, lane);
    warp_ntt_forward(v1, w.scratch, w.twd_fwd, lane);
    // 3. multiply-accumulate with this warp's half of the key row (tgsw_gpu.py:58-107)
    {
        const ulonglong2 *bk = reinterpret_cast<const ulonglong2 *>(bk_row) + (size_t)w.mi * (32 * 2 * 32) + lane;
#pragma unroll
        for (int t = 0; t < 32; t++) {
            ulonglong2 b0 = __ldg(bk + (t * 2 + 0) * 32);
            ulonglong2 b1 = __ldg(bk + (t * 2 + 1) * 32);
            u64 p0 = ff_mul2_add(v0[t], b0.x, v1[t], b1.x);      // contribution to output polynomial 0
            u64 p1 = ff_mul2_add(v0[t], b0.y, v1[t], b1.y);      // contribution to output polynomial 1
            v0[t] = w.mi ? p1 : p0;
            w.xchg_out[t * 32 + lane] = w.mi ? p0 : p1;
            if (t % 4 == 3) NB_LOCKSTEP();
        }
    }
    pair_barrier(w.bar_id);
#pragma unroll
    for (int t = 0; t < 32; t++) v0[t] = ff_add(v0[t], w.xchg_in[t * 32 + lane]);
    pair_barrier(w.bar_id);     // partner finished reading our buffer before the next step rewrites it
    // 4. inverse transform of output polynomial mi, back to Torus32 (ntt.mako:402-408)
    warp_ntt_inverse(v0, w.scratch, w.twd_inv, lane);
#pragma unroll
    for (int s = 0; s < 32; s++) {
        const int idx = ntt_in_index(lane, s);
        i32 r = ff_to_i32(v0[s]);
        if (ROTATE) w.acc[idx] = (i32)((u32)w.acc[idx] + (u32)r);
        else w.acc[idx] = r;
    }
    __syncwarp();
}

// mod-switch to [0, 2N) (numeric_functions_gpu.py:55-71)
NB_HD i32 modswitch_2n(i32 x) { return (i32)(((u32)x + (1u << 20)) >> 21); }

constexpr int BR_CT_PER_CTA = 4;
constexpr int BR_THREADS = BR_CT_PER_CTA * 64;
constexpr size_t BR_SMEM_PER_WARP = NTT_N * sizeof(i32) + TR_WORDS * sizeof(u64) + NTT_N * sizeof(u64);
constexpr size_t BR_SMEM_BYTES = 2 * NTT_N * sizeof(u64) + 2 * BR_CT_PER_CTA * BR_SMEM_PER_WARP;

struct BlindRotateArgs {
    // mode A (gate): x = c + s1 * in1 + s2 * in2 is formed on the fly (gates.py prologues), then
    //                bootstrap(mu, x) (bootstrap.py:206-229).  in2 may be null (s2 ignored).
    const i32 *in1_a, *in1_b, *in2_a, *in2_b;
    i32 c, s1, s2, mu;
    // mode B (BlindRotate_gpu, blind_rotate.py:262-281): explicit accumulator (B,2,1024) and bara (B,n)
    const i32 *accum, *bara;
    const u64 *bk;          // internal layout, n rows
    i32 *out_a, *out_b;     // extracted LWE samples (B,1024), (B,)
    i32 *accum_out;         // optional: final accumulators (B,2,1024)
    int n;                  // LWE dimension (500)
    int extract;            // write out_a/out_b
    size_t batch;
};

__global__ void __launch_bounds__(BR_THREADS, 1) blind_rotate_kernel(BlindRotateArgs p, const u64 *__restrict__ twd_fwd_g,
                                                                      const u64 *__restrict__ twd_inv_g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *twd_fwd = reinterpret_cast<u64 *>(smem_raw);
    u64 *twd_inv = twd_fwd + NTT_N;
    for (int i = threadIdx.x; i < NTT_N; i += blockDim.x) { twd_fwd[i] = twd_fwd_g[i]; twd_inv[i] = twd_inv_g[i]; }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct_local = warp >> 1, mi = warp & 1;
    unsigned char *wbase = smem_raw + 2 * NTT_N * sizeof(u64);
    auto warp_mem = [&](int wi) { return wbase + (size_t)wi * BR_SMEM_PER_WARP; };
    WarpState w;
    w.lane = lane; w.mi = mi; w.bar_id = 1 + ct_local;
    w.acc = reinterpret_cast<i32 *>(warp_mem(warp));
    w.scratch = reinterpret_cast<u64 *>(warp_mem(warp) + NTT_N * sizeof(i32));
    w.xchg_out = w.scratch + TR_WORDS;
    w.xchg_in = reinterpret_cast<u64 *>(warp_mem(warp ^ 1) + NTT_N * sizeof(i32)) + TR_WORDS;
    w.twd_fwd = twd_fwd; w.twd_inv = twd_inv;
    __syncthreads();

    const size_t ct = (size_t)blockIdx.x * BR_CT_PER_CTA + ct_local;
    // Warps of a ciphertext beyond the batch still run (on ciphertext batch-1) so that the named
    // barriers stay balanced; they just do not store anything.
    const bool live = ct < p.batch;
    const size_t c = live ? ct : p.batch - 1;
    const int n = p.n;

    // accumulator initialisation
    if (p.accum) {
        for (int s = 0; s < 32; s++) w.acc[s * 32 + lane] = p.accum[(c * 2 + mi) * NTT_N + s * 32 + lane];
    } else {
        // ACC = (0, X^(2N - barb) * [mu, ..., mu])   (bootstrap.py:177-182, 224)
        i32 xb = p.c + p.s1 * p.in1_b[c] + (p.in2_b ? p.s2 * p.in2_b[c] : 0);
        int q = 2 * NTT_N - modswitch_2n(xb);
        for (int s = 0; s < 32; s++) {
            int x = s * 32 + lane;
            i32 val;
            if (q < NTT_N) val = x < q ? (i32)(0u - (u32)p.mu) : p.mu;
            else val = x < q - NTT_N ? p.mu : (i32)(0u - (u32)p.mu);
            w.acc[x] = mi ? val : 0;
        }
    }
    __syncwarp();

    for (int i = 0; i < n; i++) {
        int a;
        if (p.bara) a = p.bara[c * n + i];
        else {
            i32 xa = p.s1 * p.in1_a[c * n + i] + (p.in2_a ? p.s2 * p.in2_a[c * n + i] : 0);
            a = modswitch_2n(xa);
        }
        external_product_step<true>(w, p.bk + (size_t)i * 8 * NTT_N, a);
    }

    if (live) {
        if (p.accum_out)
            for (int s = 0; s < 32; s++) p.accum_out[(ct * 2 + mi) * NTT_N + s * 32 + lane] = w.acc[s * 32 + lane];
        if (p.extract) {
            // sample extraction (tlwe_gpu.mako:63-82; blind_rotate.mako:213-224)
            if (mi == 0) {
                for (int s = 0; s < 32; s++) {
                    int x = s * 32 + lane;
                    p.out_a[ct * NTT_N + x] = x == 0 ? w.acc[0] : (i32)(0u - (u32)w.acc[NTT_N - x]);
                }
            } else if (lane == 0) {
                p.out_b[ct] = w.acc[0];
            }
        }
    }
}

// stand-alone external product: accum (B,2,1024) <- bk_row (x) accum   (tgsw.py:165-172)
__global__ void __launch_bounds__(BR_THREADS, 1) external_product_kernel(i32 *accum, const u64 *__restrict__ bk_row,
                                                                          size_t batch, const u64 *__restrict__ twd_fwd_g,
                                                                          const u64 *__restrict__ twd_inv_g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *twd_fwd = reinterpret_cast<u64 *>(smem_raw);
    u64 *twd_inv = twd_fwd + NTT_N;
    for (int i = threadIdx.x; i < NTT_N; i += blockDim.x) { twd_fwd[i] = twd_fwd_g[i]; twd_inv[i] = twd_inv_g[i]; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct_local = warp >> 1, mi = warp & 1;
    unsigned char *wbase = smem_raw + 2 * NTT_N * sizeof(u64);
    WarpState w;
    w.lane = lane; w.mi = mi; w.bar_id = 1 + ct_local;
    w.acc = reinterpret_cast<i32 *>(wbase + (size_t)warp * BR_SMEM_PER_WARP);
    w.scratch = reinterpret_cast<u64 *>(wbase + (size_t)warp * BR_SMEM_PER_WARP + NTT_N * sizeof(i32));
    w.xchg_out = w.scratch + TR_WORDS;
    w.xchg_in = reinterpret_cast<u64 *>(wbase + (size_t)(warp ^ 1) * BR_SMEM_PER_WARP + NTT_N * sizeof(i32)) + TR_WORDS;
    w.twd_fwd = twd_fwd; w.twd_inv = twd_inv;
    __syncthreads();
    const size_t ct = (size_t)blockIdx.x * BR_CT_PER_CTA + ct_local;
    const bool live = ct < batch;
    const size_t c = live ? ct : batch - 1;
    for (int s = 0; s < 32; s++) w.acc[s * 32 + lane] = accum[(c * 2 + mi) * NTT_N + s * 32 + lane];
    __syncwarp();
    external_product_step<false>(w, bk_row, 0);
    if (live)
        for (int s = 0; s < 32; s++) accum[(ct * 2 + mi) * NTT_N + s * 32 + lane] = w.acc[s * 32 + lane];
}

// ---- LWE key switch (lwe_gpu.mako:59-118; lwe_cpu.py:62-93) --------------------------------------
// One CTA handles KS_TILE ciphertexts; thread i owns output coefficient i (i = n: the b term) of all
// of them, so each key row is fetched once per tile.  src = src1 (+ src2) (+ (0, c)) lets gate_mux
// fold `(0,1/8) + u1 + u2` (gates.py:657-664) into the load.
constexpr int KS_TILE = 8;
constexpr int KS_THREADS = 512;

struct KeyswitchArgs {
    const i32 *src1_a, *src1_b, *src2_a, *src2_b;   // (B, in), (B,)
    i32 c;
    const i32 *ks_a, *ks_b;                          // (in, t, base, n), (in, t, base)
    const float *ks_cv;
    i32 *res_a, *res_b;                              // (B, n), (B,)
    float *res_cv;                                   // optional
    int in_size, n, t, log2_base;
    size_t batch;
};

__global__ void __launch_bounds__(KS_THREADS) keyswitch_kernel(KeyswitchArgs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    i32 *tile_a = reinterpret_cast<i32 *>(smem_raw);            // KS_TILE * in_size, with the rounding offset added
    const size_t ct0 = (size_t)blockIdx.x * KS_TILE;
    const int nct = (int)min((size_t)KS_TILE, p.batch - ct0);
    const int base = 1 << p.log2_base;
    const u32 prec_offset = 1u << (32 - (1 + p.log2_base * p.t));
    for (int idx = threadIdx.x; idx < KS_TILE * p.in_size; idx += blockDim.x) {
        int q = idx / p.in_size, j = idx % p.in_size;
        i32 v = 0;
        if (q < nct) {
            v = p.src1_a[(ct0 + q) * p.in_size + j];
            if (p.src2_a) v = (i32)((u32)v + (u32)p.src2_a[(ct0 + q) * p.in_size + j]);
        }
        tile_a[idx] = (i32)((u32)v + prec_offset);
    }
    __syncthreads();
    const int i = threadIdx.x;
    const bool is_a = i < p.n, is_b = i == p.n;
    u32 acc[KS_TILE];
    float cv[KS_TILE];
#pragma unroll
    for (int q = 0; q < KS_TILE; q++) { acc[q] = 0; cv[q] = 0.f; }
    if (is_a || is_b) {
        for (int j = 0; j < p.in_size; j++) {
            for (int k = 0; k < p.t; k++) {
                const size_t row0 = ((size_t)j * p.t + k) * base;
                const int shift = 32 - (k + 1) * p.log2_base;
#pragma unroll
                for (int q = 0; q < KS_TILE; q++) {
                    int d = (tile_a[q * p.in_size + j] >> shift) & (base - 1);
                    if (d != 0) {          // the d = 0 row is the zero padding (lwe_cpu.py:31-33)
                        if (is_a) acc[q] -= (u32)__ldg(p.ks_a + (row0 + d) * p.n + i);
                        else { acc[q] -= (u32)__ldg(p.ks_b + row0 + d); cv[q] += __ldg(p.ks_cv + row0 + d); }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KS_TILE; q++) {
            if (q < nct) {
                if (is_a) p.res_a[(ct0 + q) * p.n + i] = (i32)acc[q];
                else {
                    u32 b = (u32)p.src1_b[ct0 + q] + (p.src2_b ? (u32)p.src2_b[ct0 + q] : 0u) + (u32)p.c;
                    p.res_b[ct0 + q] = (i32)(b + acc[q]);
                    if (p.res_cv) p.res_cv[ct0 + q] = cv[q];
                }
            }
        }
    }
}

// ---- LWE linear ops (lwe_gpu.mako:123-202): res = c_b * (0,1) + s1 * x1 + s2 * x2 on dense arrays ---
__global__ void lwe_affine_kernel(i32 *res_a, i32 *res_b, const i32 *x1_a, const i32 *x1_b, const i32 *x2_a,
                                  const i32 *x2_b, i32 c, i32 s1, i32 s2, size_t batch, int n)
{
    const size_t total = batch * (size_t)(n + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (i < batch * (size_t)n) {
            u32 v = x1_a ? (u32)s1 * (u32)x1_a[i] : 0u;
            if (x2_a) v += (u32)s2 * (u32)x2_a[i];
            res_a[i] = (i32)v;
        } else {
            size_t b = i - batch * (size_t)n;
            u32 v = (u32)c + (x1_b ? (u32)s1 * (u32)x1_b[b] : 0u);
            if (x2_b) v += (u32)s2 * (u32)x2_b[b];
            res_b[b] = (i32)v;
        }
    }
}

}  // namespace nb
