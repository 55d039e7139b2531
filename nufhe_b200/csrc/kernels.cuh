// kernels.cuh -- sm_100a kernels of the TFHE gate-bootstrapping hot path.
//
// Work decomposition of the fused bootstrap: a CTA owns one or two ciphertexts whose accumulators and work
// polynomials live in shared memory; every CMux step is seven CTA-wide phases (br_phases.cuh) separated by
// __syncthreads, 16 field elements per thread and phase.  The kernel is persistent: at most one wave of CTAs, which
// pull (chain, chunk) work items from a FIFO of ready chains in global memory, so that any batch is spread over all SMs.
// The smallest batches take a second kernel that spreads ONE ciphertext over a cluster of two SMs (pair shape: the step
// divides along the two accumulator polynomials, one 8 KB exchange per step through distributed shared memory).
// The stand-alone transforms reuse the same phases behind cp.async-staged, 128-bit input / output.  The key switch is a
// TMA-fed producer / consumer pipeline.  Then the separate steps of the reference's multi-kernel bootstrap, and the
// key-generation kernels (LWE dot product, key-switch key).
//
// Reference counterparts: nufhe/blind_rotate.mako:18-226 (fused bootstrap), tgsw_gpu.py:110-169
// (external product), transform/computation.mako:18-143 (stand-alone transform), lwe_gpu.mako:59-118
// (key switch).  Nothing here is derived from those kernels' structure (128 threads/transform,
// 8*2*8*8 radix plan, 15 block barriers per step); see ntt_lane.cuh for the transform we use and DESIGN.md
// section 4 for the measurements behind each choice.
#pragma once
#include <cuda_runtime.h>
#include "ntt_lane.cuh"
#include "br_phases.cuh"

namespace nb {

constexpr int LWE_N_MAX = 1024;          // upper bound on the LWE dimension n handled by the gate kernels

NB_D u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
NB_D void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
NB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> NB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- stand-alone batched transforms (reference: transform/computation.mako `standalone_transform`) ---
// natural order in and out.  Same three passes as the fused bootstrap (br_phases.cuh), 4 polynomials per
// sweep of 256 threads, 2 persistent CTAs per SM.  The natural-order side never meets the arithmetic directly:
//   * the next sweep's input is fetched with cp.async (16 bytes per request) into a second staging buffer while the
//     current sweep computes, so no pass waits on global memory;
//   * a thread owns the natural-order pairs (2t, 2t+1) and (2t+512, 2t+513) of every polynomial: in the pass layout
//     those four elements sit in two 16-byte shared-memory words 8 rows apart (ntt_pair_position), so the u64 side is
//     moved with 128-bit shared AND 128-bit global accesses (512 contiguous bytes per warp), bank-conflict free;
//   * the int32 side is read from / written to global memory as 128-byte warp rows (x[64 j1 + j2], lanes = j2).
#ifndef NB_NTT_CTAS
#define NB_NTT_CTAS 3                      // resident CTAs per SM of the stand-alone transforms (2: -3.6 %, profiles/r2_variants.md)
#endif
constexpr int NTT_CTAS = NB_NTT_CTAS;
constexpr int NTT_RAW_I32_BYTES = NTT_SWEEP_POLYS * NTT_N * (int)sizeof(i32);
constexpr int NTT_RAW_U64_BYTES = NTT_SWEEP_POLYS * NTT_N * (int)sizeof(u64);
// staging buffers per CTA: two (prefetch one sweep ahead) unless three CTAs of u64 input have to share the SM
NB_HDC int ntt_raw_buffers(int raw_bytes) { return (NTT_CTAS > 2 && raw_bytes > NTT_RAW_I32_BYTES) ? 1 : 2; }
NB_HDC size_t ntt_smem_bytes(int raw_bytes)
{
    return (size_t)NTT_SWEEP_POLYS * POLY_STRIDE * sizeof(u64) + NTT_N * sizeof(u64) + ntt_raw_buffers(raw_bytes) * (size_t)raw_bytes;
}

// shared-memory position (in u64) of natural element k = 2 t of a polynomial, t = threadIdx.x in [0, 256): the
// elements 2t+1, 2t+512 and 2t+513 are at +8 rows, +1 and +8 rows +1 (k1 = k % 16 is even, so brev4(k1 + 1) =
// brev4(k1) + 8; adding 512 to k sets the top bit of k2 >> 2, i.e. bit 0 of the stored column index)
NB_D int ntt_pair_position(int t)
{
    const int row = brev(2 * (t & 7), 4), u = brev((t >> 3) & 3, 2), i0 = brev(t >> 5, 4);
    return row * ROW_STRIDE + col_of(u, i0);
}
NB_D u64 ff_zero_if_p(u64 x) { return x == FF_P ? 0 : x; }       // [0, p] -> canonical

// stage one sweep (4 polynomials, clamped at the end of the batch) with 16-byte cp.async requests
template <int POLY_BYTES> NB_D void ntt_stage(unsigned char *raw, const unsigned char *in, size_t p0, size_t batch, int tid)
{
    constexpr int CHUNKS = POLY_BYTES / 16;
#pragma unroll
    for (int pl = 0; pl < NTT_SWEEP_POLYS; pl++) {
        const size_t p = min(p0 + pl, batch - 1);
        for (int c = tid; c < CHUNKS; c += NTT_SWEEP_THREADS)
            cp_async16(raw + pl * POLY_BYTES + c * 16, in + p * POLY_BYTES + (size_t)c * 16);
    }
}

template <bool IN_I32>
__global__ void __launch_bounds__(NTT_SWEEP_THREADS, NTT_CTAS) ntt_forward_kernel(const void *__restrict__ in, u64 *__restrict__ out,
                                                                            const u64 *__restrict__ twd_g, size_t batch)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int POLY_BYTES = NTT_N * (IN_I32 ? 4 : 8), RAW_BYTES = NTT_SWEEP_POLYS * POLY_BYTES;
    u64 *w = reinterpret_cast<u64 *>(smem_raw);
    u64 *twd = w + NTT_SWEEP_POLYS * POLY_STRIDE;
    unsigned char *raw = reinterpret_cast<unsigned char *>(twd + NTT_N);
    const int tid = threadIdx.x;
    const size_t stride = (size_t)gridDim.x * NTT_SWEEP_POLYS;
    size_t p0 = (size_t)blockIdx.x * NTT_SWEEP_POLYS;
    if (p0 < batch) ntt_stage<POLY_BYTES>(raw, (const unsigned char *)in, p0, batch, tid);
    cp_async_commit();
    for (int i = tid; i < NTT_N; i += NTT_SWEEP_THREADS) twd[i] = twd_g[i];
    const int pos = ntt_pair_position(tid);
    constexpr int NBUF = ntt_raw_buffers(RAW_BYTES);
    int buf = 0;
    for (; p0 < batch; p0 += stride, buf ^= (NBUF - 1)) {
        if (NBUF == 2) {
            if (p0 + stride < batch) ntt_stage<POLY_BYTES>(raw + (buf ^ 1) * RAW_BYTES, (const unsigned char *)in, p0 + stride, batch, tid);
            cp_async_commit();
            cp_async_wait<1>();                   // everything but the group just committed has landed
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        {   // pass 1: thread = (poly, j2), reads x[64 j1 + j2] from the staging buffer
            const int pl = tid >> 6, j2 = tid & 63;
            const unsigned char *src = raw + buf * RAW_BYTES + pl * POLY_BYTES;
            if (IN_I32) {
                i32 x[16];
#pragma unroll
                for (int j1 = 0; j1 < 16; j1++) x[j1] = reinterpret_cast<const i32 *>(src)[64 * j1 + j2];
                phase_fwd1_i32(tid, x, w, twd);
            } else {
                u64 x[16];
#pragma unroll
                for (int j1 = 0; j1 < 16; j1++) x[j1] = ff_canon(reinterpret_cast<const u64 *>(src)[64 * j1 + j2]);
                phase_fwd1_generic(tid, x, w, twd);
            }
        }
        __syncthreads();
        { const int g = tid >> 6, x = tid & 63; phase_fwd2(x >> 4, x & 15, g, w); }
        __syncthreads();
        if (NBUF == 1 && p0 + stride < batch) {   // single staging buffer: refill it now that pass 1 has consumed it
            ntt_stage<POLY_BYTES>(raw, (const unsigned char *)in, p0 + stride, batch, tid);
            cp_async_commit();
        }
        { const int u = tid & 3, r = (tid >> 2) & 15, p = tid >> 6; phase_fwd3(p, r, u, w); }
        __syncthreads();
        // natural-order store: 16 bytes = elements (2t, 2t+1) and (2t+512, 2t+513) of each polynomial
#pragma unroll
        for (int pl = 0; pl < NTT_SWEEP_POLYS; pl++) {
            if (p0 + pl < batch) {
                u64 a0, a1, b0, b1;
                ld2(w + pl * POLY_STRIDE + pos, a0, a1);                       // k = 2t, 2t + 512
                ld2(w + pl * POLY_STRIDE + pos + 8 * ROW_STRIDE, b0, b1);      // k = 2t + 1, 2t + 513
                ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out + (p0 + pl) * NTT_N);
                dst[tid] = make_ulonglong2(ff_zero_if_p(a0), ff_zero_if_p(b0));
                dst[tid + 256] = make_ulonglong2(ff_zero_if_p(a1), ff_zero_if_p(b1));
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

template <bool OUT_I32>
__global__ void __launch_bounds__(NTT_SWEEP_THREADS, NTT_CTAS) ntt_inverse_kernel(const u64 *__restrict__ in, void *__restrict__ out,
                                                                            const u64 *__restrict__ twd_g, size_t batch)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int POLY_BYTES = NTT_N * 8, RAW_BYTES = NTT_SWEEP_POLYS * POLY_BYTES;
    u64 *w = reinterpret_cast<u64 *>(smem_raw);
    u64 *twd = w + NTT_SWEEP_POLYS * POLY_STRIDE;
    unsigned char *raw = reinterpret_cast<unsigned char *>(twd + NTT_N);
    const int tid = threadIdx.x;
    const size_t stride = (size_t)gridDim.x * NTT_SWEEP_POLYS;
    size_t p0 = (size_t)blockIdx.x * NTT_SWEEP_POLYS;
    if (p0 < batch) ntt_stage<POLY_BYTES>(raw, (const unsigned char *)in, p0, batch, tid);
    cp_async_commit();
    for (int i = tid; i < NTT_N; i += NTT_SWEEP_THREADS) twd[i] = twd_g[i];
    const int pos = ntt_pair_position(tid);
    constexpr int NBUF = ntt_raw_buffers(RAW_BYTES);
    int buf = 0;
    for (; p0 < batch; p0 += stride, buf ^= (NBUF - 1)) {
        if (NBUF == 2) {
            if (p0 + stride < batch) ntt_stage<POLY_BYTES>(raw + (buf ^ 1) * RAW_BYTES, (const unsigned char *)in, p0 + stride, batch, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        // natural order -> pass layout (the mirror image of the forward kernel's store), canonicalising on the way
#pragma unroll
        for (int pl = 0; pl < NTT_SWEEP_POLYS; pl++) {
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(raw + buf * RAW_BYTES + pl * POLY_BYTES);
            const ulonglong2 lo = src[tid], hi = src[tid + 256];
            st2(w + pl * POLY_STRIDE + pos, ff_canon(lo.x), ff_canon(hi.x));
            st2(w + pl * POLY_STRIDE + pos + 8 * ROW_STRIDE, ff_canon(lo.y), ff_canon(hi.y));
        }
        __syncthreads();
        if (NBUF == 1 && p0 + stride < batch) {   // single staging buffer: refill it now that it has been consumed
            ntt_stage<POLY_BYTES>(raw, (const unsigned char *)in, p0 + stride, batch, tid);
            cp_async_commit();
        }
        { const int u = tid & 3, r = (tid >> 2) & 15, p = tid >> 6; phase_inv3(p, r, u, w); }
        __syncthreads();
        { const int g = tid >> 6, x = tid & 63; phase_inv2(x >> 4, x & 15, g, w); }
        __syncthreads();
        {
            const int pl = tid >> 6, j2 = tid & 63;
            if (OUT_I32) {
                i32 y[16];
                phase_inv1_i32(tid, y, w, twd);
                if (p0 + pl < batch) {
#pragma unroll
                    for (int j1 = 0; j1 < 16; j1++) ((i32 *)out)[(p0 + pl) * NTT_N + 64 * j1 + j2] = y[j1];
                }
            } else {
                u64 y[16];
                phase_inv1_generic(tid, y, w, twd);
                if (p0 + pl < batch) {
#pragma unroll
                    for (int j1 = 0; j1 < 16; j1++) ((u64 *)out)[(p0 + pl) * NTT_N + 64 * j1 + j2] = ff_zero_if_p(y[j1]);
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

// ---- element-wise field ops (unit tests of the arithmetic; key generation) ----------------------
enum FfOp { FF_OP_ADD = 0, FF_OP_SUB = 1, FF_OP_MUL = 2, FF_OP_MUL_PREPARED = 3, FF_OP_PREPARE = 4, FF_OP_LSH = 5,
            FF_OP_LSH_CONST = 6 };   // 6: the compile-time-shift code paths of the transforms, for unit tests

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
        case FF_OP_LSH_CONST: {
            const int sh = (int)(y % 192);
            r = x;
            static_for<0, 192>([&](auto S) { if (sh == decltype(S)::value) r = ff_shl<decltype(S)::value>(x); });
            break;
        }
        default: r = ff_shl_var(x, (int)(y % 192)); break;
        }
        out[i] = ff_canon(r);
    }
}

// ---- bootstrap-key layout ------------------------------------------------------------------------
// reference row (nufhe/blind_rotate.py:112): [mi][j][mo][k], values NTT(bk) * 2^64 (Montgomery form).
// internal row: [m = (mi*2+j)*2+mo][row * 64 + stored column] (br_phases.cuh), plain values, so that the
// MAC phase reads, per thread, one 16-byte pair from each of the 8 planes, coalesced across the CTA.
// One thread per (row, position): 8 plain planes + the 2 correction planes (br_phases.cuh: BK_PLANES).
__global__ void bk_prepare_kernel(const u64 *__restrict__ bk_ref, u64 *__restrict__ bk_int,
                                  const u64 *__restrict__ ones512, size_t rows)
{
    const size_t total = rows * NTT_N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i >> 10;
        const int pos = (int)(i & 1023);
        const int k = w_natural_index(pos >> 6, pos & 63);
        u64 sum[2] = {0, 0};
#pragma unroll
        for (int m = 0; m < 8; m++) {
            u64 x = ff_mul(ff_canon(bk_ref[row * 8 * NTT_N + m * NTT_N + k]), FF_RINV);
            bk_int[row * BK_ROW_U64 + m * NTT_N + pos] = x;
            sum[m & 1] = ff_add(sum[m & 1], x);
        }
        bk_int[row * BK_ROW_U64 + 8 * NTT_N + pos] = ff_mul(sum[0], ones512[k]);
        bk_int[row * BK_ROW_U64 + 9 * NTT_N + pos] = ff_mul(sum[1], ones512[k]);
    }
}

// mod-switch to [0, 2N) (numeric_functions_gpu.py:55-71)
NB_HD i32 modswitch_2n(i32 x) { return (i32)(((u32)x + (1u << 20)) >> 21); }

// ---- fused bootstrap: CTA-wide phases over shared-memory-resident polynomials (br_phases.cuh) -----
template <class Cfg> constexpr size_t br_smem_bytes()
{
    return (size_t)Cfg::CT * 2 * NTT_N * sizeof(i32) + (size_t)Cfg::POLYS * POLY_STRIDE * sizeof(u64) +
           (Cfg::TWD_GLOBAL ? 0 : 2 * NTT_N * sizeof(u64)) + 64 + (Cfg::STAGE_KEY ? BK_ROW_U64 * sizeof(u64) : 0);
}
constexpr size_t BR2_SMEM_BYTES = br_smem_bytes<BrDefault>();

struct BlindRotateArgs {
    // mode A (gate): x = c + s1 * in1 + s2 * in2 is formed on the fly (gates.py prologues), then
    //                bootstrap(mu, x) (bootstrap.py:206-229).  in2 may be null (s2 ignored).
    const i32 *in1_a, *in1_b, *in2_a, *in2_b;
    i32 c, s1, s2, mu;
    // optional second job over the same launch (gate_mux's two bootstraps, gates.py:638-655): ciphertexts
    // [job_batch, 2 job_batch) use these operands / signs instead; outputs are simply the next rows
    const i32 *j2_in1_a, *j2_in1_b, *j2_in2_a, *j2_in2_b;
    i32 j2_c, j2_s1, j2_s2;
    size_t job_batch;       // 0: single job
    // mode B (BlindRotate_gpu, blind_rotate.py:262-281): explicit accumulator (B,2,1024) and bara (B,n)
    const i32 *accum, *bara;
    const u64 *bk;          // internal layout, n rows
    i32 *out_a, *out_b;     // extracted LWE samples (B,1024), (B,)
    i32 *accum_out;         // optional: final accumulators (B,2,1024)
    int n;                  // LWE dimension (500); n = 0 with `plain` = one external product
    int extract;            // write out_a/out_b
    int plain;              // 1: accum <- bk[0] (x) accum, no rotation (tgsw.py:165-172), n ignored
    size_t batch;
    // work queue (see blind_rotate_kernel): `chains` groups of Cfg::CT ciphertexts, each cut into `chunks` runs of
    // `steps_per_chunk` CMux steps.  sched == null: one chain per CTA, no queue (single wave).
    unsigned *sched;        // [0] = queue head, [1] = entries appended so far, [BR_SCHED_HEADER + t] = entry t of the ready queue
    i32 *state;             // accumulators parked between the chunks of a chain: (chains * CT, 2, 1024)
    unsigned chains, chunks;
    int steps_per_chunk;
    int sm_count, stagger_cycles;   // single-wave launches: the second CTA of every SM starts this much later
};
constexpr int BR_SCHED_HEADER = 32;      // the ticket counter has a 128-byte line of its own

struct Br2Smem {
    i32 *acc;      // [CT][2][1024]
    u64 *w;        // [16][POLY_STRIDE]
    u64 *twd_fwd;  // [16][64]
    u64 *twd_inv;
    int *rot;      // [2][CT], 64 bytes
    u64 *key;      // Cfg::STAGE_KEY: the key row of the current step, [BK_PLANES][1024]
};

template <class Cfg> NB_D Br2Smem br2_carve(unsigned char *raw)
{
    Br2Smem s;
    s.w = reinterpret_cast<u64 *>(raw);
    s.twd_fwd = s.w + Cfg::POLYS * POLY_STRIDE;
    s.twd_inv = s.twd_fwd + (Cfg::TWD_GLOBAL ? 0 : NTT_N);
    s.acc = reinterpret_cast<i32 *>(s.twd_inv + (Cfg::TWD_GLOBAL ? 0 : NTT_N));
    s.rot = reinterpret_cast<int *>(s.acc + Cfg::CT * 2 * NTT_N);
    s.key = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(s.rot) + 64);
    return s;
}

// rotation amount of ciphertext c at step i
NB_D int br2_rotation(const BlindRotateArgs &p, size_t c, int i)
{
    if (p.bara) return p.bara[c * p.n + i];
    if (p.job_batch && c >= p.job_batch) {
        const size_t d = c - p.job_batch;
        u32 xa = (u32)p.j2_s1 * (u32)p.j2_in1_a[d * p.n + i] + (p.j2_in2_a ? (u32)p.j2_s2 * (u32)p.j2_in2_a[d * p.n + i] : 0u);
        return modswitch_2n((i32)xa);
    }
    // Torus32 arithmetic wraps by design (XOR / XNOR use s = +-2): unsigned, like lwe_affine_kernel
    u32 xa = (u32)p.s1 * (u32)p.in1_a[c * p.n + i] + (p.in2_a ? (u32)p.s2 * (u32)p.in2_a[c * p.n + i] : 0u);
    return modswitch_2n((i32)xa);
}

// coefficient x of polynomial mi of ciphertext c's accumulator before the first step
NB_D i32 br2_initial_acc(const BlindRotateArgs &p, size_t c, int mi, int x)
{
    if (p.accum) return p.accum[(c * 2 + mi) * NTT_N + x];
    if (mi == 0) return 0;
    // ACC = (0, X^(2N - barb) * [mu, ..., mu])   (bootstrap.py:177-182, 224)
    u32 xb;
    if (p.job_batch && c >= p.job_batch) {
        const size_t d = c - p.job_batch;
        xb = (u32)p.j2_c + (u32)p.j2_s1 * (u32)p.j2_in1_b[d] + (p.j2_in2_b ? (u32)p.j2_s2 * (u32)p.j2_in2_b[d] : 0u);
    } else {
        xb = (u32)p.c + (u32)p.s1 * (u32)p.in1_b[c] + (p.in2_b ? (u32)p.s2 * (u32)p.in2_b[c] : 0u);
    }
    const int q = 2 * NTT_N - modswitch_2n((i32)xb);
    if (q < NTT_N) return x < q ? (i32)(0u - (u32)p.mu) : p.mu;
    return x < q - NTT_N ? p.mu : (i32)(0u - (u32)p.mu);
}

// 16-byte cp.async requests for one key row, spread over `nthreads` threads (t = 0 .. nthreads - 1); one commit group
NB_D void br2_stage_key(u64 *key, const u64 *__restrict__ bk_row, int t, int nthreads)
{
    for (int x = t; x < BK_ROW_U64 / 2; x += nthreads) cp_async16(key + 2 * x, bk_row + 2 * x);
    cp_async_commit();
}

// next_bk_row (Cfg::STAGE_KEY): the key row of the following step, or null
template <bool ROTATE, class Cfg>
NB_D void br2_step(const Br2Smem &s, const u64 *__restrict__ bk_row, const int *rot, int tid, const u64 *__restrict__ next_bk_row = nullptr)
{
    // forward transforms of the digit polynomials (CT ciphertexts x 2 polynomials x 2 digits)
    if constexpr (Cfg::SPLIT_FWD) {
        // two threads (of different warps) per task, 8 outputs each; tasks mapped as in the 256-thread shape
        using Tasks = BrCfg<Cfg::CT, 256 * Cfg::CT>;
        int h, t;
        map_split_fwd<Cfg>(tid, h, t);
        if (h) phase_fwd1_split<ROTATE, 1>(t, s.acc, s.w, s.twd_fwd, rot); else phase_fwd1_split<ROTATE, 0>(t, s.acc, s.w, s.twd_fwd, rot);
        __syncthreads();
        { int p, r, g; map_fwd2<Tasks>(t, 0, p, r, g); if (h) phase_fwd2_split<1>(p, r, g, s.w); else phase_fwd2_split<0>(p, r, g, s.w); }
        __syncthreads();
        {
            int p, r, u;
            u64 v[16];
            map_fwd3<Tasks>(t, 0, p, r, u);
            phase_fwd3_split_load(p, r, u, s.w, v);
            __syncthreads();                               // in place: all loads before any store
            if (h) phase_fwd3_split_finish<1>(p, r, u, s.w, v); else phase_fwd3_split_finish<0>(p, r, u, s.w, v);
        }
        __syncthreads();
    } else {
        if constexpr (Cfg::FWD1_BOTH_DIGITS) {
            phase_fwd1_both_digits<ROTATE>(tid, s.acc, s.w, s.twd_fwd, rot);
        } else {
#pragma unroll 1
            for (int it = 0; it < Cfg::FWD_SWEEPS; it++) phase_fwd1<ROTATE>(it * Cfg::THREADS + tid, s.acc, s.w, s.twd_fwd, rot);
        }
        __syncthreads();
#pragma unroll 1
        for (int it = 0; it < Cfg::FWD_SWEEPS; it++) { int p, r, g; map_fwd2<Cfg>(tid, it, p, r, g); phase_fwd2(p, r, g, s.w); }
        __syncthreads();
#pragma unroll 1
        for (int it = 0; it < Cfg::FWD_SWEEPS; it++) { int p, r, u; map_fwd3<Cfg>(tid, it, p, r, u); phase_fwd3(p, r, u, s.w); }
        __syncthreads();
    }
    // multiply-accumulate with the key row (tgsw_gpu.py:58-107); each key element is fetched once per CTA
    if constexpr (Cfg::STAGE_KEY) phase_mac<Cfg, true>(tid, s.w, s.key);
    else phase_mac<Cfg>(tid, s.w, bk_row);
    __syncthreads();
    // inverse transforms of the output polynomials
    if constexpr (Cfg::SPLIT_INV) {
        // twice as many threads as tasks: two threads (of different warps) per task, 8 elements each, parked values
        // exchanged through the dead work polynomials (br_phases.cuh: split inverse phases); two more barriers
        int h, t;
        // threads beyond twice the task count (the upper half of the wide2 CTA, whole warps) have no inverse work: they
        // go straight to the barrier that ends the step, and the working warps synchronise among themselves on a
        // named barrier
        if (map_split<Cfg>(tid, h, t)) {
            constexpr int WORKERS = 2 * Cfg::INV_TASKS;
            auto sync_workers = [] {
                if constexpr (WORKERS == Cfg::THREADS) __syncthreads();
                else asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");
            };
            { int p, r, u; map_inv3<Cfg>(t, p, r, u); if (h) phase_inv3_split_a<1>(p, r, u, s.w); else phase_inv3_split_a<0>(p, r, u, s.w); }
            sync_workers();
            { int p, r, u; map_inv3<Cfg>(t, p, r, u); if (h) phase_inv3_split_b<1>(p, r, u, s.w); else phase_inv3_split_b<0>(p, r, u, s.w); }
            sync_workers();
            { int p, r, g; map_inv2_split<Cfg>(t, p, r, g); if (h) phase_inv2_split<1>(p, r, g, s.w); else phase_inv2_split<0>(p, r, g, s.w); }
            sync_workers();
            if (h) phase_inv1_split_a<1>(t, s.w, s.twd_inv); else phase_inv1_split_a<0>(t, s.w, s.twd_inv);
            sync_workers();
            if (h) phase_inv1_split_b<ROTATE, 1>(t, s.acc, s.w); else phase_inv1_split_b<ROTATE, 0>(t, s.acc, s.w);
        } else if constexpr (Cfg::STAGE_KEY) {
            if (next_bk_row) {
                br2_stage_key(s.key, next_bk_row, tid - 2 * Cfg::INV_TASKS, Cfg::THREADS - 2 * Cfg::INV_TASKS);
                cp_async_wait<0>();                            // landed before this thread reaches the barrier that ends the step
            }
        }
    } else {
        // (threads beyond Cfg::INV_TASKS idle: whole warps)
        { int p, r, u; if (map_inv3<Cfg>(tid, p, r, u)) phase_inv3(p, r, u, s.w); }
        __syncthreads();
        { int p, r, g; if (map_inv2<Cfg>(tid, p, r, g)) phase_inv2(p, r, g, s.w); }
        __syncthreads();
        if (tid < Cfg::INV_TASKS) phase_inv1<ROTATE>(tid, s.acc, s.w, s.twd_inv);
    }
}

NB_D unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
NB_D void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The kernel is persistent: the grid is at most one wave of resident CTAs (Cfg::CTAS_PER_SM per SM) and every CTA
// pulls work items from a queue until none are left.  A work item is one CHUNK of one CHAIN: a chain is the whole
// blind rotation of Cfg::CT ciphertexts, a chunk `steps_per_chunk` consecutive CMux steps of it.  Cutting chains into
// chunks time-slices a batch that is not a multiple of the wave size over all SMs instead of leaving a partial last
// wave: 1024 ciphertexts take 1.73 wave-times instead of 2 (the host picks the chunk count, capi.cu: pick_chunks).
// Between chunks the accumulators (8 KB per ciphertext) are parked in global memory.
// The queue is a FIFO of READY chains in global memory: entry t (t = 0 .. chains * chunks - 1) holds 1 + chain +
// chains * chunk; the host memset leaves every entry 0 = "not written yet" and the kernel treats entries t < chains as
// the initial state (chain t, chunk 0).  A CTA takes the next entry with an atomic increment of the head and waits for
// that entry to be written; when it finishes a chunk that is not the chain's last, it parks the accumulators and
// appends the chain's next chunk at the tail (release).  Whoever pops it (acquire) finds the accumulators.  An entry
// a CTA waits for is always produced by a chunk that is running on another resident CTA, so the wait cannot deadlock,
// and a ready chain never waits behind an unready one -- with in-order tickets and per-chain progress counters, which
// this replaces, 300 chains on 296 CTAs lost 18 % to such head-of-line waits (profiles/r2_variants.md).
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::CTAS_PER_SM) blind_rotate_kernel(BlindRotateArgs p, const u64 *__restrict__ twd_fwd_g,
                                                                       const u64 *__restrict__ twd_inv_g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Br2Smem s = br2_carve<Cfg>(smem_raw);
    __shared__ unsigned s_item, s_entry;
    const int tid = threadIdx.x;
    if (Cfg::TWD_GLOBAL) {
        s.twd_fwd = const_cast<u64 *>(twd_fwd_g); s.twd_inv = const_cast<u64 *>(twd_inv_g);
    } else {
        for (int i = tid; i < NTT_N; i += Cfg::THREADS) { s.twd_fwd[i] = twd_fwd_g[i]; s.twd_inv[i] = twd_inv_g[i]; }
    }
    const unsigned total = p.chains * p.chunks;
    constexpr int ACC_WORDS = Cfg::CT * 2 * NTT_N;
    // Two CTAs share an SM.  In a single-wave launch they would march through the phases in lock step (the MAC is
    // multiplier-bound and stalls both integer pipes, the transform passes are adder-bound); started half a step apart
    // they fill each other's gaps -- what multi-wave launches drift into by themselves (12.3 vs 12.8 ms per wave,
    // profiles/r2_variants.md).
    if (Cfg::CTAS_PER_SM > 1 && !p.sched && p.stagger_cycles > 0 && blockIdx.x >= (unsigned)p.sm_count) {
        if (tid == 0) {
            const long long t0 = clock64();
            while (clock64() - t0 < p.stagger_cycles) { }
        }
        __syncthreads();
    }

    for (unsigned item = blockIdx.x;; ) {
        if (p.sched) {
            __syncthreads();                                   // everyone is done with the previous item
            if (tid == 0) s_item = atomicAdd(p.sched, 1u);
            __syncthreads();
            item = s_item;
        }
        if (item >= total) break;
        unsigned chunk = 0, chain = item;
        if (p.sched && item >= p.chains) {
            // entries beyond the initial ones are appended by the CTAs that finish chunks
            if (tid == 0) {
                unsigned e;
                while ((e = ld_acquire_u32(p.sched + BR_SCHED_HEADER + item)) == 0) __nanosleep(100);
                s_entry = e - 1;
            }
            __syncthreads();
            const unsigned e = s_entry;
            chunk = e / p.chains; chain = e - chunk * p.chains;
        }
        const size_t ct0 = (size_t)chain * Cfg::CT;
        // ciphertext slots beyond the batch replay the last ciphertext and store nothing
        auto ct_of = [&](int slot) { size_t c = ct0 + slot; return c < p.batch ? c : p.batch - 1; };
        const int step0 = (int)chunk * p.steps_per_chunk;
        const int step1 = p.plain ? 1 : min(p.n, step0 + p.steps_per_chunk);

        if (chunk == 0) {
            // accumulator initialisation: CT x 2 polynomials x 1024 coefficients
            for (int e = tid; e < ACC_WORDS; e += Cfg::THREADS) {
                const int slot = e >> 11, mi = (e >> 10) & 1, x = e & (NTT_N - 1);
                s.acc[e] = br2_initial_acc(p, ct_of(slot), mi, x);
            }
        } else {
            // resume a parked chain: its queue entry was published after the accumulators (release / acquire above)
            const int4 *src = reinterpret_cast<const int4 *>(p.state + (size_t)chain * ACC_WORDS);
            for (int e = tid; e < ACC_WORDS / 4; e += Cfg::THREADS) reinterpret_cast<int4 *>(s.acc)[e] = __ldcg(src + e);
        }

        if constexpr (Cfg::STAGE_KEY) {
            // first key row of this item (the rows after it are staged inside the steps)
            br2_stage_key(s.key, p.bk + (size_t)(p.plain ? 0 : step0) * BK_ROW_U64, tid, Cfg::THREADS);
            cp_async_wait<0>();
        }
        if (p.plain) {
            __syncthreads();
            br2_step<false, Cfg>(s, p.bk, s.rot, tid);
            __syncthreads();
        } else {
            if (tid < Cfg::CT) s.rot[(step0 & 1) * Cfg::CT + tid] = br2_rotation(p, ct_of(tid), step0);
            __syncthreads();
            for (int i = step0; i < step1; i++) {
                int next = 0;
                if (tid < Cfg::CT && i + 1 < step1) next = br2_rotation(p, ct_of(tid), i + 1);
                br2_step<true, Cfg>(s, p.bk + (size_t)i * BK_ROW_U64, s.rot + (i & 1) * Cfg::CT, tid,
                                    i + 1 < step1 ? p.bk + (size_t)(i + 1) * BK_ROW_U64 : nullptr);
                if (tid < Cfg::CT) s.rot[((i + 1) & 1) * Cfg::CT + tid] = next;
                __syncthreads();
            }
        }

        if (!p.plain && step1 < p.n) {
            // park the accumulators and publish the chain's progress
            int4 *dst = reinterpret_cast<int4 *>(p.state + (size_t)chain * ACC_WORDS);
            for (int e = tid; e < ACC_WORDS / 4; e += Cfg::THREADS) __stcg(dst + e, reinterpret_cast<const int4 *>(s.acc)[e]);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned slot = p.chains + atomicAdd(p.sched + 1, 1u);   // appended entries follow the initial ones
                st_release_u32(p.sched + BR_SCHED_HEADER + slot, 1u + chain + p.chains * (chunk + 1));
            }
        } else {
            for (int e = tid; e < ACC_WORDS; e += Cfg::THREADS) {
                const int slot = e >> 11, mi = (e >> 10) & 1, x = e & (NTT_N - 1);
                const size_t c = ct0 + slot;
                if (c >= p.batch) continue;
                if (p.accum_out) p.accum_out[(c * 2 + mi) * NTT_N + x] = s.acc[e];
                if (p.extract) {
                    // sample extraction (tlwe_gpu.mako:63-82; blind_rotate.mako:213-224)
                    const i32 *a0 = s.acc + slot * 2 * NTT_N;
                    if (mi == 0) p.out_a[c * NTT_N + x] = x == 0 ? a0[0] : (i32)(0u - (u32)a0[NTT_N - x]);
                    else if (x == 0) p.out_b[c] = a0[NTT_N];
                }
            }
        }
        if (!p.sched) break;
    }
}

// ---- pair shape: one ciphertext per cluster of two CTAs (br_phases.cuh: "pair shape") -------------------------------
// Lowest latency for batches up to 3/8 of the SM count (capi.cu: pair_max).  No work queue (one cluster per ciphertext,
// one wave or plain hardware queueing of clusters), blind rotation only (`plain` external products take the single-CTA shapes).
// Dynamic shared memory is requested well above what the shape needs so that the two CTAs of a cluster can never
// share an SM.
constexpr size_t BR_PAIR_SMEM_USED = (size_t)PAIR_POLYS * POLY_STRIDE * sizeof(u64) + 2 * NTT_N * sizeof(u64) + NTT_N * sizeof(i32) + 64 +
                                     (size_t)PAIR_KEY_PLANES * NTT_N * sizeof(u64);
constexpr size_t BR_PAIR_SMEM_BYTES = 120 * 1024;
static_assert(BR_PAIR_SMEM_USED <= BR_PAIR_SMEM_BYTES, "pair shape shared memory");

NB_D unsigned cluster_cta_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// the address of `ptr` (this CTA's shared memory, generic) in CTA `rank` of the cluster
NB_D u64 *cluster_map_shared(u64 *ptr, unsigned rank)
{
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)ptr), "r"(rank));
    return reinterpret_cast<u64 *>(out);
}
// the same for a 32-bit shared-window address (what st.async and mbarrier operands take)
NB_D unsigned cluster_map_shared_u32(unsigned saddr, unsigned rank)
{
    unsigned out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(saddr), "r"(rank));
    return out;
}
// cluster-wide barrier; orders this CTA's remote stores before the peer's loads after it
NB_D void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
NB_D void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
NB_D void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
NB_D void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "NB_MBAR_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra NB_MBAR_WAIT;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// 16 bytes into the peer's shared memory, completion counted on the peer's mbarrier
NB_D void st_async_peer(unsigned raddr, u64 x, u64 y, unsigned rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "l"(x), "l"(y), "r"(rbar) : "memory");
}

// ASYNC: the partial sums travel as st.async stores that count their bytes on an mbarrier of the receiving CTA, one
// per step parity, which the receiver arms at the top of the step and waits for after its own MAC: no cluster-wide
// barrier in the step loop.  Flow control is the data dependence itself (br_phases.cuh: pair shape).
// !ASYNC: plain remote stores and one barrier.cluster per step -- ptxas implements its release / acquire with a
// GPU-scope memory barrier and an L1 invalidation, which is why the other variant exists; kept as the reference
// implementation (NUFHE_B200_PAIR_ASYNC=0).
template <bool ASYNC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
blind_rotate_pair_kernel(BlindRotateArgs p, const u64 *__restrict__ twd_fwd_g, const u64 *__restrict__ twd_inv_g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    u64 *w = reinterpret_cast<u64 *>(smem_raw);
    u64 *twd_fwd = w + PAIR_POLYS * POLY_STRIDE, *twd_inv = twd_fwd + NTT_N;
    i32 *acc = reinterpret_cast<i32 *>(twd_inv + NTT_N);       // this CTA's accumulator polynomial
    int *rot = reinterpret_cast<int *>(acc + NTT_N);           // [2], 64 bytes reserved
    // this CTA's key planes of the current step: [j * 2 + mo] and, rank 0 only, the two correction planes
    u64 *key = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(rot) + 64);
    const int tid = threadIdx.x;
    const int rank = (int)cluster_cta_rank();
    const size_t c = blockIdx.x >> 1;                          // grid = 2 x batch
    // The key row of step i + 1 is requested (cp.async, 16 bytes per request) by the warps that have no work in the
    // middle inverse pass of step i and lands during the last one: the MAC then reads shared memory instead of waiting
    // for the L2, whose latency a step of 8 warps cannot hide -- and which every CTA of the launch asks for the same
    // lines at the same time.
    const int key_planes = rank == 0 ? PAIR_KEY_PLANES : 4;
    auto stage_key = [&](int step, int t, int nthreads) {
        const u64 *row = p.bk + (size_t)step * BK_ROW_U64;
        for (int x = t; x < key_planes * (NTT_N / 2); x += nthreads) {
            const int pl = x >> 9, e = (x & 511) * 2;
            const int src = pl < 4 ? rank * 4 + pl : 4 + pl;      // planes 8, 9: corrections
            cp_async16(key + pl * NTT_N + e, row + (size_t)src * NTT_N + e);
        }
        cp_async_commit();
    };
    if (p.n > 0) stage_key(0, tid, PAIR_THREADS);
    for (int i = tid; i < NTT_N; i += PAIR_THREADS) {
        twd_fwd[i] = twd_fwd_g[i]; twd_inv[i] = twd_inv_g[i];
        acc[i] = br2_initial_acc(p, c, rank, i);
    }
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(s_bar);
    if (tid == 0) {
        rot[0] = br2_rotation(p, c, 0);
        if (ASYNC) {
            mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    u64 *w_peer = cluster_map_shared(w, (unsigned)(rank ^ 1));
    const unsigned w_peer32 = cluster_map_shared_u32((unsigned)__cvta_generic_to_shared(w), (unsigned)(rank ^ 1));
    const unsigned bar_peer = cluster_map_shared_u32(bar0, (unsigned)(rank ^ 1));
    cp_async_wait<0>();
    cluster_barrier();                                         // the peer is running, its barriers are initialised

    const u64 *corr2 = rank == 0 ? key + 4 * NTT_N : nullptr;
    for (int i = 0; i < p.n; i++) {
        const int par = i & 1;
        const int *r = rot + par;
        int next = 0;
        if (tid == 0) {
            if (ASYNC) mbar_expect_tx(bar0 + 8 * par, PAIR_EXCHANGE_BYTES);
            if (i + 1 < p.n) next = br2_rotation(p, c, i + 1);
        }
        pair_fwd1(tid, acc, w, twd_fwd, r);
        __syncthreads();
        pair_fwd2(tid, w);
        __syncthreads();
        {
            u64 v[16];
            pair_fwd3_load(tid, w, v);
            __syncthreads();                                   // in place: all loads before any store
            pair_fwd3_finish(tid, w, v);
        }
        __syncthreads();
        if (ASYNC) {
            const unsigned rb = bar_peer + 8 * par;
            pair_mac<true>(tid, w, [=](int off, u64 x, u64 y) { st_async_peer(w_peer32 + 8u * (unsigned)off, x, y, rb); },
                           key, corr2, rank, par);
            __syncthreads();                                   // this CTA's own partial sums; the key planes are free
        } else {
            pair_mac<true>(tid, w, [=](int off, u64 x, u64 y) { st2(w_peer + off, x, y); }, key, corr2, rank, par);
            cluster_barrier();
        }
        if (ASYNC) mbar_wait(bar0 + 8 * par, (unsigned)(i >> 1) & 1u);       // the peer's partial sums have landed
        pair_inv3_a(tid, w, par);
        __syncthreads();
        pair_inv3_b(tid, w, par);
        __syncthreads();
        if (tid < PAIR_INV_WORKERS) pair_inv2(tid, w, par);
        else if (i + 1 < p.n) stage_key(i + 1, tid - PAIR_INV_WORKERS, PAIR_THREADS - PAIR_INV_WORKERS);
        __syncthreads();
        pair_inv1_a(tid, w, twd_inv, par);
        __syncthreads();
        pair_inv1_b(tid, acc, w, par);
        cp_async_wait<0>();                                    // the staged key row, before the barrier that ends the step
        if (tid == 0) rot[par ^ 1] = next;
        __syncthreads();
    }

    for (int x = tid; x < NTT_N; x += PAIR_THREADS) {
        if (p.accum_out) p.accum_out[(c * 2 + rank) * NTT_N + x] = acc[x];
        if (p.extract) {
            // sample extraction (tlwe_gpu.mako:63-82; blind_rotate.mako:213-224): a from polynomial 0, b from polynomial 1
            if (rank == 0) p.out_a[c * NTT_N + x] = x == 0 ? acc[0] : (i32)(0u - (u32)acc[NTT_N - x]);
            else if (x == 0) p.out_b[c] = acc[0];
        }
    }
    // nothing is in flight towards this CTA (it has waited for every exchange) and nothing it sent is unreceived
    // while the peer still runs, but the peer's shared memory must outlive the stores addressed to it: both leave together
    cluster_barrier();
}

// ---- the separate kernels of the reference's multi-kernel bootstrap (bootstrap.py:96-196) ---------------
// They exist so that `single_kernel_bootstrap=False` and callers of the inner seams get the same
// functions with the same results; the fused kernel above does all of this in shared memory.
//
// ShiftTorusPolynomial (polynomials_gpu.mako:18-77; polynomials_cpu.py:25-59): result = X^e * source with
// e = power (mode 2), 2N - power (mode 0, `invert_powers`), or result = (X^power - 1) * source (mode 1,
// `minus_one`).  One power per group of `polys_per_power` consecutive polynomials, read from
// powers[group * powers_stride + power_idx] (the reference's `powers_view`).
__global__ void shift_torus_polynomial_kernel(i32 *__restrict__ result, const i32 *__restrict__ source,
                                              const i32 *__restrict__ powers, size_t powers_stride, size_t power_idx,
                                              int polys_per_power, int mode, int n_log2, size_t polys)
{
    const int N = 1 << n_log2;
    const size_t total = polys << n_log2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = i >> n_log2;
        const int idx = (int)(i & (N - 1));
        int pw = powers[(poly / polys_per_power) * powers_stride + power_idx];
        if (mode == 0) pw = 2 * N - pw;
        pw &= 2 * N - 1;
        const int ar = pw & (N - 1);
        const bool flip = pw >= N;
        const i32 src = source[poly * N + ((idx - ar) & (N - 1))];
        const bool neg = (idx < ar) != flip;
        u32 v = neg ? 0u - (u32)src : (u32)src;
        if (mode == 1) v -= (u32)source[poly * N + idx];
        result[i] = (i32)v;
    }
}

// TLweNoiselessTrivial (tlwe_gpu.py:32-74; tlwe_cpu.py:26-38): acc = (0, ..., 0, mu), variances 0
__global__ void tlwe_noiseless_trivial_kernel(i32 *__restrict__ acc, float *__restrict__ cv, const i32 *__restrict__ mu,
                                              int mask_size, int n_log2, size_t batch)
{
    const int N = 1 << n_log2;
    const size_t per = (size_t)(mask_size + 1) << n_log2, total = batch * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / per, r = i % per;
        const int poly = (int)(r >> n_log2), x = (int)(r & (N - 1));
        acc[i] = poly == mask_size ? mu[c * N + x] : 0;
        if (cv && i < batch) cv[i] = 0.f;            // one variance per sample: cv is (B,) (tlwe.py:94-112)
    }
}

// TLweExtractLweSamples (tlwe_gpu.mako:54-84; tlwe_cpu.py:41-60): a[i*N] = acc_i[0], a[i*N + x] = -acc_i[N - x],
// b = acc_k[0]
__global__ void tlwe_extract_lwe_samples_kernel(i32 *__restrict__ out_a, i32 *__restrict__ out_b,
                                                const i32 *__restrict__ acc, int mask_size, int n_log2, size_t batch)
{
    const int N = 1 << n_log2;
    const size_t per = (size_t)mask_size << n_log2, total = batch * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / per, r = i % per;
        const int poly = (int)(r >> n_log2), x = (int)(r & (N - 1));
        const i32 *a = acc + ((c * (mask_size + 1) + poly) << n_log2);
        out_a[i] = x == 0 ? a[0] : (i32)(0u - (u32)a[N - x]);
        if (r == 0) out_b[c] = acc[(c * (mask_size + 1) + mask_size) << n_log2];
    }
}

// Torus32ToPhase (numeric_functions_gpu.py:39-77; numeric_functions_cpu.py:23-37): round to a multiple of
// 2^32 / mspace_size and return the multiple, in [0, mspace_size)
__global__ void t32_to_phase_kernel(i32 *__restrict__ out, const i32 *__restrict__ in, size_t n, u32 mspace_size)
{
    const u32 interv = (u32)((1ull << 32) / mspace_size), half = interv / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (i32)(((u32)in[i] + half) / interv);
}

// Gadget decomposition as its own kernel (TGswPolynomialDecompH, tgsw_gpu.py:31-54; tgsw_cpu.py:26-49):
// out[poly][j][x] = (((in[poly][x] + offset) >> (32 - (j + 1) bg_bit)) & (Bg - 1)) - Bg / 2
__global__ void tgsw_decompose_kernel(i32 *__restrict__ out, const i32 *__restrict__ in, size_t polys, int decomp_length,
                                      int bs_log2_base, i32 offset, int n_log2)
{
    const int N = 1 << n_log2;
    const size_t total = (polys * decomp_length) << n_log2;
    const u32 mask = (1u << bs_log2_base) - 1u, half = 1u << (bs_log2_base - 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i & (N - 1));
        const size_t pj = i >> n_log2;
        const int j = (int)(pj % decomp_length);
        const size_t poly = pj / decomp_length;
        const u32 t = (u32)in[(poly << n_log2) + x] + (u32)offset;
        out[i] = (i32)(((t >> (32 - (j + 1) * bs_log2_base)) & mask) - half);
    }
}

// The point-wise multiply-accumulate of the external product as its own kernel, on the reference's key layout
// (tgsw_gpu.py:58-107; tgsw_cpu.py:52-79): out[b][mo][x] = sum_{mi, j} mul_prepared(tr[b][mi][j][x], bk[mi][j][mo][x])
// for any mask size k (mi, mo <= k) and decomposition length.  bk_row: (k+1, l, k+1, N), Montgomery form.
__global__ void tgsw_mac_kernel(u64 *__restrict__ out, const u64 *__restrict__ tr, const u64 *__restrict__ bk_row,
                                size_t batch, int k1, int decomp_length)
{
    const size_t total = batch * k1 * NTT_N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i & (NTT_N - 1));
        const int mo = (int)((i >> 10) % k1);
        const size_t b = (i >> 10) / k1;
        u64 acc = 0;
        for (int mi = 0; mi < k1; mi++)
            for (int j = 0; j < decomp_length; j++) {
                const u64 a = ff_canon(tr[((b * k1 + mi) * decomp_length + j) * NTT_N + x]);
                const u64 w = ff_canon(bk_row[(((size_t)mi * decomp_length + j) * k1 + mo) * NTT_N + x]);
                acc = ff_add(acc, ff_canon(ff_mul_prepared(a, w)));
            }
        out[i] = ff_canon(acc);
    }
}

// tlwe_add_to (tlwe.py:173-175): wrap-around int32 addition, float addition of the variances
__global__ void add_to_kernel(i32 *__restrict__ res, const i32 *__restrict__ src, size_t n, float *__restrict__ res_cv,
                              const float *__restrict__ src_cv, size_t n_cv)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        res[i] = (i32)((u32)res[i] + (u32)src[i]);
        if (res_cv && i < n_cv) res_cv[i] += src_cv[i];
    }
}

// ---- LWE key switch (lwe_gpu.mako:59-118; lwe_cpu.py:62-93) --------------------------------------
// res_a[i] = - sum_{j,k} ks_a[j][k][digit(j,k)][i], res_b = b - sum ks_b[j][k][digit].
//
// The reference gathers one 2000-byte key row per (ciphertext, j, k) from global memory: ~12 MB of key
// per ciphertext.  Here a CTA owns a tile of up to KS_TILE ciphertexts and streams the key ONCE per tile:
// a producer warp moves the key, 32 000 contiguous bytes (four values of k, all four rows) at a time,
// into a ring of four shared-memory slots with cp.async.bulk (TMA, completion on an mbarrier); 8 consumer
// warps (thread t owns output coefficients 2t, 2t+1 of every ciphertext of the tile) pick the row with the
// warp-uniform 2-bit digit and subtract: 3 instructions per (ciphertext, j, k, coefficient).
// Row d = 0 of the key is zero padding (lwe_cpu.py:31-33; the reference kernel skips it, its NumPy closure
// subtracts it): it is streamed like the other rows, so d = 0 needs no branch.
// src = src1 (+ src2) (+ (0, c)) lets gate_mux fold `(0,1/8) + u1 + u2` (gates.py:657-664) into the load.
constexpr int KS_TILE = 32;                 // ciphertexts per CTA (run-time tile <= KS_TILE)
constexpr int KS_CONSUMERS = 256;           // threads 0..249: a[2t], a[2t+1] of every ciphertext of the tile
constexpr int KS_THREADS = KS_CONSUMERS + 64;   // + producer warp (TMA issue) + b/variance warp
constexpr int KS_SLOTS = 4;                 // ring of half-j slots (4 values of k each): two j in flight
constexpr int KS_IN = 1024, KS_N = 500;     // fast-path sizes (api_low_level.py:49-50)
// The variance of a result is a float32 sum of 8192 table entries.  Float addition is not associative, so its shape is
// FIXED: 128 block sums over 8 consecutive input coefficients each (64 terms, added in (j, k) order), then the block
// sums added in block order.  A CTA that owns only a slice of j (small batches, KeyswitchArgs::splits) writes its
// block sums to a scratch array and ks_cv_finalize_kernel adds them; a CTA that owns all of j does both itself.
// Either way the same additions happen in the same order: bit-identical variances for every launch shape, run to run.
constexpr int KS_CV_BLOCK_J = 8;
constexpr int KS_CV_BLOCKS = KS_IN / KS_CV_BLOCK_J;
constexpr int KS_ROW_BYTES = KS_N * 4;
constexpr int KS_SLOT_BYTES = 4 * 4 * KS_ROW_BYTES;   // 4 k x 4 rows
constexpr size_t KS_SMEM_BYTES = (size_t)KS_SLOTS * KS_SLOT_BYTES + (size_t)KS_IN * KS_TILE * 2 + 2 * KS_SLOTS * 8 + 128;

struct KeyswitchArgs {
    const i32 *src1_a, *src1_b, *src2_a, *src2_b;   // (B, in), (B,)
    i32 c;
    const i32 *ks_a, *ks_b;                          // (in, t, base, n), (in, t, base)
    const float *ks_cv;
    i32 *res_a, *res_b;                              // (B, n), (B,)
    float *res_cv;                                   // optional
    int in_size, n, t, log2_base;
    int tile;                                        // ciphertexts per CTA (fast path)
    int splits;                                      // > 1: blockIdx.y takes a slice of j and results are
                                                     // accumulated with integer atomics into zeroed outputs
    float *cv_blocks;                                // splits > 1 and res_cv: (B, KS_CV_BLOCKS) partial variances
    size_t batch;
};

NB_D void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
NB_D void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NB_D void mbar_arrive(u64 *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
NB_D void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
NB_D void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Fast path: in = 1024, n = 500, t = 8, base = 4 (the scheme's parameters, api_low_level.py:49-56).
__global__ void __launch_bounds__(KS_THREADS, 1) keyswitch_kernel(KeyswitchArgs p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;                                               // [slot][k & 3][row][2000 B]
    unsigned short *digits = reinterpret_cast<unsigned short *>(ring + KS_SLOTS * KS_SLOT_BYTES);  // [j][KS_TILE]
    u64 *full = reinterpret_cast<u64 *>(ring + KS_SLOTS * KS_SLOT_BYTES + KS_IN * KS_TILE * 2);
    u64 *empty = full + KS_SLOTS;

    const int tid = threadIdx.x;
    const size_t ct0 = (size_t)blockIdx.x * p.tile;
    const int nct = (int)min((size_t)p.tile, p.batch - ct0);
    // slice of the input coefficients handled by this CTA (small batches: split-j over blockIdx.y so that
    // one ciphertext does not stream the whole 65 MB key through a single SM); always an even number of j
    const int j_per = ((KS_IN + p.splits - 1) / p.splits + KS_CV_BLOCK_J - 1) & ~(KS_CV_BLOCK_J - 1);
    const int j_begin = blockIdx.y * j_per;
    const int j_end = min(KS_IN, j_begin + j_per);
    const bool split = p.splits > 1;

    if (tid == 0) {
        for (int s = 0; s < KS_SLOTS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], KS_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // packed digits of the tile: the 8 two-bit digits of coefficient j of ciphertext q in 16 bits
    const u32 prec_offset = 1u << (32 - (1 + 2 * 8));
    const int jn = max(j_end - j_begin, 0);
    for (int idx = tid; idx < KS_TILE * jn; idx += KS_THREADS) {
        const int q = idx / jn, j = j_begin + idx % jn;
        u32 v = 0;
        if (q < nct) {
            v = (u32)p.src1_a[(ct0 + q) * KS_IN + j];
            if (p.src2_a) v += (u32)p.src2_a[(ct0 + q) * KS_IN + j];
            v = (v + prec_offset) >> 16;              // k = 0 in the top bits
        }
        digits[j * KS_TILE + q] = (unsigned short)v;  // ciphertexts beyond nct: digits 0 (results not stored)
    }
    __syncthreads();

    if (tid >= KS_CONSUMERS + 32) {
        // ---- b / variance warp: lane q owns ciphertext q of the tile; the 16-byte (j, k) entries of ks_b and
        // ks_cv are warp-uniform loads.  Runs beside the consumers and never touches the ring.
        const int q = tid - (KS_CONSUMERS + 32);
        u32 accb = 0;
        float cv = 0.f;
        int4 kb = __ldg(reinterpret_cast<const int4 *>(p.ks_b));
        float4 kc = __ldg(reinterpret_cast<const float4 *>(p.ks_cv));
        kb = __ldg(reinterpret_cast<const int4 *>(p.ks_b) + min(j_begin * 8, KS_IN * 8 - 1));
        kc = __ldg(reinterpret_cast<const float4 *>(p.ks_cv) + min(j_begin * 8, KS_IN * 8 - 1));
        float blk = 0.f;
        for (int jk = j_begin * 8; jk < j_end * 8; jk++) {
            const int4 kb_next = __ldg(reinterpret_cast<const int4 *>(p.ks_b) + min(jk + 1, KS_IN * 8 - 1));
            const float4 kc_next = __ldg(reinterpret_cast<const float4 *>(p.ks_cv) + min(jk + 1, KS_IN * 8 - 1));
            const u32 bits = digits[(jk >> 3) * KS_TILE + q];
            const u32 d = (bits >> (14 - 2 * (jk & 7))) & 3u;
            accb -= d == 0 ? (u32)kb.x : d == 1 ? (u32)kb.y : d == 2 ? (u32)kb.z : (u32)kb.w;
            blk += d == 0 ? kc.x : d == 1 ? kc.y : d == 2 ? kc.z : kc.w;
            kb = kb_next; kc = kc_next;
            if ((jk & (8 * KS_CV_BLOCK_J - 1)) == 8 * KS_CV_BLOCK_J - 1) {      // end of a block of 8 coefficients
                if (split) { if (p.cv_blocks && q < nct) p.cv_blocks[(ct0 + q) * KS_CV_BLOCKS + (jk >> 6)] = blk; }
                else cv += blk;
                blk = 0.f;
            }
        }
        if (q < nct) {
            u32 b = (u32)p.src1_b[ct0 + q] + (p.src2_b ? (u32)p.src2_b[ct0 + q] : 0u) + (u32)p.c;
            if (!split) {
                p.res_b[ct0 + q] = (i32)(b + accb);
                if (p.res_cv) p.res_cv[ct0 + q] = cv;
            } else {
                atomicAdd(reinterpret_cast<unsigned int *>(p.res_b + ct0 + q), (blockIdx.y == 0 ? b : 0u) + accb);
            }
        }
        return;
    }
    if (tid >= KS_CONSUMERS) {
        // ---- producer warp: one lane streams the key, 32 000 bytes per slot ------------------------
        if (tid == KS_CONSUMERS) {
            for (int j = j_begin; j < j_end; j++) {
                const u32 parity = ((((j - j_begin) >> 1) & 1)) ^ 1;
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    const int slot = 2 * (j & 1) + h;
                    mbar_wait(&empty[slot], parity);  // first pass: passes immediately (phase -1 "complete")
                    mbar_expect_tx(&full[slot], KS_SLOT_BYTES);
                    bulk_g2s(ring + slot * KS_SLOT_BYTES, p.ks_a + (size_t)(j * 8 + 4 * h) * 4 * KS_N, KS_SLOT_BYTES,
                             &full[slot]);
                }
            }
        }
        return;
    }

    // ---- consumers -------------------------------------------------------------------------------
    const bool is_a = tid < KS_N / 2;
    const int lane = tid & 31;
    u32 acc0[KS_TILE], acc1[KS_TILE];
#pragma unroll
    for (int q = 0; q < KS_TILE; q++) { acc0[q] = 0; acc1[q] = 0; }
    // byte offset of this thread's pair inside a key row; idle threads read pair 0 harmlessly
    const u32 col_off = is_a ? (u32)tid * 8u : 0u;
    const u32 ring_base = smem_u32(ring);

    for (int j = j_begin; j < j_end; j++) {
        u32 w[KS_TILE / 2];                            // packed digits of the tile for this j (64 bytes)
        {
            const uint4 *dg = reinterpret_cast<const uint4 *>(digits + j * KS_TILE);
#pragma unroll
            for (int x = 0; x < KS_TILE / 8; x++) {
                uint4 t = dg[x];
                w[4 * x] = t.x; w[4 * x + 1] = t.y; w[4 * x + 2] = t.z; w[4 * x + 3] = t.w;
            }
        }
        const u32 parity = ((j - j_begin) >> 1) & 1;
        const u32 jbase = ring_base + (j & 1) * (2 * KS_SLOT_BYTES) + col_off;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if ((k & 3) == 0) mbar_wait(&full[2 * (j & 1) + (k >> 2)], parity);
            const u32 base = jbase + k * (4 * KS_ROW_BYTES);
            {
#pragma unroll
                for (int q = 0; q < KS_TILE; q++) {
                    const u32 d = (w[q >> 1] >> (16 * (q & 1) + 14 - 2 * k)) & 3u;
                    const u32 addr = d * (u32)KS_ROW_BYTES + base;
                    u32 v0, v1;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(addr));
                    acc0[q] -= v0;
                    acc1[q] -= v1;
                }
            }
            if ((k & 3) == 3) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[2 * (j & 1) + (k >> 2)]);
            }
        }
    }

#pragma unroll
    for (int q = 0; q < KS_TILE; q++) {
        if (q < nct) {
            if (is_a) {
                if (!split) {
                    p.res_a[(ct0 + q) * KS_N + 2 * tid] = (i32)acc0[q];
                    p.res_a[(ct0 + q) * KS_N + 2 * tid + 1] = (i32)acc1[q];
                } else {
                    atomicAdd(reinterpret_cast<unsigned int *>(p.res_a + (ct0 + q) * KS_N + 2 * tid), acc0[q]);
                    atomicAdd(reinterpret_cast<unsigned int *>(p.res_a + (ct0 + q) * KS_N + 2 * tid + 1), acc1[q]);
                }
            }
        }
    }
}

// second half of the fixed-shape variance sum for split launches: one thread per ciphertext adds its block sums in order
__global__ void ks_cv_finalize_kernel(float *__restrict__ res_cv, const float *__restrict__ cv_blocks, size_t batch)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= batch) return;
    const float4 *src = reinterpret_cast<const float4 *>(cv_blocks + c * KS_CV_BLOCKS);
    float cv = 0.f;
#pragma unroll 8
    for (int m = 0; m < KS_CV_BLOCKS / 4; m++) {
        const float4 v = src[m];
        cv += v.x; cv += v.y; cv += v.z; cv += v.w;
    }
    res_cv[c] = cv;
}

// Generic decomposition parameters (API parity with LweKeyswitch(…, decomp_length, log2_base)): the
// straightforward per-ciphertext gather.
__global__ void __launch_bounds__(512) keyswitch_generic_kernel(KeyswitchArgs p)
{
    const size_t ct = blockIdx.x;
    const int base = 1 << p.log2_base;
    const u32 prec_offset = 1u << (32 - (1 + p.log2_base * p.t));
    const int i = threadIdx.x;
    const bool is_a = i < p.n, is_b = i == p.n;
    if (!(is_a || is_b)) return;
    u32 acc = 0;
    float cv = 0.f;
    for (int j = 0; j < p.in_size; j++) {
        u32 v = (u32)p.src1_a[ct * p.in_size + j];
        if (p.src2_a) v += (u32)p.src2_a[ct * p.in_size + j];
        const i32 tmp = (i32)(v + prec_offset);
        for (int k = 0; k < p.t; k++) {
            const int d = (tmp >> (32 - (k + 1) * p.log2_base)) & (base - 1);
            const size_t row = ((size_t)j * p.t + k) * base + d;
            if (is_a) acc -= (u32)__ldg(p.ks_a + row * p.n + i);
            else { acc -= (u32)__ldg(p.ks_b + row); cv += __ldg(p.ks_cv + row); }
        }
    }
    if (is_a) p.res_a[ct * p.n + i] = (i32)acc;
    else {
        u32 b = (u32)p.src1_b[ct] + (p.src2_b ? (u32)p.src2_b[ct] : 0u) + (u32)p.c;
        p.res_b[ct] = (i32)(b + acc);
        if (p.res_cv) p.res_cv[ct] = cv;
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

// ---- LWE encryption / phase (lwe_gpu.py:186-284, lwe_gpu.mako:205-262; lwe_cpu.py:96-113) and the key-switch key
// (lwe_gpu.py:63-124, lwe_gpu.mako:18-56; lwe_cpu.py:26-59): wrap-around dot products with the binary key.  One warp
// per row, 16-byte loads when the row length allows; nothing is materialised in 64 bits.

// out[i] = add1[i] (+ add2[i]) + sign * <a[i, :], key>   (Torus32, mod 2^32)
//   encrypt: add1 = messages, add2 = Gaussian noise, sign = +1, a = the uniform mask      (b = mu + e + <a, s>)
//   phase:   add1 = b, sign = -1                                                          (b - <a, s>)
__global__ void lwe_dot_kernel(i32 *__restrict__ out, const i32 *__restrict__ a, const i32 *__restrict__ key,
                               const i32 *__restrict__ add1, const i32 *__restrict__ add2, i32 sign, size_t batch, int n)
{
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < batch; row += warps) {
        const i32 *ar = a + row * (size_t)n;
        u32 acc = 0;
        if ((n & 3) == 0) {
            const int4 *a4 = reinterpret_cast<const int4 *>(ar);
            const int4 *k4 = reinterpret_cast<const int4 *>(key);
            for (int j = lane; j < n / 4; j += 32) {
                const int4 x = __ldg(a4 + j), k = __ldg(k4 + j);
                acc += (u32)x.x * (u32)k.x + (u32)x.y * (u32)k.y + (u32)x.z * (u32)k.z + (u32)x.w * (u32)k.w;
            }
        } else {
            for (int j = lane; j < n; j += 32) acc += (u32)__ldg(ar + j) * (u32)__ldg(key + j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            u32 v = (u32)sign * acc;
            if (add1) v += (u32)add1[row];
            if (add2) v += (u32)add2[row];
            out[row] = (i32)v;
        }
    }
}

// key-switch key: row (i, j, h) of ks_a / ks_b / ks_cv; h = 0 is the zero padding, h >= 1 encrypts
// in_key[i] * h * 2^(32 - (j + 1) log2_base) under out_key with mask noises_a[i, j, h - 1] and noise noises_b[i, j, h - 1]
__global__ void make_keyswitch_key_kernel(i32 *__restrict__ ks_a, i32 *__restrict__ ks_b, float *__restrict__ ks_cv,
                                          const i32 *__restrict__ in_key, const i32 *__restrict__ out_key,
                                          const i32 *__restrict__ noises_a, const i32 *__restrict__ noises_b,
                                          size_t in_size, int n, int t, int log2_base, float noise_var)
{
    const int base = 1 << log2_base;
    const size_t rows = in_size * t * base;
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
        const int h = (int)(row % base);
        const size_t ij = row / base;
        const int j = (int)(ij % t);
        const size_t i = ij / t;
        i32 *dst = ks_a + row * (size_t)n;
        if (h == 0) {
            for (int x = lane; x < n; x += 32) dst[x] = 0;
            if (lane == 0) { ks_b[row] = 0; ks_cv[row] = 0.f; }
            continue;
        }
        const size_t nrow = ij * (base - 1) + (h - 1);
        const i32 *src = noises_a + nrow * (size_t)n;
        u32 acc = 0;
        for (int x = lane; x < n; x += 32) {
            const i32 v = __ldg(src + x);
            dst[x] = v;
            acc += (u32)v * (u32)__ldg(out_key + x);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const u32 message = (u32)in_key[i] * (u32)h * (1u << (32 - (j + 1) * log2_base));
            ks_b[row] = (i32)(message + (u32)noises_b[nrow] + acc);
            ks_cv[row] = noise_var;
        }
    }
}

}  // namespace nb
