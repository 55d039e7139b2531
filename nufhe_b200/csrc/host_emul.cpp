// host_emul.cpp -- runs the __host__ __device__ pass code of br_phases.cuh / ntt_lane.cuh on the CPU, one
// "thread" after the other, a phase at a time (a phase boundary = a CTA barrier).  Built with g++ into
// libnb_host_emul.so for tests/test_host_logic.py: it lets the CPU-only test suite check the index maps,
// swizzles, twiddle tables and shift constants of the GPU kernels against the oracle.
// This is a test aid for the CUDA source, not a CPU fallback: nothing in nufhe_b200/ loads it.
#include <cstring>
#include <vector>
#include "tables.h"

using namespace nb;

extern "C" {

// The stand-alone transform kernels' passes, one polynomial at a time; in/out natural order, canonical.
void emul_ntt_forward(const u64 *in, u64 *out, size_t batch)
{
    static PhaseTables T;
    std::vector<u64> w(NTT_SWEEP_POLYS * POLY_STRIDE);
    for (size_t b = 0; b < batch; b++) {
        for (int task = 0; task < 64; task++) {
            u64 x[16];
            for (int j1 = 0; j1 < 16; j1++) x[j1] = ff_canon(in[b * NTT_N + 64 * j1 + task]);
            phase_fwd1_generic(task, x, w.data(), T.fwd.data());
        }
        for (int row = 0; row < 16; row++) for (int g = 0; g < 4; g++) phase_fwd2(0, row, g, w.data());
        for (int row = 0; row < 16; row++) for (int u = 0; u < 4; u++) phase_fwd3(0, row, u, w.data());
        for (int k = 0; k < NTT_N; k++) out[b * NTT_N + k] = ff_canon(w[w_position_of_natural(k)]);
    }
}

void emul_ntt_inverse(const u64 *in, u64 *out, size_t batch)
{
    static PhaseTables T;
    std::vector<u64> w(NTT_SWEEP_POLYS * POLY_STRIDE);
    for (size_t b = 0; b < batch; b++) {
        for (int k = 0; k < NTT_N; k++) w[w_position_of_natural(k)] = ff_canon(in[b * NTT_N + k]);
        for (int row = 0; row < 16; row++) for (int u = 0; u < 4; u++) phase_inv3(0, row, u, w.data());
        for (int row = 0; row < 16; row++) for (int g = 0; g < 4; g++) phase_inv2(0, row, g, w.data());
        for (int task = 0; task < 64; task++) {
            u64 y[16];
            phase_inv1_generic(task, y, w.data(), T.inv.data());
            for (int j1 = 0; j1 < 16; j1++) out[b * NTT_N + 64 * j1 + task] = ff_canon(y[j1]);
        }
    }
}

// the same with Torus32 on the natural-order side (i32_conversion): forward reads int32, inverse writes int32
void emul_ntt_forward_i32(const i32 *in, u64 *out, size_t batch)
{
    static PhaseTables T;
    std::vector<u64> w(NTT_SWEEP_POLYS * POLY_STRIDE);
    for (size_t b = 0; b < batch; b++) {
        for (int task = 0; task < 64; task++) {
            i32 x[16];
            for (int j1 = 0; j1 < 16; j1++) x[j1] = in[b * NTT_N + 64 * j1 + task];
            phase_fwd1_i32(task, x, w.data(), T.fwd.data());
        }
        for (int row = 0; row < 16; row++) for (int g = 0; g < 4; g++) phase_fwd2(0, row, g, w.data());
        for (int row = 0; row < 16; row++) for (int u = 0; u < 4; u++) phase_fwd3(0, row, u, w.data());
        for (int k = 0; k < NTT_N; k++) out[b * NTT_N + k] = ff_canon(w[w_position_of_natural(k)]);
    }
}

void emul_ntt_inverse_i32(const u64 *in, i32 *out, size_t batch)
{
    static PhaseTables T;
    std::vector<u64> w(NTT_SWEEP_POLYS * POLY_STRIDE);
    for (size_t b = 0; b < batch; b++) {
        for (int k = 0; k < NTT_N; k++) w[w_position_of_natural(k)] = ff_canon(in[b * NTT_N + k]);
        for (int row = 0; row < 16; row++) for (int u = 0; u < 4; u++) phase_inv3(0, row, u, w.data());
        for (int row = 0; row < 16; row++) for (int g = 0; g < 4; g++) phase_inv2(0, row, g, w.data());
        for (int task = 0; task < 64; task++) {
            i32 y[16];
            phase_inv1_i32(task, y, w.data(), T.inv.data());
            for (int j1 = 0; j1 < 16; j1++) out[b * NTT_N + 64 * j1 + task] = y[j1];
        }
    }
}

// One external-product step of the phase-structured kernel (br_phases.cuh) for up to BR2_CT
// ciphertexts, phases executed in order with all "threads" of a phase run back to back.
// acc: (nct, 2, 1024) in/out; bk_ref_row: reference layout (2,2,2,1024) Montgomery; rot: rotation
// amounts per ciphertext or NULL (plain external product, overwrite).
int emul_phase_ct(void) { return BR2_CT; }

}  // extern "C"

template <class Cfg> static void phase_step(i32 *acc_io, const u64 *bk_ref_row, const int *rot, int nct)
{
    static PhaseTables T;
    std::vector<i32> acc(Cfg::CT * 2 * NTT_N, 0);
    std::vector<u64> w(Cfg::POLYS * POLY_STRIDE, 0);
    std::vector<u64> bk(BK_ROW_U64);
    for (int c = 0; c < nct; c++) memcpy(&acc[c * 2 * NTT_N], acc_io + c * 2 * NTT_N, sizeof(i32) * 2 * NTT_N);
    int rots[4] = {0, 0, 0, 0};
    if (rot) for (int c = 0; c < nct; c++) rots[c] = rot[c];
    // bk_prepare: internal [m][row*64 + scol] plain
    for (int pos = 0; pos < NTT_N; pos++) {
        const int k = w_natural_index(pos >> 6, pos & 63);
        u64 sum[2] = {0, 0};
        for (int m = 0; m < 8; m++) {
            u64 x = ff_mul(ff_canon(bk_ref_row[m * NTT_N + k]), FF_RINV);
            bk[m * NTT_N + pos] = x;
            sum[m & 1] = ff_add(sum[m & 1], x);
        }
        for (int mo = 0; mo < 2; mo++) bk[(8 + mo) * NTT_N + pos] = ff_mul(sum[mo], T.ones512[k]);
    }
    constexpr int TH = Cfg::THREADS;
    if constexpr (Cfg::SPLIT_FWD) {
        using Tasks = BrCfg<Cfg::CT, 256 * Cfg::CT>;
        auto each = [&](auto fn) { for (int tid = 0; tid < TH; tid++) { int h, t; map_split_fwd<Cfg>(tid, h, t); fn(h, t); } };
        each([&](int h, int t) {
            if (rot) { if (h) phase_fwd1_split<true, 1>(t, acc.data(), w.data(), T.fwd.data(), rots); else phase_fwd1_split<true, 0>(t, acc.data(), w.data(), T.fwd.data(), rots); }
            else { if (h) phase_fwd1_split<false, 1>(t, acc.data(), w.data(), T.fwd.data(), rots); else phase_fwd1_split<false, 0>(t, acc.data(), w.data(), T.fwd.data(), rots); }
        });
        each([&](int h, int t) { int p, r, g; map_fwd2<Tasks>(t, 0, p, r, g); if (h) phase_fwd2_split<1>(p, r, g, w.data()); else phase_fwd2_split<0>(p, r, g, w.data()); });
        {   // in place: every thread holds its 16 inputs in registers across a barrier
            std::vector<u64> held((size_t)TH * 16);
            each([&](int h, int t) { int p, r, u; map_fwd3<Tasks>(t, 0, p, r, u); phase_fwd3_split_load(p, r, u, w.data(), &held[(size_t)(h * 256 * Cfg::CT + t) * 16]); });
            each([&](int h, int t) {
                int p, r, u; map_fwd3<Tasks>(t, 0, p, r, u);
                const u64 *v = &held[(size_t)(h * 256 * Cfg::CT + t) * 16];
                if (h) phase_fwd3_split_finish<1>(p, r, u, w.data(), v); else phase_fwd3_split_finish<0>(p, r, u, w.data(), v);
            });
        }
    } else {
        for (int it = 0; it < Cfg::FWD_SWEEPS; it++)
            for (int tid = 0; tid < TH; tid++) {
                if (rot) phase_fwd1<true>(it * TH + tid, acc.data(), w.data(), T.fwd.data(), rots);
                else phase_fwd1<false>(it * TH + tid, acc.data(), w.data(), T.fwd.data(), rots);
            }
        for (int it = 0; it < Cfg::FWD_SWEEPS; it++)
            for (int tid = 0; tid < TH; tid++) { int p, r, g; map_fwd2<Cfg>(tid, it, p, r, g); phase_fwd2(p, r, g, w.data()); }
        for (int it = 0; it < Cfg::FWD_SWEEPS; it++)
            for (int tid = 0; tid < TH; tid++) { int p, r, u; map_fwd3<Cfg>(tid, it, p, r, u); phase_fwd3(p, r, u, w.data()); }
    }
    for (int tid = 0; tid < TH; tid++) phase_mac<Cfg>(tid, w.data(), bk.data());
    if constexpr (Cfg::SPLIT_INV) {
        // the split inverse phases of the wide shape, one loop per barrier-separated sub-pass (kernels.cuh: br2_step)
        auto each = [&](auto fn) { for (int tid = 0; tid < TH; tid++) { int h, t; if (map_split<Cfg>(tid, h, t)) fn(h, t); } };
        each([&](int h, int t) { int p, r, u; map_inv3<Cfg>(t, p, r, u); if (h) phase_inv3_split_a<1>(p, r, u, w.data()); else phase_inv3_split_a<0>(p, r, u, w.data()); });
        each([&](int h, int t) { int p, r, u; map_inv3<Cfg>(t, p, r, u); if (h) phase_inv3_split_b<1>(p, r, u, w.data()); else phase_inv3_split_b<0>(p, r, u, w.data()); });
        each([&](int h, int t) { int p, r, g; map_inv2_split<Cfg>(t, p, r, g); if (h) phase_inv2_split<1>(p, r, g, w.data()); else phase_inv2_split<0>(p, r, g, w.data()); });
        each([&](int h, int t) { if (h) phase_inv1_split_a<1>(t, w.data(), T.inv.data()); else phase_inv1_split_a<0>(t, w.data(), T.inv.data()); });
        each([&](int h, int t) {
            if (rot) { if (h) phase_inv1_split_b<true, 1>(t, acc.data(), w.data()); else phase_inv1_split_b<true, 0>(t, acc.data(), w.data()); }
            else { if (h) phase_inv1_split_b<false, 1>(t, acc.data(), w.data()); else phase_inv1_split_b<false, 0>(t, acc.data(), w.data()); }
        });
    } else {
        for (int tid = 0; tid < TH; tid++) { int p, r, u; if (map_inv3<Cfg>(tid, p, r, u)) phase_inv3(p, r, u, w.data()); }
        for (int tid = 0; tid < TH; tid++) { int p, r, g; if (map_inv2<Cfg>(tid, p, r, g)) phase_inv2(p, r, g, w.data()); }
        for (int tid = 0; tid < Cfg::INV_TASKS; tid++) {
            if (rot) phase_inv1<true>(tid, acc.data(), w.data(), T.inv.data());
            else phase_inv1<false>(tid, acc.data(), w.data(), T.inv.data());
        }
    }
    for (int c = 0; c < nct; c++) memcpy(acc_io + c * 2 * NTT_N, &acc[c * 2 * NTT_N], sizeof(i32) * 2 * NTT_N);
}

extern "C" {

void emul_phase_step(i32 *acc_io, const u64 *bk_ref_row, const int *rot, int nct)
{
    phase_step<BrDefault>(acc_io, bk_ref_row, rot, nct);
}
// the wide shape (one ciphertext on 256 threads)
void emul_phase_step_wide(i32 *acc_io, const u64 *bk_ref_row, const int *rot)
{
    phase_step<BrWide>(acc_io, bk_ref_row, rot, 1);
}
// the wide2 shape (one ciphertext on 512 threads: split forward and inverse phases)
void emul_phase_step_wide2(i32 *acc_io, const u64 *bk_ref_row, const int *rot)
{
    phase_step<BrWide2>(acc_io, bk_ref_row, rot, 1);
}

void emul_ff_shl_var(const u64 *in, const int *s, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_shl_var(in[i], s[i]); }
void emul_ff_mul(const u64 *a, const u64 *b, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_mul(a[i], b[i]); }
void emul_ff_add(const u64 *a, const u64 *b, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_add(a[i], b[i]); }
void emul_ff_sub(const u64 *a, const u64 *b, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_sub(a[i], b[i]); }

}  // extern "C"
