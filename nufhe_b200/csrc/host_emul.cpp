// host_emul.cpp -- runs the per-lane device code of ntt_lane.cuh on the CPU, one "warp" at a time
// (32 lanes executed in sequence, the smem transpose replaced by an array transpose).  Built with g++
// into libnb_host_emul.so for tests/test_lane_emulation.py: it lets the CPU-only test suite check the
// index maps, twiddle tables and shift constants of the GPU transform against the oracle.
// This is a test aid for the CUDA source, not a CPU fallback: nothing in nufhe_b200/ loads it.
#include <cstring>
#include <vector>
#include "tables.h"

using namespace nb;

static const NttTables &tables() { static NttTables t; return t; }

static void transpose(u64 v[32][32])
{
    for (int a = 0; a < 32; a++)
        for (int b = a + 1; b < 32; b++) { u64 t = v[a][b]; v[a][b] = v[b][a]; v[b][a] = t; }
}

extern "C" {

// in/out natural order, canonical
void emul_ntt_forward(const u64 *in, u64 *out, size_t batch)
{
    const NttTables &T = tables();
    for (size_t b = 0; b < batch; b++) {
        u64 v[32][32];
        for (int l = 0; l < 32; l++) {
            for (int s = 0; s < 32; s++) v[l][s] = ff_canon(in[b * NTT_N + ntt_in_index(l, s)]);
            ntt_fwd_pre(v[l], T.fwd.data() + l, l);
        }
        transpose(v);
        for (int l = 0; l < 32; l++) {
            ntt_fwd_post(v[l]);
            for (int s = 0; s < 32; s++) out[b * NTT_N + ntt_out_index(l, s)] = v[l][s];
        }
    }
}

void emul_ntt_inverse(const u64 *in, u64 *out, size_t batch)
{
    const NttTables &T = tables();
    for (size_t b = 0; b < batch; b++) {
        u64 v[32][32];
        for (int l = 0; l < 32; l++) {
            for (int s = 0; s < 32; s++) v[l][s] = ff_canon(in[b * NTT_N + ntt_out_index(l, s)]);
            ntt_inv_pre(v[l]);
        }
        transpose(v);
        for (int l = 0; l < 32; l++) {
            ntt_inv_post(v[l], T.inv.data() + l, l);
            for (int s = 0; s < 32; s++) out[b * NTT_N + ntt_in_index(l, s)] = v[l][s];
        }
    }
}

// One external-product step of the phase-structured kernel (br_phases.cuh) for up to BR2_CT
// ciphertexts, phases executed in order with all "threads" of a phase run back to back.
// acc: (nct, 2, 1024) in/out; bk_ref_row: reference layout (2,2,2,1024) Montgomery; rot: rotation
// amounts per ciphertext or NULL (plain external product, overwrite).
int emul_phase_ct(void) { return BR2_CT; }

void emul_phase_step(i32 *acc_io, const u64 *bk_ref_row, const int *rot, int nct)
{
    static PhaseTables T;
    std::vector<i32> acc(BR2_CT * 2 * NTT_N, 0);
    std::vector<u64> w(BR2_POLYS * POLY_STRIDE, 0);
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
    for (int it = 0; it < 2; it++)
        for (int tid = 0; tid < BR2_THREADS; tid++) {
            if (rot) phase_fwd1<true>(it * BR2_THREADS + tid, acc.data(), w.data(), T.fwd.data(), rots);
            else phase_fwd1<false>(it * BR2_THREADS + tid, acc.data(), w.data(), T.fwd.data(), rots);
        }
    for (int it = 0; it < 2; it++)
        for (int tid = 0; tid < BR2_THREADS; tid++) { int p, r, g; map_fwd2(tid, it, p, r, g); phase_fwd2(p, r, g, w.data()); }
    for (int it = 0; it < 2; it++)
        for (int tid = 0; tid < BR2_THREADS; tid++) { int p, r, u; map_fwd3(tid, it, p, r, u); phase_fwd3(p, r, u, w.data()); }
    for (int tid = 0; tid < BR2_THREADS; tid++) phase_mac(tid, w.data(), bk.data());
    for (int tid = 0; tid < BR2_THREADS; tid++) { int p, r, u; map_inv3(tid, p, r, u); phase_inv3(p, r, u, w.data()); }
    for (int tid = 0; tid < BR2_THREADS; tid++) { int p, r, g; map_inv2(tid, p, r, g); phase_inv2(p, r, g, w.data()); }
    for (int tid = 0; tid < BR2_THREADS; tid++) {
        if (rot) phase_inv1<true>(tid, acc.data(), w.data(), T.inv.data());
        else phase_inv1<false>(tid, acc.data(), w.data(), T.inv.data());
    }
    for (int c = 0; c < nct; c++) memcpy(acc_io + c * 2 * NTT_N, &acc[c * 2 * NTT_N], sizeof(i32) * 2 * NTT_N);
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
