// host_emul.cpp -- runs the per-lane device code of ntt_lane.cuh on the CPU, one "warp" at a time
// (32 lanes executed in sequence, the smem transpose replaced by an array transpose).  Built with g++
// into libnb_host_emul.so for tests/test_lane_emulation.py: it lets the CPU-only test suite check the
// index maps, twiddle tables and shift constants of the GPU transform against the oracle.
// This is a test aid for the CUDA source, not a CPU fallback: nothing in nufhe_b200/ loads it.
#include <cstring>
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

void emul_ff_shl_var(const u64 *in, const int *s, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_shl_var(in[i], s[i]); }
void emul_ff_mul(const u64 *a, const u64 *b, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_mul(a[i], b[i]); }
void emul_ff_add(const u64 *a, const u64 *b, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_add(a[i], b[i]); }
void emul_ff_sub(const u64 *a, const u64 *b, u64 *out, size_t n)
{ for (size_t i = 0; i < n; i++) out[i] = ff_sub(a[i], b[i]); }

}  // extern "C"
