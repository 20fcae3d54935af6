// DEV SIMULATION of the device code for the CPU ("not gpu") test-suite.
//
// Compiles the very same headers the CUDA kernels are built from (zkvm_pairings_b200/csrc/*.cuh)
// as plain C++ with the PTX carry flag emulated (fp.cuh, ZKP_HOST_SIM), so the limb-level
// Montgomery algorithm, the tower and the pairing control flow can be checked against the oracle
// without a GPU.  TEST INFRASTRUCTURE ONLY: it is built into tests/host_sim/libzkpair_sim.so, is
// never linked into or loaded by libzkpair.so / the zkvm_pairings_b200 package, and is not a CPU
// fallback (the product path raises when the CUDA library or a GPU is missing).
#define ZKP_HOST_SIM 1
#include <stddef.h>
#include "../../zkvm_pairings_b200/csrc/ops.cuh"

using namespace zkp;

extern "C" {
int sim_tower_op(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint8_t *status, size_t n) {
    int na, nb, nr;
    tower_op_shape(op, na, nb, nr);
    for (size_t i = 0; i < n; i++) {
        uint8_t s = tower_op_one(op, a + 6 * na * i, b ? b + 6 * nb * i : nullptr, out + 6 * nr * i);
        if (status) status[i] = s;
    }
    return 0;
}
int sim_pairing(int mode, const uint64_t *g1, const uint8_t *g1inf, const uint64_t *g2, const uint8_t *g2inf,
                size_t n, int k, const uint64_t *in12, uint64_t *out, uint8_t *is_one) {
    int bad = 0;
    for (size_t i = 0; i < n; i++) {
        size_t e = i * (size_t)k;
        bad |= pairing_one<8>(mode, g1 ? g1 + 12 * e : nullptr, g1inf ? g1inf + e : nullptr, g2 ? g2 + 24 * e : nullptr,
                           g2inf ? g2inf + e : nullptr, k, in12 ? in12 + 72 * i : nullptr, out + 72 * i,
                           is_one ? is_one + i : nullptr);
    }
    return bad ? -1 : 0;
}
uint64_t sim_splitmix64_at(uint64_t seed, uint64_t idx) { return splitmix64_at(seed, idx); }
void sim_gen_points(const uint64_t *k1, const uint64_t *k2, size_t n, uint64_t *g1, uint8_t *g1inf, uint64_t *g2, uint8_t *g2inf) {
    for (size_t i = 0; i < n; i++) {
        gen_g1_one(k1[i], g1 + 12 * i, g1inf + i);
        gen_g2_one(k2[i], g2 + 24 * i, g2inf + i);
    }
}
}

#ifdef ZKP_TRACK_BOUNDS
extern "C" void sim_max_bounds(double *out) {
    out[0] = zkp::g_max_lb; out[1] = zkp::g_max_tb; out[2] = zkp::g_max_vb; out[3] = zkp::g_max_col;
}
#endif
