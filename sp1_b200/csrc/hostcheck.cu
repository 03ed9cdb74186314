// Host execution of the DEVICE arithmetic sources (kb31.cuh / poseidon2.cuh are __host__ __device__): lets the CPU
// test-suite check the exact code the kernels run (lazy-reduction bounds, half-product Montgomery forms) against the
// oracle before any GPU time is spent.  Test support only; not part of include/sp1b200.h.
#include "poseidon2.cuh"

extern "C" {
void sp1b200_hostcheck_permute(uint32_t* states, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        uint32_t s[16];
        for (int k = 0; k < 16; k++) s[k] = states[i * 16 + k];
        p2::permute(s);
        for (int k = 0; k < 16; k++) states[i * 16 + k] = s[k];
    }
}
void sp1b200_hostcheck_ext_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        kb::Ext x{{a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]}}, y{{b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]}};
        kb::Ext r = kb::ext_mul(x, y);
        for (int k = 0; k < 4; k++) out[4 * i + k] = r.c[k];
    }
}
void sp1b200_hostcheck_ext_inv(const uint32_t* a, uint32_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        kb::Ext x{{a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]}};
        kb::Ext r = kb::ext_inv(x);
        for (int k = 0; k < 4; k++) out[4 * i + k] = r.c[k];
    }
}
void sp1b200_hostcheck_field(const uint32_t* a, const uint32_t* b, uint32_t* add, uint32_t* sub, uint32_t* mul, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) { add[i] = kb::add(a[i], b[i]); sub[i] = kb::sub(a[i], b[i]); mul[i] = kb::mul(a[i], b[i]); }
}
}
