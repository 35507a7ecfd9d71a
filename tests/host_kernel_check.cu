// host_kernel_check.cu -- runs the per-thread byte/bit permutation code of the CUDA kernels
// (PRMT selectors, 4x4 byte transposes, the 8x8 anti-diagonal bit transpose, plane scatter /
// gather, per-group bitshuffle) on the CPU and compares it with the oracle.  This is a test
// helper: it links oracle/blosc_oracle.c and is built by tests/test_host_kernel_check.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../go-blosc_b200/csrc/filter_kernels.cuh"
extern "C" {
#include "../oracle/blosc_oracle.h"
}

using namespace b2b;

static int failures = 0;
#define CHECK(cond, ...)                                   \
    do {                                                   \
        if (!(cond)) { printf("FAIL: " __VA_ARGS__); printf("\n"); failures++; } \
    } while (0)

template <int T> void check_shuffle_planes() {
    constexpr int TE = ShufCfg<T>::TE, EPV = ShufCfg<T>::EPV;
    std::vector<uint8_t> in(kTileBytes), smem(kTileBytes), want(kTileBytes), back(kTileBytes);
    for (auto &b : in) b = (uint8_t)rand();
    for (int v = 0; v < kTileBytes / 16; v++) {
        uint4 x;
        memcpy(&x, in.data() + 16 * v, 16);
        planes_from_vec<T>(smem.data(), v * EPV, x);
    }
    orc_shuffle(in.data(), want.data(), kTileBytes, T);  // one tile == a buffer of TE elements
    CHECK(memcmp(smem.data(), want.data(), kTileBytes) == 0, "planes_from_vec<%d>", T);
    for (int v = 0; v < kTileBytes / 16; v++) {
        uint4 x = vec_from_planes<T>(want.data(), v * EPV);
        memcpy(back.data() + 16 * v, &x, 16);
    }
    CHECK(memcmp(back.data(), in.data(), kTileBytes) == 0, "vec_from_planes<%d>", T);
    (void)TE;
}

template <int T> void check_bitshuffle_groups() {
    const int groups = 64, n = groups * 8 * T;
    std::vector<uint8_t> in(n), got(n), want(n), back(n);
    for (auto &b : in) b = (uint8_t)rand();
    for (int g = 0; g < groups; g++) bitshuffle_group<T>(in.data() + g * 8 * T, got.data() + g * 8 * T);
    orc_bitshuffle(in.data(), want.data(), n, T);
    CHECK(memcmp(got.data(), want.data(), n) == 0, "bitshuffle_group<%d>", T);
    for (int g = 0; g < groups; g++) bitunshuffle_group<T>(want.data() + g * 8 * T, back.data() + g * 8 * T);
    CHECK(memcmp(back.data(), in.data(), n) == 0, "bitunshuffle_group<%d>", T);
}

void check_generic_item(int T) {
    const int groups = 9, n = groups * 8 * T;
    std::vector<uint8_t> in(n), got(n), want(n), back(n);
    for (auto &b : in) b = (uint8_t)rand();
    for (int g = 0; g < groups; g++)
        for (int j = 0; j < T; j++) bitshuffle_item_generic(in.data(), got.data(), T, g, j, false);
    orc_bitshuffle(in.data(), want.data(), n, T);
    CHECK(memcmp(got.data(), want.data(), n) == 0, "bitshuffle_item_generic T=%d", T);
    for (int g = 0; g < groups; g++)
        for (int j = 0; j < T; j++) bitshuffle_item_generic(want.data(), back.data(), T, g, j, true);
    CHECK(memcmp(back.data(), in.data(), n) == 0, "bitunshuffle_item_generic T=%d", T);
}

int main() {
    srand(1234);
    // the bit transpose against the reference formula (shuffle.go:184-200)
    for (int rep = 0; rep < 2000; rep++) {
        uint8_t b[8], want[8], got[8];
        for (int m = 0; m < 8; m++) b[m] = (uint8_t)rand();
        for (int k = 0; k < 8; k++) {
            uint8_t o = 0;
            for (int m = 0; m < 8; m++) if (b[m] & (1 << (7 - k))) o |= (uint8_t)(1 << (7 - m));
            want[k] = o;
        }
        uint32_t lo, hi;
        memcpy(&lo, b, 4); memcpy(&hi, b + 4, 4);
        bit_transpose8(lo, hi);
        memcpy(got, &lo, 4); memcpy(got + 4, &hi, 4);
        CHECK(memcmp(got, want, 8) == 0, "bit_transpose8 rep %d", rep);
    }
    check_shuffle_planes<2>(); check_shuffle_planes<4>(); check_shuffle_planes<8>(); check_shuffle_planes<16>();
    check_bitshuffle_groups<2>(); check_bitshuffle_groups<4>(); check_bitshuffle_groups<8>(); check_bitshuffle_groups<16>();
    for (int T : {2, 3, 5, 7, 8, 12, 17, 255}) check_generic_item(T);
    printf(failures ? "host_kernel_check: %d FAILURES\n" : "host_kernel_check: all ok\n", failures);
    return failures ? 1 : 0;
}
