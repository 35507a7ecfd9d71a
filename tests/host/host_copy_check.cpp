// tests/host/host_copy_check.cpp -- CPU check of the host copy pool behind the pinned staging ring
// (go-blosc_b200/csrc/host_staging.hpp: CopyPool / stream_copy).  No GPU, no CUDA runtime calls: only the parts of
// the header that never touch the device are used.  Built and run by tests/test_host_copy.py.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../go-blosc_b200/csrc/host_staging.hpp"

int main() {
    unsigned long long x = 88172645463325252ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    const size_t cap = (48u << 20) + 4096;
    std::vector<unsigned char> src(cap), dst(cap), ref(cap);
    for (size_t i = 0; i < cap; i++) src[i] = (unsigned char)rnd();
    int bad = 0, cases = 0;
    for (int threads : {0, 1, 3, 7}) {
        b2b::CopyPool pool(threads);
        const size_t sizes[] = {0, 1, 31, 32, 33, 4095, 4096, 4097, 100000, (512u << 10) - 1, (512u << 10) + 1, (1u << 20) + 17,
                                (8u << 20), (8u << 20) + 4095, (40u << 20) + 123};
        for (size_t n : sizes) {
            for (int rep = 0; rep < 3; rep++) {
                const size_t so = rnd() % 67, dof = rnd() % 67;
                memset(dst.data(), 0xA5, n + 256 < cap ? n + 256 : cap);
                memset(ref.data(), 0xA5, n + 256 < cap ? n + 256 : cap);
                pool.copy(dst.data() + dof, src.data() + so, n);
                memcpy(ref.data() + dof, src.data() + so, n);
                cases++;
                if (memcmp(dst.data(), ref.data(), n + 200 < cap ? n + 200 : cap) != 0) { bad++; fprintf(stderr, "mismatch: threads %d n %zu so %zu do %zu\n", threads, n, so, dof); }
            }
        }
    }
    printf("host copy pool: %d cases, %d bad\n", cases, bad);
    return bad ? 1 : 0;
}
