// viewer_pack_quantise_check.cpp — sweeps quantise_unit<S> (viewer_pack_ops.h: the float / integer form the kernel uses)
// against the literal double expression of the reference (pack_utils.h:56-67) over EVERY float in [-1, 1] (argv[1] = stride
// between consecutive bit patterns: 1 = all 2.1e9 of them, the CPU suite uses a larger stride plus the neighbourhoods of
// every k / S boundary), and over a band outside it (the fall-back path).  Prints the number of mismatches.
// Build: g++ -O2 -fopenmp -ffp-contract=off -I divshot_b200/csrc tests/native/viewer_pack_quantise_check.cpp
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "viewer_pack_ops.h"

using namespace dvs_vp;

template <int S>
static long long sweep(uint32_t stride) {
    long long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long b = 0; b <= 0x3f900000ll; b += stride) {  // magnitudes 0 .. 1.125 (past 1: the literal fall-back)
        for (int sgn = 0; sgn < 2; sgn++) {
            const float x = f32_from_bits((uint32_t)b | (sgn ? 0x80000000u : 0u));
            const uint32_t want = trunc_wrap_u32(((double)x * 0.5 + 0.5) * (double)S);
            if (quantise_unit<S>(x) != want) bad++;
        }
    }
    // the neighbourhoods of the boundaries x = 2 k / S - 1: 64 floats either side of each
    for (int k = 0; k <= S; k++) {
        const float c = (float)(2.0 * k / S - 1.0);
        const uint32_t cb = f32_bits(c);
        for (int d = -64; d <= 64; d++) {
            const uint32_t bb = cb + (uint32_t)d;
            const float x = f32_from_bits(bb);
            if (!(x == x) || !(x >= -1.25f && x <= 1.25f)) continue;
            const uint32_t want = trunc_wrap_u32(((double)x * 0.5 + 0.5) * (double)S);
            if (quantise_unit<S>(x) != want) bad++;
        }
    }
    return bad;
}

int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)std::strtoul(argv[1], nullptr, 10) : 1u;
    const long long b11 = sweep<2047>(stride), b10 = sweep<1023>(stride);
    std::printf("stride %u mismatches_11bit %lld mismatches_10bit %lld\n", stride, b11, b10);
    return (b11 || b10) ? 1 : 0;
}
