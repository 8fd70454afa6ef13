// pattern_math.hpp -- host-side closed form for the probe pattern's checksum.
//
// The health verdict compares the checksum a pass accumulated on the GPU with the checksum of a
// clean buffer.  That reference value must not come from the device under test: it is computed
// here, on the host, in O(32 * log n) integer operations (oracle: oracle/probe.py
// expected_checksum, which sums the pattern with numpy).
//
//   word(i, seed) = (uint32(i) * K) ^ seed,   K = 2654435761 (odd)
//   checksum      = sum over i < n of word(i, seed)                      (mod 2^64)
//                 = sum over bits b of 2^b * (seed_b ? n - c_b : c_b)
//   c_b           = #{ i < n : bit b of (uint32(i) * K mod 2^32) is set }
//
// bit b of (i*K mod 2^32) = floor(i*K / 2^b) - 2*floor(i*K / 2^(b+1)), so
//   c_b(r) = F(r, 2^b) - 2*F(r, 2^(b+1)),  F(r, m) = sum_{i<r} floor(i*K / m)
// with F the Euclid-like "floor sum".  Every full period of 2^32 indices visits each 32-bit value
// once (K is odd), i.e. contributes exactly 2^31 to every c_b.
#pragma once
#include <array>
#include <cstdint>

namespace b2dp {

constexpr uint32_t kPatternMulHost = 2654435761u;

// sum_{i<n} floor((a*i + b) / m)  (mod 2^64); exact whenever the true value fits 64 bits, and a valid
// residue otherwise (only ring operations touch `ans`; every quotient below is exact).
inline unsigned long long floor_sum_u64(unsigned long long n, unsigned long long m, unsigned long long a,
                                        unsigned long long b) {
    unsigned long long ans = 0;
    for (;;) {
        if (a >= m) {
            const unsigned long long tri = (n % 2 == 0) ? (n / 2) * (n - 1) : n * ((n - 1) / 2);
            ans += tri * (a / m);
            a %= m;
        }
        if (b >= m) {
            ans += n * (b / m);
            b %= m;
        }
        const unsigned __int128 y_max = (unsigned __int128)a * n + b;
        if (y_max < m) break;
        n = (unsigned long long)(y_max / m);
        b = (unsigned long long)(y_max % m);
        const unsigned long long t = m; m = a; a = t;
    }
    return ans;
}

// c_b for b = 0..31 over word indices 0 .. n_words-1 (the index wraps at 2^32 like the kernels' uint32 cast)
inline std::array<unsigned long long, 32> pattern_bit_counts_host(unsigned long long n_words) {
    std::array<unsigned long long, 32> c{};
    const unsigned long long full = n_words >> 32, r = n_words & 0xffffffffull;
    for (int b = 0; b < 32; ++b) {
        unsigned long long v = full << 31;
        if (r) v += floor_sum_u64(r, 1ull << b, kPatternMulHost, 0) - 2ull * floor_sum_u64(r, 1ull << (b + 1), kPatternMulHost, 0);
        c[b] = v;
    }
    return c;
}

// closed-form checksum of a clean buffer of n_words words keyed with `seed`
inline unsigned long long expected_checksum_from_counts(const std::array<unsigned long long, 32>& c,
                                                        unsigned long long n_words, uint32_t seed) {
    unsigned long long s = 0;
    for (int b = 0; b < 32; ++b) {
        const unsigned long long ones = ((seed >> b) & 1u) ? n_words - c[b] : c[b];
        s += ones << b;
    }
    return s;
}

inline unsigned long long expected_checksum_host(unsigned long long n_words, uint32_t seed) {
    return expected_checksum_from_counts(pattern_bit_counts_host(n_words), n_words, seed);
}

}  // namespace b2dp
