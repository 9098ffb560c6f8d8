// exact_acc.cuh — error-free accumulation of float32 contributions (deterministic mode 2).
//
// The reference's CPU path is not bit-stable above one thread (omp critical / omp atomic, arrival order:
// src/engine/accumulator.cpp:68-89, glyph_kernels.cu:169-179), and any float accumulation — atomics, a
// sorted in-order fold, a rank-ordered merge — depends on how the points were chunked and sharded.  Here
// every additive state word is a fixed-point number wide enough for ANY finite float32:
//
//     value = sum_k limb[k] * 2^(32 k - 149),   k = 0..8,   limb[k] a signed 64-bit integer
//
// A float m * 2^(e-150) (24-bit m) lands at bit e-1 of that 288-bit number, i.e. in at most two adjacent
// limbs, each of which receives less than 2^32 per contribution: integer adds, exact, commutative and
// associative — the result does not depend on the order of the points, on the ingest chunking, or on how
// many GPUs the cloud was sharded over (ranks combine with an integer all-reduce).  One rounding happens
// at finalize (round to nearest even), so the band is the correctly rounded exact sum: always inside the
// any-order fp32 bound the parity tests use.  NaN and +-inf contributions are tracked as flags and
// reproduce float semantics (NaN wins; +inf and -inf together give NaN).  Capacity: 2^30 contributions
// per cell and word.
#pragma once

#include "common.cuh"

namespace pcrb {

constexpr int kXLimbs = 9;

struct XAcc {
    long long* limbs;      // [n_add][kXLimbs][cells]
    uint32_t*  flags;      // [n_add][cells]   byte 0: NaN seen, byte 1: +inf seen, byte 2: -inf seen
    int32_t*   ext;        // [n_max + n_min][cells]   ordered-int max / min words
    size_t     cells;
};

__device__ __forceinline__ void xacc_add(const XAcc& A, int word, size_t cell, float v)
{
    const uint32_t b = __float_as_uint(v);
    uint32_t e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
    if (e == 0xffu) {
        atomicOr(A.flags + static_cast<size_t>(word) * A.cells + cell, m ? 0x1u : ((b >> 31) ? 0x10000u : 0x100u));
        return;
    }
    if (e == 0) { if (m == 0) return; e = 1; } else m |= 0x800000u;
    const int pos = static_cast<int>(e) - 1;                   // bit of the mantissa's LSB, 0..253
    const int k = pos >> 5;
    const unsigned long long sh = static_cast<unsigned long long>(m) << (pos & 31);
    long long lo = static_cast<long long>(sh & 0xffffffffull), hi = static_cast<long long>(sh >> 32);
    if (b >> 31) { lo = -lo; hi = -hi; }
    unsigned long long* base = reinterpret_cast<unsigned long long*>(A.limbs) +
                               (static_cast<size_t>(word) * kXLimbs + k) * A.cells + cell;
    if (lo) atomicAdd(base, static_cast<unsigned long long>(lo));             // RED.E.ADD.64
    if (hi) atomicAdd(base + A.cells, static_cast<unsigned long long>(hi));
}

// limbs -> the float32 nearest to the exact sum (ties to even)
__device__ __forceinline__ float xacc_round(const long long (&L)[kXLimbs], uint32_t flags)
{
    if (flags & 0x1u) return __int_as_float(0x7fc00000);
    const bool pinf = flags & 0x100u, ninf = flags & 0x10000u;
    if (pinf && ninf) return __int_as_float(0x7fc00000);
    if (pinf) return __int_as_float(0x7f800000);
    if (ninf) return __int_as_float(0xff800000);

    uint32_t d[kXLimbs + 2];                                   // two's complement, 32-bit digits
    long long carry = 0;
#pragma unroll
    for (int k = 0; k < kXLimbs; ++k) {
        const long long t = L[k] + carry;
        d[k] = static_cast<uint32_t>(t);
        carry = t >> 32;
    }
    d[kXLimbs] = static_cast<uint32_t>(carry);
    d[kXLimbs + 1] = static_cast<uint32_t>(carry >> 32);
    const bool neg = carry < 0;
    if (neg) {                                                 // magnitude
        uint32_t c = 1;
#pragma unroll
        for (int k = 0; k < kXLimbs + 2; ++k) {
            const uint32_t t = ~d[k];
            d[k] = t + c;
            c = (c && t == 0xffffffffu) ? 1u : 0u;
        }
    }
    int top = -1;
#pragma unroll
    for (int k = 0; k < kXLimbs + 2; ++k) if (d[k]) top = k;
    if (top < 0) return 0.0f;
    const int P = 32 * top + 31 - __clz(d[top]);               // most significant bit of the magnitude
    uint32_t bits;
    if (P <= 23) {
        bits = d[0];                                           // subnormal or first binade: exact as it is
    } else {
        const int shift = P - 23;                              // keep bits [shift, shift + 24)
        auto digit = [&](int k) -> uint32_t {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < kXLimbs + 2; ++j) if (j == k) v = d[j];
            return v;
        };
        const int w = shift >> 5, o = shift & 31;
        const unsigned long long pair = (static_cast<unsigned long long>(digit(w + 1)) << 32) | digit(w);
        uint32_t top24 = static_cast<uint32_t>((pair >> o) & 0xffffffu);
        // round bit = bit shift-1, sticky = anything below it
        const int rb = shift - 1, rw = rb >> 5, ro = rb & 31;
        const uint32_t rdig = digit(rw);
        const bool round = (rdig >> ro) & 1u;
        bool sticky = (rdig & ((1u << ro) - 1u)) != 0;
#pragma unroll
        for (int j = 0; j < kXLimbs + 2; ++j) if (j < rw && d[j]) sticky = true;
        if (round && (sticky || (top24 & 1u))) ++top24;
        const unsigned long long bb = (static_cast<unsigned long long>(shift) << 23) + top24;   // carries into the exponent
        bits = bb >= 0x7f800000ull ? 0x7f800000u : static_cast<uint32_t>(bb);
    }
    return __uint_as_float(bits | (neg ? 0x80000000u : 0u));
}

}  // namespace pcrb
