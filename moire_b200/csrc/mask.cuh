// Allele sets (observed alleles, latent genotypes) as bit masks, one bit per allele of a locus.
//
// The reference keeps them as std::vector<int8_t> of any length (src/genotyping_data.cpp:17-48,
// src/chain.h:96-97).  Here the width is a compile-time choice of the whole library:
//   MOIRE_MASK_WORDS == 1 (default, libmoire_b200.so): one 64-bit word, up to 64 alleles per locus --
//     every benchmarked shape, every list entry is 16 bytes;
//   MOIRE_MASK_WORDS == 2 (libmoire_b200_w128.so): two words, up to 128 alleles per locus -- the same
//     sources compiled once more, same entry points, masks cross the C ABI as [..][2] uint64 (word 0 =
//     alleles 0..63).  moire_engine_mask_words() says which library one is talking to; the Python
//     host side picks by the widest locus of the data (moire_b200/_lib.py).
// Everything that looks inside a mask goes through the helpers below, so the two builds differ in
// this header and in the list-entry layout only.
#pragma once
#include <cstdint>

#ifndef MOIRE_MASK_WORDS
#define MOIRE_MASK_WORDS 1
#endif

namespace moire
{
#if MOIRE_MASK_WORDS == 1
typedef unsigned long long mask_t;
constexpr int kMaskBits = 64;

__host__ __device__ __forceinline__ int mpopc(mask_t x)
{
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// 1-based position of the lowest set bit, 0 for the empty set (ffs convention)
__host__ __device__ __forceinline__ int mffs(mask_t x)
{
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x);
#else
    return __builtin_ffsll((long long)x);
#endif
}
__host__ __device__ __forceinline__ unsigned long long mword(mask_t x, int w) { return w == 0 ? x : 0ull; }
__host__ __device__ __forceinline__ mask_t mfrom_words(unsigned long long lo, unsigned long long) { return lo; }
#elif MOIRE_MASK_WORDS == 2
typedef unsigned __int128 mask_t;
constexpr int kMaskBits = 128;

__host__ __device__ __forceinline__ int mpopc(mask_t x)
{
#ifdef __CUDA_ARCH__
    return __popcll((unsigned long long)x) + __popcll((unsigned long long)(x >> 64));
#else
    return __builtin_popcountll((unsigned long long)x) + __builtin_popcountll((unsigned long long)(x >> 64));
#endif
}
__host__ __device__ __forceinline__ int mffs(mask_t x)
{
    const unsigned long long lo = (unsigned long long)x, hi = (unsigned long long)(x >> 64);
#ifdef __CUDA_ARCH__
    if (lo) return __ffsll((long long)lo);
    return hi ? 64 + __ffsll((long long)hi) : 0;
#else
    if (lo) return __builtin_ffsll((long long)lo);
    return hi ? 64 + __builtin_ffsll((long long)hi) : 0;
#endif
}
__host__ __device__ __forceinline__ unsigned long long mword(mask_t x, int w)
{
    return (unsigned long long)(w == 0 ? x : x >> 64);
}
__host__ __device__ __forceinline__ mask_t mfrom_words(unsigned long long lo, unsigned long long hi)
{
    return ((mask_t)hi << 64) | (mask_t)lo;
}
#else
#error "MOIRE_MASK_WORDS must be 1 or 2"
#endif

__host__ __device__ __forceinline__ mask_t mbit(int a) { return (mask_t)1 << a; }
__host__ __device__ __forceinline__ bool mtest(mask_t x, int a) { return (unsigned)(x >> a) & 1u; }
// the alleles 0 .. A-1 of a locus
__host__ __device__ __forceinline__ mask_t mall(int A) { return A >= kMaskBits ? ~(mask_t)0 : (mbit(A) - 1); }
// drop the lowest set bit / isolate it
__host__ __device__ __forceinline__ mask_t mdrop_low(mask_t x) { return x & (x - 1); }
__host__ __device__ __forceinline__ mask_t mlowest(mask_t x) { return x & ((mask_t)0 - x); }

// ---- an entry of the evaluator's work list ----------------------------------------------------------
// {latent genotype, (t << 16) | j, i}: 16 bytes with one-word masks (one 128-bit load), 32 with two.
#if MOIRE_MASK_WORDS == 1
typedef uint4 ListEntry;
__host__ __device__ __forceinline__ ListEntry make_entry(mask_t lat, unsigned tj, unsigned i)
{
    return make_uint4((unsigned)lat, (unsigned)(lat >> 32), tj, i);
}
__host__ __device__ __forceinline__ mask_t entry_lat(const ListEntry &e)
{
    return ((unsigned long long)e.y << 32) | e.x;
}
// every latent allele below 32 (the register path reads the genotype from e.x alone)
__host__ __device__ __forceinline__ bool entry_low32(const ListEntry &e) { return e.y == 0; }
#else
struct alignas(16) ListEntry
{
    unsigned x, y, z, w;  // as the 16-byte entry: genotype words 0 and 1 (alleles 0..63), (t << 16) | j, i
    unsigned long long hi, pad;  // alleles 64..127
};
__host__ __device__ __forceinline__ ListEntry make_entry(mask_t lat, unsigned tj, unsigned i)
{
    ListEntry e;
    e.x = (unsigned)lat;
    e.y = (unsigned)(lat >> 32);
    e.z = tj;
    e.w = i;
    e.hi = (unsigned long long)(lat >> 64);
    e.pad = 0;
    return e;
}
__host__ __device__ __forceinline__ mask_t entry_lat(const ListEntry &e)
{
    return mfrom_words(((unsigned long long)e.y << 32) | e.x, e.hi);
}
__host__ __device__ __forceinline__ bool entry_low32(const ListEntry &e) { return e.y == 0 && e.hi == 0; }
#endif
// the entry of a lane without a cell: one latent allele (allele 0), so that a stray evaluation is harmless
__host__ __device__ __forceinline__ ListEntry idle_entry() { return make_entry((mask_t)1, 0u, 0u); }
}  // namespace moire
