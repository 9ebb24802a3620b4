// Counter-based random numbers for the engine: Philox4x32-10 (Salmon et al., SC'11), written out
// here so that host and device draw from the very same function.
//
// The reference seeds a std::ranlux24_base from std::random_device per Sampler and has no seed
// argument anywhere (src/sampler.cpp:14-19), so there is no stream to be faithful to.
// The engine instead gives every (chain, update kind, iteration, sample-or-locus, cell) its own
// stream, which is what makes the per-sample / per-locus Metropolis problems run in parallel
// and makes results independent of how the temperature ladder is sharded over GPUs.
//
// Stream addressing (one Philox key per engine seed):
//   counter[0] = iteration                      (initialisation attempt t uses 0xFFFF0000 + t)
//   counter[1] = (kind << 16) | global chain id
//   counter[2] = unit: sample index for per-sample updates, locus for update_p, 0 for mean_coi
//   counter[3] = (sub << 16) | block
//       sub   = 0 for the unit's own proposal draws; locus + 1 for the latent-genotype draws of
//               one cell; proposal index for update_p
//       block = running index of the 128-bit block within the stream
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MOIRE_HD __host__ __device__ __forceinline__
#define MOIRE_HD_OUTLINE __host__ __device__ __noinline__
#else
#define MOIRE_HD inline
#define MOIRE_HD_OUTLINE inline
#endif

namespace moire
{
enum StreamKind : uint32_t
{
    KIND_EPS_NEG = 0,
    KIND_EPS_POS = 1,
    KIND_P = 2,
    KIND_M = 3,
    KIND_R = 4,
    KIND_EFF_COI = 5,
    KIND_SAMPLES = 6,
    KIND_MEAN_COI = 7,
    KIND_SWAP = 8,
    KIND_INIT_SAMPLE = 12,
    KIND_INIT_P = 13,
    KIND_INIT_MEAN_COI = 14
};

MOIRE_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

MOIRE_HD void philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
#pragma unroll
    for (int round = 0; round < 10; ++round)
    {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

#if defined(__CUDACC__)
static __device__ __noinline__ uint4 philox_block_device(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1)
{
    const uint32_t ctr[4] = {c0, c1, c2, c3};
    uint32_t out[4];
    philox4x32_10(ctr, k0, k1, out);
    return make_uint4(out[0], out[1], out[2], out[3]);
}
#endif

// A lazily refilled stream of 32-bit words.
struct Stream
{
    uint32_t ctr[4];
    uint32_t k0, k1;
    uint32_t buf[4];
    int have;

    MOIRE_HD void open(uint64_t seed, uint32_t iteration, uint32_t kind, uint32_t chain,
                       uint32_t unit, uint32_t sub)
    {
        k0 = static_cast<uint32_t>(seed);
        k1 = static_cast<uint32_t>(seed >> 32);
        ctr[0] = iteration;
        ctr[1] = (kind << 16) | (chain & 0xFFFFu);
        ctr[2] = unit;
        ctr[3] = sub << 16;
        have = 0;
    }

    MOIRE_HD void refill()
    {
#if defined(__CUDA_ARCH__)
        // one out-of-line copy of the ten rounds (the kernels are instruction-fetch bound and a
        // stream is refilled from dozens of call sites); by value, so the stream stays in registers
        const uint4 o = philox_block_device(ctr[0], ctr[1], ctr[2], ctr[3], k0, k1);
        buf[0] = o.x;
        buf[1] = o.y;
        buf[2] = o.z;
        buf[3] = o.w;
#else
        philox4x32_10(ctr, k0, k1, buf);
#endif
        ctr[3] += 1;
        have = 4;
    }

    MOIRE_HD uint32_t next_u32()
    {
        if (have == 0) refill();
        // selects, not an indexed load: keeps buf[] in registers on the device
        const uint32_t v = have == 4 ? buf[0] : have == 3 ? buf[1] : have == 2 ? buf[2] : buf[3];
        --have;
        return v;
    }

    // Uniform on (0,1): never 0, never 1 (log U and Bernoulli comparisons are always defined).
    MOIRE_HD double next_uniform() { return (static_cast<double>(next_u32()) + 0.5) * 0x1p-32; }
};
}  // namespace moire
