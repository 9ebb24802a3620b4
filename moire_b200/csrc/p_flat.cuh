// update_p as a flat pipeline (src/chain.cpp:466-565).
//
// The fused kernel (k_update_p, kernels.cuh) keeps a (chain, locus) in one block for all A_j
// proposals.  ncu and a per-warp trace showed what that costs on B200: 38 % of warp time at the
// barrier after the evaluation (a block's few 63..255-subset cells gate it), 18.7 of 32 lanes per
// instruction (1000 cells per block are too few to sort into uniform warps) and a 10 K-instruction
// kernel on an instruction-fetch-bound path (profiles/README.md).  Here the work is cut the other
// way: the latent genotypes do not change during update_p, so ALL cells of ALL chains are sorted
// ONCE by cost class (device-wide counting sort), and every proposal round is
//   k_pf_round  one block per (chain, locus): sum and decide the previous proposal, commit, draw
//               the next SALT proposal                        (one latency chain per locus)
//   k_flat_eval a persistent grid over the sorted cell list: warps with identical (k, m - k + 1) in
//               every lane, no block barrier, only the evaluator's code   (all the arithmetic)
// Rounds = max_j A_j; loci with fewer alleles (or that hit the 1e-12 floor) sit out.  Streams,
// proposals and per-cell arithmetic are those of the fused kernel (the locus sums are taken in a
// different fixed order), so the two produce the same chain up to rounding-level ties
// (tests/test_gpu_shapes.py::test_flat_and_fused_pipelines_agree_...).
#pragma once
#include <algorithm>
#include <vector>
#include "kernels.cuh"

namespace moire
{
// What the evaluator needs of a cell's sample, one 32-byte sector per gather.
struct alignas(32) SampleScal
{
    double log_r, log_1mr, odds;
    int m, pad;
};

struct PFlat
{
    ListEntry *list;     // [cells] {latent genotype lo, hi, (t << 16) | j, i}, sorted by cost class
                         // (t: chain of this engine, j: locus, i: sample)
    SampleScal *ss;      // [T * N] current (m, r) of every sample, for the duration of update_p
    unsigned int *hist;  // [kSortBins] counts, [kSortBins] running offsets, 4 scalars, task counters
    double *Tcur, *Tnew; // [T][L][N] transmission terms, locus-major, for the duration of the update
    double *pprop;       // [T * L][AP] proposed frequencies
    int eval_threads;   // threads of one evaluator launch (sets the cooperative threshold)
    double *round;       // [T * L][4] logAdj, log u, allele index, flags
    const int *nonmiss;  // [L] non-missing cells of a locus
    mask_t *latLM;              // [T][L][N] latent genotypes, locus-major, 0 at missing cells
    // update_p's PAM cache (null unless the update runs in cache mode, see k_pc_sweep)
    double *Wcur, *Wnew;        // [max_coi][T * L * N] PAM vectors: accepted state / current proposal
    size_t wstride;             // T * L * N
};
constexpr unsigned kPfActive = 1u, kPfBroken = 2u;
// cache mode: the previous proposal of the locus was accepted (bit 2) and which allele it moved
// (bits 8..15) -- k_pc_sweep commits the PAM vectors of the cells that held it
constexpr unsigned kPfAccepted = 4u;
// The two term arrays of update_p swap roles per locus instead of being copied when a proposal is
// accepted: bit 3 set = the accepted terms of the locus live in PFlat::Tnew and the next proposal's
// go to PFlat::Tcur (the 8 KB copy sat on the one-warp-per-locus latency chain of k_pf_round).
constexpr unsigned kPfFlip = 8u;
__device__ __forceinline__ double *pf_terms(const PFlat &f, unsigned flags, bool proposed)
{
    return (((flags & kPfFlip) != 0) != proposed) ? f.Tnew : f.Tcur;
}
constexpr int kPfScalars = 2 * kSortBins;  // hist[kPfScalars + 0..3] = n_egf, coop_end, nwork, mid_end
constexpr int kPfCounters = kPfScalars + 4;  // one task counter per evaluator launch of an update
constexpr int kPfHistWords = kPfCounters + kMaskBits + 8;  // one per proposal round (as many as a locus has alleles) + the cache fill

// ---- once per update: locus-major copy of the current terms, class histogram, sorted list -----
// ---- cost classes of the flat pipelines, ranked by work ----------------------------------------
// The persistent evaluator is as slow as its longest single task.  A one-thread cell costs
// ~(2^k - 1) (k + 2 (m - k + 1)) dependent instructions: 127 subsets x 34 mixture terms (k = 7 in
// a hot chain that wandered to COI 40) keep one thread busy for ~70 us while a whole launch of
// the c3 ladder takes ~200 us, and a 5-chain shard of it took 180 us instead of ~35
// (profiles/README.md, r01n).  So the classes (k, m - k + 1) are ranked by estimated one-thread
// work, heaviest first (c_flat_class[kk * 64 + cnt - 1] is the rank, c_flat_est[rank] the
// estimate; ranks < 64 are the EGF classes, k > 10, by m), and every launch splits the list where
// the estimate falls below a threshold: above it a cell goes to the warp-cooperative evaluator
// whatever its k.  The threshold follows the launch's work per thread (k_pf_scan): the cooperative
// path has the shorter critical path but the lower throughput (lanes idle in its sums), so a full
// ladder wants it for the few heaviest cells only (8000 instructions ~ k >= 8, the tuned optimum
// at 40 chains) and a small shard for many more.
#ifndef MOIRE_COOP_COST_SLOPE
#define MOIRE_COOP_COST_SLOPE .25f
#endif
#ifndef MOIRE_COOP_COST_FIXED
#define MOIRE_COOP_COST_FIXED 300.f
#endif
#ifndef MOIRE_COOP_WORK_MAX
#define MOIRE_COOP_WORK_MAX 8000.f
#endif
#ifndef MOIRE_COOP_WORK_SHARE
#define MOIRE_COOP_WORK_SHARE 1.2f
#endif
__constant__ unsigned short c_flat_class[12 * 64];
__constant__ float c_flat_est[kSortBins];

// EGF cells (k > 10) cost ~k cnt^2 / 2 (the band recurrence, likelihood.cuh): 16 widths x 4 ranges
// of k share the 64 slots of row 11, so that the lanes of a chunk run (nearly) the same loops.
__host__ __device__ inline int egf_k_range(int k) { return k <= 13 ? 0 : k <= 17 ? 1 : k <= 23 ? 2 : 3; }

__device__ __forceinline__ int work_class_flat(int k, int m)
{
    int cnt = m - k + 1;
    if (k > kIeMax)
    {
        cnt = cnt < 1 ? 1 : (cnt > 16 ? 16 : cnt);
        return c_flat_class[(kIeMax + 1) * 64 + egf_k_range(k) * 16 + cnt - 1];
    }
    int kk = k;
    cnt = cnt < 1 ? 1 : (cnt > 64 ? 64 : cnt);
    if (kk < MOIRE_MERGE_K) kk = MOIRE_MERGE_K;
    return c_flat_class[kk * 64 + cnt - 1];
}

// host: the ranking.  est = subsets x (loop overhead + leading multiplications + two operations
// per accumulator of the register tier), summed over the passes of 16 mixture terms.
inline void build_flat_classes(std::vector<unsigned short> &rank, std::vector<float> &est_of_rank)
{
    struct C { int kk, cnt; double est; };
    std::vector<C> cs;
    for (int kk = MOIRE_MERGE_K; kk <= kIeMax; ++kk)
        for (int cnt = 1; cnt <= 64; ++cnt)
        {
            const double subsets = (double)((1 << kk) - 1);
            double est = 0.0;
            for (int t0 = 0; t0 < cnt; t0 += 16)
            {
                const int left = cnt - t0;
                const int nacc = left <= 2 ? 2 : left <= 4 ? 4 : left <= 6 ? 6 : left <= 8 ? 8 : left <= 12 ? 12 : 16;
                est += subsets * (24.0 + (kk - 1 + t0) + 2.0 * nacc);
            }
            cs.push_back({kk, cnt, est});
        }
    std::stable_sort(cs.begin(), cs.end(), [](const C &a, const C &b) { return a.est > b.est; });
    rank.assign(12 * 64, 0);
    est_of_rank.assign(kSortBins, 0.f);
    {
        // EGF classes: ranks 0..63 (their own section of the list, ahead of the inclusion-exclusion
        // classes), heaviest first among themselves
        struct E { int slot; double est; };
        std::vector<E> es;
        const double kmid[4] = {12.0, 15.5, 20.0, 30.0};
        for (int kr = 0; kr < 4; ++kr)
            for (int cnt = 1; cnt <= 16; ++cnt)
            {
                const double w = cnt < 16 ? cnt : 24.0;  // the last slot also holds everything wider
                es.push_back({kr * 16 + cnt - 1, kmid[kr] * (w * (w + 1.0) / 2.0 * 7.0 + 20.0) + 40.0 * w + 600.0});
            }
        std::stable_sort(es.begin(), es.end(), [](const E &a, const E &b) { return a.est > b.est; });
        for (size_t i = 0; i < es.size(); ++i)
        {
            rank[(kIeMax + 1) * 64 + es[i].slot] = (unsigned short)i;
            est_of_rank[i] = (float)es[i].est;
        }
    }
    for (size_t i = 0; i < cs.size(); ++i)
    {
        rank[cs[i].kk * 64 + cs[i].cnt - 1] = (unsigned short)(64 + i);
        est_of_rank[64 + i] = (float)cs[i].est;
    }
}

// The cache is sample-major ([t][i][j], like the reference's), update_p works locus-major
// ([t][j][i]): 32 x 32 tiles go through shared memory so that both sides are read / written along
// their fast index (ncu, thread-per-cell version: every 8-byte load of the strided side cost a
// 32-byte sector, 101 us at the benchmark ladder).  One block per (chain, tile of samples, tile of loci).
__host__ __device__ inline unsigned pf_tiles(int N, int L) { return (unsigned)(((N + 31) / 32) * ((L + 31) / 32)); }
struct PfTile
{
    int t, i0, j0;
};
__device__ __forceinline__ PfTile pf_tile_of_block(const View &v)
{
    const int tiles_j = (v.L + 31) / 32, tiles_i = (v.N + 31) / 32;
    int b = blockIdx.x;
    PfTile r;
    r.j0 = (b % tiles_j) * 32;
    b /= tiles_j;
    r.i0 = (b % tiles_i) * 32;
    r.t = b / tiles_i;
    return r;
}

__global__ void __launch_bounds__(256) k_pf_prepare(const View v, const PFlat f)
{
    pdl_enter();
    __shared__ unsigned int s_hist[kSortBins];
    __shared__ double s_T[32][33];
    __shared__ mask_t s_lat[32][33];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int b = tid; b < kSortBins; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    const PfTile tile = pf_tile_of_block(v);
    const int N = v.N, L = v.L, t = tile.t;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int li = ty + 8 * r, i = tile.i0 + li, j = tile.j0 + tx;
        double T = 0.0;  // missing cells: never listed, never evaluated: a zero term on both sides
        mask_t lat = 0;
        if (i < N && j < L)
        {
            const size_t tn = (size_t)t * N + i;
            if (j == 0)
            {
                SampleScal sc;
                sc.log_r = v.log_r[tn];
                sc.log_1mr = v.log_1mr[tn];
                sc.odds = v.odds[tn];
                sc.m = v.m[tn];
                sc.pad = 0;
                f.ss[tn] = sc;
            }
            if (v.missing[(size_t)i * L + j] == 0)
            {
                const size_t cell = tn * L + j;
                lat = v.latent[cell];
                T = v.cellT[cell];
                atomicAdd(&s_hist[work_class_flat(mpopc(lat), v.m[tn])], 1u);
            }
        }
        s_T[li][tx] = T;
        s_lat[li][tx] = lat;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int lj = ty + 8 * r, j = tile.j0 + lj, i = tile.i0 + tx;
        if (i < N && j < L)
        {
            const size_t loc = ((size_t)t * L + j) * N + i;
            const mask_t lat = s_lat[tx][lj];
            f.Tcur[loc] = s_T[tx][lj];
            f.latLM[loc] = lat;  // 0 marks a missing cell (a latent genotype is never empty)
            if (lat == 0) f.Tnew[loc] = 0.0;
        }
    }
    for (int b = tid; b < kSortBins; b += blockDim.x)
        if (s_hist[b]) atomicAdd(&f.hist[b], s_hist[b]);
}

// One block: the exclusive scan of the class counts (the list offsets) and the section bounds.
// Where the one-thread / cooperative boundary goes is decided per launch from a cost model of the
// evaluator (calibrated on its per-block trace, scripts/eval_trace_digest.py; unit = one dependent
// instruction of a resident warp, ~11 ns with eight warps per scheduler):
//   a chunk of 32 one-thread cells of a class costs its warp  est           (and is one latency chain)
//   a cooperative cell costs its warp                         est / 4 + 300
// With the boundary in front of class b (classes are ranked by est, heaviest first) the launch takes
//   T(b) = max( (sum_{c < b} n_c (est_c / 4 + 300) + sum_{c >= b} n_c est_c / 32) / warps ,  est_b )
// -- throughput against the longest one-thread chain -- and the populated b with the smallest T(b)
// wins, capped at MOIRE_COOP_WORK_MAX.  A full benchmark ladder lands where round 1 tuned it by hand
// (~8000: the work per warp is that large); a 5-chain shard of it near 3000; the Namibia shape
// (work per warp ~400, a few hundred cells of 1000..3000) moves down to ~1000, where the fixed 3000
// left 30 us one-thread chains in a launch with 5 us of work (profiles/README.md, round 2).
__global__ void __launch_bounds__(256) k_pf_scan(const PFlat f)
{
    pdl_enter();
    __shared__ unsigned int s_part[9];
    __shared__ float s_work[8], s_lw[9], s_cw[9], s_bestT[8];
    __shared__ int s_bestB[8];
    const int tid = threadIdx.x;
    constexpr int per = kSortBins / 256;  // 3
    unsigned int local = 0, h[per];
    float work = 0.f, lw[per], cw[per], lsum = 0.f, csum = 0.f;
    for (int q = 0; q < per; ++q)
    {
        const int b = tid * per + q;
        h[q] = f.hist[b];
        local += h[q];
        const float est = c_flat_est[b];
        work += (float)h[q] * est;
        const bool ie = b >= kClassCoopBegin;  // the EGF classes (ranks < 64) have their own section
        lw[q] = ie ? (float)h[q] * est * (1.f / 32.f) : 0.f;
        cw[q] = ie ? (float)h[q] * (est * MOIRE_COOP_COST_SLOPE + MOIRE_COOP_COST_FIXED) : 0.f;
        lsum += lw[q];
        csum += cw[q];
    }
    unsigned int x = local;
    float xl = lsum, xc = csum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        const float yl = __shfl_up_sync(0xffffffffu, xl, o), yc = __shfl_up_sync(0xffffffffu, xc, o);
        if ((tid & 31) >= o)
        {
            x += y;
            xl += yl;
            xc += yc;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) work += __shfl_xor_sync(0xffffffffu, work, o);
    if ((tid & 31) == 31)
    {
        s_part[tid >> 5] = x;
        s_lw[tid >> 5] = xl;
        s_cw[tid >> 5] = xc;
    }
    if ((tid & 31) == 0) s_work[tid >> 5] = work;
    __syncthreads();
    if (tid == 0)
    {
        unsigned int run = 0;
        float rl = 0.f, rc = 0.f;
        for (int w = 0; w < 8; ++w)
        {
            const unsigned int c = s_part[w];
            const float cl = s_lw[w], cc = s_cw[w];
            s_part[w] = run;
            s_lw[w] = rl;
            s_cw[w] = rc;
            run += c;
            rl += cl;
            rc += cc;
        }
        s_part[8] = run;
        s_lw[8] = rl;
        s_cw[8] = rc;
    }
    __syncthreads();
    float total_work = 0.f;
    for (int w = 0; w < 8; ++w) total_work += s_work[w];
    const float warps = (float)f.eval_threads * (1.f / 32.f);
    // a cell whose one-thread work exceeds the launch's work per thread would outlast everything else
    // on its own (a small shard of a high-COI ladder: one k = 10 cell with 30 mixture terms is ~10^5
    // dependent instructions, ~1 ms, while the launch's share per thread is a tenth of that), so it
    // ALWAYS goes to a warp: the classes above `share`
    const float share = MOIRE_COOP_WORK_SHARE * total_work / (float)f.eval_threads;
    // -- T(b) of this thread's populated classes
    float bestT = 3.0e38f;
    int bestB = kSortBins;
    {
        float coop_before = s_cw[tid >> 5] + xc - csum, light_before = s_lw[tid >> 5] + xl - lsum;
        const float light_total = s_lw[8];
        for (int q = 0; q < per; ++q)
        {
            const int b = tid * per + q;
            if (b >= kClassCoopBegin && h[q] && c_flat_est[b] <= MOIRE_COOP_WORK_MAX)
            {
                const float T = fmaxf((coop_before + (light_total - light_before)) / warps, c_flat_est[b]);
                if (T < bestT)  // ties: the first (heaviest) class, i.e. fewer cooperative cells
                {
                    bestT = T;
                    bestB = b;
                }
            }
            coop_before += cw[q];
            light_before += lw[q];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const float oT = __shfl_xor_sync(0xffffffffu, bestT, o);
        const int oB = __shfl_xor_sync(0xffffffffu, bestB, o);
        if (oT < bestT || (oT == bestT && oB < bestB))
        {
            bestT = oT;
            bestB = oB;
        }
    }
    if ((tid & 31) == 0)
    {
        s_bestT[tid >> 5] = bestT;
        s_bestB[tid >> 5] = bestB;
    }
    __syncthreads();
    for (int w = 0; w < 8; ++w)
        if (s_bestT[w] < bestT || (s_bestT[w] == bestT && s_bestB[w] < bestB))
        {
            bestT = s_bestT[w];
            bestB = s_bestB[w];
        }
    // bestB: the first class that runs one cell per thread (kSortBins: every populated class is above
    // MOIRE_COOP_WORK_MAX or the list holds EGF cells only)
    if (tid == 0)
    {
        // defaults (no class below the boundary): everything is cooperative
        f.hist[kPfScalars + 1] = s_part[8];
        f.hist[kPfScalars + 2] = s_part[8];
        f.hist[kPfScalars + 3] = s_part[8];
    }
    __syncthreads();
    unsigned int base = s_part[tid >> 5] + x - local;
    for (int q = 0; q < per; ++q)
    {
        const int b = tid * per + q;
        f.hist[kSortBins + b] = base;
        if (b == kClassCoopBegin) f.hist[kPfScalars + 0] = base;
        // ranks >= 64 are sorted by estimate, descending.  [n_egf, coop_end): est >= share and above the
        // boundary: always a warp per cell; [coop_end, mid_end): above the boundary but below the
        // launch's share: a warp per cell while rare, chunks when plentiful
        if (b >= kClassCoopBegin && (c_flat_est[b] < share || b >= bestB) &&
            (b == kClassCoopBegin || !(c_flat_est[b - 1] < share || b - 1 >= bestB)))
            f.hist[kPfScalars + 1] = base;
        if (b >= kClassCoopBegin && b == max(bestB, kClassCoopBegin) && b < kSortBins) f.hist[kPfScalars + 3] = base;
        base += h[q];
    }
}

// Scatter of one block's cells into the sorted list: ranks inside the block come from a
// shared-memory histogram, and the block reserves its range of every class with ONE global atomic
// per class (a global atomic per cell on ~50 hot counters cost 200 us per 800 K cells).
// cell_fn(idx, cls, entry) returns false for cells that are not listed.
#ifndef MOIRE_SCATTER_CPT
#define MOIRE_SCATTER_CPT 4
#endif
constexpr int kScatterCellsPerThread = MOIRE_SCATTER_CPT;
template <class CellFn>
__device__ __forceinline__ void block_scatter(size_t total, CellFn cell_fn, unsigned int *offsets,
                                              ListEntry *list)
{
    __shared__ unsigned int s_cnt[kSortBins];
    const int tid = threadIdx.x;
    for (int b = tid; b < kSortBins; b += 256) s_cnt[b] = 0;
    __syncthreads();
    const size_t first = (size_t)blockIdx.x * 256 * kScatterCellsPerThread;
    int cls[kScatterCellsPerThread];
    unsigned int rank[kScatterCellsPerThread];
    ListEntry entry[kScatterCellsPerThread];
#pragma unroll
    for (int c = 0; c < kScatterCellsPerThread; ++c)
    {
        const size_t idx = first + (size_t)c * 256 + tid;
        cls[c] = -1;
        if (idx < total && cell_fn(idx, cls[c], entry[c])) rank[c] = atomicAdd(&s_cnt[cls[c]], 1u);
        else cls[c] = -1;
    }
    __syncthreads();
    for (int b = tid; b < kSortBins; b += 256)
        if (s_cnt[b]) s_cnt[b] = atomicAdd(&offsets[b], s_cnt[b]);  // count -> base of the block's range
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kScatterCellsPerThread; ++c)
        if (cls[c] >= 0) list[s_cnt[cls[c]] + rank[c]] = entry[c];
}
__host__ __device__ inline unsigned scatter_blocks(size_t total)
{
    return (unsigned)((total + 256 * kScatterCellsPerThread - 1) / (256 * kScatterCellsPerThread));
}

__global__ void __launch_bounds__(256) k_pf_scatter(const View v, const PFlat f)
{
    pdl_enter();
    const size_t total = (size_t)v.T * v.L * v.N;
    auto cell_fn = [&](size_t loc, int &cls, ListEntry &entry) -> bool {
        const mask_t lat = f.latLM[loc];  // locus-major copy of k_pf_prepare: coalesced
        if (lat == 0) return false;                // missing
        const unsigned tl = (unsigned)(loc / v.N), i = (unsigned)(loc - (size_t)tl * v.N);
        const unsigned t = tl / v.L, j = tl - t * v.L;
        cls = work_class_flat(mpopc(lat), v.m[t * v.N + i]);
        entry = make_entry(lat, (t << 16) | j, i);
        return true;
    };
    block_scatter(total, cell_fn, f.hist + kSortBins, f.list);
}

// The SALT proposal keeps the alleles of a locus across the lanes of one warp, kAPL per lane (two with
// one-word masks, as in round 1; four with two words).  x[slot], slot < kAPL, without dynamic indexing.
constexpr int kAPL = kMaskBits / 32;
__device__ __forceinline__ double lane_slot(const double (&x)[kAPL], int slot)
{
    double r = x[0];
#pragma unroll
    for (int s = 1; s < kAPL; ++s) r = slot == s ? x[s] : r;
    return r;
}

// ---- once per proposal round: decide the previous proposal, draw the next -----------------------
// One block per (chain, locus); the block size follows the shape (engine.cu: about as many threads
// as the device holds, spread over T * L loci: 64 at the benchmark ladder, 1024 for a few loci with
// thousands of samples) and the registers are capped at 64, so every locus of a round is resident at
// once -- the kernel is one latency chain per locus (sum, decision, commit, proposal), not
// throughput.  Tcur / Tnew hold 0 at missing cells (k_pf_prepare), so the sum needs no mask.
__global__ void __launch_bounds__(1024, 1) k_pf_round(const View v, const PFlat f, const int iteration_arg,
                                                    const int rep)
{
    pdl_enter();
    __shared__ double s_part[32];
    __shared__ int s_accept;
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
    const int N = v.N, L = v.L, AP = v.AP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, B = blockDim.x;
    const int tl = blockIdx.x;
    const int t = tl / L, j = tl - t * L;
    const int A = v.nall[j];
    const size_t prow = (size_t)tl * AP;
    double *st = f.round + (size_t)tl * 4;
    unsigned flags = rep == 0 ? 0u : (unsigned)st[3];
    const double *tn = pf_terms(f, flags, true) + (size_t)tl * N, *tc = pf_terms(f, flags, false) + (size_t)tl * N;
    // the scalars of the decision, in flight together with the sum's loads (thread 0)
    double e_adj = 0.0, e_logu = 0.0, e_temp = 0.0, e_sd = 0.0;
    int e_idx = 0, e_att = 0, e_acc = 0;
    if ((flags & kPfActive) && tid == 0)
    {
        e_idx = (int)st[2];
        e_adj = st[0];
        e_logu = st[1];
        e_temp = v.temp[t];
        e_att = v.p_att[prow + e_idx];
        e_acc = v.p_acc[prow + e_idx];
        e_sd = v.p_sd[prow + e_idx];
    }
    // the draws of proposal `rep` depend on nothing the decision changes: before the sum, so that
    // their arithmetic runs while its loads are in flight
    const bool will_propose = rep < A && !(flags & kPfBroken) && warp == 0;
    int n_idx = 0;
    double n_z = 0.0, n_logu = 0.0;
    if (will_propose)
    {
        Stream rng;
        rng.open(v.seed, (uint32_t)iteration, KIND_P, (unsigned)(v.chain_offset + t), (uint32_t)j,
                 (uint32_t)rep);
        n_idx = (int)mulhi32(rng.next_u32(), (uint32_t)A);
        n_z = next_normal(rng);
        n_logu = cold_log(rng.next_uniform());
    }
    // cache mode: every round sorts its own list (k_pc_sweep counts into these bins)
    if (f.Wcur && blockIdx.x == 0)
        for (int b = tid; b < kSortBins; b += B) f.hist[b] = 0;

    if (flags & kPfActive)
    {
        // -- the change of the locus' log-likelihood under proposal rep - 1: thread-strided partial
        // sums (four loads in flight per array), the shuffle tree, then the warp sums in warp order
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        int i = tid;
        for (; i + 3 * B < N; i += 4 * B)
        {
            const double a0 = tn[i], a1 = tn[i + B], a2 = tn[i + 2 * B], a3 = tn[i + 3 * B];
            const double b0 = tc[i], b1 = tc[i + B], b2 = tc[i + 2 * B], b3 = tc[i + 3 * B];
            d0 += a0 - b0;
            d1 += a1 - b1;
            d2 += a2 - b2;
            d3 += a3 - b3;
        }
        for (; i < N; i += B) d0 += tn[i] - tc[i];
        const double dw = warp_sum((d0 + d1) + (d2 + d3));
        if (lane == 0) s_part[warp] = dw;
        __syncthreads();
        if (tid == 0)
        {
            double d = 0.0;
            for (int w = 0; w < (B >> 5); ++w) d += s_part[w];
            const int idx = e_idx;
            const double dpost = e_temp * d;  // the priors do not depend on p
            const double ratio = dpost + e_adj;
            const int accept = !(isnan(dpost) || e_logu > ratio);
            if (accept) v.p_acc[prow + idx] = ++e_acc;
            if (iteration < v.burnin && e_att > 15)
            {
                // src/chain.cpp:555-562
                const double att = (double)e_att;
                const double rate = ((double)e_acc + 1.0) / (att + 1.0);
                const double sd = e_sd + (rate - .23) / sqrt(att + 1.0);
                v.p_sd[prow + idx] = sd > .01 ? sd : .01;
            }
            atomicAdd(v.evals + KIND_P, (unsigned long long)f.nonmiss[j]);
            s_accept = accept;
        }
        __syncthreads();
        if (s_accept)
            for (int a = tid; a < A; a += B) v.p[prow + a] = f.pprop[prow + a];
        __syncthreads();
    }
    {
        const bool accepted = (flags & kPfActive) && s_accept;
        flags &= ~(kPfActive | kPfAccepted | 0xff00u);
        if (accepted) flags ^= kPfFlip;  // the proposal's terms are the locus' terms now: no copy
        // cache mode: what k_pc_sweep has to commit before the next proposal is evaluated
        if (f.Wcur && accepted) flags |= kPfAccepted | ((unsigned)e_idx << 8);  // thread 0's word is the one stored
    }
    if (rep >= A || (flags & kPfBroken))
    {
        if (tid == 0) st[3] = (double)flags;
        return;
    }
    if (warp != 0) return;
    // -- proposal `rep`, lanes across alleles (same arithmetic as k_update_p)
    {
        const int idx = n_idx;
        const double z = n_z, logu = n_logu;
        double lg[kAPL], lp[kAPL], lq[kAPL];
        bool rest[kAPL];
#pragma unroll
        for (int s = 0; s < kAPL; ++s)
        {
            const int a = lane + 32 * s;
            const bool has = a < A;
            lg[s] = has ? logit_d(v.p[prow + a]) : 0.0;
            lp[s] = lq[s] = 0.0;
            if (has) log_pq_d(lg[s], lp[s], lq[s]);
            rest[s] = has && a != idx;
        }
        const double lg_sel = lane_slot(lg, idx >> 5);
        const double logit_curr = __shfl_sync(0xffffffffu, lg_sel, idx & 31);
        const double logit_prop = logit_curr + v.p_sd[prow + idx] * z;
        double clp, clq, plp, plq;
        log_pq_d(logit_curr, clp, clq);
        log_pq_d(logit_prop, plp, plq);
        double bv = -INFINITY;
        int ba = 1 << 30;
#pragma unroll
        for (int s = 0; s < kAPL; ++s)
            if (rest[s] && (lg[s] > bv)) { bv = lg[s]; ba = lane + 32 * s; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oa = __shfl_xor_sync(0xffffffffu, ba, o);
            if (ov > bv || (ov == bv && oa < ba)) { bv = ov; ba = oa; }
        }
        const double lp_sel = lane_slot(lp, ba >> 5);
        const double lq_sel = lane_slot(lq, ba >> 5);
        const double lp1 = __shfl_sync(0xffffffffu, lp_sel, ba & 31);
        const double lq1 = __shfl_sync(0xffffffffu, lq_sel, ba & 31);
        double cum = 0.0;
#pragma unroll
        for (int s = 0; s < kAPL; ++s)
            if (rest[s] && (lane + 32 * s) != ba) cum += (bv < 0.0) ? cold_exp(lp[s] - lp1) : cold_exp(lp[s]);
        cum = warp_sum(cum);
        const double lsum = (bv < 0.0) ? lp1 + cold_log1p(cum) : cold_log1p(-cold_exp(lq1) + cum);
        const double ls = plq - lsum;
        bool below = false;
#pragma unroll
        for (int s = 0; s < kAPL; ++s)
        {
            const int a = lane + 32 * s;
            if (a < A)
            {
                const double np = expit_d(a == idx ? logit_prop : logit_scale_d(lg[s], ls));
                f.pprop[prow + a] = np;
                below |= np < 1e-12;
            }
        }
        below = __any_sync(0xffffffffu, below);
        if (lane == 0)
        {
            ++v.p_att[prow + idx];
            st[0] = (clp - plp) + (double)(A - 1) * (clq - plq);
            st[1] = logu;
            st[2] = (double)idx;
            st[3] = (double)(below ? (flags | kPfBroken) : (flags | kPfActive));  // :503-518
        }
    }
}

// ---- cache mode: a round evaluates from scratch only the cells that hold the proposed allele -----
// SALT moves ONE frequency and rescales all others by a common factor (src/mcmc_utils.h:252-300),
// so for a cell whose latent genotype does not contain the proposed allele the normalised
// frequencies q_a = p_a / sum -- and with them every PAM -- are unchanged (to the last bits of the
// rescaling); only the sum enters the mixture anew.  With A_j alleles per locus and k of them in a
// cell that is a fraction 1 - k / A_j of the cells of every round.  In cache mode update_p keeps the
// PAM vector of every cell (Wcur: accepted state, Wnew: the proposal's, for the cells evaluated from
// scratch; committed by the next round's k_pc_sweep when k_pf_round accepted the proposal), filled by one evaluation of the current
// state at the start of the update, and each round is
//   k_pf_round   decide / commit the previous proposal, draw the next          (as before)
//   k_pc_sweep   every cell of an active locus, in memory order: cells without the proposed allele
//                get their term from the cached PAMs -- the fresh sum of the proposed frequencies
//                through the SAME mixture function the evaluators end in -- the others are counted
//                by cost class
//   k_pf_scan, k_pc_scatter   the round's own sorted list of from-scratch cells
//   k_flat_eval<true>         evaluates that list, terms to Tnew, PAM vectors to Wnew
// What changes numerically: a cached PAM was computed from q of an earlier p of the same update
// (rescaled since, so equal to rounding); the from-scratch value (k_eval_cells) stays the truth the
// tests bound the difference against (tests/test_gpu_shapes.py).
__device__ __forceinline__ bool pc_fresh(mask_t S, int idx, int k, int m)
{
    return mtest(S, idx) || k > m;
}

// relatedness_mixture for at most eight mixture terms with the PAMs in registers: the loop of
// small_cell_tail (likelihood.cuh), i.e. the same operations in the same order as the generic
// function's fast branch; its rare fallbacks go through the generic function.
__device__ __forceinline__ double mixture_regs8(const double (&pam)[8], int m, int cnt, double total,
                                                const Rel rel, const double *lgam)
{
    const double lead = __dmul_rn((double)(m - 1), rel.log_1mr);
    if (cnt == 1)
        return __dadd_rn(lead, log_total_pow_times(total, m, __dsub_rn(1.0, clamp_pam(pam[0]))));
    const double qr = __ddiv_rn(rel.odds, total);
    const int e2 = ((__double2hiint(qr) >> 20) & 0x7ff) - 1022;
    if ((cnt - 1) * max(e2, 0) > 850)
    {
        double w[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) w[t] = pam[t];
        return relatedness_mixture(w, 1, m, cnt, total, rel, lgam, nullptr);
    }
    const double *cb = g_binom + (m - 1) * (kMaxCoi + 1);
    double qi = 1.0, S = 0.0;
#pragma unroll
    for (int t = 7; t >= 0; --t)
        if (t < cnt)
        {
            S = __dadd_rn(S, __dmul_rn(__dmul_rn(cb[cnt - 1 - t], qi), __dsub_rn(1.0, clamp_pam(pam[t]))));
            qi = __dmul_rn(qi, qr);
        }
    return __dadd_rn(lead, log_total_pow_times(total, m, S));
}

__global__ void __launch_bounds__(256) k_pc_sweep(const View v, const PFlat f)
{
    pdl_enter();
    __shared__ unsigned int s_hist[kSortBins];
    for (int b = threadIdx.x; b < kSortBins; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    const size_t total = (size_t)v.T * v.L * v.N;
    const size_t loc = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (loc < total)
    {
        const unsigned tl = (unsigned)(loc / v.N), i = (unsigned)(loc - (size_t)tl * v.N);
        const double *st = f.round + (size_t)tl * 4;
        const mask_t S = f.latLM[loc];
        const unsigned fl = (unsigned)st[3];
        if ((fl & kPfActive) && S)
        {
            const unsigned t = tl / v.L;
            const size_t tn = (size_t)t * v.N + i;
            const int idx = (int)st[2], k = mpopc(S);
            const SampleScal sc = f.ss[tn];
            const int m = sc.m;
            const bool rel_on = v.allow_rel != 0;
            const int nw = rel_on ? (k >= 2 ? m - k + 1 : 0) : 1;  // cached PAMs of the cell
            // the previous proposal of this locus was accepted and this cell held its allele: the
            // PAM vector the evaluator left in Wnew is the accepted state's now
            const bool commit = (fl & kPfAccepted) && mtest(S, (int)((fl >> 8) & 0xffu)) && k <= m;
            const bool fresh = pc_fresh(S, idx, k, m);
            const double *wsrc = (commit ? f.Wnew : f.Wcur) + loc;
            if (fresh)
            {
                atomicAdd(&s_hist[work_class_flat(k, m)], 1u);
                if (commit)
                    for (int q = 0; q < nw; ++q) f.Wcur[(size_t)q * f.wstride + loc] = wsrc[(size_t)q * f.wstride];
            }
            else
            {
                // the sum of the proposed frequencies of the latent alleles, ascending allele order
                // (the evaluators' order)
                const double *prow = f.pprop + (size_t)tl * v.AP;
                double tot = 0.0;
                mask_t rest = S;
                while (rest)
                {
                    const int al = mffs(rest) - 1;
                    rest &= rest - 1;
                    tot = __dadd_rn(tot, prow[al]);
                }
                double T;
                if (!rel_on)
                {
                    const double pam = wsrc[0];
                    if (commit) f.Wcur[loc] = pam;
                    T = norel_term(pam, tot, m);
                }
                else
                {
                    const Rel rel{sc.log_r, sc.log_1mr, sc.odds};
                    if (k == 1)
                        T = relatedness_mixture(nullptr, 0, m, m, tot, rel, v.lgam, nullptr);
                    else if (nw <= 8)
                    {
                        double w[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) w[q] = q < nw ? wsrc[(size_t)q * f.wstride] : 0.0;
                        if (commit)
                        {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                if (q < nw) f.Wcur[(size_t)q * f.wstride + loc] = w[q];
                        }
                        T = mixture_regs8(w, m, nw, tot, rel, v.lgam);
                    }
                    else
                    {
                        double w[kMaxCoi];
                        for (int q = 0; q < nw; ++q) w[q] = wsrc[(size_t)q * f.wstride];
                        if (commit)
                            for (int q = 0; q < nw; ++q) f.Wcur[(size_t)q * f.wstride + loc] = w[q];
                        T = relatedness_mixture(w, 1, m, nw, tot, rel, v.lgam, nullptr);
                    }
                }
                pf_terms(f, fl, true)[loc] = T;
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSortBins; b += blockDim.x)
        if (s_hist[b]) atomicAdd(&f.hist[b], s_hist[b]);
}

__global__ void __launch_bounds__(256) k_pc_scatter(const View v, const PFlat f)
{
    pdl_enter();
    const size_t total = (size_t)v.T * v.L * v.N;
    auto cell_fn = [&](size_t loc, int &cls, ListEntry &entry) -> bool {
        const unsigned tl = (unsigned)(loc / v.N), i = (unsigned)(loc - (size_t)tl * v.N);
        const double *st = f.round + (size_t)tl * 4;
        const mask_t S = f.latLM[loc];
        if (((unsigned)st[3] & kPfActive) == 0 || S == 0) return false;
        const unsigned t = tl / v.L, j = tl - t * v.L;
        const int m = v.m[(size_t)t * v.N + i], k = mpopc(S);
        if (!pc_fresh(S, (int)st[2], k, m)) return false;
        cls = work_class_flat(k, m);
        entry = make_entry(S, (t << 16) | j, i);
        return true;
    };
    block_scatter(total, cell_fn, f.hist + kSortBins, f.list);
}

// ---- the flat evaluator: every listed cell, from the sorted list ---------------------------------
// Sections of the list: [0, n_egf) k > 10, [n_egf, coop_end) the classes above the launch's
// cooperative threshold (k_pf_scan), the rest light.
// A task is a few consecutive units of the list -- cooperative cells or chunks of 32 one-thread
// cells of (nearly) one cost class.  The grid is persistent (a few blocks per SM); blocks draw
// tasks from a device counter, heaviest first, so the section sizes never travel to the host.
// Heavy sections go one cell per thread too where they hold a large share of all cells (high-COI
// shapes); otherwise a chunk of 1023-subset cells would outlast everything else.
struct FlatSrc
{
    const ListEntry *list;                 // {latent genotype (two words), (t << 16) | j, i}: the
                                        //  genotype travels in the list, so a unit's loads are
                                        //  list -> frequencies / sample scalars only
    const unsigned int *bounds;         // [4] n_egf, coop_end, nwork, mid_end
    unsigned int *counter;              // task counter, zero at launch
    const double *p;                    // row t * L + j, p_stride doubles apart
    int p_stride;
    const SampleScal *ss;               // [t * N + i]
    const double *round_flags;          // null, or [(t * L + j) * 4 + 3] & kPfActive says "evaluate"
    double *out;                        // locus-major [t][j][i] (p rounds) or sample-major [t][i][j]
    double *out_flip;                   // p rounds: where the terms of a locus whose arrays are swapped go
                                        // (kPfFlip in its round flags); null otherwise
    int out_locus_major;
    int binom_rows;                     // rows of g_binom staged in shared memory (= max_coi)
    double *pam_out;                    // CACHE evaluators: PAM vectors, [t][pam_stride] + locus-major cell
    size_t pam_stride;
};

#ifndef MOIRE_PF_THREADS
#define MOIRE_PF_THREADS 256
#endif
constexpr int kEvalThreads = MOIRE_PF_THREADS, kEvalWarps = kEvalThreads / 32;
#ifndef MOIRE_PF_LIGHT_UPB
#define MOIRE_PF_LIGHT_UPB kEvalWarps
#endif
#ifndef MOIRE_PF_MIN_BLOCKS
#define MOIRE_PF_MIN_BLOCKS 4
#endif
// Units per task: one per warp for cooperative cells and for chunks of heavy cells (each is long:
// up to 1023 subsets), MOIRE_PF_LIGHT_UPB / 8 light chunks per warp.
__host__ __device__ inline int pf_units_per_task(bool light) { return light ? MOIRE_PF_LIGHT_UPB : kEvalWarps; }
__host__ __device__ inline int pf_tasks(int cells, bool chunked, bool light)
{
    const int units = chunked ? (cells + 31) / 32 : cells;
    const int upb = pf_units_per_task(light);
    return (units + upb - 1) / upb;
}
__host__ __device__ inline size_t pf_eval_smem_doubles(int binom_rows)
{
    return 128 + (size_t)binom_rows * kBinomCols + scratch_doubles(kEvalThreads);
}

// One unit of the list: a cooperative cell (all lanes of the warp on one cell) or a chunk of 32
// one-thread cells starting at lo + 32 * unit.  CACHE: the evaluator also leaves every cell's PAM
// vector in update_p's cache (likelihood.cuh: PamOut), at the cell's locus-major index.
template <bool CACHE, bool STASH>
__device__ __forceinline__ void flat_unit(const View &v, const FlatSrc &f, const int lo, const int hi,
                                          const bool coop, const bool egf, const int unit, double *s_scr,
                                          const double *s_lgam, const double *s_binom, const bool allow_rel,
                                          const bool locus_major)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int ST = kEvalThreads + 1;
    const unsigned N = (unsigned)v.N, L = (unsigned)v.L;
    // STASH: where a one-thread cell's term goes is parked in shared memory for the length of the
    // evaluation.  Held in registers it was spilled to local memory around every chunk (the
    // evaluators need all 64), and the reload at the end of a chunk found its line gone from L1 half
    // the time -- an extra trip to L2 per chunk, and half of the kernel's L2 traffic (ncu,
    // profiles/README.md): -3 % of a step at the benchmark shapes.  The register allocation it leaves
    // is worse for the long single-thread walks of heavy cells (a high-COI shard of 2 chains: 20 -> 30
    // ms per step), so the engine picks the variant from the data (engine.cu: eval_stash).
    __shared__ double *s_dst[STASH ? kEvalThreads : 1];
    if (coop)
    {
        const int w = lo + unit;
        if (w >= hi) return;
        const ListEntry e = f.list[w];
        const unsigned t = e.z >> 16, j = e.z & 0xffffu, tl = t * L + j, tn = t * N + e.w;
        const unsigned fl = f.round_flags ? (unsigned)f.round_flags[(size_t)tl * 4 + 3] : kPfActive;
        if ((fl & kPfActive) == 0) return;
        double *const outp = (f.out_flip && (fl & kPfFlip)) ? f.out_flip : f.out;
        const mask_t lat = entry_lat(e);
        const SampleScal sc = f.ss[tn];
        const Rel rel{sc.log_r, sc.log_1mr, sc.odds};
        const int m = sc.m, k = mpopc(lat);
        const double *prow = f.p + (size_t)tl * f.p_stride;
        const PamOut po{CACHE ? f.pam_out + ((size_t)tl * N + e.w) : nullptr, f.pam_stride};
        double T;
        if (k > m)
            T = NAN;
        else if (egf)
            T = transmission_ll_egf_coop<CACHE>(lat, prow, m, rel, allow_rel, s_scr + warp * 32, ST, s_lgam, s_binom, po);
        else
            T = transmission_ll_coop<CACHE>(lat, prow, m, rel, allow_rel, s_scr + warp * 32, ST, s_lgam, s_binom, po);
        if (lane == 0) outp[locus_major ? (size_t)tl * N + e.w : (size_t)tn * L + j] = T;
        __syncwarp();
        return;
    }
    const int w0 = lo + unit * 32;
    if (w0 >= hi) return;
    const int w = w0 + lane;
    bool go = w < hi;
    ListEntry e = idle_entry();
    if (go) e = f.list[w];
#ifndef MOIRE_PF_PREFETCH
#define MOIRE_PF_PREFETCH 1
#endif
#if MOIRE_PF_PREFETCH
    // the entries this block is likely to meet next (tasks are handed out in list order, a grid's worth
    // of tasks ahead is about one round of claims away): pulled into L2 now, so that the load at the
    // head of that chunk -- the longest single wait of a chunk -- does not go to DRAM
    {
        const size_t ahead = (size_t)w + (size_t)gridDim.x * (kEvalThreads * MOIRE_PF_PREFETCH);
        if (ahead < (size_t)hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(f.list + ahead));
    }
#endif
    // -- class-uniform chunk of small cells: the register path (likelihood.cuh)
    {
        const unsigned in = __ballot_sync(0xffffffffu, go);
        const int k = __popc(e.x);
        const bool small = allow_rel && in != 0 &&
                           __all_sync(0xffffffffu, !go || (entry_low32(e) && k <= kSmallK && k >= 1)) &&
                           __reduce_max_sync(0xffffffffu, go ? k : 0) ==
                               __reduce_min_sync(0xffffffffu, go ? k : kSmallK + 1);
        if (small)
        {
            const int K = __shfl_sync(0xffffffffu, k, __ffs(in) - 1);
            const unsigned tl = (e.z >> 16) * L + (e.z & 0xffffu), tn = (e.z >> 16) * N + e.w;
            const double flagword = f.round_flags ? f.round_flags[(size_t)tl * 4 + 3] : (double)kPfActive;
            const SampleScal sc = f.ss[tn];
            const double *prow = f.p + (size_t)tl * f.p_stride;
            const int cnt = sc.m - K + 1;
            const int cnt_hi = __reduce_max_sync(0xffffffffu, go ? cnt : 0);
            const int cnt_lo = __reduce_min_sync(0xffffffffu, go ? cnt : 1);
            if (cnt_hi <= kSmallCnt && cnt_lo >= 1)
            {
                if (go && ((unsigned)flagword & kPfActive))
                {
                    const Rel rel{sc.log_r, sc.log_1mr, sc.odds};
                    const PamOut po{CACHE ? f.pam_out + ((size_t)tl * N + e.w) : nullptr, f.pam_stride};
                    if constexpr (STASH)
                    {
                        double *const outp0 = (f.out_flip && ((unsigned)flagword & kPfFlip)) ? f.out_flip : f.out;
                        s_dst[tid] = outp0 + (locus_major ? (size_t)tl * N + e.w : (size_t)tn * L + (e.z & 0xffffu));
                    }
                    double T;
                    switch (K)
                    {
                        case 1: T = small_cell<1, CACHE>(e.x, prow, sc.m, cnt_hi, rel, s_scr + tid, ST, s_lgam, s_binom, po); break;
                        case 2: T = small_cell<2, CACHE>(e.x, prow, sc.m, cnt_hi, rel, s_scr + tid, ST, s_lgam, s_binom, po); break;
                        case 3: T = small_cell<3, CACHE>(e.x, prow, sc.m, cnt_hi, rel, s_scr + tid, ST, s_lgam, s_binom, po); break;
                        case 4: T = small_cell<4, CACHE>(e.x, prow, sc.m, cnt_hi, rel, s_scr + tid, ST, s_lgam, s_binom, po); break;
                        case 5: T = small_cell<5, CACHE>(e.x, prow, sc.m, cnt_hi, rel, s_scr + tid, ST, s_lgam, s_binom, po); break;
                        default: T = small_cell<6, CACHE>(e.x, prow, sc.m, cnt_hi, rel, s_scr + tid, ST, s_lgam, s_binom, po); break;
                    }
                    if constexpr (STASH)
                        *(*(double *volatile *)(s_dst + tid)) = T;
                    else
                    {
                        double *const outp = (f.out_flip && ((unsigned)flagword & kPfFlip)) ? f.out_flip : f.out;
                        outp[locus_major ? (size_t)tl * N + e.w : (size_t)tn * L + (e.z & 0xffffu)] = T;
                    }
                }
                __syncwarp();
                return;
            }
        }
    }
    // everything the cell needs hangs off the list entry: issue it all at once -- the
    // round flag, the sample's scalars, the gathered frequencies -- and only then look
    // at the flag (two trips to L2 per cell instead of 3 + k)
    const unsigned tl = (e.z >> 16) * L + (e.z & 0xffffu), tn = (e.z >> 16) * N + e.w;
    const mask_t lat = entry_lat(e);
    const double *prow = f.p + (size_t)tl * f.p_stride;
    const double flagword = f.round_flags ? f.round_flags[(size_t)tl * 4 + 3] : (double)kPfActive;
    const SampleScal sc = f.ss[tn];
    double pv[kGather];
    gather_freqs(lat, prow, pv);
    go = go && ((unsigned)flagword & kPfActive);
    const unsigned callers = __ballot_sync(0xffffffffu, go);
    if (go)
    {
        const Rel rel{sc.log_r, sc.log_1mr, sc.odds};
        const PamOut po{CACHE ? f.pam_out + ((size_t)tl * N + e.w) : nullptr, f.pam_stride};
        double *const outp = (f.out_flip && ((unsigned)flagword & kPfFlip)) ? f.out_flip : f.out;
        if constexpr (STASH)
        {
            s_dst[tid] = outp + (locus_major ? (size_t)tl * N + e.w : (size_t)tn * L + (e.z & 0xffffu));
            const double T = transmission_ll<CACHE>(lat, prow, pv, sc.m, rel, allow_rel, s_scr + tid, ST, s_lgam, s_binom, callers, po);
            *(*(double *volatile *)(s_dst + tid)) = T;
        }
        else
            outp[locus_major ? (size_t)tl * N + e.w : (size_t)tn * L + (e.z & 0xffffu)] =
                transmission_ll<CACHE>(lat, prow, pv, sc.m, rel, allow_rel, s_scr + tid, ST, s_lgam, s_binom, callers, po);
    }
    __syncwarp();
}

// Sections of the list, heaviest first (k_pf_scan): [0, n_egf) k > 10; [n_egf, coop_end) cells that
// would outlast the launch on one thread: always one warp per cell; [coop_end, mid_end) heavy
// classes: a warp per cell while they are rare, chunks of 32 when they hold a large share of all
// cells (high-COI shapes); [mid_end, nwork) light cells in chunks.
template <bool CACHE, bool STASH>
__global__ void __launch_bounds__(kEvalThreads, MOIRE_PF_MIN_BLOCKS) k_flat_eval(const View v, const FlatSrc f)
{
    extern __shared__ double smem[];
    double *s_lgam = smem;
    double *s_scr = smem + 128;                                   // fixed offsets for the hot arrays
    double *s_binom = s_scr + scratch_doubles(kEvalThreads);
    const int tid = threadIdx.x, warp = tid >> 5;
    // the tables staged here are constants of the engine (written once at creation), so the staging
    // runs while the kernel in front of this one is still finishing (pdl_enter, state.cuh); everything
    // below the wait may depend on it
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    for (int i = tid; i < 128; i += kEvalThreads) s_lgam[i] = v.lgam[i];
    for (int i = tid; i < f.binom_rows * kBinomCols; i += kEvalThreads)
        s_binom[i] = g_binom[(i / kBinomCols) * (kMaxCoi + 1) + (i % kBinomCols)];
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    const bool allow_rel = v.allow_rel != 0;
    const bool locus_major = f.out_locus_major != 0;
    const int n_egf = (int)f.bounds[0], coop_end = (int)f.bounds[1], nwork = (int)f.bounds[2],
              mid_end = (int)f.bounds[3];
    // a section is cut into chunks of 32 one-thread cells once it holds enough cells for the chunks
    // to be class-uniform and to spread over the grid; a handful of heavy cells get a warp each
    const size_t enough = (size_t)gridDim.x * kEvalThreads / 8;
    const bool chunked_egf = (size_t)n_egf >= enough;
    const bool chunked_mid = (size_t)(mid_end - coop_end) >= enough;
    __shared__ int s_task;
#ifdef MOIRE_EVAL_TRACE
    // tuning aid: per block {entry, after the prologue, then per task: claimed at, task id}
    unsigned long long *tr = (v.trace && blockIdx.x < 1024) ? v.trace + (size_t)blockIdx.x * 256 : nullptr;
    int ntr = 2;
    auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (tr && tid == 0) tr[0] = now();
#endif
    const int t_egf = pf_tasks(n_egf, chunked_egf, false);
    const int t_strag = pf_tasks(coop_end - n_egf, false, false);
    const int t_mid = pf_tasks(mid_end - coop_end, chunked_mid, false);
    const int tasks = t_egf + t_strag + t_mid + pf_tasks(nwork - mid_end, true, true);
    for (;;)
    {
        __syncthreads();
#ifdef MOIRE_EVAL_TRACE
        if (tr && tid == 0 && ntr + 3 <= 256) tr[ntr++] = now();  // the block is done with its previous task
#endif
        if (tid == 0) s_task = (int)atomicAdd(f.counter, 1u);
        __syncthreads();
        int b = s_task;
#ifdef MOIRE_EVAL_TRACE
        if (tr && tid == 0 && ntr + 2 <= 256)
        {
            tr[ntr++] = now();  // claimed
            tr[ntr++] = (unsigned long long)b;
            tr[1] = (unsigned long long)ntr;
        }
#endif
        if (b >= tasks) break;
        // Claim order: the cooperative tasks (heaviest first) and the one-thread tasks (heaviest
        // first) ALTERNATE while both last.  The heaviest one-thread chunks are the longest single
        // chains of the launch (the trace of a 5-chain shard: 37 us, started 16 us late behind 660
        // cooperative tasks -- the launch ended when they did), so they start with the first wave.
        // (Not when a heavy section runs as one-thread chunks itself -- high-COI shapes: those chunks
        // are longer than anything in the light section and must not wait: c5 lost 4 % with it.)
        const int t_heavy = t_egf + t_strag + t_mid;
        const int t_pair = (chunked_egf || chunked_mid) ? 0 : min(t_heavy, tasks - t_heavy);
        if (t_pair == 0)
            ;
        else if (b < 2 * t_pair)
            b = (b & 1) ? t_heavy + (b >> 1) : (b >> 1);
        else if (t_heavy > t_pair)
            b -= t_pair;  // the one-thread tasks are used up: b - 2 t_pair cooperative tasks beyond t_pair
        // else: the cooperative tasks are used up and b already is the task's own index
        int lo, hi;
        bool coop, egf = false, light = false;
        if (b < t_egf)
        {
            lo = 0;
            hi = n_egf;
            coop = !chunked_egf;
            egf = true;
        }
        else if (b < t_egf + t_strag)
        {
            b -= t_egf;
            lo = n_egf;
            hi = coop_end;
            coop = true;
        }
        else if (b < t_heavy)
        {
            b -= t_egf + t_strag;
            lo = coop_end;
            hi = mid_end;
            coop = !chunked_mid;
        }
        else
        {
            b -= t_heavy;
            lo = mid_end;
            hi = nwork;
            coop = false;
            light = true;
        }
        const int upb = pf_units_per_task(light);
        for (int u = warp; u < upb; u += kEvalWarps)
            flat_unit<CACHE, STASH>(v, f, lo, hi, coop, egf, b * upb + u, s_scr, s_lgam, s_binom, allow_rel, locus_major);
    }
}

// ---- once per update: the accepted terms back into the sample-major cache -----------------------
__global__ void __launch_bounds__(256) k_pf_finish(const View v, const PFlat f)
{
    pdl_enter();
    __shared__ double s_T[32][33];
    __shared__ unsigned char s_in[32][33];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const PfTile tile = pf_tile_of_block(v);
    const int N = v.N, L = v.L, t = tile.t;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int lj = ty + 8 * r, j = tile.j0 + lj, i = tile.i0 + tx;
        bool in = false;
        double T = 0.0;
        if (i < N && j < L)
        {
            const size_t loc = ((size_t)t * L + j) * N + i;
            in = f.latLM[loc] != 0;
            T = pf_terms(f, (unsigned)f.round[((size_t)t * L + j) * 4 + 3], false)[loc];
        }
        s_T[lj][tx] = T;
        s_in[lj][tx] = in;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int li = ty + 8 * r, i = tile.i0 + li, j = tile.j0 + tx;
        if (i < N && j < L && s_in[tx][li]) v.cellT[((size_t)t * N + i) * L + j] = s_T[tx][li];
    }
}
}  // namespace moire
