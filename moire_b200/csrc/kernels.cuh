// The Metropolis kernels of the engine (sm_100a).
//
// Parallel decomposition: given (p, mean_coi) the per-sample updates of the
// reference are conditionally independent across samples, and given the sample state the
// allele-frequency updates are independent across loci.  So
//   * the six per-sample updates run as ONE templated kernel, one thread block per
//     (chain, tile of S consecutive samples), threads across the S*L contiguous cells of the
//     tile (coalesced), one warp owning each sample's accept decision (shuffle reductions);
//   * update_p runs one block per (chain, locus), the A_j SALT proposals sequential inside the
//     block, threads across samples, latent genotypes and likelihood terms staged in shared
//     memory for all A_j proposals;
//   * update_mean_coi and the per-chain reductions run one block per chain.
// The Metropolis ratio uses the change of the affected cells only, summed in fp64 in a fixed
// order (the reference re-reduces all N*L cells in fp32 per proposal, src/chain.cpp:978-1026;
// the two are the same number up to the reference's rounding noise).
#pragma once
#include <cfloat>

#include "likelihood.cuh"
#include "state.cuh"

namespace moire
{
#ifndef MOIRE_MIN_BLOCKS_P
#define MOIRE_MIN_BLOCKS_P 3
#endif
#ifndef MOIRE_MIN_BLOCKS_S
#define MOIRE_MIN_BLOCKS_S 2
#endif
#ifndef MOIRE_MERGE_K
#define MOIRE_MERGE_K 1
#endif
#ifndef MOIRE_SAMPLE_THREADS
#define MOIRE_SAMPLE_THREADS 256
#endif
#ifndef MOIRE_P_THREADS
#define MOIRE_P_THREADS 256
#endif
constexpr int kThreads = MOIRE_SAMPLE_THREADS;
constexpr int kMaxTileSamples = 32;
#ifndef MOIRE_TILE_CELLS
#define MOIRE_TILE_CELLS 1024
#endif
constexpr int kMaxTileCells = MOIRE_TILE_CELLS;  // cells of one (chain, sample-tile) block
constexpr int kSortBins = 12 * 64;   // (min(k, 11), m - k + 1) work classes

struct Proposal
{
    int valid, adapt, m_new, accept;
    double r_new, en_new, ep_new, log_r, log_1mr, odds;
    double adj;
    double pr_en_new, pr_ep_new, pr_coi_new, pr_r_new, dprior;
    double logu;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum; every thread gets the total.  s_red: >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double *s_red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double tot = 0.0;
        for (int i = 0; i < nw; ++i) tot += s_red[i];
        s_red[32] = tot;
    }
    __syncthreads();
    return s_red[32];
}

// Robbins-Monro proposal-scale adaptation (e.g. src/chain.cpp:618-627).
__device__ __forceinline__ double adapt_sd(double sd, double accepts, double denom, double root_arg,
                                           double floor_)
{
    const double rate = accepts / denom;
    const double upd = (rate - .23) / sqrt(root_arg);
    const double nv = sd + upd;
    return nv > floor_ ? nv : floor_;  // std::max(sd + update, floor)
}

// ---- work lists: cells sorted by cost class ---------------------------------------------------
// The cost of one transmission term is ~2^k (k + m - k + 1) (inclusion-exclusion over the subsets
// of the k latent alleles, vectorised over the m - k + 1 mixture terms), so cells of one block
// differ by three orders of magnitude.  A thread-per-cell mapping in natural order leaves ~3 of
// 32 lanes active (ncu, profiles/r01a).  Instead every block counting-sorts its cells by
// (k, m - k + 1), heaviest first, and its warps walk the sorted list round-robin: the lanes of a
// warp then run the same subset sequence and the same register tier, and the heavy classes are
// spread over all warps.  Which thread evaluates a cell does not change the cell's value, and all
// sums over cells are taken in cell order afterwards, so results do not depend on the
// (atomics-ordered) position of a cell inside its class.
__device__ __forceinline__ int work_class(int k, int m)
{
    // reversed so that ascending class = heaviest first
    const int kk = k > kIeMax ? kIeMax + 1 : k;
    int cnt = m - k + 1;
    if (k > kIeMax) cnt = m;  // the EGF recurrence costs ~k m^2: order those cells by m
    cnt = cnt < 1 ? 1 : (cnt > 64 ? 64 : cnt);
    // cells with at most MOIRE_MERGE_K latent alleles (<= 2^k - 1 cheap subsets) form one class
    // per mixture count: there the cnt terms of log1p + exp outweigh the subset sums, so equal
    // counts matter more than equal k inside a warp
    if (kk <= MOIRE_MERGE_K) return kSortBins - 1 - (MOIRE_MERGE_K * 64 + (cnt - 1));
    return kSortBins - 1 - (kk * 64 + (cnt - 1));
}

constexpr int kClassCoopBegin = 64;                       // first class with k <= 10
constexpr int kClassLightBegin = (12 - kCoopMinK) * 64;   // first class with k < kCoopMinK
// The fused kernels serve tiny problems only (a few blocks, one per SM: latency, not throughput,
// decides), so they hand cells to the cooperative evaluator from a lower k.
#ifndef MOIRE_FUSED_COOP_MIN_K
#define MOIRE_FUSED_COOP_MIN_K 6
#endif
constexpr int kClassLightBeginFused = (12 - MOIRE_FUSED_COOP_MIN_K) * 64;

// hist: kSortBins counters; order: out; s_tmp: >= 40 words.  cls(cell) < 0 leaves the cell out.
// Returns the number of listed cells (same value in every thread); s_tmp[33] / s_tmp[34] get the
// list positions where the k <= 10 cells and the k < kCoopMinK cells begin.  Ends with a barrier.
template <class ClassFn>
__device__ __forceinline__ int block_sort_cells(int ncell, ClassFn cls, unsigned int *hist,
                                                unsigned short *order, unsigned int *s_tmp)
{
    const int tid = threadIdx.x, B = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = (B + 31) >> 5;
    for (int b = tid; b < kSortBins; b += B) hist[b] = 0;
    __syncthreads();
    for (int c = tid; c < ncell; c += B)
    {
        const int key = cls(c);
        if (key >= 0) atomicAdd(&hist[key], 1u);
    }
    __syncthreads();
    const int per = (kSortBins + B - 1) / B, b0 = tid * per;
    unsigned int local = 0;
    for (int q = 0; q < per; ++q)
        if (b0 + q < kSortBins) local += hist[b0 + q];
    unsigned int x = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_tmp[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
        const unsigned int mine = lane < nw ? s_tmp[lane] : 0u;
        unsigned int w = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_tmp[lane] = w - mine;
        if (lane == 31) s_tmp[32] = w;
    }
    __syncthreads();
    unsigned int base = s_tmp[warp] + x - local;
    for (int q = 0; q < per; ++q)
    {
        if (b0 + q < kSortBins)
        {
            const unsigned int h = hist[b0 + q];
            hist[b0 + q] = base;
            if (b0 + q == kClassCoopBegin) s_tmp[33] = base;
            if (b0 + q == kClassLightBeginFused) s_tmp[34] = base;
            base += h;
        }
    }
    const int total = (int)s_tmp[32];
    __syncthreads();
    for (int c = tid; c < ncell; c += B)
    {
        const int key = cls(c);
        if (key >= 0) order[atomicAdd(&hist[key], 1u)] = (unsigned short)c;
    }
    __syncthreads();
    return total;
}

// Transmission terms of the listed cells.  The list is heaviest first: [0, n_egf) cells with
// k > 10 (EGF recurrence), [n_egf, coop_end) cells with kCoopMinK <= k <= 10, the rest light
// cells.  Work is handed out dynamically, heaviest first, in units of one chunk of 32
// one-thread cells or one cooperatively evaluated cell (one warp: draws resp. subsets across
// lanes): `next` is a shared counter the caller has zeroed before its last barrier.
// fetch(cell, lat, prow, m, rel) supplies a cell's inputs; out[cell] gets the term.
// scr: kScratchSlots rows of blockDim.x + 1 doubles.  No block barrier inside.
template <class Fetch>
__device__ __forceinline__ void block_eval_transmissions(const unsigned short *order, int nwork,
                                                         int n_egf, int coop_end, Fetch fetch,
                                                         double *out, double *scr,
                                                         const double *lgam, bool allow_rel,
                                                         unsigned int *next)
{
    const int tid = threadIdx.x, ST = blockDim.x + 1;
    const int lane = tid & 31, warp = tid >> 5;
    // A cooperative evaluation costs 3-4x the issue slots of a thread-per-cell one, so whole
    // warps' worth of heavy cells (plentiful on high-COI data) still go one per thread; only
    // the leftover of each heavy range -- the few cells that would otherwise occupy a warp alone
    // and gate the block -- is evaluated cooperatively.
    // ... and only where a heavy range is long enough to give every warp such a
    // chunk; a lone chunk of 1023-subset cells would gate the block just the same.
    const int nw = blockDim.x >> 5;
    const int heavy = coop_end - n_egf;
    const int egf_chunks = (n_egf >> 5) >= nw ? n_egf >> 5 : 0;
    const int heavy_chunks = (heavy >> 5) >= nw ? heavy >> 5 : 0;
    const int egf_left = n_egf - 32 * egf_chunks, heavy_left = heavy - 32 * heavy_chunks;
    const int light_chunks = (nwork - coop_end + 31) >> 5;
    const int u1 = egf_chunks, u2 = u1 + egf_left, u3 = u2 + heavy_chunks, u4 = u3 + heavy_left;
    const int units = u4 + light_chunks;
    for (;;)
    {
        int u = 0;
        if (lane == 0) u = (int)atomicAdd(next, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= units) break;
        mask_t lat;
        const double *prow;
        int m;
        Rel rel;
        const bool coop_egf = u >= u1 && u < u2, coop_ie = u >= u3 && u < u4;
        if (coop_egf || coop_ie)
        {
            const int cell = order[coop_egf ? 32 * egf_chunks + (u - u1)
                                            : n_egf + 32 * heavy_chunks + (u - u3)];
            fetch(cell, lat, prow, m, rel);
            const int k = mpopc(lat);
            double T;
            if (k > m)
                T = NAN;
            else if (coop_egf)
                T = transmission_ll_egf_coop<false>(lat, prow, m, rel, allow_rel, scr + warp * 32, ST, lgam, nullptr, PamOut{});
            else
                T = transmission_ll_coop<false>(lat, prow, m, rel, allow_rel, scr + warp * 32, ST, lgam, nullptr, PamOut{});
            if (lane == 0) out[cell] = T;
        }
        else
        {
            int w, end;
            if (u < u1)
            {
                w = 32 * u + lane;
                end = 32 * egf_chunks;
            }
            else if (u < u3)
            {
                w = n_egf + 32 * (u - u2) + lane;
                end = n_egf + 32 * heavy_chunks;
            }
            else
            {
                w = coop_end + 32 * (u - u4) + lane;
                end = nwork;
            }
            const unsigned callers = __ballot_sync(0xffffffffu, w < end);
            if (w < end)
            {
                const int cell = order[w];
                fetch(cell, lat, prow, m, rel);
                double pv[kGather];
                gather_freqs(lat, prow, pv);
                out[cell] = transmission_ll<false>(lat, prow, pv, m, rel, allow_rel, scr + tid, ST, lgam, nullptr, callers, PamOut{});
            }
        }
        __syncwarp();
    }
}

__host__ __device__ inline size_t align8(size_t b) { return (b + 7) & ~(size_t)7; }
__host__ __device__ inline size_t scratch_doubles(int B) { return (size_t)kScratchSlots * (B + 1); }

// One proposal for sample i of chain t: the draws, validity checks and prior terms of
// update_eps_neg / update_eps_pos / update_m / update_r / update_eff_coi / update_samples
// (src/chain.cpp:165-313, 315-374, 567-807) before any cell is evaluated.
template <int KIND>
__device__ __forceinline__ Proposal draw_sample_proposal(const View &v, int t, int i, int iteration,
                                                         const double *s_lgam)
{
    const int N = v.N;
    const unsigned gchain = (unsigned)(v.chain_offset + t);
    const size_t si = (size_t)t * N + i;
    Stream rng;
    rng.open(v.seed, (uint32_t)iteration, KIND, gchain, (uint32_t)i, 0);
    Proposal P;
    P.valid = 1;
    P.adapt = 0;
    P.accept = 0;
    P.m_new = v.m[si];
    P.r_new = v.r[si];
    P.en_new = v.eps_neg[si];
    P.ep_new = v.eps_pos[si];
    P.log_r = v.log_r[si];
    P.log_1mr = v.log_1mr[si];
    P.odds = v.odds[si];
    P.adj = 0.0;
    P.dprior = 0.0;
    P.pr_en_new = P.pr_ep_new = P.pr_coi_new = P.pr_r_new = 0.0;
    const double min_sampled = DBL_MIN;  // std::numeric_limits<double>::min()
    const double lambda = v.mean_coi[t];
    const double log_geom = -0.40546510810816438;  // log(1 - 1/3): sample_coi_delta(2)

    if (KIND == KIND_EPS_NEG || KIND == KIND_SAMPLES)
    {
        const double z = v.sd_eps_neg[si] * next_normal(rng);
        double prop, adj;
        constrained_proposal(z, P.en_new, min_sampled, 1.0, &prop, &adj);
        const bool ok = prop < 1.0 && prop > 1e-32;
        P.valid &= ok;
        P.en_new = prop;
        P.adj += adj;
        P.pr_en_new = dbeta_log(prop, v.eps_neg_a, v.eps_neg_b, v.eps_neg_lbeta);
        P.dprior += P.pr_en_new - v.pr_eps_neg[si];
    }
    if (KIND == KIND_EPS_POS || KIND == KIND_SAMPLES)
    {
        const double z = v.sd_eps_pos[si] * next_normal(rng);
        double prop, adj;
        constrained_proposal(z, P.ep_new, min_sampled, 1.0, &prop, &adj);
        const bool ok = prop < 1.0 && prop > 1e-32;
        P.valid &= ok;
        P.ep_new = prop;
        P.adj += adj;
        P.pr_ep_new = dbeta_log(prop, v.eps_pos_a, v.eps_pos_b, v.eps_pos_lbeta);
        P.dprior += P.pr_ep_new - v.pr_eps_pos[si];
    }
    if (KIND == KIND_R || (KIND == KIND_SAMPLES && v.allow_rel))
    {
        const double z = v.sd_r[si] * next_normal(rng);
        double prop, adj;
        constrained_proposal(z, P.r_new, min_sampled, .99, &prop, &adj);
        if (KIND == KIND_SAMPLES) P.valid &= (prop < 1.0 && prop > 1e-32);
        P.r_new = prop;
        P.adj += adj;
    }
    if (KIND == KIND_SAMPLES && !v.allow_rel) P.r_new = 0.0;  // src/chain.cpp:719
    if (KIND == KIND_EFF_COI)
    {
        const double curr_eff = (double)(P.m_new - 1) * (1.0 - P.r_new) + 1.0;
        const double z = v.sd_mr[si] * next_normal(rng);
        double prop_eff, adj;
        constrained_proposal(z, curr_eff, 1.0, (double)v.max_coi, &prop_eff, &adj);
        const int prop_m = P.m_new + next_coi_delta(rng, log_geom);
        const double prop_r = 1.0 - (prop_eff - 1.0) / ((double)prop_m - 1.0);
        // src/chain.cpp:244-246: skipped proposals neither evaluate nor adapt
        if (prop_m <= 0 || prop_m > v.max_coi || prop_r > 1.0 || prop_r < 1e-32) P.valid = 0;
        P.m_new = prop_m;
        P.r_new = prop_r;
        P.adj += adj;
    }
    if (KIND == KIND_M || KIND == KIND_SAMPLES)
    {
        const int prop_m = P.m_new + next_coi_delta(rng, log_geom);
        P.valid &= (prop_m > 0 && prop_m <= v.max_coi);
        P.m_new = prop_m;
    }
    if (KIND == KIND_R || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES)
    {
        P.pr_r_new = dbeta_log(P.r_new, v.r_a, v.r_b, v.r_lbeta);
        P.dprior += P.pr_r_new - v.pr_r[si];
        P.log_r = cold_log(P.r_new);
        P.log_1mr = cold_log(1.0 - P.r_new);
        P.odds = P.r_new / (1.0 - P.r_new);
    }
    if (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES)
    {
        const int mm = P.valid ? P.m_new : 1;  // keep the LUT index in range
        P.pr_coi_new = dztpois_log(mm, cold_log(lambda), cold_log(cold_exp(lambda) - 1.0), s_lgam);
        P.dprior += P.pr_coi_new - v.pr_coi[si];
    }
    // which updates adapt, and when (src/chain.cpp:303-311, 364-372, 618-627, 683-692)
    if (KIND == KIND_EPS_NEG || KIND == KIND_EPS_POS || KIND == KIND_EFF_COI)
        P.adapt = P.valid;
    if (KIND == KIND_R) P.adapt = 1;
    P.logu = cold_log(rng.next_uniform());
    return P;
}

// The Metropolis decision for one sample given the summed change d of its cells' log-likelihood
// and a of its latent genotypes' log proposal densities: acceptance rule of each update, state
// and prior caches on acceptance, counters, Robbins-Monro adaptation.  Sets P.accept.
template <int KIND>
__device__ __forceinline__ void decide_sample(const View &v, Proposal &P, int t, int i, int iteration,
                                              double d, double a)
{
    const size_t si = (size_t)t * v.N + i;
    const double temp = v.temp[t];
    bool accept = false;
    if (P.valid)
    {
        const double dpost = temp * d + P.dprior;  // new_post - old_post
        const double ratio = dpost + (P.adj + a);
        if (KIND == KIND_M || KIND == KIND_SAMPLES)
            accept = !isnan(dpost) && P.logu <= ratio;
        else if (KIND == KIND_EFF_COI)
            accept = !(isnan(dpost) || isnan(ratio) || P.logu > ratio);
        else
            accept = !(isnan(dpost) || P.logu > ratio);
    }
    P.accept = accept;
    int nacc = 0;
    if (accept)
    {
        if (KIND == KIND_EPS_NEG || KIND == KIND_SAMPLES)
        {
            v.eps_neg[si] = P.en_new;
            v.pr_eps_neg[si] = P.pr_en_new;
        }
        if (KIND == KIND_EPS_POS || KIND == KIND_SAMPLES)
        {
            v.eps_pos[si] = P.ep_new;
            v.pr_eps_pos[si] = P.pr_ep_new;
        }
        if (KIND == KIND_R || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES)
        {
            v.r[si] = P.r_new;
            v.log_r[si] = P.log_r;
            v.log_1mr[si] = P.log_1mr;
            v.odds[si] = P.odds;
            v.pr_r[si] = P.pr_r_new;
        }
        if (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES)
        {
            v.m[si] = P.m_new;
            v.pr_coi[si] = P.pr_coi_new;
        }
    }
    int *accp = KIND == KIND_EPS_NEG   ? v.acc_eps_neg
                : KIND == KIND_EPS_POS ? v.acc_eps_pos
                : KIND == KIND_M       ? v.acc_m
                : KIND == KIND_R       ? v.acc_r
                : KIND == KIND_EFF_COI ? v.acc_mr
                                       : v.acc_samples;
    nacc = accp[si] + (accept ? 1 : 0);
    if (accept) accp[si] = nacc;
    if (P.adapt && iteration < v.burnin && iteration > 15)
    {
        double *sdp = KIND == KIND_EPS_NEG   ? v.sd_eps_neg
                      : KIND == KIND_EPS_POS ? v.sd_eps_pos
                      : KIND == KIND_R       ? v.sd_r
                                             : v.sd_mr;
        sdp[si] = adapt_sd(sdp[si], (double)nacc, (double)iteration,
                           (double)iteration + 1.0, .01);
    }
}

// Shared memory of the per-sample kernel.  The two error-rate updates evaluate no transmission
// term: no scratch columns, no sort, one array per cell -- which lets several of their blocks
// share an SM (they are plain streaming kernels).
__host__ __device__ inline size_t sample_kernel_smem(int S, int L, int nA, bool need_t)
{
    const size_t C = (size_t)S * L;
    size_t b = 0;
    b += 128 * sizeof(double);                                      // lgam
    b += (need_t ? scratch_doubles(kThreads) : 0) * sizeof(double);  // per-thread scratch columns
    b += (need_t ? 4 : 2) * C * sizeof(double);  // proposed latent, [its log density, T,] O
    b += (size_t)S * nA * sizeof(EpsLogs);
    b += (size_t)S * sizeof(Proposal);
    b += kSortBins * sizeof(unsigned int) + 40 * sizeof(unsigned int);
    b += align8(C * sizeof(unsigned short));
    return b;
}

// ---- per-sample updates: eps_neg, eps_pos, m, r, eff_coi, samples ----------------------------
template <int KIND>
__global__ void __launch_bounds__(kThreads, MOIRE_MIN_BLOCKS_S * (256 / kThreads))
    k_update_sample(const View v, const int iteration_arg, const int S, const int tiles_per_chain)
{
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
    constexpr bool RESAMPLE = (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES);
    constexpr bool NEED_T = RESAMPLE || KIND == KIND_R;
    constexpr bool NEED_O = RESAMPLE || KIND == KIND_EPS_NEG || KIND == KIND_EPS_POS;

    extern __shared__ double smem[];
    const int C = S * v.L;
    double *s_lgam = smem;
    double *s_qs = s_lgam + 128;
    mask_t *s_nlat =
        reinterpret_cast<mask_t *>(s_qs + (NEED_T ? scratch_doubles(kThreads) : 0));
    double *s_nadj = reinterpret_cast<double *>(s_nlat + C);
    double *s_T = s_nadj + (NEED_T ? C : 0);  // the error-rate updates keep only latent and O
    double *s_O = s_T + (NEED_T ? C : 0);
    EpsLogs *s_elogs = reinterpret_cast<EpsLogs *>(s_O + C);
    Proposal *s_prop = reinterpret_cast<Proposal *>(s_elogs + S * v.nA);
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(s_prop + S);
    unsigned int *s_tmp = s_hist + kSortBins;
    unsigned short *s_order = reinterpret_cast<unsigned short *>(s_tmp + 40);
    __shared__ unsigned int s_evals, s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const int t = blockIdx.x / tiles_per_chain;
    const int s0 = (blockIdx.x - t * tiles_per_chain) * S;
    const int ns = min(S, v.N - s0);
    const int N = v.N, L = v.L, nA = v.nA;
    const unsigned gchain = (unsigned)(v.chain_offset + t);

    for (int i = tid; i < 128; i += kThreads) s_lgam[i] = v.lgam[i];
    if (tid == 0) s_evals = s_next = 0;
    __syncthreads();

    // ---- phase A: one proposal per sample -----------------------------------------------------
    if (tid < ns) s_prop[tid] = draw_sample_proposal<KIND>(v, t, s0 + tid, iteration, s_lgam);
    __syncthreads();

    if (NEED_O)
    {
        for (int e = tid; e < ns * nA; e += kThreads)
        {
            const int sl = e / nA, a = e - sl * nA;
            build_eps_logs(s_elogs[e], s_prop[sl].en_new, s_prop[sl].ep_new, v.max_eps_neg,
                           v.max_eps_pos, v.avals[a]);
        }
        __syncthreads();
    }

    // ---- phase B1: per cell, natural (coalesced) order: proposed latent genotype, observation
    // term -----------------------------------------------------------------------------------
    const size_t tile_base = ((size_t)t * N + s0) * L;  // first cell of the tile in [T][N][L]
    const size_t data_base = (size_t)s0 * L;            // same in [N][L]
    const int ncell = ns * L;
    unsigned my_evals = 0;
    for (int cell = tid; cell < ncell; cell += kThreads)
    {
        const int sl = cell / L, j = cell - sl * L;
        const Proposal &P = s_prop[sl];
        if (!P.valid) continue;
        const mask_t obs = v.obs[data_base + cell];
        const bool miss = v.missing[data_base + cell] != 0;
        const int A = v.nall[j];
        mask_t lat;
        const EpsLogs *el = NEED_O ? &s_elogs[sl * nA + v.aclass[j]] : nullptr;
        if (RESAMPLE)
        {
            Stream crng;
            crng.open(v.seed, (uint32_t)iteration, KIND, gchain, (uint32_t)(s0 + sl),
                      (uint32_t)(j + 1));
            double lp;
            lat = sample_latent_genotype(crng, obs, A, P.m_new, *el, v.logC, &lp);
            s_nadj[cell] = lp;
        }
        else
        {
            lat = v.latent[tile_base + cell];
        }
        s_nlat[cell] = lat;
        if (!miss)
        {
            if (NEED_O) s_O[cell] = observation_ll(lat, obs, A, *el);
            ++my_evals;
        }
    }
    if (my_evals) atomicAdd(&s_evals, my_evals);
    __syncthreads();

    // ---- phase B2: transmission terms, heaviest class first -----------------------------------
    if (NEED_T)
    {
        auto cls = [&](int cell) -> int {
            const int sl = cell / L;
            const Proposal &P = s_prop[sl];
            if (!P.valid || v.missing[data_base + cell] != 0) return -1;
            return work_class(mpopc(s_nlat[cell]), P.m_new);
        };
        const int nwork = block_sort_cells(ncell, cls, s_hist, s_order, s_tmp);
        auto fetch = [&](int cell, mask_t &lat, const double *&prow, int &m, Rel &rel) {
            const int sl = cell / L, j = cell - sl * L;
            const Proposal &P = s_prop[sl];
            lat = s_nlat[cell];
            prow = v.p + ((size_t)t * L + j) * v.AP;
            m = P.m_new;
            rel.log_r = P.log_r;
            rel.log_1mr = P.log_1mr;
            rel.odds = P.odds;
        };
        block_eval_transmissions(s_order, nwork, (int)s_tmp[33], (int)s_tmp[34], fetch, s_T, s_qs,
                                 s_lgam, v.allow_rel != 0, &s_next);
        __syncthreads();
    }

    // ---- phase C: one warp per sample sums its cells (in locus order) and decides -------------
    for (int sl = warp; sl < ns; sl += kWarps)
    {
        const Proposal &Pc = s_prop[sl];
        double d = 0.0, a = 0.0;
        if (Pc.valid)
        {
            for (int j = lane; j < L; j += 32)
            {
                const int cell = sl * L + j;
                const size_t gi = tile_base + cell;
                if (RESAMPLE) a += s_nadj[cell] - v.lgadj[gi];
                if (v.missing[data_base + cell] == 0)
                {
                    const double oldT = v.cellT[gi], oldO = v.cellO[gi];
                    const double T = NEED_T ? s_T[cell] : oldT;
                    const double O = NEED_O ? s_O[cell] : oldO;
                    d += __dadd_rn(T, O) - __dadd_rn(oldT, oldO);
                }
            }
        }
        d = warp_sum(d);
        if (RESAMPLE) a = warp_sum(a);
        if (lane == 0) decide_sample<KIND>(v, s_prop[sl], t, s0 + sl, iteration, d, a);
    }
    __syncthreads();

    // ---- phase D: commit the accepted samples' cells -----------------------------------------
    for (int cell = tid; cell < ncell; cell += kThreads)
    {
        const int sl = cell / L;
        if (!s_prop[sl].accept) continue;
        const size_t gi = tile_base + cell;
        if (RESAMPLE)
        {
            v.latent[gi] = s_nlat[cell];
            v.lgadj[gi] = s_nadj[cell];
        }
        if (v.missing[data_base + cell] == 0)
        {
            if (NEED_T) v.cellT[gi] = s_T[cell];
            if (NEED_O) v.cellO[gi] = s_O[cell];
        }
    }
    if (tid == 0 && s_evals) atomicAdd(v.evals + KIND, (unsigned long long)s_evals);
}

// ---- SALT proposal math (src/mcmc_utils.h:150-313) --------------------------------------------
__device__ __noinline__ double logit_d(double x)
{
    return x < .5 ? cold_log(x) - cold_log1p(-x) : cold_log(x / (1.0 - x));
}

__device__ __noinline__ void log_pq_d(double x, double &lp, double &lq)
{
    const double ex = cold_exp(x);
    if (x < 0.0)
    {
        lq = -cold_log1p(ex);
        lp = lq + x;
    }
    else
    {
        lp = -cold_log1p(1.0 / ex);
        lq = lp - x;
    }
}

__device__ __noinline__ double logit_scale_d(double x, double scale)
{
    const bool ok = (scale < 0.69314718055994530942) && (fabs(scale) < fabs(x + scale));
    const double u = -scale - (ok ? 0.0 : x);
    const double w = -scale - (ok ? x : 0.0);
    const double ev = cold_exp(w);
    const double eumo = cold_expm1(u);
    double l2;
    if (isinf(eumo))
        l2 = fmax(u, w) + cold_log1p(cold_exp(-fabs(u - w)));
    else
        l2 = cold_log(eumo + ev);
    if (w > cold_log(2.0 * fabs(eumo))) return -(w + cold_log1p(eumo / ev));
    return -l2;
}

__device__ __noinline__ double expit_d(double x) { return 1.0 / (1.0 + cold_exp(-x)); }

__host__ __device__ inline int p_kernel_threads(int N) { return N > 2048 ? 512 : MOIRE_P_THREADS; }

__host__ __device__ inline size_t p_kernel_smem(int N, int AP, int B)
{
    size_t b = 0;
    b += 128 * sizeof(double);                  // lgam
    b += scratch_doubles(B) * sizeof(double);   // per-thread scratch columns
    b += (size_t)3 * N * sizeof(double);        // latent, Told, Tnew
    b += (size_t)3 * AP * sizeof(double);       // p_cur, p_prop, p_sd
    b += 40 * sizeof(double);                   // reductions + scalars
    b += (size_t)2 * AP * sizeof(int);          // acc, att
    b += kSortBins * sizeof(unsigned int) + 40 * sizeof(unsigned int);
    b += align8((size_t)N * sizeof(unsigned short));  // work list
    b += align8((size_t)N);                     // COI per sample, 0 = cell missing
    return b;
}

// ---- update_p: SALT sampler, one block per (chain, locus) (src/chain.cpp:466-565) -------------
// The locus' N cells are staged once (latent genotype, current transmission term, COI) and
// sorted once by cost class; each of the A_j sequential proposals then re-evaluates them from
// shared memory under the proposed frequencies.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK == 512 ? 1 : MOIRE_MIN_BLOCKS_P * (256 / BLOCK))
    k_update_p(const View v, const int iteration_arg)
{
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
    extern __shared__ double smem[];
    const int N = v.N, L = v.L, AP = v.AP, B = blockDim.x;
    double *s_lgam = smem;
    double *s_qs = s_lgam + 128;
    mask_t *s_lat = reinterpret_cast<mask_t *>(s_qs + scratch_doubles(B));
    double *s_Told = reinterpret_cast<double *>(s_lat + N);
    double *s_Tnew = s_Told + N;
    double *s_pcur = s_Tnew + N;
    double *s_pprop = s_pcur + AP;
    double *s_psd = s_pprop + AP;
    double *s_red = s_psd + AP;  // 40 doubles: [0..32] reduction, 33 logAdj, 34 flag, 35 accept
    int *s_acc = reinterpret_cast<int *>(s_red + 40);
    int *s_att = s_acc + AP;
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(s_att + AP);
    unsigned int *s_tmp = s_hist + kSortBins;
    unsigned short *s_order = reinterpret_cast<unsigned short *>(s_tmp + 40);
    unsigned char *s_m = reinterpret_cast<unsigned char *>(s_order) + align8((size_t)N * 2);

    __shared__ unsigned int s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x / L, j = blockIdx.x - t * L;
    const int A = v.nall[j];
    const unsigned gchain = (unsigned)(v.chain_offset + t);
    const double temp = v.temp[t];
    const size_t prow = ((size_t)t * L + j) * AP;

    for (int i = tid; i < 128; i += B) s_lgam[i] = v.lgam[i];
    for (int a = tid; a < AP; a += B)
    {
        s_pcur[a] = v.p[prow + a];
        s_psd[a] = v.p_sd[prow + a];
        s_acc[a] = v.p_acc[prow + a];
        s_att[a] = v.p_att[prow + a];
    }
    for (int i = tid; i < N; i += B)
    {
        const size_t gi = ((size_t)t * N + i) * L + j;
        const bool miss = v.missing[(size_t)i * L + j] != 0;
        s_lat[i] = v.latent[gi];
        s_Told[i] = v.cellT[gi];
        s_Tnew[i] = 0.0;
        s_m[i] = miss ? 0 : (unsigned char)v.m[(size_t)t * N + i];
    }
    __syncthreads();
    auto cls = [&](int i) -> int {
        const int m = s_m[i];
        return m == 0 ? -1 : work_class(mpopc(s_lat[i]), m);
    };
    const int nwork = block_sort_cells(N, cls, s_hist, s_order, s_tmp);
    const int n_egf = (int)s_tmp[33], coop_end = (int)s_tmp[34];
    auto fetch = [&](int i, mask_t &lat, const double *&prow_, int &m, Rel &rel) {
        const size_t si = (size_t)t * N + i;
        lat = s_lat[i];
        prow_ = s_pprop;
        m = s_m[i];
        rel.log_r = v.log_r[si];
        rel.log_1mr = v.log_1mr[si];
        rel.odds = v.odds[si];
    };

    unsigned long long evals = 0;
    for (int rep = 0; rep < A; ++rep)
    {
        // -- the proposal, warp 0, lanes across alleles
        if (warp == 0)
        {
            Stream rng;
            rng.open(v.seed, (uint32_t)iteration, KIND_P, gchain, (uint32_t)j, (uint32_t)rep);
            const int idx = (int)mulhi32(rng.next_u32(), (uint32_t)A);
            const double z = next_normal(rng);
            const double logu = cold_log(rng.next_uniform());
            double lg[2], lp[2], lq[2];
            bool rest[2];
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                const int a = lane + 32 * s;
                const bool has = a < A;
                lg[s] = has ? logit_d(s_pcur[a]) : 0.0;
                lp[s] = lq[s] = 0.0;
                if (has) log_pq_d(lg[s], lp[s], lq[s]);
                rest[s] = has && a != idx;
            }
            const double lg_sel = (idx >> 5) ? lg[1] : lg[0];
            const double logit_curr = __shfl_sync(0xffffffffu, lg_sel, idx & 31);
            const double logit_prop = logit_curr + s_psd[idx] * z;
            double clp, clq, plp, plq;
            log_pq_d(logit_curr, clp, clq);
            log_pq_d(logit_prop, plp, plq);
            // logitSum over the other alleles (src/mcmc_utils.h:219-248): the largest logit
            // leads; on ties the lowest allele index does.
            double bv = -INFINITY;
            int ba = 1 << 30;
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (rest[s] && (lg[s] > bv)) { bv = lg[s]; ba = lane + 32 * s; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oa = __shfl_xor_sync(0xffffffffu, ba, o);
                if (ov > bv || (ov == bv && oa < ba)) { bv = ov; ba = oa; }
            }
            const double lp_sel = (ba >> 5) ? lp[1] : lp[0];
            const double lq_sel = (ba >> 5) ? lq[1] : lq[0];
            const double lp1 = __shfl_sync(0xffffffffu, lp_sel, ba & 31);
            const double lq1 = __shfl_sync(0xffffffffu, lq_sel, ba & 31);
            double cum = 0.0;
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (rest[s] && (lane + 32 * s) != ba) cum += (bv < 0.0) ? cold_exp(lp[s] - lp1) : cold_exp(lp[s]);
            cum = warp_sum(cum);
            const double lsum = (bv < 0.0) ? lp1 + cold_log1p(cum) : cold_log1p(-cold_exp(lq1) + cum);
            const double ls = plq - lsum;
            bool below = false;
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                const int a = lane + 32 * s;
                if (a < A)
                {
                    const double np = expit_d(a == idx ? logit_prop : logit_scale_d(lg[s], ls));
                    s_pprop[a] = np;
                    below |= np < 1e-12;
                }
            }
            below = __any_sync(0xffffffffu, below);
            if (lane == 0)
            {
                s_next = 0;
                ++s_att[idx];
                s_red[33] = (clp - plp) + (double)(A - 1) * (clq - plq);
                s_red[34] = below ? 1.0 : 0.0;
                s_red[36] = logu;
                s_red[37] = (double)idx;
            }
        }
        __syncthreads();
        if (s_red[34] != 0.0) break;  // src/chain.cpp:503-518: leaves the locus

        // -- all cells of the locus under the proposed frequencies, heaviest class first
        const long long tr0 = v.trace ? clock64() : 0;
        block_eval_transmissions(s_order, nwork, n_egf, coop_end, fetch, s_Tnew, s_qs, s_lgam,
                                 v.allow_rel != 0, &s_next);
        if (v.trace && blockIdx.x < 64 && rep < 16 && lane == 0)
        {
            // [block][rep][warp] -> (clock at start of the evaluation phase, clock at its end)
            unsigned long long *tr = v.trace + (((size_t)blockIdx.x * 16 + rep) * 16 + warp) * 2;
            tr[0] = (unsigned long long)tr0;
            tr[1] = (unsigned long long)clock64();
        }
        evals += tid == 0 ? (unsigned long long)nwork : 0ull;
        __syncthreads();
        // -- the change of the locus' log-likelihood, summed in sample order
        double dsum = 0.0;
        for (int i = tid; i < N; i += B)
            if (s_m[i] != 0) dsum += s_Tnew[i] - s_Told[i];
        const double d = block_sum(dsum, s_red);
        if (tid == 0)
        {
            const int idx = (int)s_red[37];
            const double dpost = temp * d;  // the priors do not depend on p
            const double ratio = dpost + s_red[33];
            const bool accept = !(isnan(dpost) || s_red[36] > ratio);
            if (accept) ++s_acc[idx];
            if (iteration < v.burnin && s_att[idx] > 15)
            {
                // src/chain.cpp:555-562
                const double rate = ((double)s_acc[idx] + 1.0) / ((double)s_att[idx] + 1.0);
                double sd = s_psd[idx] + (rate - .23) / sqrt((double)s_att[idx] + 1.0);
                s_psd[idx] = sd > .01 ? sd : .01;
            }
            s_red[35] = accept ? 1.0 : 0.0;
        }
        __syncthreads();
        if (s_red[35] != 0.0)
        {
            for (int i = tid; i < N; i += B)
                if (s_m[i] != 0) s_Told[i] = s_Tnew[i];
            for (int a = tid; a < A; a += B) s_pcur[a] = s_pprop[a];
        }
        __syncthreads();
    }
    __syncthreads();

    for (int i = tid; i < N; i += B)
    {
        if (s_m[i] != 0) v.cellT[((size_t)t * N + i) * L + j] = s_Told[i];
    }
    for (int a = tid; a < AP; a += B)
    {
        v.p[prow + a] = s_pcur[a];
        v.p_sd[prow + a] = s_psd[a];
        v.p_acc[prow + a] = s_acc[a];
        v.p_att[prow + a] = s_att[a];
    }
    if (tid == 0 && evals) atomicAdd(v.evals + KIND_P, evals);
}

// ---- update_mean_coi: one block per chain (src/chain.cpp:809-856) ----------------------------
__global__ void __launch_bounds__(kThreads) k_update_mean_coi(const View v, const int iteration_arg)
{
    pdl_enter();
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
    __shared__ double s_lgam[128];
    __shared__ double s_red[40];
    const int tid = threadIdx.x, t = blockIdx.x, N = v.N;
    for (int i = tid; i < 128; i += kThreads) s_lgam[i] = v.lgam[i];
    if (tid == 0)
    {
        Stream rng;
        rng.open(v.seed, (uint32_t)iteration, KIND_MEAN_COI, (unsigned)(v.chain_offset + t), 0, 0);
        const double z = v.mean_coi_sd[t] * next_normal(rng);
        double prop, adj;
        constrained_proposal(z, v.mean_coi[t], 1e-5, (double)v.max_coi, &prop, &adj);
        s_red[33] = prop;
        s_red[34] = adj;
        s_red[35] = cold_log(rng.next_uniform());
    }
    __syncthreads();
    const double prop = s_red[33];
    const double ll = cold_log(prop), lem = cold_log(cold_exp(prop) - 1.0);
    double dsum = 0.0;
    for (int i = tid; i < N; i += kThreads)
    {
        const size_t si = (size_t)t * N + i;
        dsum += dztpois_log(v.m[si], ll, lem, s_lgam) - v.pr_coi[si];
    }
    const double d = block_sum(dsum, s_red);
    if (tid == 0)
    {
        const double hyper_new =
            dgamma_log(prop, v.mean_coi_shape, v.mean_coi_scale, v.lg_shape, v.l_scale);
        const double dpost = d + (hyper_new - v.hyper[t]);  // the likelihood does not change
        const double ratio = dpost + s_red[34];
        const bool accept = !isnan(dpost) && s_red[35] <= ratio;
        if (accept)
        {
            v.mean_coi[t] = prop;
            v.hyper[t] = hyper_new;
            v.mean_coi_acc[t] += 1.0;
        }
        if (iteration < v.burnin && iteration > 15)
        {
            v.mean_coi_sd[t] = adapt_sd(v.mean_coi_sd[t], v.mean_coi_acc[t],
                                        (double)(iteration + 1), (double)iteration + 1.0, .1);
        }
        s_red[36] = accept ? 1.0 : 0.0;
    }
    __syncthreads();
    if (s_red[36] != 0.0)
    {
        for (int i = tid; i < N; i += kThreads)
        {
            const size_t si = (size_t)t * N + i;
            v.pr_coi[si] = dztpois_log(v.m[si], ll, lem, s_lgam);
        }
    }
}

// ---- per-chain sums (calc_new_likelihood / calc_new_prior, src/chain.cpp:978-1026) ------------
// kReduceParts blocks per chain (one block per chain left most of the device idle: 24 us for the
// 40-chain ladder, as long again for a 5-chain shard): part g sums the cells c = g B + tid (mod
// kReduceParts B) in a fixed order, and the last part to finish adds the parts in index order -- the
// result does not depend on which block that is.
constexpr int kReduceParts = 8;
__global__ void __launch_bounds__(1024) k_reduce(const View v, double *part, unsigned int *done)
{
    pdl_enter();
    __shared__ double s_red[40];
    const int tid = threadIdx.x, t = blockIdx.x / kReduceParts, g = blockIdx.x % kReduceParts;
    const size_t ncell = (size_t)v.N * v.L;
    const double *cT = v.cellT + (size_t)t * ncell;
    const double *cO = v.cellO + (size_t)t * ncell;
    double s = 0.0;
    {
        // a thread's cells in its fixed order, eight loads in flight
        const size_t B = (size_t)blockDim.x * kReduceParts;
        size_t c = (size_t)g * blockDim.x + tid;
        for (; c + 3 * B < ncell; c += 4 * B)
        {
            const double a0 = cT[c], a1 = cT[c + B], a2 = cT[c + 2 * B], a3 = cT[c + 3 * B];
            const double b0 = cO[c], b1 = cO[c + B], b2 = cO[c + 2 * B], b3 = cO[c + 3 * B];
            s += __dadd_rn(a0, b0);
            s += __dadd_rn(a1, b1);
            s += __dadd_rn(a2, b2);
            s += __dadd_rn(a3, b3);
        }
        for (; c < ncell; c += B) s += __dadd_rn(cT[c], cO[c]);
    }
    const double llik = block_sum(s, s_red);
    double prior = 0.0;
    if (g == 0)
    {
        double q = 0.0;
        for (int i = tid; i < v.N; i += blockDim.x)
        {
            const size_t si = (size_t)t * v.N + i;
            q += ((v.pr_eps_neg[si] + v.pr_eps_pos[si]) + v.pr_r[si]) + v.pr_coi[si];
        }
        prior = block_sum(q, s_red);
    }
    if (tid == 0)
    {
        double *mine = part + ((size_t)t * kReduceParts + g) * 2;
        mine[0] = llik;
        if (g == 0) mine[1] = prior;
        __threadfence();
        if (atomicAdd(&done[t], 1u) == kReduceParts - 1)
        {
            __threadfence();
            double l = 0.0;
            for (int q = 0; q < kReduceParts; ++q) l += __ldcg(part + ((size_t)t * kReduceParts + q) * 2);
            v.llik[t] = l;
            v.prior[t] = __ldcg(part + (size_t)t * kReduceParts * 2 + 1) + v.hyper[t];
            done[t] = 0;  // ready for the next launch
        }
    }
}

// Chain::get_llik(sample) for one chain (src/chain.cpp:1083-1097).
__global__ void k_sample_llik(const View v, const int t, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.N) return;
    const size_t base = ((size_t)t * v.N + i) * v.L;
    double s = 0.0;
    for (int j = 0; j < v.L; ++j) s += __dadd_rn(v.cellT[base + j], v.cellO[base + j]);
    out[i] = s;
}

// ---- rebuild every cache from the parameters (initialize_likelihood, chain.cpp:1166-1223) ----
// chain_sel < 0: all chains (optionally only the `active` ones); else one chain.
__global__ void __launch_bounds__(kThreads)
    k_eval_cells(const View v, const int chain_sel, const int use_active)
{
    __shared__ double s_lgam[128];
    __shared__ double s_qs[kScratchSlots * (kThreads + 1)];
    const int tid = threadIdx.x;
    for (int i = tid; i < 128; i += kThreads) s_lgam[i] = v.lgam[i];
    __syncthreads();
    const size_t ncell = (size_t)v.N * v.L;
    const size_t total = (chain_sel < 0 ? (size_t)v.T : 1) * ncell;
    const size_t g = (size_t)blockIdx.x * kThreads + tid;
    const bool in_range = g < total;
    const int t = chain_sel < 0 ? (int)((in_range ? g : 0) / ncell) : chain_sel;
    const bool live = in_range && !(use_active && !v.active[t]);
    const size_t cell = (in_range ? g : 0) % ncell;
    const int i = (int)(cell / v.L), j = (int)(cell - (size_t)i * v.L);
    const size_t gi = (size_t)t * ncell + cell;
    const bool eval = live && !v.missing[cell];
    const unsigned callers = __ballot_sync(0xffffffffu, eval);
    if (!live) return;
    double T = 0.0, O = 0.0;
    if (eval)
    {
        const size_t si = (size_t)t * v.N + i;
        EpsLogs el;
        build_eps_logs(el, v.eps_neg[si], v.eps_pos[si], v.max_eps_neg, v.max_eps_pos, v.nall[j]);
        const Rel rel{v.log_r[si], v.log_1mr[si], v.odds[si]};
        const double *prow = v.p + ((size_t)t * v.L + j) * v.AP;
        double pv[kGather];
        gather_freqs(v.latent[gi], prow, pv);
        T = transmission_ll<false>(v.latent[gi], prow, pv, v.m[si], rel, v.allow_rel != 0, s_qs + tid,
                                   kThreads + 1, s_lgam, nullptr, callers, PamOut{});
        O = observation_ll(v.latent[gi], v.obs[cell], v.nall[j], el);
    }
    v.cellT[gi] = T;
    v.cellO[gi] = O;
}

__global__ void k_eval_priors(const View v, const int chain_sel, const int use_active)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (chain_sel < 0 ? (size_t)v.T : 1) * v.N;
    if (g >= total) return;
    const int t = chain_sel < 0 ? (int)(g / v.N) : chain_sel;
    if (use_active && !v.active[t]) return;
    const int i = (int)(g % v.N);
    const size_t si = (size_t)t * v.N + i;
    const double lambda = v.mean_coi[t];
    v.log_r[si] = log(v.r[si]);
    v.log_1mr[si] = log(1.0 - v.r[si]);
    v.odds[si] = v.r[si] / (1.0 - v.r[si]);
    v.pr_eps_neg[si] = dbeta_log(v.eps_neg[si], v.eps_neg_a, v.eps_neg_b, v.eps_neg_lbeta);
    v.pr_eps_pos[si] = dbeta_log(v.eps_pos[si], v.eps_pos_a, v.eps_pos_b, v.eps_pos_lbeta);
    v.pr_r[si] = dbeta_log(v.r[si], v.r_a, v.r_b, v.r_lbeta);
    const int mm = v.m[si];
    v.pr_coi[si] = (mm >= 0 && mm < 128)
                       ? dztpois_log(mm, log(lambda), log(exp(lambda) - 1.0), v.lgam)
                       : NAN;
    if (i == 0)
        v.hyper[t] = dgamma_log(lambda, v.mean_coi_shape, v.mean_coi_scale, v.lg_shape, v.l_scale);
}

// ---- starting state (Chain::initialize_parameters, src/chain.cpp:21-163, 1294-1310) ----------
// Gamma(shape, 1) by Marsaglia-Tsang, clamped like Sampler::rgamma (src/sampler.cpp:57-72).
__device__ inline double next_gamma(Stream &rng, double shape)
{
    double boost = 1.0;
    if (shape < 1.0)
    {
        boost = pow(rng.next_uniform(), 1.0 / shape);
        shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double x = 1.0;
    for (int guard = 0; guard < 1000; ++guard)
    {
        const double z = next_normal(rng);
        double w = 1.0 + c * z;
        if (w <= 0.0) continue;
        w = w * w * w;
        const double u = rng.next_uniform();
        if (log(u) < 0.5 * z * z + d - d * w + d * log(w))
        {
            x = d * w;
            break;
        }
    }
    x *= boost;
    return fmin(fmax(x, 1e-100), 1e100);
}

__global__ void k_init_samples(const View v, const int attempt)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)v.T * v.N) return;
    const int t = (int)(g / v.N), i = (int)(g % v.N);
    if (!v.active[t]) return;
    const size_t si = g;
    Stream rng;
    rng.open(v.seed, 0xFFFF0000u + (uint32_t)attempt, KIND_INIT_SAMPLE,
             (unsigned)(v.chain_offset + t), (uint32_t)i, 0);
    // initialize_m (src/chain.cpp:77-96): observed COI + geometric(1/4) jitter, never below it
    const int coi = v.obs_coi[i];
    const int lower = min(max(coi, 1), v.max_coi);
    const int jit = coi + next_coi_delta(rng, -0.28768207245178093 /* log(3/4) */);
    v.m[si] = min(max(jit, lower), v.max_coi);
    v.eps_neg[si] = rng.next_uniform() * .1;
    v.eps_pos[si] = rng.next_uniform() * .1;
    v.r[si] = v.allow_rel ? rng.next_uniform() * .99 : 0.0;
    v.sd_eps_neg[si] = v.sd_eps_pos[si] = v.sd_r[si] = v.sd_mr[si] = 1.0;
    v.acc_eps_neg[si] = v.acc_eps_pos[si] = v.acc_m[si] = 0;
    v.acc_r[si] = v.acc_mr[si] = v.acc_samples[si] = 0;
}

__global__ void k_init_latent(const View v, const int attempt)
{
    const size_t ncell = (size_t)v.N * v.L;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)v.T * ncell) return;
    const int t = (int)(g / ncell);
    if (!v.active[t]) return;
    const size_t cell = g % ncell;
    const int i = (int)(cell / v.L), j = (int)(cell - (size_t)i * v.L);
    Stream rng;
    rng.open(v.seed, 0xFFFF0000u + (uint32_t)attempt, KIND_INIT_SAMPLE,
             (unsigned)(v.chain_offset + t), (uint32_t)i, (uint32_t)(j + 1));
    EpsLogs el;
    build_eps_logs(el, .2, .2, v.max_eps_neg, v.max_eps_pos, v.nall[j]);  // src/chain.cpp:39
    double lp;
    v.latent[g] = sample_latent_genotype(rng, v.obs[cell], v.nall[j], v.m[(size_t)t * v.N + i], el,
                                         v.logC, &lp);
    v.lgadj[g] = lp;
}

__global__ void k_init_p(const View v, const int attempt)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)v.T * v.L) return;
    const int t = (int)(g / v.L), j = (int)(g % v.L);
    if (!v.active[t]) return;
    const int A = v.nall[j];
    const size_t row = g * v.AP;
    // initialize_p: Dirichlet(10 * 1) via gamma draws (src/chain.cpp:51-75, sampler.cpp:79-98)
    double tot = 0.0;
    for (int a = 0; a < v.AP; ++a)
    {
        double x = 0.0;
        if (a < A)
        {
            Stream rng;
            rng.open(v.seed, 0xFFFF0000u + (uint32_t)attempt, KIND_INIT_P,
                     (unsigned)(v.chain_offset + t), (uint32_t)j, (uint32_t)a);
            x = next_gamma(rng, 10.0);
        }
        v.p[row + a] = x;
        tot += x;
        v.p_sd[row + a] = 1.0;
        v.p_acc[row + a] = 0;
        v.p_att[row + a] = 0;
    }
    const double inv = 1.0 / tot;
    for (int a = 0; a < A; ++a) v.p[row + a] *= inv;
    if (j == 0 && attempt == 0)
    {
        // Chain::Chain: mean_coi = Gamma(shape, scale) + 1 (src/chain.cpp:1300-1304) -- drawn once
        // in the constructor; the retries of MCMC::MCMC (src/mcmc.cpp:92) call
        // initialize_parameters(), which leaves it alone
        Stream rng;
        rng.open(v.seed, 0xFFFF0000u + (uint32_t)attempt, KIND_INIT_MEAN_COI,
                 (unsigned)(v.chain_offset + t), 0, 0);
        const double g0 = next_gamma(rng, v.mean_coi_shape) * v.mean_coi_scale;
        v.mean_coi[t] = fmin(fmax(g0, 1e-100), 1e100) + 1.0;
        v.mean_coi_sd[t] = 1.0;
        v.mean_coi_acc[t] = 0.0;
    }
}
}  // namespace moire
