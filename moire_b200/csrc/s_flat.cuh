// The per-sample updates that need transmission terms (update_m, update_r, update_eff_coi,
// update_samples; src/chain.cpp:165-374, 697-807) as a flat pipeline, like update_p (p_flat.cuh):
//   k_sf_propose  one block per (chain, sample tile): the proposals, the resampled latent
//                 genotypes (own Philox stream per cell), the observation terms, the cost-class
//                 histogram
//   k_pf_scan / k_sf_scatter   device-wide counting sort of the cells of valid proposals
//   k_flat_eval   the transmission terms, uniform warps, no barriers            (p_flat.cuh)
//   k_sf_decide   one block per (chain, sample tile): per-sample sums in locus order, the
//                 Metropolis decisions, commit
// Streams, arithmetic and summation order are those of the fused k_update_sample, which stays in
// use for the two error-rate updates (no transmission term) and for tiny problems.
#pragma once
#include "p_flat.cuh"

namespace moire
{
struct SFlat
{
    Proposal *prop;                  // [T * N]
    int *m_new;                      // [T * N] proposed COI
    double *log_r_new, *log_1mr_new; // [T * N]
    unsigned long long *lat;         // [T][N][L] proposed latent genotypes
    double *adj;                     // [T][N][L] their log proposal densities
    double *O, *Tn;                  // [T][N][L] observation / transmission terms of the proposal
};

__host__ __device__ inline size_t sf_smem(int S, int nA)
{
    return 128 * sizeof(double) + (size_t)S * nA * sizeof(EpsLogs) + (size_t)S * sizeof(Proposal) +
           kSortBins * sizeof(unsigned int);
}

template <int KIND>
__global__ void __launch_bounds__(256)
    k_sf_propose(const View v, const SFlat sf, const PFlat pf, const int iteration, const int S,
                 const int tiles_per_chain)
{
    constexpr bool RESAMPLE = (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES);
    extern __shared__ double smem[];
    double *s_lgam = smem;
    EpsLogs *s_elogs = reinterpret_cast<EpsLogs *>(s_lgam + 128);
    Proposal *s_prop = reinterpret_cast<Proposal *>(s_elogs + S * v.nA);
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(s_prop + S);
    __shared__ unsigned int s_evals;
    const int tid = threadIdx.x;
    const int t = blockIdx.x / tiles_per_chain;
    const int s0 = (blockIdx.x - t * tiles_per_chain) * S;
    const int ns = min(S, v.N - s0);
    const int N = v.N, L = v.L, nA = v.nA;
    const unsigned gchain = (unsigned)(v.chain_offset + t);
    for (int i = tid; i < 128; i += 256) s_lgam[i] = v.lgam[i];
    for (int b = tid; b < kSortBins; b += 256) s_hist[b] = 0;
    if (tid == 0) s_evals = 0;
    __syncthreads();
    if (tid < ns)
    {
        const Proposal P = draw_sample_proposal<KIND>(v, t, s0 + tid, iteration, s_lgam);
        const size_t si = (size_t)t * N + s0 + tid;
        s_prop[tid] = P;
        sf.prop[si] = P;
        sf.m_new[si] = P.m_new;
        sf.log_r_new[si] = P.log_r;
        sf.log_1mr_new[si] = P.log_1mr;
    }
    __syncthreads();
    if (RESAMPLE)
    {
        for (int e = tid; e < ns * nA; e += 256)
        {
            const int sl = e / nA, a = e - sl * nA;
            build_eps_logs(s_elogs[e], s_prop[sl].en_new, s_prop[sl].ep_new, v.max_eps_neg,
                           v.max_eps_pos, v.avals[a]);
        }
        __syncthreads();
    }
    const size_t tile_base = ((size_t)t * N + s0) * L;
    const size_t data_base = (size_t)s0 * L;
    const int ncell = ns * L;
    unsigned my_evals = 0;
    for (int cell = tid; cell < ncell; cell += 256)
    {
        const int sl = cell / L, j = cell - sl * L;
        const Proposal &P = s_prop[sl];
        if (!P.valid) continue;
        const size_t gi = tile_base + cell;
        const bool miss = v.missing[data_base + cell] != 0;
        unsigned long long lat;
        if (RESAMPLE)
        {
            const unsigned long long obs = v.obs[data_base + cell];
            const int A = v.nall[j];
            const EpsLogs &el = s_elogs[sl * nA + v.aclass[j]];
            Stream crng;
            crng.open(v.seed, (uint32_t)iteration, KIND, gchain, (uint32_t)(s0 + sl), (uint32_t)(j + 1));
            double lp;
            lat = sample_latent_genotype(crng, obs, A, P.m_new, el, v.logC, &lp);
            sf.lat[gi] = lat;
            sf.adj[gi] = lp;
            if (!miss) sf.O[gi] = observation_ll(lat, obs, A, el);
        }
        else
        {
            lat = v.latent[gi];
        }
        if (!miss)
        {
            atomicAdd(&s_hist[work_class(__popcll(lat), P.m_new)], 1u);
            ++my_evals;
        }
    }
    if (my_evals) atomicAdd(&s_evals, my_evals);
    __syncthreads();
    for (int b = tid; b < kSortBins; b += 256)
        if (s_hist[b]) atomicAdd(&pf.hist[b], s_hist[b]);
    if (tid == 0 && s_evals) atomicAdd(v.evals + KIND, (unsigned long long)s_evals);
}

template <int KIND>
__global__ void __launch_bounds__(256) k_sf_scatter(const View v, const SFlat sf, const PFlat pf)
{
    constexpr bool RESAMPLE = (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES);
    const size_t total = (size_t)v.T * v.N * v.L;
    auto cell_fn = [&](size_t gi, int &cls, uint4 &entry) -> bool {
        const unsigned tn = (unsigned)(gi / v.L), j = (unsigned)(gi - (size_t)tn * v.L);
        const unsigned t = tn / v.N, i = tn - t * v.N;
        if (!sf.prop[tn].valid || v.missing[(size_t)i * v.L + j] != 0) return false;
        const unsigned long long lat = RESAMPLE ? sf.lat[gi] : v.latent[gi];
        cls = work_class(__popcll(lat), sf.m_new[tn]);
        entry = make_uint4((unsigned)gi, t * v.L + j, tn, (unsigned)gi);
        return true;
    };
    block_scatter(total, cell_fn, pf.hist + kSortBins, pf.list);
}

template <int KIND>
__global__ void __launch_bounds__(256)
    k_sf_decide(const View v, const SFlat sf, const int iteration, const int S,
                const int tiles_per_chain)
{
    constexpr bool RESAMPLE = (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES);
    extern __shared__ double smem[];
    Proposal *s_prop = reinterpret_cast<Proposal *>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x / tiles_per_chain;
    const int s0 = (blockIdx.x - t * tiles_per_chain) * S;
    const int ns = min(S, v.N - s0);
    const int N = v.N, L = v.L;
    if (tid < ns) s_prop[tid] = sf.prop[(size_t)t * N + s0 + tid];
    __syncthreads();
    const size_t tile_base = ((size_t)t * N + s0) * L;
    const size_t data_base = (size_t)s0 * L;
    // one warp per sample sums its cells (in locus order) and decides
    for (int sl = warp; sl < ns; sl += 8)
    {
        const Proposal &Pc = s_prop[sl];
        double d = 0.0, a = 0.0;
        if (Pc.valid)
        {
            for (int j = lane; j < L; j += 32)
            {
                const int cell = sl * L + j;
                const size_t gi = tile_base + cell;
                if (RESAMPLE) a += sf.adj[gi] - v.lgadj[gi];
                if (v.missing[data_base + cell] == 0)
                {
                    const double oldT = v.cellT[gi], oldO = v.cellO[gi];
                    const double T = sf.Tn[gi];
                    const double O = RESAMPLE ? sf.O[gi] : oldO;
                    d += __dadd_rn(T, O) - __dadd_rn(oldT, oldO);
                }
            }
        }
        d = warp_sum(d);
        if (RESAMPLE) a = warp_sum(a);
        if (lane == 0) decide_sample<KIND>(v, s_prop[sl], t, s0 + sl, iteration, d, a);
    }
    __syncthreads();
    // commit the accepted samples' cells
    const int ncell = ns * L;
    for (int cell = tid; cell < ncell; cell += 256)
    {
        const int sl = cell / L;
        if (!s_prop[sl].accept) continue;
        const size_t gi = tile_base + cell;
        if (RESAMPLE)
        {
            v.latent[gi] = sf.lat[gi];
            v.lgadj[gi] = sf.adj[gi];
        }
        if (v.missing[data_base + cell] == 0)
        {
            v.cellT[gi] = sf.Tn[gi];
            if (RESAMPLE) v.cellO[gi] = sf.O[gi];
        }
    }
}
}  // namespace moire
