// The per-sample updates that need transmission terms (update_m, update_r, update_eff_coi,
// update_samples; src/chain.cpp:165-374, 697-807) as a flat pipeline, like update_p (p_flat.cuh):
//   k_sf_propose  one block per (chain, sample tile): the proposals, the resampled latent
//                 genotypes (own Philox stream per cell), the observation terms, the cost-class
//                 histogram
//   k_pf_scan / k_sf_scatter   device-wide counting sort of the cells of valid proposals
//   k_flat_eval   the transmission terms, uniform warps, no barriers            (p_flat.cuh)
//   k_sf_decide   one warp per (chain, sample): the sample's sums, the Metropolis decision, commit
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
        entry = make_uint4((unsigned)lat, (unsigned)(lat >> 32), t * v.L + j, tn);
        return true;
    };
    block_scatter(total, cell_fn, pf.hist + kSortBins, pf.list);
}

// One warp per (chain, sample), no block barriers: the warp sums the sample's cells (lanes across
// loci, then the shuffle tree -- the order of the fused kernel), lane 0 decides, the warp commits.
template <int KIND>
__global__ void __launch_bounds__(256) k_sf_decide(const View v, const SFlat sf, const int iteration)
{
    constexpr bool RESAMPLE = (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES);
    const int lane = threadIdx.x & 31;
    const size_t si = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (si >= (size_t)v.T * v.N) return;
    const int N = v.N, L = v.L;
    const int t = (int)(si / N), i = (int)(si - (size_t)t * N);
    const size_t base = si * L, data_base = (size_t)i * L;
    const bool valid = sf.prop[si].valid != 0;
    double d = 0.0, a = 0.0;
    if (valid)
    {
        for (int j = lane; j < L; j += 32)
        {
            const size_t gi = base + j;
            if (RESAMPLE) a += sf.adj[gi] - v.lgadj[gi];
            if (v.missing[data_base + j] == 0)
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
    int accept = 0;
    if (lane == 0)
    {
        Proposal P = sf.prop[si];
        decide_sample<KIND>(v, P, t, i, iteration, d, a);
        accept = P.accept;
    }
    accept = __shfl_sync(0xffffffffu, accept, 0);
    if (!accept) return;
    for (int j = lane; j < L; j += 32)
    {
        const size_t gi = base + j;
        if (RESAMPLE)
        {
            v.latent[gi] = sf.lat[gi];
            v.lgadj[gi] = sf.adj[gi];
        }
        if (v.missing[data_base + j] == 0)
        {
            v.cellT[gi] = sf.Tn[gi];
            if (RESAMPLE) v.cellO[gi] = sf.O[gi];
        }
    }
}
// ---- update_eps_neg / update_eps_pos (src/chain.cpp:267-374): no transmission term, so the whole
// update is one kernel; one warp per (chain, sample), no block barriers.  Lane 0 draws the
// proposal, lanes across the distinct allele counts build the observation logs of the proposed
// rates (per-warp shared memory), lanes across loci recompute the observation terms.  Streams,
// arithmetic and summation order are those of the fused k_update_sample<KIND>.
constexpr int kEpsWarps = 4;
__host__ __device__ inline size_t eps_smem(int nA) { return (size_t)kEpsWarps * nA * sizeof(EpsLogs); }

template <int KIND>
__global__ void __launch_bounds__(kEpsWarps * 32) k_eps_update(const View v, const int iteration)
{
    static_assert(KIND == KIND_EPS_NEG || KIND == KIND_EPS_POS, "error-rate updates only");
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t si = (size_t)blockIdx.x * kEpsWarps + warp;
    if (si >= (size_t)v.T * v.N) return;
    const int N = v.N, L = v.L, nA = v.nA;
    const int t = (int)(si / N), i = (int)(si - (size_t)t * N);
    EpsLogs *el = reinterpret_cast<EpsLogs *>(smem) + (size_t)warp * nA;
    Proposal P = {};
    if (lane == 0) P = draw_sample_proposal<KIND>(v, t, i, iteration, v.lgam);
    const int valid = __shfl_sync(0xffffffffu, lane == 0 ? P.valid : 0, 0);
    double d = 0.0;
    const size_t base = si * L, data_base = (size_t)i * L;
    if (valid)
    {
        const double en = __shfl_sync(0xffffffffu, P.en_new, 0);
        const double ep = __shfl_sync(0xffffffffu, P.ep_new, 0);
        for (int a = lane; a < nA; a += 32)
            build_eps_logs(el[a], en, ep, v.max_eps_neg, v.max_eps_pos, v.avals[a]);
        __syncwarp();
        unsigned n_eval = 0;
        for (int j = lane; j < L; j += 32)
        {
            if (v.missing[data_base + j] != 0) continue;
            const size_t gi = base + j;
            const double oldT = v.cellT[gi], oldO = v.cellO[gi];
            const double O = observation_ll(v.latent[gi], v.obs[data_base + j], v.nall[j], el[v.aclass[j]]);
            d += __dadd_rn(oldT, O) - __dadd_rn(oldT, oldO);
            ++n_eval;
        }
        n_eval = __reduce_add_sync(0xffffffffu, n_eval);
        if (lane == 0 && n_eval) atomicAdd(v.evals + KIND, (unsigned long long)n_eval);
    }
    d = warp_sum(d);
    int accept = 0;
    if (lane == 0)
    {
        decide_sample<KIND>(v, P, t, i, iteration, d, 0.0);
        accept = P.accept;
    }
    accept = __shfl_sync(0xffffffffu, accept, 0);
    if (!accept) return;
    for (int j = lane; j < L; j += 32)
    {
        if (v.missing[data_base + j] != 0) continue;
        const size_t gi = base + j;
        v.cellO[gi] = observation_ll(v.latent[gi], v.obs[data_base + j], v.nall[j], el[v.aclass[j]]);
    }
}
}  // namespace moire
