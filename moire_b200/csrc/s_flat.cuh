// The per-sample updates (update_eps_neg, update_eps_pos, update_m, update_r, update_eff_coi,
// update_samples; src/chain.cpp:165-374, 697-807) as flat pipelines, like update_p (p_flat.cuh):
//   k_sf_draw     one thread per (chain, sample): the proposal (serial scalar math, so it runs
//                 lanes-across-samples and all samples at once), then the observation logs of the
//                 proposed error rates per (sample, distinct allele count)
//   k_sf_cells    one thread per cell: the resampled latent genotype (own Philox stream per
//                 cell), the observation term, the cost-class histogram
//   k_pf_scan / k_sf_scatter   device-wide counting sort of the cells of valid proposals
//   k_flat_eval   the transmission terms, uniform warps, no barriers            (p_flat.cuh)
//   k_sf_decide   one warp per (chain, sample): the sample's sums, the Metropolis decision, commit
//   k_eps_cells   the error-rate updates have no transmission term: cells, decision and commit in
//                 one warp-per-sample kernel after k_sf_draw
// No kernel has a serial phase behind a block barrier (ncu, tile-per-block version: the ten
// proposal lanes of a 256-thread block kept the other warps at the barrier for ~20 us per tile,
// 37 % of the stall samples).  Streams, arithmetic and summation order are those of the fused
// k_update_sample, which stays in use for tiny problems.
#pragma once
#include "p_flat.cuh"

namespace moire
{
struct SFlat
{
    Proposal *prop;                  // [T * N]
    int *m_new;                      // [T * N] proposed COI
    SampleScal *ss;                  // [T * N] the proposal's (m, r) as the evaluator gathers it
    EpsLogs *elogs;                  // [T * N][nA] logs of the proposed error rates
    mask_t *lat;                     // [T][N][L] proposed latent genotypes
    double *adj;                     // [T][N][L] their log proposal densities
    double *O, *Tn;                  // [T][N][L] observation / transmission terms of the proposal
};

constexpr int kDrawThreads = 128;

// A block draws the proposals of `spb` consecutive (chain, sample) units -- one thread each, a serial
// chain of ~15 double-precision transcendentals -- and then builds the observation logs of the
// proposed error rates, one thread per (unit, distinct allele count).  spb = kDrawThreads / nA
// (sf_draw_units_per_block), so the second phase is one pair per thread whatever the number of
// distinct allele counts (the Namibia shape has 20: with 128 units per block a thread built 20 sets of
// logs in turn, 65 us per launch where the benchmark shape, one allele count, takes 15).
__host__ __device__ inline int sf_draw_units_per_block(int nA, bool logs)
{
    return logs ? max(1, kDrawThreads / max(1, nA)) : kDrawThreads;
}
template <int KIND>
__global__ void __launch_bounds__(kDrawThreads) k_sf_draw(const View v, const SFlat sf, const int iteration_arg)
{
    pdl_enter();
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
    constexpr bool LOGS = KIND != KIND_R;  // the kinds that recompute observation terms
    __shared__ double s_en[kDrawThreads], s_ep[kDrawThreads];
    __shared__ int s_valid[kDrawThreads];
    const int tid = threadIdx.x, nA = v.nA;
    const int spb = sf_draw_units_per_block(nA, LOGS);
    const size_t total = (size_t)v.T * v.N;
    const size_t si0 = (size_t)blockIdx.x * spb, si = si0 + tid;
    if (tid < spb && si < total)
    {
        const int t = (int)(si / v.N), i = (int)(si - (size_t)t * v.N);
        const Proposal P = draw_sample_proposal<KIND>(v, t, i, iteration, v.lgam);
        sf.prop[si] = P;
        sf.m_new[si] = P.m_new;
        SampleScal sc;
        sc.log_r = P.log_r;
        sc.log_1mr = P.log_1mr;
        sc.odds = P.odds;
        sc.m = P.m_new;
        sc.pad = 0;
        sf.ss[si] = sc;
        s_en[tid] = P.en_new;
        s_ep[tid] = P.ep_new;
        s_valid[tid] = P.valid;
    }
    if (!LOGS) return;
    __syncthreads();
    const int n = (int)min((size_t)spb, total - si0);
    for (int e = tid; e < n * nA; e += kDrawThreads)
    {
        const int sl = e / nA, a = e - sl * nA;
        if (s_valid[sl])
            build_eps_logs(sf.elogs[(si0 + sl) * nA + a], s_en[sl], s_ep[sl], v.max_eps_neg,
                           v.max_eps_pos, v.avals[a]);
    }
}

constexpr int kCellsPerThread = 4;
__host__ __device__ inline unsigned sf_cell_blocks(size_t total)
{
    return (unsigned)((total + 256 * kCellsPerThread - 1) / (256 * kCellsPerThread));
}

template <int KIND>
__global__ void __launch_bounds__(256, 4) k_sf_cells(const View v, const SFlat sf, const PFlat pf, const int iteration_arg)
{
    pdl_enter();
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
    constexpr bool RESAMPLE = (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES);
    __shared__ unsigned int s_hist[kSortBins];
    __shared__ unsigned int s_evals;
    const int tid = threadIdx.x;
    for (int b = tid; b < kSortBins; b += 256) s_hist[b] = 0;
    if (tid == 0) s_evals = 0;
    __syncthreads();
    const unsigned N = (unsigned)v.N, L = (unsigned)v.L, nA = (unsigned)v.nA;
    const size_t total = (size_t)v.T * N * L;
    unsigned my_evals = 0;
#pragma unroll 1
    for (int c = 0; c < kCellsPerThread; ++c)
    {
        const size_t gi = ((size_t)blockIdx.x * kCellsPerThread + c) * 256 + tid;
        if (gi >= total) break;
        const unsigned tn = (unsigned)(gi / L), j = (unsigned)(gi - (size_t)tn * L);
        if (!sf.prop[tn].valid) continue;
        const unsigned t = tn / N, i = tn - t * N;
        const size_t di = (size_t)i * L + j;
        const bool miss = v.missing[di] != 0;
        const int m_new = sf.m_new[tn];
        mask_t lat;
        if (RESAMPLE)
        {
            const mask_t obs = v.obs[di];
            const int A = v.nall[j];
            const EpsLogs el = sf.elogs[(size_t)tn * nA + v.aclass[j]];
            Stream crng;
            crng.open(v.seed, (uint32_t)iteration, KIND, (unsigned)(v.chain_offset + t), i, j + 1u);
            double lp;
            lat = sample_latent_genotype(crng, obs, A, m_new, el, v.logC, &lp);
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
            atomicAdd(&s_hist[work_class_flat(mpopc(lat), m_new)], 1u);
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
    pdl_enter();
    constexpr bool RESAMPLE = (KIND == KIND_M || KIND == KIND_EFF_COI || KIND == KIND_SAMPLES);
    const size_t total = (size_t)v.T * v.N * v.L;
    auto cell_fn = [&](size_t gi, int &cls, ListEntry &entry) -> bool {
        const unsigned tn = (unsigned)(gi / v.L), j = (unsigned)(gi - (size_t)tn * v.L);
        const unsigned t = tn / v.N, i = tn - t * v.N;
        if (!sf.prop[tn].valid || v.missing[(size_t)i * v.L + j] != 0) return false;
        const mask_t lat = RESAMPLE ? sf.lat[gi] : v.latent[gi];
        cls = work_class_flat(mpopc(lat), sf.m_new[tn]);
        entry = make_entry(lat, (t << 16) | j, i);
        return true;
    };
    block_scatter(total, cell_fn, pf.hist + kSortBins, pf.list);
}

// One warp per (chain, sample), no block barriers: the warp sums the sample's cells (lanes across
// loci, then the shuffle tree -- the order of the fused kernel), lane 0 decides, the warp commits.
template <int KIND>
__global__ void __launch_bounds__(256) k_sf_decide(const View v, const SFlat sf, const int iteration_arg)
{
    pdl_enter();
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
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
        // four loci per lane in flight (the loads of a chunk are issued before its sums).  Tried:
        // every load issued unconditionally, ahead of the validity and missing flags -- 74 registers
        // instead of 48, fewer resident warps, 68 -> 91 us (profiles/README.md, r3k)
        for (int j0 = 0; j0 < L; j0 += 128)
        {
            double vT[4], vO[4], vTn[4], vOn[4], va[4], vb[4];
            bool use[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                const int j = j0 + 32 * u + lane;
                const bool in = j < L;
                use[u] = in && v.missing[data_base + (in ? j : 0)] == 0;
                const size_t gi = base + (in ? j : 0);
                va[u] = (RESAMPLE && in) ? sf.adj[gi] : 0.0;
                vb[u] = (RESAMPLE && in) ? v.lgadj[gi] : 0.0;
                vT[u] = use[u] ? v.cellT[gi] : 0.0;
                vO[u] = use[u] ? v.cellO[gi] : 0.0;
                vTn[u] = use[u] ? sf.Tn[gi] : 0.0;
                vOn[u] = (RESAMPLE && use[u]) ? sf.O[gi] : vO[u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                if (RESAMPLE && j0 + 32 * u + lane < L) a += va[u] - vb[u];
                if (use[u]) d += __dadd_rn(vTn[u], vOn[u]) - __dadd_rn(vT[u], vO[u]);
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
// ---- update_eps_neg / update_eps_pos (src/chain.cpp:267-374) after k_sf_draw: one warp per
// (chain, sample) recomputes the observation terms under the proposed rate (lanes across loci, four
// loci per lane in flight), lane 0 decides, the warp commits.
template <int KIND>
__global__ void __launch_bounds__(256) k_eps_cells(const View v, const SFlat sf, const int iteration_arg)
{
    pdl_enter();
    const int iteration = iteration_arg >= 0 ? iteration_arg : *v.iter;
    static_assert(KIND == KIND_EPS_NEG || KIND == KIND_EPS_POS, "error-rate updates only");
    const int lane = threadIdx.x & 31;
    const size_t si = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (si >= (size_t)v.T * v.N) return;
    const int N = v.N, L = v.L, nA = v.nA;
    const int t = (int)(si / N), i = (int)(si - (size_t)t * N);
    const size_t base = si * L, data_base = (size_t)i * L;
    const int valid_word = sf.prop[si].valid;
    const EpsLogs *el = sf.elogs + si * nA;
    double d = 0.0;
    unsigned n_eval = 0;
    {
        // as in k_sf_decide: every load of a chunk is issued before anything that depends on one
        for (int j0 = 0; j0 < L; j0 += 128)
        {
            double vT[4], vO[4];
            mask_t vl[4], vo[4];
            int vA[4], vc[4];
            unsigned char vm[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                const int j = j0 + 32 * u + lane;
                const int jj = j < L ? j : 0;
                vm[u] = v.missing[data_base + jj];
                vT[u] = v.cellT[base + jj];
                vO[u] = v.cellO[base + jj];
                vl[u] = v.latent[base + jj];
                vo[u] = v.obs[data_base + jj];
                vA[u] = v.nall[jj];
                vc[u] = v.aclass[jj];
            }
            const bool valid = valid_word != 0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (valid && j0 + 32 * u + lane < L && vm[u] == 0)
                {
                    const double O = observation_ll(vl[u], vo[u], vA[u], el[vc[u]]);
                    d += __dadd_rn(vT[u], O) - __dadd_rn(vT[u], vO[u]);
                    ++n_eval;
                }
        }
        n_eval = __reduce_add_sync(0xffffffffu, n_eval);
        if (lane == 0 && n_eval) atomicAdd(v.evals + KIND, (unsigned long long)n_eval);
    }
    d = warp_sum(d);
    int accept = 0;
    if (lane == 0)
    {
        Proposal P = sf.prop[si];
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
