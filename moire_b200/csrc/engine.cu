// libmoire_b200.so: engine object + C ABI (include/moire_engine.h).
//
// Host-side responsibilities only: own the device allocations, lay the data out, launch the
// kernels of kernels.cuh on one stream, move small results.  No arithmetic of the likelihood path
// runs on the host: there is no CPU fallback (a missing device is an error).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/moire_engine.h"
#include "kernels.cuh"
#include "p_flat.cuh"
#include "s_flat.cuh"

namespace
{
thread_local std::string g_last_error;

struct CudaError
{
    std::string msg;
};

#define CUDA_TRY(expr)                                                                        \
    do                                                                                        \
    {                                                                                         \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess)                                                             \
            throw CudaError{std::string(#expr) + ": " + cudaGetErrorString(err__) + " (" +    \
                            __FILE__ + ":" + std::to_string(__LINE__) + ")"};                 \
    } while (0)

struct InvalidArg
{
    std::string msg;
};

// Device memory comes from a small caching layer: an engine's ~60 buffers (0.7 GB at C3) are
// handed back to a per-process free list instead of cudaFree -- which synchronises the device and
// took 5..300 ms for the lot, inside every run_mcmc call -- and the next engine of the same shape
// (the next of `num_chains` runs, the next call) takes them from there.  Blocks are matched by
// (device, exact size); moire_engine_trim_memory() returns the list to the driver.
struct DeviceBlockCache
{
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_blocks;
    size_t cached_bytes = 0;
    static constexpr size_t kMaxCachedBytes = (size_t)32 << 30;

    void *take(size_t bytes)
    {
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = free_blocks.find({dev, bytes});
            if (it != free_blocks.end())
            {
                void *p = it->second;
                free_blocks.erase(it);
                cached_bytes -= bytes;
                return p;
            }
        }
        void *p = nullptr;
        const cudaError_t rc = cudaMalloc(&p, bytes);
        if (rc == cudaErrorMemoryAllocation)
        {
            cudaGetLastError();
            trim();  // cached blocks of other shapes may be what is in the way
            CUDA_TRY(cudaMalloc(&p, bytes));
        }
        else
            CUDA_TRY(rc);
        return p;
    }
    void give(void *p, size_t bytes)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cached_bytes + bytes > kMaxCachedBytes)
        {
            cudaFree(p);
            return;
        }
        std::lock_guard<std::mutex> lock(mu);
        free_blocks.insert({{dev, bytes}, p});
        cached_bytes += bytes;
    }
    void trim()
    {
        std::lock_guard<std::mutex> lock(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto &kv : free_blocks)
        {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second);
        }
        cudaSetDevice(cur);
        free_blocks.clear();
        cached_bytes = 0;
    }
};
DeviceBlockCache g_block_cache;

template <class T>
struct DevBuf
{
    T *ptr = nullptr;
    size_t n = 0;
    size_t bytes() const { return std::max<size_t>(n, 1) * sizeof(T); }
    void alloc(size_t count)
    {
        n = count;
        ptr = static_cast<T *>(g_block_cache.take(bytes()));
        CUDA_TRY(cudaMemset(ptr, 0, bytes()));
    }
    void upload(const T *src, size_t count, cudaStream_t s)
    {
        CUDA_TRY(cudaMemcpyAsync(ptr, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void release()
    {
        if (ptr) g_block_cache.give(ptr, bytes());
        ptr = nullptr;
    }
};

// The reference's subset order for every k <= 10 (CombinationIndicesGenerator driven by
// src/prob_any_missing.cpp:20-52): by size, then lexicographic index tuples.  See likelihood.cuh.
void build_subset_tables(std::vector<unsigned short> &tail, std::vector<unsigned short> &mask)
{
    tail.assign(moire::kSubsetEntries, 0);
    mask.assign(moire::kSubsetEntries, 0);
    for (int k = 1; k <= moire::kIeMax; ++k)
    {
        int idx = moire::subset_block(k);
        for (int s = 1; s <= k; ++s)
        {
            std::vector<int> c(s), prev;
            for (int i = 0; i < s; ++i) c[i] = i;
            for (;;)
            {
                int d = 0;
                if (!prev.empty())
                    while (d < s - 1 && c[d] == prev[d]) ++d;
                unsigned t = 0, full = 0;
                for (int i = 0; i < s; ++i)
                {
                    full |= 1u << c[i];
                    if (i >= d) t |= 1u << c[i];
                }
                tail[idx] = (unsigned short)(t | (unsigned)d << 10 | (s % 2 == 0 ? 0x4000u : 0u));
                mask[idx] = (unsigned short)full;
                ++idx;
                prev = c;
                int i = s - 1;  // rightmost position that can still move right
                while (i >= 0 && c[i] == k - s + i) --i;
                if (i < 0) break;
                ++c[i];
                for (int j = i + 1; j < s; ++j) c[j] = c[j - 1] + 1;
            }
        }
    }
}

// Row m of the EGF recurrence's binomials exactly as the reference produces them
// (src/prob_any_missing.cpp:81-88): comb = m; comb *= (m - j) / (j + 1).  [j][m] layout.
void build_egf_binomials(std::vector<double> &tab)
{
    const int D = moire::kMaxCoi + 1;
    tab.assign((size_t)D * D, 0.0);
    for (int m = 1; m < D; ++m)
    {
        volatile double comb = (double)m;
        for (int j = 1; j <= m; ++j)
        {
            tab[(size_t)j * D + m] = comb;
            const volatile double ratio = (double)(m - j) / (double)(j + 1);
            comb = comb * ratio;
        }
    }
}

double log_binom_exact(unsigned n, unsigned k)
{
    // boost::math::binomial_coefficient<double> returns C(n,k) as an integer-valued double
    // (src/sampler.cpp:270,305,327,330 take its log).
    if (k > n) return -INFINITY;
    if (k > n - k) k = n - k;
    long double c = 1.0L;
    for (unsigned i = 1; i <= k; ++i)
    {
        c = c * (long double)(n - k + i) / (long double)i;
        c = floorl(c + 0.5L);
    }
    return std::log((double)c);
}
}  // namespace

__global__ void k_set_iteration(int *dst, const int iteration) { *dst = iteration; }

// The subset order of k latent alleles as the kernels see it, decoded on the device from each of
// the three places it lives: 0 = c_sub_tail (the one-thread table walk: depth + tail against a
// stack of prefixes, exactly the control flow of ie_vectorized with element masks in place of the
// partial differences), 1 = g_sub_mask (the warp-cooperative evaluator), 2 = SmallSubsets<K> (the
// compile-time order of the unrolled walks, k <= MOIRE_SMALL_K).  out[i] = element mask of the
// i-th subset, bit 15 set when the table gives its term a minus sign.
// What MCMC::sample stores of the hot chain (src/mcmc.cpp:327-347), gated and addressed by
// device-resident words so that it can be enqueued every sampling step without the host knowing
// which chain (if any) is hot: sel[0] = local chain index or -1, sel[1] = slot of the store.
__global__ void __launch_bounds__(256) k_record_draw(const moire::View v, const int *sel, const moire_draw_store st)
{
    const int t = sel[0], slot = sel[1];
    if (t < 0 || t >= v.T || slot < 0 || slot >= st.capacity) return;
    const size_t N = v.N, L = v.L, LA = L * v.AP, NL = N * L;
    const size_t g0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = g0; i < LA; i += stride) st.p[slot * LA + i] = v.p[t * LA + i];
    for (size_t i = g0; i < N; i += stride)
    {
        const size_t si = t * N + i;
        st.m[slot * N + i] = v.m[si];
        st.eps_neg[slot * N + i] = v.eps_neg[si];
        st.eps_pos[slot * N + i] = v.eps_pos[si];
        st.r[slot * N + i] = v.r[si];
        // Chain::get_llik(sample), src/chain.cpp:1083-1097: the sample's cells in locus order
        const size_t base = si * L;
        double sum = 0.0;
        for (size_t j = 0; j < L; ++j) sum += __dadd_rn(v.cellT[base + j], v.cellO[base + j]);
        st.sample_llik[slot * N + i] = sum;
    }
    if (g0 == 0) st.mean_coi[slot] = v.mean_coi[t];
    if (st.latent)
    {
        moire::mask_t *dst = reinterpret_cast<moire::mask_t *>(st.latent);  // [capacity][N][L] masks of W words
        for (size_t i = g0; i < NL; i += stride) dst[slot * NL + i] = v.latent[t * NL + i];
    }
}

// ---- synthetic data on the device (row f2; R/simulator.R:154-212) ---------------------------------
// One thread per (sample, locus) cell, its own Philox stream: related strains ~ Binom(coi - 1, r)
// (:63-65), the coi - related unrelated strains each draw an allele from the locus' frequencies
// (the multinomial of :66-69, one categorical draw per strain by inverse CDF), then the observation
// process (:85-107): a present allele is seen w.p. 1 - eps_neg / A, an absent one w.p. eps_pos / A,
// the whole cell is blanked w.p. `missingness`; is_missing = nothing observed (:195-199).
__global__ void __launch_bounds__(256) k_simulate_cells(const int N, const int L, const int AP, const int *nall,
                                                        const double *freqs, const int *coi, const double *rel,
                                                        const double eps_pos, const double eps_neg,
                                                        const double missingness, const unsigned long long seed,
                                                        moire::mask_t *obs, unsigned char *missing,
                                                        moire::mask_t *truth)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)N * L) return;
    const int i = (int)(g / L), j = (int)(g - (size_t)i * L);
    const int A = nall[j];
    const double *f = freqs + (size_t)j * AP;
    moire::Stream rng;
    rng.open(seed, 0x51D0u, moire::KIND_INIT_SAMPLE, 0u, (uint32_t)i, (uint32_t)(j + 1));
    int related = 0;
    const double r = rel[i];
    for (int s = 0; s < coi[i] - 1; ++s) related += rng.next_uniform() < r;
    moire::mask_t present = 0;
    for (int s = 0; s < coi[i] - related; ++s)
    {
        const double u = rng.next_uniform();
        double c = 0.0;
        int a = 0;
        for (; a < A - 1; ++a)
        {
            c += f[a];
            if (u < c) break;
        }
        present |= moire::mbit(a);
    }
    moire::mask_t seen = 0;
    const double p_hit = 1.0 - eps_neg / (double)A, p_false = eps_pos / (double)A;
    for (int a = 0; a < A; ++a)
        if (rng.next_uniform() < (moire::mtest(present, a) ? p_hit : p_false)) seen |= moire::mbit(a);
    if (rng.next_uniform() < missingness) seen = 0;
    obs[g] = seen;
    missing[g] = seen == 0;
    if (truth) truth[g] = present;
}

template <int K>
__device__ void dump_small_subsets(unsigned short *out)
{
    constexpr moire::SmallSubsets<K> ord{};
    for (int i = 0; i < (1 << K) - 1; ++i)
    {
        int size = 0;
        for (int e = 0; e < K; ++e) size += (ord.mask[i] >> e) & 1;
        out[i] = (unsigned short)(ord.mask[i] | ((size & 1) ? 0 : 0x8000));  // ie_small's sign rule
    }
}

__global__ void k_dump_subset_order(const int which, const int k, unsigned short *out)
{
    using namespace moire;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nsub = (1 << k) - 1;
    if (which == 0)
    {
        const unsigned short *tab = c_sub_tail + subset_block(k);
        unsigned int prefix[kIeMax + 1];
        for (int idx = 0; idx < nsub; ++idx)
        {
            const unsigned int e = tab[idx];
            unsigned int tail = e & 1023u;
            int pos = (e >> 10) & 15;
            unsigned int base = pos ? prefix[pos - 1] : 0u;
            for (;;)
            {
                const int el = __ffs(tail) - 1;
                tail &= tail - 1u;
                base |= 1u << el;
                if (!tail) break;
                prefix[pos] = base;
                ++pos;
            }
            out[idx] = (unsigned short)(base | ((e & 0x4000u) ? 0x8000u : 0u));
        }
    }
    else if (which == 1)
    {
        const unsigned short *masks = g_sub_mask + subset_block(k);
        for (int idx = 0; idx < nsub; ++idx)
        {
            const unsigned int msk = masks[idx];
            out[idx] = (unsigned short)(msk | (((__popc(msk) & 1) == 0) ? 0x8000u : 0u));  // coop's sign rule
        }
    }
    else
    {
        switch (k)
        {
            case 1: dump_small_subsets<1>(out); break;
            case 2: dump_small_subsets<2>(out); break;
            case 3: dump_small_subsets<3>(out); break;
            case 4: dump_small_subsets<4>(out); break;
            case 5: dump_small_subsets<5>(out); break;
            case 6: dump_small_subsets<6>(out); break;
            default: break;
        }
    }
}

struct moire_engine
{
    static constexpr size_t kTraceWords = 1024 * 256;  // MOIRE_B200_TRACE: 256 words per evaluator block
    moire::View v{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool initialised = false;
    bool timing = false;
    std::string last_error;
    uint64_t launches = 0;
    int tile_S = 1, tiles_per_chain = 1;
    size_t sample_smem = 0, sample_smem_o = 0, p_smem = 0;
    int p_threads = 256;

    // allocations
    DevBuf<moire::mask_t> obs, latent;
    DevBuf<unsigned long long> evals, trace;
    // flat update_p pipeline (p_flat.cuh)
    moire::PFlat pf{};
    DevBuf<moire::ListEntry> pf_list;
    DevBuf<unsigned int> pf_hist;
    DevBuf<double> pf_Tcur, pf_Tnew, pf_pprop, pf_round;
    DevBuf<int> pf_nonmiss;
    int amax_alleles = 0;
    bool p_fused = false;
    // update_p's PAM cache (p_flat.cuh: k_pc_sweep)
    bool p_cache = false;
    DevBuf<moire::mask_t> pf_latLM;
    DevBuf<double> red_part;          // k_reduce: per-part sums
    DevBuf<unsigned int> red_done;    //           parts finished, per chain
    DevBuf<double> pf_Wcur, pf_Wnew;
    // the step as a CUDA graph: ~55 launches replayed with one call, iteration read from iter_dev
    DevBuf<int> iter_dev;
    cudaGraphExec_t step_graph = nullptr;
    uint64_t step_graph_launches = 0;
    bool use_graph = true;
    moire::SFlat sf{};
    DevBuf<moire::Proposal> sf_prop;
    DevBuf<int> sf_m;
    DevBuf<double> sf_adj;
    DevBuf<moire::SampleScal> sf_ss, pf_ss;
    DevBuf<moire::mask_t> sf_lat;
    DevBuf<moire::EpsLogs> sf_elogs;
    DevBuf<unsigned char> missing;
    DevBuf<int> nall, aclass, avals, obs_coi, m, acc6, p_acc, p_att, active;
    DevBuf<double> logC, lgam, lgadj, cellT, cellO, persample, p, p_sd, perchain, scratchN;

    struct TimedLaunch
    {
        cudaEvent_t a, b;
        int kind;
    };
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    std::vector<TimedLaunch> pending;
    std::vector<int32_t> init_log;  // (chain, sample, locus) per ill-conditioned start
    double kind_ms[10] = {0};  // 8 update kinds, the reduction, [9] = the flat evaluator alone
    uint64_t kind_launches[10] = {0};

    ~moire_engine()
    {
        cudaSetDevice(device);
        // The buffers go back to a per-process cache, not to cudaFree (which would wait for the
        // device): nothing of this engine may still be running when the next engine picks them up.
        if (stream) cudaStreamSynchronize(stream);
        for (auto &t : pending)
        {
            cudaEventDestroy(t.a);
            cudaEventDestroy(t.b);
        }
        if (step_graph) cudaGraphExecDestroy(step_graph);
        iter_dev.release();
        obs.release(); latent.release(); evals.release(); missing.release(); nall.release();
        aclass.release(); avals.release(); obs_coi.release(); m.release(); acc6.release();
        p_acc.release(); p_att.release(); active.release(); logC.release(); lgam.release();
        pf_list.release(); pf_hist.release(); pf_Tcur.release(); pf_Tnew.release(); pf_pprop.release();
        red_part.release(); red_done.release();
        pf_round.release(); pf_nonmiss.release(); pf_latLM.release(); pf_Wcur.release(); pf_Wnew.release();
        sf_prop.release(); sf_m.release(); sf_ss.release(); pf_ss.release(); sf_lat.release(); sf_adj.release(); sf_elogs.release();
        lgadj.release(); cellT.release(); cellO.release(); persample.release(); p.release();
        p_sd.release(); perchain.release(); scratchN.release();
        if (v.trace && std::getenv("MOIRE_B200_TRACE"))
        {
            std::vector<unsigned long long> h(kTraceWords);
            if (cudaMemcpy(h.data(), v.trace, h.size() * sizeof(h[0]), cudaMemcpyDeviceToHost) == cudaSuccess)
                if (FILE *f = std::fopen(std::getenv("MOIRE_B200_TRACE"), "wb"))
                {
                    std::fwrite(h.data(), sizeof(h[0]), h.size(), f);
                    std::fclose(f);
                }
            trace.release();
        }
        if (timer_a) cudaEventDestroy(timer_a);
        if (timer_b) cudaEventDestroy(timer_b);
        if (stream) cudaStreamDestroy(stream);
    }

    void bind() { CUDA_TRY(cudaSetDevice(device)); }

    void create(const moire_data *d, const moire_params *prm, int T, int chain_offset,
                const double *temps, int device_id, uint64_t seed)
    {
        if (!d || !prm || !temps) throw InvalidArg{"null argument"};
        // Programmatic dependent launch pays where the step is launch latency (the Namibia shape: -1.8 %)
        // and costs where kernels are long and a round's (chain, locus) blocks number thousands: the
        // blocks of the kernel behind become resident early and wait on the SMs the running kernel
        // needs (full C3 ladder, 4000 blocks per round: +3.5 %; 2000: +1.2 %; <= 1000: nothing either
        // way; profiles/r02z_pdl_by_shape.log).  MOIRE_B200_PDL=0 / 1 overrides.
        pdl = (size_t)T * (size_t)d->num_loci <= 1024;
        if (const char *env = std::getenv("MOIRE_B200_PDL")) pdl = std::atoi(env) != 0;
        const int N = d->num_samples, L = d->num_loci;
        if (N < 1 || L < 1 || T < 1) throw InvalidArg{"num_samples, num_loci, num_chains must be >= 1"};
        if (!d->num_alleles || !d->obs_mask || !d->is_missing) throw InvalidArg{"null data array"};
        if (prm->max_coi < 1 || prm->max_coi > MOIRE_MAX_COI)
            throw InvalidArg{"max_coi must be in 1.." + std::to_string(MOIRE_MAX_COI)};
        if (chain_offset < 0 || chain_offset + T > 65535) throw InvalidArg{"chain id out of range"};
        if (L > 65534) throw InvalidArg{"num_loci > 65534 unsupported"};
        int amax = 0;
        std::vector<int> avals_h, aclass_h(L);
        for (int j = 0; j < L; ++j)
        {
            const int A = d->num_alleles[j];
            if (A < 2 || A > moire::kMaskBits)
                throw InvalidArg{"num_alleles[" + std::to_string(j) + "] = " + std::to_string(A) +
                                 " outside 2.." + std::to_string(moire::kMaskBits) +
                                 (MOIRE_MASK_WORDS == 1 ? " (one-word masks: loci with up to " +
                                                              std::to_string(MOIRE_MAX_ALLELES) +
                                                              " alleles need libmoire_b200_w128)"
                                                        : std::string())};
            amax = std::max(amax, A);
            auto it = std::find(avals_h.begin(), avals_h.end(), A);
            if (it == avals_h.end())
            {
                avals_h.push_back(A);
                it = avals_h.end() - 1;
            }
            aclass_h[j] = (int)(it - avals_h.begin());
        }
        std::vector<int> coi_h(N, 0);
        for (int i = 0; i < N; ++i)
        {
            for (int j = 0; j < L; ++j)
            {
                const uint64_t *w = d->obs_mask + ((size_t)i * L + j) * MOIRE_MASK_WORDS;
                const moire::mask_t o = moire::mfrom_words(w[0], MOIRE_MASK_WORDS > 1 ? w[MOIRE_MASK_WORDS - 1] : 0);
                const int A = d->num_alleles[j];
                if (A < moire::kMaskBits && (o >> A) != 0) throw InvalidArg{"obs_mask has bits above num_alleles"};
                coi_h[i] = std::max(coi_h[i], moire::mpopc(o));
            }
            if (d->observed_coi) coi_h[i] = d->observed_coi[i];
        }
        {
            // observed COI >= 8 in more than a tenth of the samples: a high-COI data set (c5: 56 %; the
            // other named shapes: <= 6 %)
            size_t high = 0;
            for (int i = 0; i < N; ++i) high += coi_h[i] >= 8;
            eval_stash = high * 10 <= (size_t)N;
            if (const char *env = std::getenv("MOIRE_B200_EVAL_STASH")) eval_stash = std::atoi(env) != 0;
        }

        int ndev = 0;
        cudaError_t e0 = cudaGetDeviceCount(&ndev);
        if (e0 != cudaSuccess || ndev == 0)
            throw CudaError{std::string("no CUDA device: ") +
                            (e0 == cudaSuccess ? "device count is 0" : cudaGetErrorString(e0)) +
                            " -- libmoire_b200 has no CPU fallback"};
        if (device_id < 0 || device_id >= ndev) throw InvalidArg{"device_id out of range"};
        device = device_id;
        bind();
        CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

        v.N = N; v.L = L; v.T = T;
        v.AP = (amax + 3) / 4 * 4;
        v.nA = (int)avals_h.size();
        v.allow_rel = prm->allow_relatedness != 0;
        v.burnin = prm->burnin;
        v.max_coi = prm->max_coi;
        v.chain_offset = chain_offset;
        v.seed = seed;
        v.mean_coi_shape = prm->mean_coi_shape;
        v.mean_coi_scale = prm->mean_coi_scale;
        v.lg_shape = std::lgamma(prm->mean_coi_shape);
        v.l_scale = std::log(prm->mean_coi_scale);
        v.max_eps_pos = prm->max_eps_pos; v.eps_pos_a = prm->eps_pos_alpha; v.eps_pos_b = prm->eps_pos_beta;
        v.max_eps_neg = prm->max_eps_neg; v.eps_neg_a = prm->eps_neg_alpha; v.eps_neg_b = prm->eps_neg_beta;
        v.r_a = prm->r_alpha; v.r_b = prm->r_beta;
        auto lbeta = [](double a, double b) { return std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b); };
        v.eps_pos_lbeta = lbeta(v.eps_pos_a, v.eps_pos_b);
        v.eps_neg_lbeta = lbeta(v.eps_neg_a, v.eps_neg_b);
        v.r_lbeta = lbeta(v.r_a, v.r_b);

        const size_t NL = (size_t)N * L, TNL = (size_t)T * NL, TN = (size_t)T * N;
        const size_t TLA = (size_t)T * L * v.AP;
        obs.alloc(NL); missing.alloc(NL); nall.alloc(L); aclass.alloc(L); avals.alloc(v.nA);
        obs_coi.alloc(N); logC.alloc(moire::kLogCDim * moire::kLogCDim); lgam.alloc(128);
        latent.alloc(TNL); lgadj.alloc(TNL); cellT.alloc(TNL); cellO.alloc(TNL);
        m.alloc(TN); acc6.alloc(6 * TN); persample.alloc(14 * TN);
        p.alloc(TLA); p_sd.alloc(TLA); p_acc.alloc(TLA); p_att.alloc(TLA);
        perchain.alloc(7 * (size_t)T); active.alloc(T); evals.alloc(16); iter_dev.alloc(1); scratchN.alloc(N);
        red_part.alloc((size_t)T * moire::kReduceParts * 2); red_done.alloc(T);
        CUDA_TRY(cudaMemsetAsync(red_done.ptr, 0, (size_t)T * sizeof(unsigned int), stream));

        std::vector<double> logC_h(moire::kLogCDim * moire::kLogCDim, -INFINITY), lgam_h(128);
        for (int n = 0; n < moire::kLogCDim; ++n)
            for (int k = 0; k <= n; ++k) logC_h[n * moire::kLogCDim + k] = log_binom_exact(n, k);
        for (int i = 0; i < 128; ++i) lgam_h[i] = std::lgamma((double)(i + 1));  // sampler.cpp:22-25

        {
            std::vector<unsigned short> tail_h, mask_h;
            build_subset_tables(tail_h, mask_h);
            CUDA_TRY(cudaMemcpyToSymbol(moire::c_sub_tail, tail_h.data(), tail_h.size() * sizeof(unsigned short)));
            CUDA_TRY(cudaMemcpyToSymbol(moire::g_sub_mask, mask_h.data(), mask_h.size() * sizeof(unsigned short)));
            std::vector<double> comb_h;
            build_egf_binomials(comb_h);
            CUDA_TRY(cudaMemcpyToSymbol(moire::g_egf_comb, comb_h.data(), comb_h.size() * sizeof(double)));
            std::vector<unsigned short> rank_h;
            std::vector<float> est_h;
            moire::build_flat_classes(rank_h, est_h);
            CUDA_TRY(cudaMemcpyToSymbol(moire::c_flat_class, rank_h.data(), rank_h.size() * sizeof(unsigned short)));
            CUDA_TRY(cudaMemcpyToSymbol(moire::c_flat_est, est_h.data(), est_h.size() * sizeof(float)));
            const int D = moire::kMaxCoi + 1;
            std::vector<double> binom_h((size_t)D * D, 0.0);
            for (int n = 0; n < D; ++n)
            {
                long double c = 1.0L;  // exact: C(n, i) < 2^64 for n <= 64
                for (int i = 0; i <= n; ++i)
                {
                    binom_h[(size_t)n * D + i] = (double)c;
                    c = c * (long double)(n - i) / (long double)(i + 1);
                }
            }
            CUDA_TRY(cudaMemcpyToSymbol(moire::g_binom, binom_h.data(), binom_h.size() * sizeof(double)));
        }
        obs.upload((const moire::mask_t *)d->obs_mask, NL, stream);  // [N][L][words], word 0 = alleles 0..63
        missing.upload(d->is_missing, NL, stream);
        nall.upload(d->num_alleles, L, stream);
        aclass.upload(aclass_h.data(), L, stream);
        avals.upload(avals_h.data(), v.nA, stream);
        obs_coi.upload(coi_h.data(), N, stream);
        logC.upload(logC_h.data(), logC_h.size(), stream);
        lgam.upload(lgam_h.data(), 128, stream);

        v.obs = obs.ptr; v.missing = missing.ptr; v.nall = nall.ptr; v.aclass = aclass.ptr;
        v.avals = avals.ptr; v.obs_coi = obs_coi.ptr; v.logC = logC.ptr; v.lgam = lgam.ptr;
        v.latent = latent.ptr; v.lgadj = lgadj.ptr; v.cellT = cellT.ptr; v.cellO = cellO.ptr;
        v.m = m.ptr;
        double *ps = persample.ptr;
        v.r = ps + 0 * TN; v.log_r = ps + 1 * TN; v.log_1mr = ps + 2 * TN;
        v.eps_pos = ps + 3 * TN; v.eps_neg = ps + 4 * TN;
        v.sd_eps_neg = ps + 5 * TN; v.sd_eps_pos = ps + 6 * TN; v.sd_r = ps + 7 * TN; v.sd_mr = ps + 8 * TN;
        v.pr_eps_neg = ps + 9 * TN; v.pr_eps_pos = ps + 10 * TN; v.pr_coi = ps + 11 * TN; v.pr_r = ps + 12 * TN;
        v.odds = ps + 13 * TN;
        int *ac = acc6.ptr;
        v.acc_eps_neg = ac + 0 * TN; v.acc_eps_pos = ac + 1 * TN; v.acc_m = ac + 2 * TN;
        v.acc_r = ac + 3 * TN; v.acc_mr = ac + 4 * TN; v.acc_samples = ac + 5 * TN;
        v.p = p.ptr; v.p_sd = p_sd.ptr; v.p_acc = p_acc.ptr; v.p_att = p_att.ptr;
        double *pc = perchain.ptr;
        v.mean_coi = pc + 0 * T; v.mean_coi_sd = pc + 1 * T; v.mean_coi_acc = pc + 2 * T;
        v.hyper = pc + 3 * T; v.temp = pc + 4 * T; v.llik = pc + 5 * T; v.prior = pc + 6 * T;
        v.active = active.ptr; v.evals = evals.ptr; v.iter = iter_dev.ptr;
        v.trace = nullptr;
        if (std::getenv("MOIRE_B200_TRACE"))  // tuning aid: per-warp clocks of update_p
        {
            trace.alloc(kTraceWords);
            CUDA_TRY(cudaMemsetAsync(trace.ptr, 0, kTraceWords * sizeof(unsigned long long), stream));
#ifdef MOIRE_EVAL_TRACE
            {
                // cooperative cells' phase clocks go behind the per-block records (block 600 on)
                unsigned long long *coop = trace.ptr + 600 * 256;
                CUDA_TRY(cudaMemcpyToSymbol(moire::g_coop_trace, &coop, sizeof(coop)));
            }
#endif
            v.trace = trace.ptr;
        }
        CUDA_TRY(cudaMemcpyAsync(v.temp, temps, T * sizeof(double), cudaMemcpyHostToDevice, stream));

        // tile shape of the per-sample kernel: S consecutive samples, about kMaxTileCells cells
        tile_S = std::max(1, std::min({moire::kMaxTileSamples, moire::kMaxTileCells / L, N}));
        tiles_per_chain = (N + tile_S - 1) / tile_S;
        sample_smem = moire::sample_kernel_smem(tile_S, L, v.nA, true);
        sample_smem_o = moire::sample_kernel_smem(tile_S, L, v.nA, false);
        p_threads = moire::p_kernel_threads(N);
        p_smem = moire::p_kernel_smem(N, v.AP, p_threads);
        // update_p: the flat pipeline unless MOIRE_B200_P_FUSED is set (A/B runs) or the cell
        // index does not fit 32 bits
        amax_alleles = amax;
        // ... or the problem is tiny (the flat pipeline is 2 A_max + 6 launches; below ~10^5 cells
        // they cost more than the fused kernel's barriers)
        // MOIRE_B200_FLAT_MIN_CELLS moves that threshold (tests run both pipelines on one shape)
        size_t flat_min_cells = 100000;
        if (const char *env = std::getenv("MOIRE_B200_FLAT_MIN_CELLS")) flat_min_cells = (size_t)std::atoll(env);
        p_fused = std::getenv("MOIRE_B200_P_FUSED") != nullptr || TNL < flat_min_cells;
        // moire_params::pipeline overrides the size rule (tests run both generations on one shape)
        const int pipeline = prm->pipeline & 0xff;
        if (pipeline == MOIRE_PIPELINE_FLAT) p_fused = false;
        if (pipeline == MOIRE_PIPELINE_FUSED) p_fused = true;
#if MOIRE_MASK_WORDS > 1
        // two-word masks: the flat pipelines only (the fused tiny-problem kernels keep one-word
        // per-block genotype arrays and two alleles per lane)
        if (pipeline == MOIRE_PIPELINE_FUSED || std::getenv("MOIRE_B200_P_FUSED"))
            throw InvalidArg{"the fused kernels are built for one-word masks only (libmoire_b200.so)"};
        p_fused = false;
#endif
        if (prm->pipeline < 0 || pipeline > MOIRE_PIPELINE_FUSED ||
            (prm->pipeline & ~(0xff | MOIRE_NUMERICS_REF32 | MOIRE_P_CACHE_ON | MOIRE_P_CACHE_OFF)) ||
            ((prm->pipeline & MOIRE_P_CACHE_ON) && (prm->pipeline & MOIRE_P_CACHE_OFF)))
            throw InvalidArg{"unknown pipeline / numerics flags"};
        {
            const int pam_float = (prm->pipeline & MOIRE_NUMERICS_REF32) != 0;
            const double max_pam = pam_float ? 1.0 - (double)FLT_EPSILON : 1.0 - DBL_EPSILON;
            CUDA_TRY(cudaMemcpyToSymbol(moire::c_max_pam, &max_pam, sizeof(double)));
            CUDA_TRY(cudaMemcpyToSymbol(moire::c_pam_float, &pam_float, sizeof(int)));
        }
        if (!p_fused && TNL >= (1ull << 32))
            throw InvalidArg{"num_chains * num_samples * num_loci must stay below 2^32 per engine "
                             "(32-bit cell lists): shard the ladder over more engines"};
        if (p_fused && N > 65535)
            throw InvalidArg{"the fused (tiny-problem) kernels keep 16-bit work lists: num_samples > 65535 "
                             "needs the flat pipeline"};
        if (!p_fused) sample_smem = sample_smem_o = 0;  // the fused kernels' per-block arrays are not needed
        if (!p_fused)
        {
            pf_list.alloc(TNL); pf_hist.alloc(moire::kPfHistWords);
            pf_Tcur.alloc(TNL); pf_Tnew.alloc(TNL); pf_latLM.alloc(TNL); pf.latLM = pf_latLM.ptr;
            pf_pprop.alloc(TLA); pf_round.alloc((size_t)T * L * 4); pf_nonmiss.alloc(L);
            std::vector<int> nm(L, 0);
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < L; ++j) nm[j] += d->is_missing[(size_t)i * L + j] == 0;
            pf_nonmiss.upload(nm.data(), L, stream);
            pf.list = pf_list.ptr; pf.hist = pf_hist.ptr; pf.Tcur = pf_Tcur.ptr; pf.Tnew = pf_Tnew.ptr;
            pf.pprop = pf_pprop.ptr; pf.round = pf_round.ptr; pf.nonmiss = pf_nonmiss.ptr;
            p_smem = 0;  // the fused kernel's per-block arrays are not needed
            {
                // update_p's PAM cache: two [max_coi][cells] arrays + a locus-major genotype copy
                const size_t wbytes = 2 * (size_t)v.max_coi * TNL * sizeof(double);
                size_t budget = (size_t)48 << 30;
                if (const char *env = std::getenv("MOIRE_B200_P_CACHE_MAX_GB")) budget = (size_t)std::atoll(env) << 30;
                const bool on = (prm->pipeline & MOIRE_P_CACHE_ON) != 0, off = (prm->pipeline & MOIRE_P_CACHE_OFF) != 0;
                if (on && wbytes > budget)
                    throw InvalidArg{"MOIRE_P_CACHE_ON needs " + std::to_string(wbytes >> 20) + " MiB for the PAM cache; budget " +
                                     std::to_string(budget >> 20) + " MiB (MOIRE_B200_P_CACHE_MAX_GB)"};
                // off unless asked for: a cached PAM comes from frequencies that were rescaled since, and
                // a cell whose coverage probability is tiny amplifies that last-bit difference past the
                // 1e-9 the from-scratch evaluation holds against the reference
                (void)off;
                p_cache = on;
                if (p_cache)
                {
                    pf_Wcur.alloc((size_t)v.max_coi * TNL); pf_Wnew.alloc((size_t)v.max_coi * TNL);
                    pf.Wcur = pf_Wcur.ptr; pf.Wnew = pf_Wnew.ptr; pf.wstride = TNL;
                }
            }
            {
                // k_pf_round: one block per (chain, locus), about as many threads in all as the device
                // holds at 64 registers each (one warp per locus at the benchmark ladder, 512 threads
                // for a few dozen loci with thousands of samples)
                cudaDeviceProp pr;
                CUDA_TRY(cudaGetDeviceProperties(&pr, device));
                const size_t want = (size_t)pr.multiProcessorCount * 1024 / ((size_t)T * L);
                pf_round_threads = 32;
                while (pf_round_threads < 1024 && (size_t)pf_round_threads * 2 <= want && pf_round_threads * 4 <= N)
                    pf_round_threads *= 2;
                if (const char *env = std::getenv("MOIRE_B200_PF_ROUND_THREADS"))  // A/B runs
                    pf_round_threads = std::max(32, std::min(1024, std::atoi(env) / 32 * 32));
            }
            // the per-sample updates share the list, the histogram and the two term arrays
            sf_prop.alloc(TN); sf_m.alloc(TN); sf_ss.alloc(TN); pf_ss.alloc(TN);
            pf.ss = pf_ss.ptr;
            sf_lat.alloc(TNL); sf_adj.alloc(TNL); sf_elogs.alloc(TN * (size_t)v.nA);
            sf.prop = sf_prop.ptr; sf.m_new = sf_m.ptr; sf.ss = sf_ss.ptr;
            sf.lat = sf_lat.ptr; sf.adj = sf_adj.ptr; sf.elogs = sf_elogs.ptr;
            sf.O = pf_Tcur.ptr; sf.Tn = pf_Tnew.ptr;
        }
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        const size_t smem_cap = prop.sharedMemPerBlockOptin;
        if (sample_smem > smem_cap || p_smem > smem_cap)
            throw InvalidArg{"problem shape needs " + std::to_string(std::max(sample_smem, p_smem)) +
                             " B of shared memory per block; device allows " + std::to_string(smem_cap)};
        use_graph = std::getenv("MOIRE_B200_NO_GRAPH") == nullptr;
        set_smem_attr();
        CUDA_TRY(cudaStreamSynchronize(stream));
    }

    void set_smem_attr()
    {
        using namespace moire;
        const int ss = (int)sample_smem;
        if (p_fused)
        {
            CUDA_TRY(cudaFuncSetAttribute(k_update_sample<KIND_EPS_NEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sample_smem_o));
            CUDA_TRY(cudaFuncSetAttribute(k_update_sample<KIND_EPS_POS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sample_smem_o));
            CUDA_TRY(cudaFuncSetAttribute(k_update_sample<KIND_M>, cudaFuncAttributeMaxDynamicSharedMemorySize, ss));
            CUDA_TRY(cudaFuncSetAttribute(k_update_sample<KIND_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, ss));
            CUDA_TRY(cudaFuncSetAttribute(k_update_sample<KIND_EFF_COI>, cudaFuncAttributeMaxDynamicSharedMemorySize, ss));
            CUDA_TRY(cudaFuncSetAttribute(k_update_sample<KIND_SAMPLES>, cudaFuncAttributeMaxDynamicSharedMemorySize, ss));
            CUDA_TRY(cudaFuncSetAttribute(k_update_p<MOIRE_P_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p_smem));
            CUDA_TRY(cudaFuncSetAttribute(k_update_p<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p_smem));
        }
        CUDA_TRY(cudaFuncSetAttribute(k_flat_eval<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf_eval_smem()));
        CUDA_TRY(cudaFuncSetAttribute(k_flat_eval<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf_eval_smem()));
        CUDA_TRY(cudaFuncSetAttribute(k_flat_eval<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf_eval_smem()));
        {
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_flat_eval<false, true>, kEvalThreads, pf_eval_smem()));
            cudaDeviceProp prop1;
            CUDA_TRY(cudaGetDeviceProperties(&prop1, device));
            flat_grid = std::max(1, per_sm) * prop1.multiProcessorCount;
            pf.eval_threads = flat_grid * kEvalThreads;
        }
    }

    struct Timed
    {
        moire_engine *e;
        cudaEvent_t a = nullptr, b = nullptr;
        int kind;
        Timed(moire_engine *e_, int kind_) : e(e_), kind(kind_)
        {
            ++e->launches;
            if (e->timing)
            {
                CUDA_TRY(cudaEventCreate(&a));
                CUDA_TRY(cudaEventCreate(&b));
                CUDA_TRY(cudaEventRecord(a, e->stream));
            }
        }
        void done()
        {
            CUDA_TRY(cudaGetLastError());
            if (e->timing)
            {
                CUDA_TRY(cudaEventRecord(b, e->stream));
                e->pending.push_back({a, b, kind});
            }
        }
    };

    void resolve_timing()
    {
        if (pending.empty()) return;
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (auto &t : pending)
        {
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, t.a, t.b));
            kind_ms[t.kind] += ms;
            kind_launches[t.kind] += 1;
            cudaEventDestroy(t.a);
            cudaEventDestroy(t.b);
        }
        pending.clear();
    }

    size_t pf_eval_smem() const { return moire::pf_eval_smem_doubles(v.max_coi) * sizeof(double); }
    int flat_grid = 0;  // persistent evaluator grid: resident blocks of the device

    // A kernel of the step chain: launched with the programmatic-stream-serialization attribute, so that
    // it may become resident while the kernel in front of it finishes (every such kernel starts with
    // pdl_enter(), state.cuh).  MOIRE_B200_PDL=0 launches them the ordinary way (A/B runs).
    bool pdl = true;
    // which evaluator runs (p_flat.cuh: STASH): the one that parks a cell's output address in shared
    // memory unless heavy cells (k >= 8 latent alleles: long single-thread walks) are common in the data
    bool eval_stash = true;
    template <class... P, class... A>
    void launch_chain(void (*kernel)(P...), unsigned grid, unsigned block, size_t smem, A &&...args)
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = pdl ? 1 : 0;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, P(std::forward<A>(args))...));
    }

    void launch_flat_eval(const moire::FlatSrc &src, bool cache = false)
    {
        cudaEvent_t a = nullptr, b = nullptr;
        if (timing)
        {
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            CUDA_TRY(cudaEventRecord(a, stream));
        }
        if (cache)
            launch_chain(moire::k_flat_eval<true, false>, flat_grid, moire::kEvalThreads, pf_eval_smem(), v, src);
        else if (eval_stash)
            launch_chain(moire::k_flat_eval<false, true>, flat_grid, moire::kEvalThreads, pf_eval_smem(), v, src);
        else
            launch_chain(moire::k_flat_eval<false, false>, flat_grid, moire::kEvalThreads, pf_eval_smem(), v, src);
        if (timing)
        {
            CUDA_TRY(cudaEventRecord(b, stream));
            pending.push_back({a, b, 9});
        }
    }

    // update_p as sort-once + (round, eval) per proposal (p_flat.cuh); nothing returns to the host
    int pf_round_threads = 32;  // k_pf_round: threads of the block that owns a (chain, locus)
    void update_p_flat(int iteration)
    {
        using namespace moire;
        const size_t total = (size_t)v.T * v.L * v.N;
        const unsigned cell_blocks = (unsigned)((total + 255) / 256);
        CUDA_TRY(cudaMemsetAsync(pf.hist, 0, kPfHistWords * sizeof(unsigned int), stream));
        const unsigned tile_blocks = (unsigned)v.T * pf_tiles(v.N, v.L);
        launch_chain(k_pf_prepare, tile_blocks, 256, 0, v, pf);
        launch_chain(k_pf_scan, 1, 256, 0, pf);
        launch_chain(k_pf_scatter, scatter_blocks(total), 256, 0, v, pf);
        FlatSrc src{};
        src.list = pf.list; src.bounds = pf.hist + kPfScalars;
        src.p = pf.pprop; src.p_stride = v.AP; src.out_locus_major = 1;
        src.ss = pf.ss; src.binom_rows = v.max_coi;
        src.round_flags = pf.round; src.out = pf.Tnew; src.out_flip = pf.Tcur;
        launches += 2;  // prepare, scan, scatter (the caller counted one)
        if (p_cache)
        {
            // the PAM vectors of the current state: one evaluation of every cell at the accepted p
            // (its terms land in Tnew, which every round overwrites; no locus is swapped yet)
            src.p = v.p; src.round_flags = nullptr; src.out_flip = nullptr; src.counter = pf.hist + kPfCounters;
            src.pam_out = pf.Wcur; src.pam_stride = pf.wstride;
            launch_flat_eval(src, true);
            src.p = pf.pprop; src.pam_out = pf.Wnew; src.round_flags = pf.round; src.out_flip = pf.Tcur;
            const unsigned list_blocks = scatter_blocks(total);
            for (int rep = 0; rep < amax_alleles; ++rep)
            {
                launch_chain(k_pf_round, (unsigned)(v.T * v.L), pf_round_threads, 0, v, pf, iteration, rep);
                launch_chain(k_pc_sweep, cell_blocks, 256, 0, v, pf);
                launch_chain(k_pf_scan, 1, 256, 0, pf);
                launch_chain(k_pc_scatter, list_blocks, 256, 0, v, pf);
                src.counter = pf.hist + kPfCounters + 1 + rep;
                launch_flat_eval(src, true);
                launches += 5;
            }
            launch_chain(k_pf_round, (unsigned)(v.T * v.L), pf_round_threads, 0, v, pf, iteration, amax_alleles);
            launch_chain(k_pf_finish, tile_blocks, 256, 0, v, pf);
            launches += 3;
            CUDA_TRY(cudaGetLastError());
            return;
        }
        for (int rep = 0; rep < amax_alleles; ++rep)
        {
            launch_chain(k_pf_round, (unsigned)(v.T * v.L), pf_round_threads, 0, v, pf, iteration, rep);
            src.counter = pf.hist + kPfCounters + rep;
            launch_flat_eval(src);
            launches += 2;
        }
        launch_chain(k_pf_round, (unsigned)(v.T * v.L), pf_round_threads, 0, v, pf, iteration, amax_alleles);
        launch_chain(k_pf_finish, tile_blocks, 256, 0, v, pf);
        launches += 2;
        CUDA_TRY(cudaGetLastError());
    }

    // update_m / update_r / update_eff_coi / update_samples as a flat pipeline (s_flat.cuh)
    template <int KIND>
    void update_sample_flat(int iteration)
    {
        using namespace moire;
        const size_t total = (size_t)v.T * v.N * v.L;
        const size_t draw_spb = (size_t)sf_draw_units_per_block(v.nA, KIND != KIND_R);
        const unsigned draw_blocks = (unsigned)(((size_t)v.T * v.N + draw_spb - 1) / draw_spb);
        CUDA_TRY(cudaMemsetAsync(pf.hist, 0, kPfHistWords * sizeof(unsigned int), stream));
        launch_chain(k_sf_draw<KIND>, draw_blocks, kDrawThreads, 0, v, sf, iteration);
        launch_chain(k_sf_cells<KIND>, sf_cell_blocks(total), 256, 0, v, sf, pf, iteration);
        launch_chain(k_pf_scan, 1, 256, 0, pf);
        launch_chain(k_sf_scatter<KIND>, scatter_blocks(total), 256, 0, v, sf, pf);
        FlatSrc src{};
        src.list = pf.list; src.bounds = pf.hist + kPfScalars; src.counter = pf.hist + kPfCounters;
        src.p = v.p; src.p_stride = v.AP; src.out_locus_major = 0;
        src.ss = sf.ss; src.binom_rows = v.max_coi;
        src.round_flags = nullptr; src.out = sf.Tn;
        launch_flat_eval(src);
        launch_chain(k_sf_decide<KIND>, (unsigned)(((size_t)v.T * v.N + 7) / 8), 256, 0, v, sf, iteration);
        launches += 5;  // draw, cells, scan, scatter, eval, decide (the caller counted one)
        CUDA_TRY(cudaGetLastError());
    }

    template <int KIND>
    void launch_fused_or_eps(int iteration)
    {
        if constexpr (KIND == moire::KIND_EPS_NEG || KIND == moire::KIND_EPS_POS)
        {
            if (!p_fused)
            {
                const size_t units = (size_t)v.T * v.N;
                const size_t spb = (size_t)moire::sf_draw_units_per_block(v.nA, true);
                launch_chain(moire::k_sf_draw<KIND>, (unsigned)((units + spb - 1) / spb), moire::kDrawThreads, 0, v, sf, iteration);
                launch_chain(moire::k_eps_cells<KIND>, (unsigned)((units + 7) / 8), 256, 0, v, sf, iteration);
                launches += 1;  // two kernels; the caller counts one
                return;
            }
        }
        const size_t smem = (KIND == moire::KIND_EPS_NEG || KIND == moire::KIND_EPS_POS) ? sample_smem_o : sample_smem;
        moire::k_update_sample<KIND><<<v.T * tiles_per_chain, moire::kThreads, smem, stream>>>(
            v, iteration, tile_S, tiles_per_chain);
    }

    template <int KIND>
    void launch_sample(int iteration)
    {
        Timed tm(this, KIND);
        if (!p_fused && KIND != moire::KIND_EPS_NEG && KIND != moire::KIND_EPS_POS)
        {
            update_sample_flat<KIND>(iteration);
            tm.done();
            return;
        }
        launch_fused_or_eps<KIND>(iteration);
        tm.done();
    }

    void update(int kind, int iteration)
    {
        using namespace moire;
        switch (kind)
        {
            case MOIRE_UPDATE_EPS_NEG: launch_sample<KIND_EPS_NEG>(iteration); break;
            case MOIRE_UPDATE_EPS_POS: launch_sample<KIND_EPS_POS>(iteration); break;
            case MOIRE_UPDATE_M: launch_sample<KIND_M>(iteration); break;
            case MOIRE_UPDATE_R: launch_sample<KIND_R>(iteration); break;
            case MOIRE_UPDATE_EFF_COI: launch_sample<KIND_EFF_COI>(iteration); break;
            case MOIRE_UPDATE_SAMPLES: launch_sample<KIND_SAMPLES>(iteration); break;
            case MOIRE_UPDATE_P:
            {
                Timed tm(this, kind);
                if (!p_fused)
                    update_p_flat(iteration);
                else if (p_threads == MOIRE_P_THREADS)
                    k_update_p<MOIRE_P_THREADS><<<v.T * v.L, MOIRE_P_THREADS, p_smem, stream>>>(v, iteration);
                else
                    k_update_p<512><<<v.T * v.L, 512, p_smem, stream>>>(v, iteration);
                tm.done();
                break;
            }
            case MOIRE_UPDATE_MEAN_COI:
            {
                Timed tm(this, kind);
                launch_chain(k_update_mean_coi, v.T, kThreads, 0, v, iteration);
                tm.done();
                break;
            }
            default: throw InvalidArg{"unknown update kind"};
        }
    }

    void reduce()
    {
        Timed tm(this, 8);
        launch_chain(moire::k_reduce, v.T * moire::kReduceParts, 1024, 0, v, red_part.ptr, red_done.ptr);
        tm.done();
    }

    // MCMC::burnin / MCMC::sample loop body (src/mcmc.cpp:162-173, 311-322)
    void step_launches(int iteration)
    {
        update(MOIRE_UPDATE_EPS_NEG, iteration);
        update(MOIRE_UPDATE_EPS_POS, iteration);
        update(MOIRE_UPDATE_P, iteration);
        update(MOIRE_UPDATE_M, iteration);
        if (v.allow_rel)
        {
            update(MOIRE_UPDATE_R, iteration);
            update(MOIRE_UPDATE_EFF_COI, iteration);
        }
        update(MOIRE_UPDATE_SAMPLES, iteration);
        update(MOIRE_UPDATE_MEAN_COI, iteration);
        reduce();
    }

    // The launches of a step never change (grids, pointers, round numbers; the list bounds and
    // task counters live on the device), only the iteration does -- and the kernels read that
    // from device memory when handed kIterFromDevice.  So the step is captured once and replayed:
    // one host call per step instead of ~55 launches and a handful of memsets, back-to-back on
    // the device.  Per-kind timing (events between the kernels) uses the plain launches.
    uint64_t replay_step_launches = 0;  // launches of one step issued into a caller's capture

    void step(int iteration)
    {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        CUDA_TRY(cudaStreamIsCapturing(stream, &cap));
        if (cap != cudaStreamCaptureStatusNone)
        {
            // the caller (the driver) is capturing this step together with what follows it: plain
            // launches into its graph; they execute when it replays (moire_engine_note_replay)
            const uint64_t before = launches;
            step_launches(iteration);
            replay_step_launches = launches - before;
            launches = before;
            return;
        }
        if (iteration < 0)
        {
            step_launches(moire::kIterFromDevice);
            return;
        }
        if (!use_graph || timing)
        {
            step_launches(iteration);
            return;
        }
        if (!step_graph)
        {
            const uint64_t before = launches;
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            try
            {
                step_launches(moire::kIterFromDevice);
            }
            catch (...)
            {
                cudaStreamEndCapture(stream, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            CUDA_TRY(cudaStreamEndCapture(stream, &graph));
            step_graph_launches = launches - before;
            launches = before;
            const cudaError_t rc = cudaGraphInstantiate(&step_graph, graph, 0);
            cudaGraphDestroy(graph);
            CUDA_TRY(rc);
        }
        k_set_iteration<<<1, 1, 0, stream>>>(iter_dev.ptr, iteration);  // by value: steps may queue up
        CUDA_TRY(cudaGraphLaunch(step_graph, stream));
        launches += step_graph_launches + 1;
    }

    void eval_all(int chain_sel, int use_active)
    {
        using namespace moire;
        const size_t nchains = chain_sel < 0 ? v.T : 1;
        {
            const size_t total = nchains * v.N;
            ++launches;
            k_eval_priors<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(v, chain_sel, use_active);
            CUDA_TRY(cudaGetLastError());
        }
        {
            const size_t total = nchains * (size_t)v.N * v.L;
            ++launches;
            k_eval_cells<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, stream>>>(
                v, chain_sel, use_active);
            CUDA_TRY(cudaGetLastError());
        }
        reduce();
    }

    void get_scalars(double *llik, double *prior)
    {
        if (llik) CUDA_TRY(cudaMemcpyAsync(llik, v.llik, v.T * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (prior) CUDA_TRY(cudaMemcpyAsync(prior, v.prior, v.T * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }

    void init_chains(int max_tries, int32_t *ill_count, int32_t *still_ill, int32_t *bad_sample,
                     int32_t *bad_locus, int32_t *prior_nonfinite)
    {
        using namespace moire;
        const int T = v.T, N = v.N, L = v.L;
        std::vector<int> act(T, 1), ill(T, 0);
        std::vector<double> llik(T);
        init_log.clear();
        if (bad_sample) std::fill(bad_sample, bad_sample + T, 0);
        if (bad_locus) std::fill(bad_locus, bad_locus + T, 0);
        if (prior_nonfinite) std::fill(prior_nonfinite, prior_nonfinite + 5 * T, 0);
        for (int attempt = 0; attempt <= max_tries; ++attempt)
        {
            CUDA_TRY(cudaMemcpyAsync(v.active, act.data(), T * sizeof(int), cudaMemcpyHostToDevice, stream));
            const size_t TN = (size_t)T * N, TNL = TN * L, TL = (size_t)T * L;
            launches += 3;
            k_init_samples<<<(unsigned)((TN + 255) / 256), 256, 0, stream>>>(v, attempt);
            k_init_latent<<<(unsigned)((TNL + 255) / 256), 256, 0, stream>>>(v, attempt);
            k_init_p<<<(unsigned)((TL + 127) / 128), 128, 0, stream>>>(v, attempt);
            CUDA_TRY(cudaGetLastError());
            eval_all(-1, 1);
            get_scalars(llik.data(), nullptr);
            bool any = false;
            for (int t = 0; t < T; ++t)
            {
                if (!act[t]) continue;
                if (std::isfinite(llik[t]))
                {
                    act[t] = 0;
                    continue;
                }
                ++ill[t];
                any = true;
                int32_t bs = 0, bl = 0;
                diagnose(t, &bs, &bl, prior_nonfinite ? &prior_nonfinite[5 * t] : nullptr);
                if (bad_sample) bad_sample[t] = bs;
                if (bad_locus) bad_locus[t] = bl;
                init_log.push_back(t);
                init_log.push_back(bs);
                init_log.push_back(bl);
            }
            if (!any) break;
        }
        for (int t = 0; t < T; ++t)
        {
            if (ill_count) ill_count[t] = ill[t];
            if (still_ill) still_ill[t] = act[t];
        }
        initialised = true;
    }

    // first_nonfinite_genotyping_term / collect_nonfinite_prior_terms (src/chain.cpp:1032-1081);
    // failure path only, so the scan runs on dumped arrays.
    void diagnose(int t, int32_t *bad_sample, int32_t *bad_locus, int32_t *prior_counts)
    {
        const size_t NL = (size_t)v.N * v.L;
        std::vector<double> T_(NL), O_(NL), pr(4 * (size_t)v.N);
        double hyper = 0;
        CUDA_TRY(cudaMemcpyAsync(T_.data(), v.cellT + t * NL, NL * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(O_.data(), v.cellO + t * NL, NL * sizeof(double), cudaMemcpyDeviceToHost, stream));
        const double *src[4] = {v.pr_eps_neg, v.pr_eps_pos, v.pr_r, v.pr_coi};
        for (int q = 0; q < 4; ++q)
            CUDA_TRY(cudaMemcpyAsync(pr.data() + (size_t)q * v.N, src[q] + (size_t)t * v.N,
                                     v.N * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(&hyper, v.hyper + t, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (bad_sample && bad_locus)
        {
            *bad_sample = *bad_locus = 0;
            for (size_t c = 0; c < NL; ++c)
            {
                if (!std::isfinite(T_[c] + O_[c]))
                {
                    *bad_sample = (int)(c / v.L) + 1;
                    *bad_locus = (int)(c % v.L) + 1;
                    break;
                }
            }
        }
        if (prior_counts)
        {
            for (int q = 0; q < 4; ++q)
                for (int i = 0; i < v.N; ++i)
                    if (!std::isfinite(pr[(size_t)q * v.N + i])) ++prior_counts[q];
            if (!std::isfinite(hyper)) ++prior_counts[4];
        }
    }

    template <class T>
    void d2h(T *dst, const T *src, size_t n)
    {
        if (dst) CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
    }
    template <class T>
    void h2d(T *dst, const T *src, size_t n)
    {
        if (src) CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, stream));
    }

    void dump_state(int t, moire_chain_state *o)
    {
        const size_t N = v.N, NL = N * v.L, LA = (size_t)v.L * v.AP;
        d2h(o->m, v.m + t * N, N);
        d2h(o->r, v.r + t * N, N);
        d2h(o->eps_pos, v.eps_pos + t * N, N);
        d2h(o->eps_neg, v.eps_neg + t * N, N);
        d2h(o->p, v.p + t * LA, LA);
        d2h((moire::mask_t *)o->latent, v.latent + t * NL, NL);
        d2h(o->lg_adj, v.lgadj + t * NL, NL);
        d2h(o->mean_coi, v.mean_coi + t, 1);
        d2h(o->cell_t, v.cellT + t * NL, NL);
        d2h(o->cell_o, v.cellO + t * NL, NL);
        d2h(o->prior_eps_neg, v.pr_eps_neg + t * N, N);
        d2h(o->prior_eps_pos, v.pr_eps_pos + t * N, N);
        d2h(o->prior_coi, v.pr_coi + t * N, N);
        d2h(o->prior_r, v.pr_r + t * N, N);
        d2h(o->mean_coi_hyper_prior, v.hyper + t, 1);
        CUDA_TRY(cudaStreamSynchronize(stream));
    }

    void load_state(int t, const moire_chain_state *in)
    {
        const size_t N = v.N, NL = N * v.L, LA = (size_t)v.L * v.AP;
        h2d(v.m + t * N, in->m, N);
        h2d(v.r + t * N, in->r, N);
        h2d(v.eps_pos + t * N, in->eps_pos, N);
        h2d(v.eps_neg + t * N, in->eps_neg, N);
        h2d(v.p + t * LA, in->p, LA);
        h2d(v.latent + t * NL, (const moire::mask_t *)in->latent, NL);
        h2d(v.lgadj + t * NL, in->lg_adj, NL);
        h2d(v.mean_coi + t, in->mean_coi, 1);
        CUDA_TRY(cudaStreamSynchronize(stream));  // host buffers may go away after return
        eval_all(t, 0);
        initialised = true;
    }

    void read_diag(int chain, moire_chain_diagnostics *o)
    {
        // chain < 0: every chain of the engine in one pass ([T][...] arrays are contiguous)
        const size_t nc = chain < 0 ? (size_t)v.T : 1, t = chain < 0 ? 0 : (size_t)chain;
        const size_t N = v.N * nc, LA = (size_t)v.L * v.AP * nc;
        const size_t sN = v.N, sLA = (size_t)v.L * v.AP;
        d2h(o->eps_neg_accept, v.acc_eps_neg + t * sN, N);
        d2h(o->eps_pos_accept, v.acc_eps_pos + t * sN, N);
        d2h(o->m_accept, v.acc_m + t * sN, N);
        d2h(o->r_accept, v.acc_r + t * sN, N);
        d2h(o->m_r_accept, v.acc_mr + t * sN, N);
        d2h(o->sample_accept, v.acc_samples + t * sN, N);
        d2h(o->eps_neg_var, v.sd_eps_neg + t * sN, N);
        d2h(o->eps_pos_var, v.sd_eps_pos + t * sN, N);
        d2h(o->r_var, v.sd_r + t * sN, N);
        d2h(o->m_r_var, v.sd_mr + t * sN, N);
        d2h(o->p_accept, v.p_acc + t * sLA, LA);
        d2h(o->p_attempt, v.p_att + t * sLA, LA);
        d2h(o->p_prop_var, v.p_sd + t * sLA, LA);
        d2h(o->mean_coi_accept, v.mean_coi_acc + t, nc);
        d2h(o->mean_coi_var, v.mean_coi_sd + t, nc);
        CUDA_TRY(cudaStreamSynchronize(stream));
    }

    void sample_llik_async(int t)
    {
        ++launches;
        moire::k_sample_llik<<<(v.N + 127) / 128, 128, 0, stream>>>(v, t, scratchN.ptr);
        CUDA_TRY(cudaGetLastError());
    }
};

// ---- C ABI -------------------------------------------------------------------------------------
namespace
{
template <class F>
int guarded(moire_engine *e, F &&f)
{
    try
    {
        if (e) e->bind();
        f();
        return MOIRE_OK;
    }
    catch (const CudaError &ce)
    {
        g_last_error = ce.msg;
        if (e) e->last_error = ce.msg;
        return MOIRE_ECUDA;
    }
    catch (const InvalidArg &ia)
    {
        g_last_error = ia.msg;
        if (e) e->last_error = ia.msg;
        return MOIRE_EINVAL;
    }
    catch (const std::bad_alloc &)
    {
        g_last_error = "host allocation failed";
        if (e) e->last_error = g_last_error;
        return MOIRE_ENOMEM;
    }
    catch (...)
    {
        g_last_error = "unknown error";
        if (e) e->last_error = g_last_error;
        return MOIRE_EINVAL;
    }
}

void need_chain(moire_engine *e, int chain)
{
    if (!e) throw InvalidArg{"null engine"};
    if (chain < 0 || chain >= e->v.T) throw InvalidArg{"chain index out of range"};
}
}  // namespace

extern "C"
{
const char *moire_last_error(void) { return g_last_error.c_str(); }

const char *moire_engine_last_error(const moire_engine *e) { return e ? e->last_error.c_str() : ""; }

int moire_engine_create(const moire_data *data, const moire_params *params, int32_t num_chains,
                        int32_t chain_offset, const double *temps, int32_t device_id,
                        uint64_t seed, moire_engine **out)
{
    if (!out)
    {
        g_last_error = "null out pointer";
        return MOIRE_EINVAL;
    }
    *out = nullptr;
    moire_engine *e = nullptr;
    const int rc = guarded(nullptr, [&] {
        e = new moire_engine();
        e->create(data, params, num_chains, chain_offset, temps, device_id, seed);
    });
    if (rc != MOIRE_OK)
    {
        delete e;
        return rc;
    }
    *out = e;
    return MOIRE_OK;
}

int moire_engine_destroy(moire_engine *e)
{
    delete e;
    return MOIRE_OK;
}

int32_t moire_engine_a_pad(const moire_engine *e) { return e ? e->v.AP : 0; }
int32_t moire_engine_mask_words(void) { return MOIRE_MASK_WORDS; }
int32_t moire_engine_max_alleles(void) { return moire::kMaskBits; }
int32_t moire_engine_num_chains(const moire_engine *e) { return e ? e->v.T : 0; }

int moire_engine_init_chains(moire_engine *e, int32_t max_tries, int32_t *ill_count,
                             int32_t *still_ill, int32_t *first_bad_sample,
                             int32_t *first_bad_locus, int32_t *prior_nonfinite)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->init_chains(max_tries, ill_count, still_ill, first_bad_sample, first_bad_locus,
                       prior_nonfinite);
    });
}

int moire_engine_read_init_log(moire_engine *e, int32_t *records, int32_t max_records,
                               int32_t *count)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        const int32_t n = (int32_t)(e->init_log.size() / 3);
        if (count) *count = n;
        if (records)
            std::copy(e->init_log.begin(), e->init_log.begin() + 3 * (size_t)std::min(n, std::max(max_records, 0)),
                      records);
    });
}

int moire_engine_step(moire_engine *e, int32_t iteration)
{
    if (e && !e->initialised)
    {
        e->last_error = g_last_error = "step before init_chains/load_state";
        return MOIRE_ESTATE;
    }
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->step(iteration);
    });
}

int moire_engine_update(moire_engine *e, int32_t kind, int32_t iteration)
{
    if (e && !e->initialised)
    {
        e->last_error = g_last_error = "update before init_chains/load_state";
        return MOIRE_ESTATE;
    }
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->update(kind, iteration);
    });
}

int moire_engine_eval_all(moire_engine *e)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->eval_all(-1, 0);
    });
}

int moire_engine_reduce(moire_engine *e)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->reduce();
    });
}

int moire_engine_get_scalars(moire_engine *e, double *llik, double *prior)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->get_scalars(llik, prior);
    });
}

int moire_engine_device_scalars(moire_engine *e, void **dev_llik, void **dev_prior)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        if (dev_llik) *dev_llik = e->v.llik;
        if (dev_prior) *dev_prior = e->v.prior;
    });
}

int moire_engine_trim_memory(void)
{
    g_block_cache.trim();
    return MOIRE_OK;
}

int moire_engine_stream(moire_engine *e, void **stream)
{
    return guarded(e, [&] {
        if (!e || !stream) throw InvalidArg{"null argument"};
        *stream = (void *)e->stream;
    });
}

int moire_engine_set_temps(moire_engine *e, const double *temps)
{
    return guarded(e, [&] {
        if (!e || !temps) throw InvalidArg{"null argument"};
        CUDA_TRY(cudaMemcpyAsync(e->v.temp, temps, e->v.T * sizeof(double), cudaMemcpyHostToDevice,
                                 e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    });
}

int moire_engine_get_temps(moire_engine *e, double *temps)
{
    return guarded(e, [&] {
        if (!e || !temps) throw InvalidArg{"null argument"};
        CUDA_TRY(cudaMemcpyAsync(temps, e->v.temp, e->v.T * sizeof(double), cudaMemcpyDeviceToHost,
                                 e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    });
}

int moire_engine_dump_state(moire_engine *e, int32_t chain, moire_chain_state *out)
{
    return guarded(e, [&] {
        need_chain(e, chain);
        if (!out) throw InvalidArg{"null state"};
        e->dump_state(chain, out);
    });
}

int moire_engine_load_state(moire_engine *e, int32_t chain, const moire_chain_state *in)
{
    return guarded(e, [&] {
        need_chain(e, chain);
        if (!in) throw InvalidArg{"null state"};
        e->load_state(chain, in);
    });
}

int moire_engine_read_diagnostics(moire_engine *e, int32_t chain, moire_chain_diagnostics *out)
{
    return guarded(e, [&] {
        if (chain >= 0) need_chain(e, chain);
        if (!e) throw InvalidArg{"null engine"};
        if (!out) throw InvalidArg{"null diagnostics"};
        e->read_diag(chain, out);
    });
}

int moire_engine_read_sample_llik(moire_engine *e, int32_t chain, double *out)
{
    return guarded(e, [&] {
        need_chain(e, chain);
        if (!out) throw InvalidArg{"null output"};
        e->sample_llik_async(chain);
        e->d2h(out, e->scratchN.ptr, (size_t)e->v.N);
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    });
}

int moire_engine_record_draw(moire_engine *e, int32_t chain, double *p, int32_t *m,
                             double *eps_neg, double *eps_pos, double *r, double *sample_llik,
                             double *mean_coi, uint64_t *latent)
{
    return guarded(e, [&] {
        need_chain(e, chain);
        const moire::View &v = e->v;
        const size_t N = v.N, NL = N * v.L, LA = (size_t)v.L * v.AP, t = chain;
        if (sample_llik) e->sample_llik_async(chain);
        e->d2h(p, v.p + t * LA, LA);
        e->d2h(m, v.m + t * N, N);
        e->d2h(eps_neg, v.eps_neg + t * N, N);
        e->d2h(eps_pos, v.eps_pos + t * N, N);
        e->d2h(r, v.r + t * N, N);
        e->d2h(sample_llik, e->scratchN.ptr, N);
        e->d2h(mean_coi, v.mean_coi + t, 1);
        e->d2h((moire::mask_t *)latent, v.latent + t * NL, NL);
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    });
}

int moire_engine_counters(moire_engine *e, uint64_t *cell_evals, uint64_t *kernel_launches)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        if (cell_evals)
        {
            unsigned long long ev[16];
            CUDA_TRY(cudaMemcpyAsync(ev, e->v.evals, sizeof(ev), cudaMemcpyDeviceToHost, e->stream));
            CUDA_TRY(cudaStreamSynchronize(e->stream));
            *cell_evals = 0;
            for (int k = 0; k < 16; ++k) *cell_evals += ev[k];
        }
        if (kernel_launches) *kernel_launches = e->launches;
    });
}

int moire_engine_kind_evals(moire_engine *e, uint64_t *evals)
{
    return guarded(e, [&] {
        if (!e || !evals) throw InvalidArg{"null argument"};
        unsigned long long ev[16];
        CUDA_TRY(cudaMemcpyAsync(ev, e->v.evals, sizeof(ev), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        for (int k = 0; k < 8; ++k) evals[k] = ev[k];
    });
}

int moire_engine_timer_start(moire_engine *e)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        if (!e->timer_a)
        {
            CUDA_TRY(cudaEventCreate(&e->timer_a));
            CUDA_TRY(cudaEventCreate(&e->timer_b));
        }
        CUDA_TRY(cudaEventRecord(e->timer_a, e->stream));
    });
}

int moire_engine_timer_stop(moire_engine *e, double *ms)
{
    return guarded(e, [&] {
        if (!e || !ms) throw InvalidArg{"null argument"};
        if (!e->timer_a) throw InvalidArg{"timer_stop without timer_start"};
        CUDA_TRY(cudaEventRecord(e->timer_b, e->stream));
        CUDA_TRY(cudaEventSynchronize(e->timer_b));
        float f = 0;
        CUDA_TRY(cudaEventElapsedTime(&f, e->timer_a, e->timer_b));
        *ms = f;
    });
}

int moire_engine_set_timing(moire_engine *e, int32_t on)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->resolve_timing();
        e->timing = on != 0;
    });
}

int moire_engine_read_timing(moire_engine *e, double *ms, uint64_t *launches)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->resolve_timing();
        for (int k = 0; k < 9; ++k)
        {
            if (ms) ms[k] = e->kind_ms[k];
            if (launches) launches[k] = e->kind_launches[k];
        }
    });
}

int moire_engine_read_evaluator_timing(moire_engine *e, double *ms, uint64_t *launches)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->resolve_timing();
        if (ms) *ms = e->kind_ms[9];
        if (launches) *launches = e->kind_launches[9];
    });
}

int moire_engine_synchronize(moire_engine *e)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    });
}

int moire_simulate_data(int32_t num_samples, int32_t num_loci, const int32_t *num_alleles,
                        const double *allele_freqs, int32_t a_pad, const int32_t *sample_cois,
                        const double *sample_relatedness, double epsilon_pos, double epsilon_neg,
                        double missingness, uint64_t seed, int32_t device_id, uint64_t *obs_mask,
                        uint8_t *is_missing, uint64_t *true_mask)
{
    return guarded(nullptr, [&] {
        if (!num_alleles || !allele_freqs || !sample_cois || !sample_relatedness || !obs_mask || !is_missing)
            throw InvalidArg{"null argument"};
        if (num_samples < 1 || num_loci < 1 || a_pad < 1) throw InvalidArg{"empty shape"};
        for (int j = 0; j < num_loci; ++j)
            if (num_alleles[j] < 2 || num_alleles[j] > moire::kMaskBits || num_alleles[j] > a_pad)
                throw InvalidArg{"num_alleles outside 2.." + std::to_string(moire::kMaskBits) + " / a_pad"};
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw CudaError{"no CUDA device -- libmoire_b200 has no CPU fallback"};
        if (device_id < 0 || device_id >= ndev) throw InvalidArg{"device_id out of range"};
        CUDA_TRY(cudaSetDevice(device_id));
        const size_t N = num_samples, L = num_loci, NL = N * L;
        DevBuf<int> nall, coi;
        DevBuf<double> freqs, rel;
        DevBuf<moire::mask_t> obs, truth;
        DevBuf<unsigned char> miss;
        auto release = [&] { nall.release(); coi.release(); freqs.release(); rel.release(); obs.release(); truth.release(); miss.release(); };
        try
        {
            nall.alloc(L); coi.alloc(N); freqs.alloc(L * a_pad); rel.alloc(N); obs.alloc(NL); truth.alloc(NL); miss.alloc(NL);
            nall.upload(num_alleles, L, nullptr);
            coi.upload(sample_cois, N, nullptr);
            freqs.upload(allele_freqs, L * a_pad, nullptr);
            rel.upload(sample_relatedness, N, nullptr);
            k_simulate_cells<<<(unsigned)((NL + 255) / 256), 256>>>(num_samples, num_loci, a_pad, nall.ptr, freqs.ptr,
                                                                     coi.ptr, rel.ptr, epsilon_pos, epsilon_neg,
                                                                     missingness, seed, obs.ptr, miss.ptr, truth.ptr);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpy(obs_mask, obs.ptr, NL * sizeof(moire::mask_t), cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(is_missing, miss.ptr, NL, cudaMemcpyDeviceToHost));
            if (true_mask) CUDA_TRY(cudaMemcpy(true_mask, truth.ptr, NL * sizeof(moire::mask_t), cudaMemcpyDeviceToHost));
        }
        catch (...)
        {
            release();
            throw;
        }
        release();
    });
}

int moire_engine_diagnose(moire_engine *e, int32_t chain, int32_t *first_bad_sample,
                          int32_t *first_bad_locus, int32_t *prior_nonfinite)
{
    return guarded(e, [&] {
        need_chain(e, chain);
        if (prior_nonfinite) std::fill(prior_nonfinite, prior_nonfinite + 5, 0);
        int32_t bs = 0, bl = 0;
        e->diagnose(chain, &bs, &bl, prior_nonfinite);
        if (first_bad_sample) *first_bad_sample = bs;
        if (first_bad_locus) *first_bad_locus = bl;
    });
}

int moire_engine_device_iteration(moire_engine *e, int32_t **dev_iteration)
{
    return guarded(e, [&] {
        if (!e || !dev_iteration) throw InvalidArg{"null argument"};
        *dev_iteration = e->iter_dev.ptr;
    });
}

int moire_engine_device_temps(moire_engine *e, double **dev_temps)
{
    return guarded(e, [&] {
        if (!e || !dev_temps) throw InvalidArg{"null argument"};
        *dev_temps = e->v.temp;
    });
}

int moire_engine_captured_step_launches(moire_engine *e, int64_t *launches)
{
    return guarded(e, [&] {
        if (!e || !launches) throw InvalidArg{"null argument"};
        *launches = (int64_t)e->replay_step_launches;
    });
}

int moire_engine_add_launches(moire_engine *e, int64_t n)
{
    return guarded(e, [&] {
        if (!e) throw InvalidArg{"null engine"};
        e->launches = (uint64_t)((int64_t)e->launches + n);
    });
}

int moire_engine_get_timing(moire_engine *e, int32_t *on)
{
    return guarded(e, [&] {
        if (!e || !on) throw InvalidArg{"null argument"};
        *on = e->timing ? 1 : 0;
    });
}

int moire_engine_enqueue_record_draw(moire_engine *e, const int32_t *dev_sel, const moire_draw_store *store)
{
    return guarded(e, [&] {
        if (!e || !dev_sel || !store) throw InvalidArg{"null argument"};
        if (!store->p || !store->m || !store->eps_neg || !store->eps_pos || !store->r ||
            !store->sample_llik || !store->mean_coi)
            throw InvalidArg{"draw store: only `latent` may be null"};
        const size_t most = std::max({(size_t)e->v.N, (size_t)e->v.L * e->v.AP,
                                      store->latent ? (size_t)e->v.N * e->v.L : (size_t)0});
        const unsigned blocks = (unsigned)std::min<size_t>((most + 255) / 256, 592);
        ++e->launches;
        k_record_draw<<<blocks, 256, 0, e->stream>>>(e->v, dev_sel, *store);
        CUDA_TRY(cudaGetLastError());
    });
}

int moire_engine_subset_order(moire_engine *e, int32_t which, int32_t k, uint16_t *out)
{
    return guarded(e, [&] {
        if (!e || !out) throw InvalidArg{"null argument"};
        if (which < 0 || which > 2) throw InvalidArg{"which must be 0 (tail table), 1 (mask table) or 2 (unrolled)"};
        if (k < 1 || k > (which == 2 ? MOIRE_SMALL_K : moire::kIeMax)) throw InvalidArg{"k out of range for this table"};
        const size_t n = ((size_t)1 << k) - 1;
        DevBuf<unsigned short> buf;
        buf.alloc(n);
        k_dump_subset_order<<<1, 32, 0, e->stream>>>(which, k, buf.ptr);
        const cudaError_t rc = cudaMemcpyAsync(out, buf.ptr, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, e->stream);
        const cudaError_t rs = cudaStreamSynchronize(e->stream);
        buf.release();
        CUDA_TRY(rc);
        CUDA_TRY(rs);
    });
}

void moire_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4])
{
    moire::philox4x32_10(counter, key[0], key[1], out);
}
}  // extern "C"
