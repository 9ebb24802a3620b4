// The parallel-tempering ladder on the device (row f3 of SURVEY.md section 8).
//
// MCMC::swap_chains (src/mcmc.cpp:190-240) needs, per step, every chain's scalar log-likelihood and
// temperature and a handful of counters.  Keeping that state on the host cost a collective ->
// device-to-host copy -> cudaStreamSynchronize -> host pass -> host-to-device copy of the
// temperatures after EVERY step (7 % of an 8-GPU step) and made it impossible to queue steps.
// Here the ladder lives in device memory (a replica per rank; every rank runs the identical pass
// from the identical gathered scalars and the shared Philox stream, so no decision is ever
// communicated), the pass is a one-warp kernel that follows the all-gather on the engine's stream,
// and the host reads traces / accumulators back in batches -- only where MCMC::adapt_temp is due
// (src/mcmc.cpp:244-304: the spline fit stays on the host, every temp_adapt_steps steps) or a
// result is asked for.
//
// The per-pair arithmetic is ONE function (swap_pair) compiled for host and device: the host
// ladder of mcmc_driver.cpp (ladder-only mode, callback exchanges, the golden-vector tests against
// the reference's own MCMC::swap_chains) and the device kernel run the same code.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "rng.cuh"

namespace moire
{
// control words (int64) of the device ladder
enum LadderCtl
{
    LC_NUM_SWAPS = 0,   // MCMC::num_swaps
    LC_EVEN_SWAP = 1,   // MCMC::even_swap
    LC_SWAP_PASSES = 2, // pass counter keying the uniform stream
    LC_STEP = 3,        // step the next enqueued step will run as
    LC_PHASE = 4,       // 0 burn-in, 1 sampling
    LC_TRACE_COUNT = 5, // steps recorded since the host's last read-back
    LC_NDRAWS = 6,      // draws stored since the host's last read-back
    LC_STOP = 7,        // some rank asked to stop (sticky)
    LC_WORDS = 8
};

struct LadderDev
{
    int T, world, first, count, cmax, per;  // per = 2 * cmax + 2 doubles per rank in the exchange
    int pre_adapt_steps, thin, burnin;
    int trace_cap, draw_cap;
    unsigned long long seed;
    // engine
    const double *eng_llik, *eng_prior;  // [count] per-chain sums of this rank (k_reduce)
    double *eng_temp;                    // [count] Chain::temp of this rank's chains
    int *eng_iter;                       // the word the engine's kernels read the iteration from
    // exchange
    double *send;                // [per]  llik[cmax], prior[cmax], minutes, stop
    const double *recv;          // [world][per]; world == 1: the send buffer itself
    const double *host_flags;    // pinned, device-visible: [0] minutes on this rank's clock, [1] stop
    // ladder, by chain id unless noted
    double *chain_llik, *chain_prior, *chain_temp;
    int *chain_hot;
    int *swap_indices;           // ladder position -> chain id
    long long *swap_acc;         // [T - 1]
    double *swap_barriers;       // [T - 1]
    long long *ctl;              // [LC_WORDS]
    double *minutes;             // [1] rank 0's clock as of the last gathered step
    double *trace;               // [trace_cap][3] llik, prior, posterior of the chain at position 0
    int *swap_store;             // [trace_cap]
    int *draw_sel;               // [2] local chain to record next (or -1), slot
    int *draw_step;              // [draw_cap]
};

// One candidate pair (i, i + 1) of a swap pass (src/mcmc.cpp:194-228).  log_u = log of the pair's
// uniform.  Returns true when the pair swapped.  Pairs of one pass are disjoint, so the device
// runs them across the lanes of a warp.
MOIRE_HD bool swap_pair(int i, bool burnin, bool accumulate_barrier, double log_u, const double *chain_llik,
                        double *chain_temp, int *chain_hot, int *swap_indices, double *swap_barriers,
                        long long *swap_acc)
{
    const int a = swap_indices[i], b = swap_indices[i + 1];
    const double V_a = -chain_llik[a], temp_a = chain_temp[a];
    const double V_b = -chain_llik[b], temp_b = chain_temp[b];
    const double ratio = (temp_b - temp_a) * (V_b - V_a);
    const double e = exp(ratio);
    const double rate = e < 1.0 ? e : 1.0;  // std::min(1.0, exp(ratio)): NaN -> 1.0
    if (accumulate_barrier) swap_barriers[i] += 1.0 - rate;
    if ((ratio > 0 || log_u < ratio) && !(ratio != ratio))
    {
        swap_indices[i] = b;
        swap_indices[i + 1] = a;
        chain_temp[a] = temp_b;
        chain_temp[b] = temp_a;
        if (i == 0)
        {
            chain_hot[b] = 1;
            chain_hot[a] = 0;
        }
        if (!burnin) swap_acc[i] += 1;
        return true;
    }
    return false;
}

// The uniform of pair `pair` in pass `pass` (the reference draws R::runif, src/mcmc.cpp:212): every
// rank must draw the same number, so the stream is keyed by (seed, pass counter, pair, step) only.
MOIRE_HD double swap_uniform(unsigned long long seed, uint32_t pass, int pair, int step)
{
    Stream s;
    s.open(seed, pass, KIND_SWAP, 0xFFFFu, (uint32_t)pair, (uint32_t)(step & 0xFFFF));
    return s.next_uniform();
}

// ladder.cu; each returns the cudaError_t of its launch
cudaError_t launch_ladder_begin(const LadderDev &ld, cudaStream_t stream);
cudaError_t launch_ladder_pack(const LadderDev &ld, cudaStream_t stream);
cudaError_t launch_ladder_step(const LadderDev &ld, cudaStream_t stream);
}  // namespace moire
