// Device side of the parallel-tempering ladder: see ladder.cuh.
//
// Per step, after the engine's kernels (which end in k_reduce), the stream carries
//   k_ladder_pack   this rank's (llik, prior) sums + its clock / stop words -> the exchange buffer
//   ncclAllGather   (world > 1; enqueued by the driver)
//   k_ladder_step   one warp: unpack, trace of the chain at ladder position 0, bookkeeping of the
//                   draw the engine just stored, the even/odd swap pass with the candidate pairs
//                   across lanes, the new temperatures into the engine, the next step's iteration
// and nothing returns to the host.
#include "ladder.cuh"

namespace moire
{
namespace
{
// Which chain of this rank the NEXT step has to store a draw of (MCMC::sample, src/mcmc.cpp:327-347:
// the chain flagged hot, every thin-th step), or -1.  At most one chain is hot: the flag is set only
// on the chain at ladder position 0 and cleared when a swap moves it away (src/mcmc.cpp:221-225).
__device__ void select_next_draw(const LadderDev &ld, long long step, long long phase)
{
    int sel = -1;
    if (phase == 1 && (ld.thin == 0 || step % ld.thin == 0))
        for (int k = 0; k < ld.count && sel < 0; ++k)
            if (ld.chain_hot[ld.first + k] || ld.T == 1) sel = k;
    ld.draw_sel[0] = sel;
    ld.draw_sel[1] = (int)ld.ctl[LC_NDRAWS];
}

__global__ void k_ladder_begin(const LadderDev ld)
{
    // after the host changed the ladder (start of a run, adapt_temp, a test hook): temperatures
    // into the engine, the iteration word, the draw selector of the first step
    const int lane = threadIdx.x;
    for (int i = lane; i < ld.count; i += 32) ld.eng_temp[i] = ld.chain_temp[ld.first + i];
    __syncwarp();
    if (lane == 0)
    {
        const long long step = ld.ctl[LC_STEP], phase = ld.ctl[LC_PHASE];
        *ld.eng_iter = (int)(phase == 0 ? step : ld.burnin + step);
        select_next_draw(ld, step, phase);
    }
}

__global__ void k_ladder_pack(const LadderDev ld)
{
    const int i = threadIdx.x;
    if (i < ld.cmax)
    {
        ld.send[i] = i < ld.count ? ld.eng_llik[i] : 0.0;
        ld.send[ld.cmax + i] = i < ld.count ? ld.eng_prior[i] : 0.0;
    }
    if (i == 0)
    {
        ld.send[2 * ld.cmax] = ld.host_flags[0];
        ld.send[2 * ld.cmax + 1] = ld.host_flags[1];
    }
}

__global__ void k_ladder_step(const LadderDev ld)
{
    const int lane = threadIdx.x;
    const int T = ld.T;
    const long long step = ld.ctl[LC_STEP], phase = ld.ctl[LC_PHASE];
    const bool burnin = phase == 0;
    // -- every rank's scalars, by chain id (moire_mcmc_partition: first_r = r T / world)
    bool stop = false;
    for (int r = 0; r < ld.world; ++r)
    {
        const int f = (int)((long long)r * T / ld.world), c = (int)((long long)(r + 1) * T / ld.world) - f;
        const double *src = ld.recv + (size_t)r * ld.per;
        for (int i = lane; i < c; i += 32)
        {
            ld.chain_llik[f + i] = src[i];
            ld.chain_prior[f + i] = src[ld.cmax + i];
        }
        stop |= src[2 * ld.cmax + 1] != 0.0;
    }
    __syncwarp();
    if (lane == 0)
    {
        *ld.minutes = ld.recv[2 * ld.cmax];  // rank 0's clock decides max_runtime for every rank
        if (stop) ld.ctl[LC_STOP] = 1;
        // -- traces of the chain at ladder position 0 (src/mcmc.cpp:175-177, 349-351)
        const long long k = ld.ctl[LC_TRACE_COUNT];
        if (k < ld.trace_cap)
        {
            const int c = ld.swap_indices[0];
            ld.trace[3 * k + 0] = ld.chain_llik[c];
            ld.trace[3 * k + 1] = ld.chain_prior[c];
            ld.trace[3 * k + 2] = ld.chain_llik[c] * ld.chain_temp[c] + ld.chain_prior[c];  // Chain::get_posterior
        }
        // -- the draw k_record_draw stored during this step
        if (ld.draw_sel[0] >= 0 && ld.draw_sel[1] < ld.draw_cap)
        {
            ld.draw_step[ld.draw_sel[1]] = (int)step;
            ld.ctl[LC_NDRAWS] = ld.draw_sel[1] + 1;
        }
    }
    __syncwarp();
    // -- MCMC::swap_chains (src/mcmc.cpp:190-240): the candidate pairs of a pass are disjoint.
    // MCMC::burnin skips the pass for a single chain (src/mcmc.cpp:179), MCMC::sample does not
    // (:355: swap_store grows, even_swap toggles)
    const bool run_pass = !(burnin && T == 1);
    const bool even = ld.ctl[LC_EVEN_SWAP] != 0;
    const uint32_t pass = (uint32_t)ld.ctl[LC_SWAP_PASSES];
    const bool acc_barrier = run_pass && burnin && step > ld.pre_adapt_steps;
    for (int i = (even ? 1 : 0) + 2 * lane; run_pass && i < T - 1; i += 64)
    {
        const double u = log(swap_uniform(ld.seed, pass, i, (int)step));
        swap_pair(i, burnin, acc_barrier, u, ld.chain_llik, ld.chain_temp, ld.chain_hot, ld.swap_indices,
                  ld.swap_barriers, ld.swap_acc);
    }
    __syncwarp();
    for (int i = lane; i < ld.count; i += 32) ld.eng_temp[i] = ld.chain_temp[ld.first + i];
    if (lane == 0)
    {
        const long long k = ld.ctl[LC_TRACE_COUNT];
        if (k < ld.trace_cap) ld.swap_store[k] = ld.swap_indices[0];
        ld.ctl[LC_TRACE_COUNT] = k + 1;
        if (acc_barrier) ld.ctl[LC_NUM_SWAPS] += 1;
        if (run_pass)
        {
            ld.ctl[LC_EVEN_SWAP] = even ? 0 : 1;
            ld.ctl[LC_SWAP_PASSES] = (long long)pass + 1;
        }
        ld.ctl[LC_STEP] = step + 1;
        *ld.eng_iter = (int)(burnin ? step + 1 : ld.burnin + step + 1);
        select_next_draw(ld, step + 1, phase);
    }
}
}  // namespace

cudaError_t launch_ladder_begin(const LadderDev &ld, cudaStream_t stream)
{
    k_ladder_begin<<<1, 32, 0, stream>>>(ld);
    return cudaGetLastError();
}

cudaError_t launch_ladder_pack(const LadderDev &ld, cudaStream_t stream)
{
    k_ladder_pack<<<1, ld.cmax < 32 ? 32 : (ld.cmax + 31) / 32 * 32, 0, stream>>>(ld);
    return cudaGetLastError();
}

cudaError_t launch_ladder_step(const LadderDev &ld, cudaStream_t stream)
{
    k_ladder_step<<<1, 32, 0, stream>>>(ld);
    return cudaGetLastError();
}
}  // namespace moire
