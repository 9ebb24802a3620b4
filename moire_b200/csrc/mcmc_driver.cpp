// Parallel-tempering driver (include/moire_mcmc.h).
//
// What the reference keeps on its master thread, in double: the even/odd temperature swaps
// (MCMC::swap_chains, src/mcmc.cpp:190-240), the adaptive ladder (MCMC::adapt_temp,
// src/mcmc.cpp:244-304, with the monotone cubic spline of src/include/spline/src/spline.h), the
// traces and the stores of the temp = 1 chain (src/mcmc.cpp:175-177, 327-355) and the two loops
// of run_mcmc (src/main.cpp:52-93).  Everything per-chain goes through the engine's C ABI; no
// likelihood arithmetic runs here.
//
// Two homes for the ladder.  With an engine and a device-side exchange (one rank, or the NCCL
// all-gather) the ladder lives ON THE DEVICE (ladder.cuh): a step is one graph launch -- the
// engine's kernels, the pack, the all-gather, the one-warp swap pass -- and the host neither waits
// for it nor copies anything; traces, accumulators and draws are read back in batches, where
// adapt_temp is due or a result is asked for.  Otherwise (ladder-only mode for the CPU tests, a
// callback exchange such as gloo, MOIRE_B200_HOST_LADDER=1) the same pass runs on the host after a
// per-step read-back, as in round 1.  Both call the one swap_pair() of ladder.cuh.
//
// The spline below restates, in double, the parts of the reference's vendored tk::spline
// (src/include/spline/src/spline.h, (c) Tino Kluge, GPL-2+; moire itself is GPL-3) that
// MCMC::adapt_temp exercises -- cspline with natural boundaries, make_monotonic, solve_one and its
// cubic solver -- closely enough to reproduce the reference's adapted ladders (the golden vectors
// of tests/golden/make_golden_driver.py): that is third-party arithmetic this file credits, not
// an original design.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>

#include "../../include/moire_mcmc.h"
#include "ladder.cuh"
#include "rng.cuh"

namespace
{
// ---- monotone cubic spline + first-root solve ---------------------------------------------------
// Restates what adapt_temp uses of tk::spline: a C2 cubic spline with zero second derivative at
// both ends, slopes then limited so that the interpolant is monotone wherever the data are, and
// "first x with s(x) = y" by scanning the segments left to right.
struct MonotoneSpline
{
    std::vector<double> x, y, b, c, d;

    void coefficients_from_slopes()
    {
        const size_t n = x.size();
        for (size_t i = 0; i + 1 < n; ++i)
        {
            const double h = x[i + 1] - x[i];
            c[i] = (3.0 * (y[i + 1] - y[i]) / h - (2.0 * b[i] + b[i + 1])) / h;
            d[i] = ((b[i + 1] - b[i]) / (3.0 * h) - 2.0 / 3.0 * c[i]) / h;
        }
    }

    void fit(const double *xs, const double *ys, int n)
    {
        x.assign(xs, xs + n);
        y.assign(ys, ys + n);
        b.assign(n, 0.0);
        c.assign(n, 0.0);
        d.assign(n, 0.0);
        // natural spline: tridiagonal system for the quadratic coefficients, c[0] = c[n-1] = 0
        std::vector<double> diag(n, 2.0), up(n, 0.0), lo(n, 0.0), rhs(n, 0.0);
        for (int i = 1; i + 1 < n; ++i)
        {
            lo[i] = (x[i] - x[i - 1]) / 3.0;
            diag[i] = 2.0 / 3.0 * (x[i + 1] - x[i - 1]);
            up[i] = (x[i + 1] - x[i]) / 3.0;
            rhs[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
        }
        for (int i = 1; i < n; ++i)  // forward elimination
        {
            const double w = lo[i] / diag[i - 1];
            diag[i] -= w * up[i - 1];
            rhs[i] -= w * rhs[i - 1];
        }
        c[n - 1] = rhs[n - 1] / diag[n - 1];
        for (int i = n - 2; i >= 0; --i) c[i] = (rhs[i] - up[i] * c[i + 1]) / diag[i];
        for (int i = 0; i + 1 < n; ++i)
        {
            const double h = x[i + 1] - x[i];
            d[i] = (c[i + 1] - c[i]) / (3.0 * h);
            b[i] = (y[i + 1] - y[i]) / h - (2.0 * c[i] + c[i + 1]) * h / 3.0;
        }
        if (n >= 2)
        {
            const double h = x[n - 1] - x[n - 2];
            b[n - 1] = 3.0 * d[n - 2] * h * h + 2.0 * c[n - 2] * h + b[n - 2];
        }
        make_monotone();
    }

    void make_monotone()
    {
        const int n = (int)x.size();
        bool changed = false;
        for (int i = 0; i < n; ++i)
        {
            const double yl = y[std::max(i - 1, 0)], yr = y[std::min(i + 1, n - 1)];
            const bool rising = yl <= y[i] && y[i] <= yr, falling = yl >= y[i] && y[i] >= yr;
            if ((rising && b[i] < 0.0) || (falling && b[i] > 0.0))
            {
                b[i] = 0.0;
                changed = true;
            }
        }
        for (int i = 0; i + 1 < n; ++i)
        {
            const double secant = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
            if (secant == 0.0)
            {
                if (b[i] != 0.0 || b[i + 1] != 0.0)
                {
                    b[i] = b[i + 1] = 0.0;
                    changed = true;
                }
            }
            else if ((secant > 0.0 && b[i] >= 0.0 && b[i + 1] >= 0.0) ||
                     (secant < 0.0 && b[i] <= 0.0 && b[i + 1] <= 0.0))
            {
                const double ratio = std::sqrt(b[i] * b[i] + b[i + 1] * b[i + 1]) / std::fabs(secant);
                if (ratio > 3.0)
                {
                    b[i] *= 3.0 / ratio;
                    b[i + 1] *= 3.0 / ratio;
                    changed = true;
                }
            }
        }
        if (changed) coefficients_from_slopes();
    }

    // real roots of a0 + a1 t + a2 t^2 + a3 t^3, ascending
    static int cubic_roots(double a0, double a1, double a2, double a3, double out[3])
    {
        int nr = 0;
        const double eps = std::numeric_limits<double>::epsilon();
        if (a3 == 0.0)
        {
            if (a2 == 0.0)
            {
                if (a1 == 0.0)
                {
                    if (a0 == 0.0) out[nr++] = 0.0;
                    return nr;
                }
                out[nr++] = -a0 / a1;
                return nr;
            }
            const double p = 0.5 * a1 / a2, q = a0 / a2, disc = p * p - q;
            const double tol = (6.0 * p * p + 3.0 * std::fabs(q) + std::fabs(disc)) * 0.5 * eps;
            if (std::fabs(disc) <= tol)
                out[nr++] = -p;
            else if (disc > 0.0)
            {
                out[nr++] = -p - std::sqrt(disc);
                out[nr++] = -p + std::sqrt(disc);
            }
            for (int i = 0; i < nr; ++i)
            {
                const double f = (a2 * out[i] + a1) * out[i] + a0, f1 = 2.0 * a2 * out[i] + a1;
                if (std::fabs(f1) > 1e-8) out[i] -= f / f1;
            }
            return nr;
        }
        // monic: t^3 + A t^2 + B t + C; depressed with t = z - A/3: z^3 - 3 P z - 2 Q = 0
        const double C = a0 / a3, B = a1 / a3, A = a2 / a3;
        const double P = -B / 3.0 + A * A / 9.0;
        const double R = 2.0 * A * A - 9.0 * B;
        const double Q = -0.5 * C - A * R / 54.0;
        const double disc = P * P * P - Q * Q;
        const double P_err = eps * (std::fabs(B) + 4.0 / 9.0 * A * A + std::fabs(P));
        const double R_err = eps * (6.0 * A * A + 18.0 * std::fabs(B) + std::fabs(R));
        const double Q_err = 0.5 * std::fabs(C) * eps +
                             std::fabs(A) / 54.0 * (R_err + std::fabs(R) * 3.0 * eps) +
                             std::fabs(Q) * eps;
        const double disc_err = P * P * (3.0 * P_err + std::fabs(P) * 2.0 * eps) +
                                std::fabs(Q) * (2.0 * Q_err + std::fabs(Q) * eps) +
                                std::fabs(disc) * eps;
        double z[3];
        if (std::fabs(disc) <= disc_err)
        {
            if (std::fabs(P) <= P_err)
                z[nr++] = 0.0;
            else
            {
                z[nr++] = 2.0 * Q / P;
                z[nr++] = -Q / P;
            }
        }
        else if (disc > 0.0)
        {
            const double phi = std::acos(Q / (P * std::sqrt(P))) / 3.0, amp = 2.0 * std::sqrt(P);
            const double third = 2.0943951023931954923;  // 2 pi / 3
            z[nr++] = amp * std::cos(phi);
            z[nr++] = amp * std::cos(phi - third);
            z[nr++] = amp * std::cos(phi - 2.0 * third);
        }
        else
        {
            const double u = std::copysign(std::cbrt(std::fabs(Q) + std::sqrt(-disc)), Q >= 0 ? 1.0 : -1.0);
            z[nr++] = u + P / u;
        }
        for (int i = 0; i < nr; ++i)
        {
            double t = z[i] - A / 3.0;
            const double f = ((t + A) * t + B) * t + C, f1 = (3.0 * t + 2.0 * A) * t + B;
            if (std::fabs(f1) > 1e-8) t -= f / f1;
            out[i] = t;
        }
        if (a0 == 0.0)  // an exact root at the knot stays exact
        {
            int best = 0;
            for (int i = 1; i < nr; ++i)
                if (std::fabs(out[i]) < std::fabs(out[best])) best = i;
            out[best] = 0.0;
        }
        std::sort(out, out + nr);
        return nr;
    }

    double solve_first(double target) const
    {
        const size_t n = x.size();
        for (size_t i = 0; i + 1 < n; ++i)
        {
            double roots[3];
            const int nr = cubic_roots(y[i] - target, b[i], c[i], d[i], roots);
            const double hprev = i > 0 ? x[i] - x[i - 1] : 0.0;
            const double slack = std::numeric_limits<double>::epsilon() * 512.0 * std::min(hprev, 1.0);
            for (int k = 0; k < nr; ++k)
                if (-slack <= roots[k] && roots[k] < x[i + 1] - x[i]) return x[i] + roots[k];
        }
        return std::numeric_limits<double>::quiet_NaN();
    }
};

struct DriverError
{
    int code;
    std::string msg;
};

thread_local std::string g_mcmc_error;
}  // namespace

namespace
{
// ---- NCCL, bound at run time ---------------------------------------------------------------------
// The ladder's one exchange (T scalars per step) runs as an ncclAllGather enqueued on the engine's
// stream straight from the device-resident sums: no host staging and no callback into the
// embedding language inside the step loop.  The library is the one the process already holds
// (torch's bundled libnccl.so.2, or the path the caller names); only the four entry points below
// are used, through their stable C signatures (nccl.h: ncclGetUniqueId, ncclCommInitRank,
// ncclAllGather, ncclCommDestroy).
struct NcclApi
{
    struct UniqueId { char internal[128]; };
    void *lib = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    static constexpr int kFloat64 = 8;  // ncclFloat64 / ncclDouble

    bool load(const char *path, std::string &err)
    {
        if (lib) return true;
        const char *names[] = {path && path[0] ? path : nullptr, "libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
        {
            if (!n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib)
        {
            err = std::string("cannot load NCCL: ") + dlerror();
            return false;
        }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllGather || !CommDestroy)
        {
            err = "NCCL library lacks an entry point";
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
// one communicator per process, kept for its lifetime: creating one costs a few hundred ms
void *g_nccl_comm = nullptr;
int g_nccl_world = 0, g_nccl_rank = -1;
int g_nccl_users = 0;  // live drivers holding g_nccl_comm
}  // namespace

struct moire_mcmc
{
    moire_engine *engine = nullptr;
    moire_mcmc_params prm{};
    int N = 0, L = 0, AP = 0;
    int T = 0, rank = 0, world = 1, first = 0, count = 0, cmax = 0;
    uint64_t seed = 0;
    bool allow_rel = true;
    bool initialised = false;
    std::string last_error;

    moire_allgather_fn gather = nullptr;
    void *gather_ctx = nullptr;
    // native exchange (moire_mcmc_init_nccl)
    void *nccl_comm = nullptr;
    double *send_d = nullptr, *recv_d = nullptr;   // [per], [world * per]
    double *send_pin = nullptr, *recv_pin = nullptr;
    cudaStream_t stream = nullptr;

    // ladder (src/mcmc.h:47-52) -- identical on every rank
    std::vector<int32_t> swap_indices;     // ladder position -> chain id
    std::vector<int64_t> swap_acceptances;  // [T-1]
    std::vector<double> swap_barriers;      // [T-1]
    std::vector<double> temp_gradient;      // [T]
    std::vector<double> chain_temp;         // [T] by chain id (Chain::temp)
    std::vector<int> chain_hot;             // [T] by chain id (Chain::hot)
    std::vector<double> chain_llik, chain_prior;  // [T] by chain id, after the last step
    int64_t num_swaps = 0;
    bool even_swap = true;
    uint32_t swap_passes = 0;
    std::vector<double> forced_u;

    // traces and stores (src/mcmc.h:25-45)
    std::vector<double> llik_burnin, prior_burnin, posterior_burnin;
    std::vector<double> llik_sample, prior_sample, posterior_sample;
    std::vector<int32_t> swap_store;
    std::vector<int32_t> draw_step, m_store;
    std::vector<double> p_store, eps_neg_store, eps_pos_store, r_store, data_llik_store,
        mean_coi_store;
    std::vector<uint64_t> latent_store;

    std::vector<int32_t> init_ill_count, init_prior_nonfinite;  // per local chain, [count] / [count][5]
    std::vector<double> send, recv;
    std::chrono::steady_clock::time_point t_start;
    double elapsed_minutes_rank0 = 0.0;
    bool local_stop = false, stop_any = false;  // the progress callback's request, here / on any rank

    // ---- the ladder on the device (ladder.cuh) ----
    bool dev_mode = false;
    moire::LadderDev ld{};
    std::vector<void *> dev_allocs;
    double *host_flags = nullptr;       // pinned + mapped: minutes, stop
    moire_draw_store store{};           // device-resident draws since the last read-back
    cudaGraphExec_t step_graph[2] = {nullptr, nullptr};  // burn-in step, sampling step
    int64_t graph_launches[2] = {0, 0};  // kernel launches one replay stands for
    bool graph_ok = true;
    bool batch_open = false, host_dirty = true;
    int next_step = -1, next_phase = -1;
    int pending = 0, pending_phase = 0, draws_bound = 0;
    static constexpr int kTraceCap = 256;

    ~moire_mcmc()
    {
        if (stream) cudaStreamSynchronize(stream);  // queued steps still write the ladder and the draw store
        for (auto &g : step_graph)
            if (g) cudaGraphExecDestroy(g);
        for (void *p : dev_allocs) cudaFree(p);
        if (host_flags) cudaFreeHost(host_flags);
        if (send_d) cudaFree(send_d);
        if (recv_d) cudaFree(recv_d);
        if (send_pin) cudaFreeHost(send_pin);
        if (recv_pin) cudaFreeHost(recv_pin);
        if (nccl_comm && nccl_comm == g_nccl_comm) --g_nccl_users;
        if (engine) moire_engine_destroy(engine);
    }

    void cu(cudaError_t rc, const char *what)
    {
        if (rc != cudaSuccess)
            throw DriverError{MOIRE_ECUDA, std::string(what) + ": " + cudaGetErrorString(rc)};
    }
    void nc(int rc, const char *what)
    {
        if (rc != 0)
            throw DriverError{MOIRE_ECUDA, std::string(what) + ": " +
                                               (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error")};
    }

    // recv[r * per + i] = send_r[i] from host buffers, through whichever exchange is installed
    void host_allgather(const double *snd, int per, double *rcv)
    {
        if (nccl_comm)
        {
            std::memcpy(send_pin, snd, sizeof(double) * per);
            cu(cudaMemcpyAsync(send_d, send_pin, sizeof(double) * per, cudaMemcpyHostToDevice, stream), "h2d");
            nc(g_nccl.AllGather(send_d, recv_d, (size_t)per, NcclApi::kFloat64, nccl_comm, stream), "ncclAllGather");
            cu(cudaMemcpyAsync(recv_pin, recv_d, sizeof(double) * per * world, cudaMemcpyDeviceToHost, stream), "d2h");
            cu(cudaStreamSynchronize(stream), "sync");
            std::memcpy(rcv, recv_pin, sizeof(double) * per * world);
            return;
        }
        if (!gather) throw DriverError{MOIRE_ESTATE, "world_size > 1 needs moire_mcmc_init_nccl or moire_mcmc_set_allgather"};
        if (gather(gather_ctx, snd, per, rcv) != 0)
            throw DriverError{MOIRE_ESTATE, "all-gather callback failed"};
    }

    void ck(int rc, const char *what)
    {
        if (rc != MOIRE_OK)
            throw DriverError{rc, std::string(what) + ": " + moire_engine_last_error(engine)};
    }

    bool owns(int chain) const { return chain >= first && chain < first + count; }

    // ladder-only mode (device_id < 0): no engine; the local scalars come from
    // moire_mcmc_set_local_scalars.  Host-logic tests and CPU multi-process tests use it.
    std::vector<double> injected_llik, injected_prior;

    void push_local_temps()
    {
        if (engine) ck(moire_engine_set_temps(engine, chain_temp.data() + first), "set_temps");
    }

    void engine_step(int iteration)
    {
        if (engine) ck(moire_engine_step(engine, iteration), "step");
    }

    void local_scalars(double *llik, double *prior)
    {
        if (engine)
        {
            ck(moire_engine_get_scalars(engine, llik, prior), "get_scalars");
            return;
        }
        for (int i = 0; i < count; ++i)
        {
            llik[i] = i < (int)injected_llik.size() ? injected_llik[i] : 0.0;
            prior[i] = i < (int)injected_prior.size() ? injected_prior[i] : 0.0;
        }
    }

    // every rank's (llik, prior) of its chains -> chain_llik / chain_prior of all T chains
    void exchange_scalars()
    {
        const double minutes =
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() / 60.0;
        const int per = 2 * cmax + 2;
        if (world > 1 && nccl_comm && engine)
        {
            // device to device: the sums k_reduce left on the engine's stream, the collective on
            // the same stream, one read-back of all T pairs
            void *dl = nullptr, *dp = nullptr;
            ck(moire_engine_device_scalars(engine, &dl, &dp), "device_scalars");
            send_pin[0] = minutes;
            send_pin[1] = local_stop ? 1.0 : 0.0;
            cu(cudaMemcpyAsync(send_d, dl, sizeof(double) * count, cudaMemcpyDeviceToDevice, stream), "d2d");
            cu(cudaMemcpyAsync(send_d + cmax, dp, sizeof(double) * count, cudaMemcpyDeviceToDevice, stream), "d2d");
            cu(cudaMemcpyAsync(send_d + 2 * cmax, send_pin, 2 * sizeof(double), cudaMemcpyHostToDevice, stream), "h2d");
            nc(g_nccl.AllGather(send_d, recv_d, (size_t)per, NcclApi::kFloat64, nccl_comm, stream), "ncclAllGather");
            cu(cudaMemcpyAsync(recv_pin, recv_d, sizeof(double) * per * world, cudaMemcpyDeviceToHost, stream), "d2h");
            cu(cudaStreamSynchronize(stream), "sync");
            std::memcpy(recv.data(), recv_pin, sizeof(double) * per * world);
        }
        else
        {
            std::fill(send.begin(), send.end(), 0.0);
            local_scalars(send.data(), send.data() + cmax);
            send[2 * cmax] = minutes;
            send[2 * cmax + 1] = local_stop ? 1.0 : 0.0;
            if (world == 1)
                std::copy(send.begin(), send.end(), recv.begin());
            else
                host_allgather(send.data(), per, recv.data());
        }
        for (int r = 0; r < world; ++r)
        {
            int32_t f, c;
            moire_mcmc_partition(T, r, world, &f, &c);
            for (int i = 0; i < c; ++i)
            {
                chain_llik[f + i] = recv[(size_t)r * per + i];
                chain_prior[f + i] = recv[(size_t)r * per + cmax + i];
            }
        }
        elapsed_minutes_rank0 = recv[2 * cmax];
        // the stop request of any rank's progress callback reaches every rank with the same step's
        // exchange, so all ranks leave the loop together (nobody is left waiting in a collective)
        for (int r = 0; r < world; ++r) stop_any = stop_any || recv[(size_t)r * per + 2 * cmax + 1] != 0.0;
    }

    double uniform_for_pair(int step, int pair)
    {
        if (!forced_u.empty())
        {
            const double u = forced_u.front();
            forced_u.erase(forced_u.begin());
            return u;
        }
        // the reference draws R::runif here (src/mcmc.cpp:212); every rank must draw the same
        // number, so the stream is keyed by (seed, pass counter, pair) only
        return moire::swap_uniform(seed, swap_passes, pair, step);
    }

    // MCMC::swap_chains (src/mcmc.cpp:190-240)
    bool swap_pass(int step, bool burnin)
    {
        bool moved = false;
        static_assert(sizeof(long long) == sizeof(int64_t), "swap counters");
        for (int i = even_swap ? 1 : 0; i < T - 1; i += 2)
        {
            const double u = std::log(uniform_for_pair(step, i));
            moved |= moire::swap_pair(i, burnin, burnin && step > prm.pre_adapt_steps, u, chain_llik.data(),
                                      chain_temp.data(), chain_hot.data(), swap_indices.data(),
                                      swap_barriers.data(), reinterpret_cast<long long *>(swap_acceptances.data()));
        }
        swap_store.push_back(swap_indices[0]);
        if (burnin && step > prm.pre_adapt_steps) num_swaps++;
        even_swap = !even_swap;
        ++swap_passes;
        return moved;
    }

    // MCMC::adapt_temp (src/mcmc.cpp:244-304)
    bool adapt_ladder()
    {
        std::vector<double> new_gradient(temp_gradient);
        if (T >= 3)
        {
            std::vector<double> cum(T, 0.0), grad_up(temp_gradient.rbegin(), temp_gradient.rend());
            for (int i = 1; i < T; ++i)
                cum[i] = cum[i - 1] + swap_barriers[T - 1 - i] / ((double)num_swaps / 2.0);
            for (int i = 0; i + 1 < T; ++i)
                if (!(grad_up[i] < grad_up[i + 1])) return false;  // the spline needs increasing x
            MonotoneSpline s;
            s.fit(grad_up.data(), cum.data(), T);
            const double grid_step = cum.back() / (double)(T - 1);
            std::vector<double> up(T);
            up[0] = temp_gradient.back();
            up[T - 1] = 1.0;
            for (int i = 1; i + 1 < T; ++i) up[i] = s.solve_first((double)i * grid_step);
            for (int i = 0; i < T; ++i)
                if (std::isnan(up[i])) return false;
            std::reverse(up.begin(), up.end());
            new_gradient = up;
        }
        chain_hot[swap_indices[0]] = 1;
        for (int i = 1; i + 1 < T; ++i)
        {
            chain_temp[swap_indices[i]] = new_gradient[i];
            chain_hot[swap_indices[i]] = 0;
        }
        temp_gradient = new_gradient;
        num_swaps = 0;
        std::fill(swap_barriers.begin(), swap_barriers.end(), 0.0);
        return true;
    }

    void push_traces(std::vector<double> &llik, std::vector<double> &prior, std::vector<double> &post)
    {
        const int c = swap_indices[0];
        llik.push_back(chain_llik[c]);
        prior.push_back(chain_prior[c]);
        post.push_back(chain_llik[c] * chain_temp[c] + chain_prior[c]);  // Chain::get_posterior
    }

    void record_draw(int chain, int step)
    {
        const size_t D = draw_step.size(), LA = (size_t)L * AP, NL = (size_t)N * L;
        draw_step.push_back(step);
        p_store.resize((D + 1) * LA);
        m_store.resize((D + 1) * N);
        eps_neg_store.resize((D + 1) * N);
        eps_pos_store.resize((D + 1) * N);
        r_store.resize((D + 1) * N);
        data_llik_store.resize((D + 1) * N);
        mean_coi_store.resize(D + 1);
        uint64_t *lat = nullptr;
        if (prm.record_latent_genotypes)
        {
            const size_t NLW = NL * (size_t)moire_engine_mask_words();  // a genotype is W mask words
            latent_store.resize((D + 1) * NLW);
            lat = latent_store.data() + D * NLW;
        }
        if (!engine) return;
        ck(moire_engine_record_draw(engine, chain - first, p_store.data() + D * LA,
                                    m_store.data() + D * N, eps_neg_store.data() + D * N,
                                    eps_pos_store.data() + D * N, r_store.data() + D * N,
                                    data_llik_store.data() + D * N, mean_coi_store.data() + D, lat),
           "record_draw");
    }

    // ================= the ladder on the device =================================================
    template <class T>
    T *dev_alloc(size_t n)
    {
        void *p = nullptr;
        cu(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)), "cudaMalloc");
        cu(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)), "cudaMemset");
        dev_allocs.push_back(p);
        return static_cast<T *>(p);
    }

    // Called at the end of moire_mcmc_init: from here on the steps of this driver run without the
    // host (see the header comment).  Needs an engine and an exchange that lives on the device.
    void setup_device_ladder()
    {
        if (!engine || (world > 1 && !nccl_comm) || std::getenv("MOIRE_B200_HOST_LADDER")) return;
        if (!stream)
        {
            void *st = nullptr;
            ck(moire_engine_stream(engine, &st), "engine_stream");
            stream = (cudaStream_t)st;
        }
        graph_ok = std::getenv("MOIRE_B200_NO_GRAPH") == nullptr;
        ld.T = T; ld.world = world; ld.first = first; ld.count = count; ld.cmax = cmax;
        ld.per = 2 * cmax + 2;
        ld.pre_adapt_steps = prm.pre_adapt_steps; ld.thin = prm.thin; ld.burnin = prm.burnin;
        ld.trace_cap = kTraceCap;
        ld.seed = seed;
        void *dl = nullptr, *dp = nullptr;
        ck(moire_engine_device_scalars(engine, &dl, &dp), "device_scalars");
        ld.eng_llik = (const double *)dl;
        ld.eng_prior = (const double *)dp;
        double *dt = nullptr;
        int32_t *di = nullptr;
        ck(moire_engine_device_temps(engine, &dt), "device_temps");
        ck(moire_engine_device_iteration(engine, &di), "device_iteration");
        ld.eng_temp = dt;
        ld.eng_iter = di;
        if (world > 1)
        {
            ld.send = send_d;  // moire_mcmc_init_nccl sized them for `per`
            ld.recv = recv_d;
        }
        else
        {
            ld.send = dev_alloc<double>(ld.per);
            ld.recv = ld.send;
        }
        cu(cudaHostAlloc((void **)&host_flags, 2 * sizeof(double), cudaHostAllocMapped), "cudaHostAlloc");
        host_flags[0] = host_flags[1] = 0.0;
        double *hf_dev = nullptr;
        cu(cudaHostGetDevicePointer((void **)&hf_dev, host_flags, 0), "cudaHostGetDevicePointer");
        ld.host_flags = hf_dev;
        ld.chain_llik = dev_alloc<double>(T);
        ld.chain_prior = dev_alloc<double>(T);
        ld.chain_temp = dev_alloc<double>(T);
        ld.chain_hot = dev_alloc<int>(T);
        ld.swap_indices = dev_alloc<int>(T);
        ld.swap_acc = dev_alloc<long long>(T);
        ld.swap_barriers = dev_alloc<double>(T);
        ld.ctl = dev_alloc<long long>(moire::LC_WORDS);
        ld.minutes = dev_alloc<double>(1);
        ld.trace = dev_alloc<double>((size_t)3 * kTraceCap);
        ld.swap_store = dev_alloc<int>(kTraceCap);
        ld.draw_sel = dev_alloc<int>(2);
        // the draw store: as many draws as the run can produce, within ~1 GiB (then it is read
        // back when full)
        const size_t LA = (size_t)L * AP, NL = (size_t)N * L;
        const size_t per_draw = LA * 8 + (size_t)N * (4 + 4 * 8) + 8 + (prm.record_latent_genotypes ? NL * 8 * (size_t)moire_engine_mask_words() : 0);
        const size_t thin = (size_t)std::max(prm.thin, 1);
        const size_t want = ((size_t)std::max(prm.samples, 0) + thin - 1) / thin + 1;
        const size_t cap = std::max<size_t>(1, std::min(want, ((size_t)1 << 30) / std::max<size_t>(per_draw, 1)));
        ld.draw_cap = (int)cap;
        ld.draw_step = dev_alloc<int>(cap);
        store.capacity = (int32_t)cap;
        store.p = dev_alloc<double>(cap * LA);
        store.m = dev_alloc<int32_t>(cap * N);
        store.eps_neg = dev_alloc<double>(cap * N);
        store.eps_pos = dev_alloc<double>(cap * N);
        store.r = dev_alloc<double>(cap * N);
        store.sample_llik = dev_alloc<double>(cap * N);
        store.mean_coi = dev_alloc<double>(cap);
        store.latent = prm.record_latent_genotypes ? dev_alloc<uint64_t>(cap * NL * (size_t)moire_engine_mask_words()) : nullptr;
        dev_mode = true;
        host_dirty = true;
        batch_open = false;
    }

    template <class T>
    void h2d(T *dst, const T *src, size_t n)
    {
        cu(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, stream), "h2d");
    }
    template <class T>
    void d2h(T *dst, const T *src, size_t n)
    {
        cu(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, stream), "d2h");
    }

    // host ladder -> device, and the control words of the batch that starts at (step, phase)
    void begin_batch(int step, int phase)
    {
        long long ctl[moire::LC_WORDS] = {(long long)num_swaps, even_swap ? 1 : 0, (long long)swap_passes,
                                          step, phase, 0, 0, stop_any ? 1 : 0};
        h2d(ld.ctl, ctl, moire::LC_WORDS);
        if (host_dirty)
        {
            h2d(ld.chain_llik, chain_llik.data(), T);
            h2d(ld.chain_prior, chain_prior.data(), T);
            h2d(ld.chain_temp, chain_temp.data(), T);
            h2d(ld.chain_hot, chain_hot.data(), T);
            h2d(ld.swap_indices, swap_indices.data(), T);
            if (T > 1)
            {
                h2d(ld.swap_acc, reinterpret_cast<const long long *>(swap_acceptances.data()), T - 1);
                h2d(ld.swap_barriers, swap_barriers.data(), T - 1);
            }
            host_dirty = false;
        }
        cu(moire::launch_ladder_begin(ld, stream), "k_ladder_begin");
        // the copies above read pageable host memory: done with it before the caller moves on
        cu(cudaStreamSynchronize(stream), "sync");
        batch_open = true;
        next_step = step;
        next_phase = phase;
        pending = 0;
        pending_phase = phase;
        draws_bound = 0;
    }

    // device ladder -> host: the state, the traces of the steps enqueued since the last read-back
    // and the draws stored since then.  The one place the host waits for the device.
    void flush()
    {
        if (!dev_mode || !batch_open) return;
        long long ctl[moire::LC_WORDS];
        std::vector<double> tr((size_t)3 * std::max(pending, 1));
        std::vector<int> ss(std::max(pending, 1));
        double minutes = 0.0;
        d2h(ctl, ld.ctl, moire::LC_WORDS);
        d2h(&minutes, ld.minutes, 1);
        d2h(chain_llik.data(), ld.chain_llik, T);
        d2h(chain_prior.data(), ld.chain_prior, T);
        d2h(chain_temp.data(), ld.chain_temp, T);
        d2h(chain_hot.data(), ld.chain_hot, T);
        d2h(swap_indices.data(), ld.swap_indices, T);
        if (T > 1)
        {
            d2h(reinterpret_cast<long long *>(swap_acceptances.data()), ld.swap_acc, T - 1);
            d2h(swap_barriers.data(), ld.swap_barriers, T - 1);
        }
        if (pending > 0)
        {
            d2h(tr.data(), ld.trace, (size_t)3 * pending);
            d2h(ss.data(), ld.swap_store, pending);
        }
        cu(cudaStreamSynchronize(stream), "sync");
        batch_open = false;
        if (pending > 0)
        {
            if (ctl[moire::LC_TRACE_COUNT] != pending)
                throw DriverError{MOIRE_ESTATE, "device ladder: step count out of step with the host"};
            num_swaps = ctl[moire::LC_NUM_SWAPS];
            even_swap = ctl[moire::LC_EVEN_SWAP] != 0;
            swap_passes = (uint32_t)ctl[moire::LC_SWAP_PASSES];
            stop_any = stop_any || ctl[moire::LC_STOP] != 0;
            elapsed_minutes_rank0 = minutes;
            auto &vl = pending_phase == 0 ? llik_burnin : llik_sample;
            auto &vp = pending_phase == 0 ? prior_burnin : prior_sample;
            auto &vq = pending_phase == 0 ? posterior_burnin : posterior_sample;
            for (int k = 0; k < pending; ++k)
            {
                vl.push_back(tr[3 * k]);
                vp.push_back(tr[3 * k + 1]);
                vq.push_back(tr[3 * k + 2]);
                if (!(pending_phase == 0 && T == 1)) swap_store.push_back(ss[k]);  // src/mcmc.cpp:179
            }
        }
        const int nd = (int)ctl[moire::LC_NDRAWS];
        if (nd > 0)
        {
            const size_t D = draw_step.size(), LA = (size_t)L * AP, NL = (size_t)N * L;
            draw_step.resize(D + nd);
            p_store.resize((D + nd) * LA);
            m_store.resize((D + nd) * N);
            eps_neg_store.resize((D + nd) * N);
            eps_pos_store.resize((D + nd) * N);
            r_store.resize((D + nd) * N);
            data_llik_store.resize((D + nd) * N);
            mean_coi_store.resize(D + nd);
            d2h(draw_step.data() + D, ld.draw_step, nd);
            d2h(p_store.data() + D * LA, store.p, nd * LA);
            d2h(m_store.data() + D * N, store.m, (size_t)nd * N);
            d2h(eps_neg_store.data() + D * N, store.eps_neg, (size_t)nd * N);
            d2h(eps_pos_store.data() + D * N, store.eps_pos, (size_t)nd * N);
            d2h(r_store.data() + D * N, store.r, (size_t)nd * N);
            d2h(data_llik_store.data() + D * N, store.sample_llik, (size_t)nd * N);
            d2h(mean_coi_store.data() + D, store.mean_coi, nd);
            if (store.latent)
            {
                const size_t NLW = NL * (size_t)moire_engine_mask_words();
                latent_store.resize((D + nd) * NLW);
                d2h(latent_store.data() + D * NLW, store.latent, nd * NLW);
            }
            cu(cudaStreamSynchronize(stream), "sync");
        }
        pending = 0;
        draws_bound = 0;
    }

    // the launches of one step on the stream: what a replayed graph holds
    void step_body(int phase)
    {
        ck(moire_engine_step(engine, MOIRE_ITERATION_FROM_DEVICE), "step");
        if (phase == 1) ck(moire_engine_enqueue_record_draw(engine, ld.draw_sel, &store), "record_draw");
        cu(moire::launch_ladder_pack(ld, stream), "k_ladder_pack");
        if (world > 1)
            nc(g_nccl.AllGather(ld.send, recv_d, (size_t)ld.per, NcclApi::kFloat64, nccl_comm, stream), "ncclAllGather");
        cu(moire::launch_ladder_step(ld, stream), "k_ladder_step");
    }

    void enqueue_step(int phase)
    {
        host_flags[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() / 60.0;
        host_flags[1] = local_stop ? 1.0 : 0.0;
        int32_t timing_on = 0;
        ck(moire_engine_get_timing(engine, &timing_on), "get_timing");
        if (graph_ok && !timing_on && !step_graph[phase])
        {
            // capture the step once; from then on it is one cudaGraphLaunch
            uint64_t l0 = 0, l1 = 0;
            ck(moire_engine_counters(engine, nullptr, &l0), "counters");
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok)
            {
                try
                {
                    step_body(phase);
                }
                catch (const DriverError &)
                {
                    ok = false;
                }
                ok = cudaStreamEndCapture(stream, &graph) == cudaSuccess && ok && graph;
            }
            if (ok) ok = cudaGraphInstantiate(&step_graph[phase], graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (!ok)
            {
                // a library in the step that cannot be captured: plain launches from here on
                cudaGetLastError();
                step_graph[phase] = nullptr;
                graph_ok = false;
            }
            else
            {
                ck(moire_engine_counters(engine, nullptr, &l1), "counters");
                int64_t captured = 0;
                ck(moire_engine_captured_step_launches(engine, &captured), "captured_step_launches");
                graph_launches[phase] = captured + (int64_t)(l1 - l0) + 2;
                ck(moire_engine_add_launches(engine, -(int64_t)(l1 - l0)), "add_launches");
            }
        }
        if (graph_ok && !timing_on && step_graph[phase])
        {
            cu(cudaGraphLaunch(step_graph[phase], stream), "cudaGraphLaunch");
            ck(moire_engine_add_launches(engine, graph_launches[phase]), "add_launches");
        }
        else
        {
            step_body(phase);
            ck(moire_engine_add_launches(engine, 2), "add_launches");
        }
        ++pending;
        ++next_step;
    }

    void device_step(int step, int phase)
    {
        const bool thinned = phase == 1 && (prm.thin == 0 || step % prm.thin == 0);
        if (batch_open && (next_step != step || next_phase != phase || pending >= kTraceCap ||
                           (thinned && draws_bound >= ld.draw_cap)))
            flush();
        if (!batch_open) begin_batch(step, phase);
        enqueue_step(phase);
        if (thinned) ++draws_bound;
        if (phase == 0 && T > 1)
        {
            // MCMC::adapt_temp is due on a schedule the host can follow without reading anything:
            // num_swaps counts the burn-in steps past pre_adapt_steps since the last adaptation
            if (step > prm.pre_adapt_steps) ++num_swaps;
            if (prm.temp_adapt_steps > 0 && num_swaps % prm.temp_adapt_steps == 0 &&
                step > prm.pre_adapt_steps && prm.adapt_temp)
            {
                flush();  // num_swaps, the barriers and the index map as the device left them
                if (adapt_ladder()) host_dirty = true;
            }
        }
    }

    // ================= the ladder on the host ===================================================
    void burnin_step(int step)
    {
        if (dev_mode)
        {
            device_step(step, 0);
            return;
        }
        engine_step(step);
        exchange_scalars();
        push_traces(llik_burnin, prior_burnin, posterior_burnin);
        if (T > 1)
        {
            bool moved = swap_pass(step, true);
            if (prm.temp_adapt_steps > 0 && num_swaps % prm.temp_adapt_steps == 0 &&
                step > prm.pre_adapt_steps && prm.adapt_temp)
                moved |= adapt_ladder();
            if (moved) push_local_temps();
        }
    }

    void sample_step(int step)
    {
        if (dev_mode)
        {
            device_step(step, 1);
            return;
        }
        engine_step(prm.burnin + step);
        if (prm.thin == 0 || step % prm.thin == 0)
        {
            for (int c = first; c < first + count; ++c)
                if (chain_hot[c] || T == 1) record_draw(c, step);
        }
        exchange_scalars();
        push_traces(llik_sample, prior_sample, posterior_sample);
        // the reference runs swap_chains unconditionally here (src/mcmc.cpp:355); with one chain
        // its loop body never executes but swap_store still grows
        if (swap_pass(step, false)) push_local_temps();
    }

    void finalize()
    {
        flush();
        const double half = (double)(num_swaps / 2);  // integer division, as the reference
        for (auto &b : swap_barriers) b /= half;
    }
};

namespace
{
template <class F>
int guarded(moire_mcmc *h, F &&f)
{
    try
    {
        if (!h) throw DriverError{MOIRE_EINVAL, "null driver"};
        f();
        return MOIRE_OK;
    }
    catch (const DriverError &e)
    {
        g_mcmc_error = e.msg;
        if (h) h->last_error = e.msg;
        return e.code;
    }
    catch (const std::bad_alloc &)
    {
        g_mcmc_error = "host allocation failed";
        if (h) h->last_error = g_mcmc_error;
        return MOIRE_ENOMEM;
    }
    catch (...)
    {
        g_mcmc_error = "unknown error";
        if (h) h->last_error = g_mcmc_error;
        return MOIRE_EINVAL;
    }
}

template <class V, class P>
void copy_out(const V &v, P *dst)
{
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
}
}  // namespace

extern "C"
{
void moire_mcmc_partition(int32_t T, int32_t rank, int32_t world_size, int32_t *first,
                          int32_t *count)
{
    const int64_t lo = (int64_t)rank * T / world_size, hi = (int64_t)(rank + 1) * T / world_size;
    if (first) *first = (int32_t)lo;
    if (count) *count = (int32_t)(hi - lo);
}

const char *moire_mcmc_last_error(const moire_mcmc *h)
{
    return h ? h->last_error.c_str() : g_mcmc_error.c_str();
}

int moire_mcmc_create(const moire_data *data, const moire_params *model,
                      const moire_mcmc_params *mcmc, const double *pt_chains, int32_t T,
                      int32_t rank, int32_t world_size, int32_t device_id, uint64_t seed,
                      moire_mcmc **out)
{
    if (!out || !data || !model || !mcmc || !pt_chains)
    {
        g_mcmc_error = "null argument";
        return MOIRE_EINVAL;
    }
    *out = nullptr;
    if (T < 1 || world_size < 1 || rank < 0 || rank >= world_size || world_size > T)
    {
        g_mcmc_error = "need 1 <= world_size <= num_chains and 0 <= rank < world_size";
        return MOIRE_EINVAL;
    }
    moire_mcmc *h = new moire_mcmc();
    h->prm = *mcmc;
    h->T = T;
    h->rank = rank;
    h->world = world_size;
    h->seed = seed;
    h->N = data->num_samples;
    h->L = data->num_loci;
    h->allow_rel = model->allow_relatedness != 0;
    moire_mcmc_partition(T, rank, world_size, &h->first, &h->count);
    for (int r = 0; r < world_size; ++r)
    {
        int32_t f, c;
        moire_mcmc_partition(T, r, world_size, &f, &c);
        h->cmax = std::max(h->cmax, (int)c);
    }
    if (device_id >= 0)
    {
        moire_params mp = *model;
        mp.burnin = mcmc->burnin;
        const int rc = moire_engine_create(data, &mp, h->count, h->first, pt_chains + h->first,
                                           device_id, seed, &h->engine);
        if (rc != MOIRE_OK)
        {
            g_mcmc_error = moire_last_error();
            delete h;
            return rc;
        }
        h->AP = moire_engine_a_pad(h->engine);
    }
    else
    {
        h->AP = 4;  // ladder-only mode: nothing is stored per allele
    }
    h->swap_indices.resize(T);
    for (int i = 0; i < T; ++i) h->swap_indices[i] = i;
    h->swap_acceptances.assign(T > 0 ? T - 1 : 0, 0);
    h->swap_barriers.assign(T > 0 ? T - 1 : 0, 0.0);
    h->temp_gradient.assign(pt_chains, pt_chains + T);
    h->chain_temp.assign(pt_chains, pt_chains + T);
    h->chain_hot.assign(T, 0);
    h->chain_llik.assign(T, 0.0);
    h->chain_prior.assign(T, 0.0);
    h->send.assign(2 * h->cmax + 2, 0.0);
    h->recv.assign((size_t)world_size * (2 * h->cmax + 2), 0.0);
    h->t_start = std::chrono::steady_clock::now();
    *out = h;
    return MOIRE_OK;
}

int moire_mcmc_destroy(moire_mcmc *h)
{
    delete h;
    return MOIRE_OK;
}

int moire_mcmc_set_allgather(moire_mcmc *h, moire_allgather_fn fn, void *ctx)
{
    return guarded(h, [&] {
        h->gather = fn;
        h->gather_ctx = ctx;
    });
}

int moire_mcmc_nccl_unique_id(char id[128], const char *library_path)
{
    std::string err;
    if (!id || !g_nccl.load(library_path, err))
    {
        g_mcmc_error = id ? err : "null id";
        return MOIRE_EINVAL;
    }
    NcclApi::UniqueId u;
    if (g_nccl.GetUniqueId(&u) != 0)
    {
        g_mcmc_error = "ncclGetUniqueId failed";
        return MOIRE_ECUDA;
    }
    std::memcpy(id, u.internal, 128);
    return MOIRE_OK;
}

int moire_mcmc_init_nccl(moire_mcmc *h, const char id[128], const char *library_path)
{
    return guarded(h, [&] {
        if (!h->engine) throw DriverError{MOIRE_ESTATE, "the NCCL exchange needs an engine (device_id >= 0)"};
        if (h->world < 2) return;
        std::string err;
        if (!g_nccl.load(library_path, err)) throw DriverError{MOIRE_EINVAL, err};
        void *st = nullptr;
        h->ck(moire_engine_stream(h->engine, &st), "engine_stream");  // also binds the device
        h->stream = (cudaStream_t)st;
        const int per = 2 * h->cmax + 2;
        h->cu(cudaMalloc(&h->send_d, sizeof(double) * per), "cudaMalloc");
        h->cu(cudaMalloc(&h->recv_d, sizeof(double) * per * h->world), "cudaMalloc");
        h->cu(cudaMemset(h->send_d, 0, sizeof(double) * per), "cudaMemset");
        h->cu(cudaMallocHost(&h->send_pin, sizeof(double) * per), "cudaMallocHost");
        h->cu(cudaMallocHost(&h->recv_pin, sizeof(double) * per * h->world), "cudaMallocHost");
        if (!id)
        {
            if (!g_nccl_comm || g_nccl_world != h->world || g_nccl_rank != h->rank)
                throw DriverError{MOIRE_ESTATE, "no NCCL communicator of this shape to reuse"};
        }
        else
        {
            NcclApi::UniqueId u;
            std::memcpy(u.internal, id, 128);
            void *comm = nullptr;
            h->nc(g_nccl.CommInitRank(&comm, h->world, u, h->rank), "ncclCommInitRank");
            // the communicator this replaces goes away unless a live driver still holds it
            if (g_nccl_comm && g_nccl_users == 0 && g_nccl.CommDestroy) g_nccl.CommDestroy(g_nccl_comm);
            g_nccl_comm = comm;
            g_nccl_world = h->world;
            g_nccl_rank = h->rank;
            g_nccl_users = 0;
        }
        h->nccl_comm = g_nccl_comm;
        ++g_nccl_users;
    });
}

moire_engine *moire_mcmc_engine(moire_mcmc *h) { return h ? h->engine : nullptr; }

int moire_mcmc_init(moire_mcmc *h, int32_t *failed)
{
    return guarded(h, [&] {
        std::vector<int32_t> still(h->count, 0);
        h->init_ill_count.assign(h->count, 0);
        h->init_prior_nonfinite.assign((size_t)5 * h->count, 0);
        if (h->engine)
            h->ck(moire_engine_init_chains(h->engine, h->prm.max_initialization_tries,
                                           h->init_ill_count.data(), still.data(), nullptr, nullptr,
                                           h->init_prior_nonfinite.data()),
                  "init_chains");
        double bad = 0.0;
        for (int v : still) bad += v != 0;
        // share the failure flag through the same exchange the steps use
        std::fill(h->send.begin(), h->send.end(), 0.0);
        h->send[0] = bad;
        double any = bad;
        if (h->world > 1)
        {
            const int per = 2 * h->cmax + 2;
            h->host_allgather(h->send.data(), per, h->recv.data());
            any = 0.0;
            for (int r = 0; r < h->world; ++r) any += h->recv[(size_t)r * per];
        }
        if (failed) *failed = any > 0.0;
        h->initialised = any == 0.0;
        h->t_start = std::chrono::steady_clock::now();
        if (h->initialised) h->exchange_scalars();
        h->stop_any = false;
        if (h->initialised && !h->dev_mode) h->setup_device_ladder();
    });
}

int moire_mcmc_get_init_counts(moire_mcmc *h, int32_t *ill_count, int32_t *prior_nonfinite)
{
    return guarded(h, [&] {
        copy_out(h->init_ill_count, ill_count);
        copy_out(h->init_prior_nonfinite, prior_nonfinite);
    });
}

int moire_mcmc_burnin(moire_mcmc *h, int32_t step)
{
    return guarded(h, [&] {
        if (!h->initialised) throw DriverError{MOIRE_ESTATE, "burnin before a successful moire_mcmc_init"};
        h->burnin_step(step);
    });
}

int moire_mcmc_sample(moire_mcmc *h, int32_t step)
{
    return guarded(h, [&] {
        if (!h->initialised) throw DriverError{MOIRE_ESTATE, "sample before a successful moire_mcmc_init"};
        h->sample_step(step);
    });
}

int moire_mcmc_finalize(moire_mcmc *h)
{
    return guarded(h, [&] { h->finalize(); });
}

int moire_mcmc_run(moire_mcmc *h, moire_progress_fn progress, void *ctx,
                   int32_t *max_runtime_reached, double *runtime_minutes)
{
    return guarded(h, [&] {
        if (!h->initialised) throw DriverError{MOIRE_ESTATE, "run before a successful moire_mcmc_init"};
        // negative or non-finite: no limit; 0 stops at once, as `timer.elapsed() < 0` would
        const double limit = (h->prm.max_runtime >= 0 && std::isfinite(h->prm.max_runtime))
                                 ? h->prm.max_runtime
                                 : std::numeric_limits<double>::infinity();
        h->t_start = std::chrono::steady_clock::now();
        h->elapsed_minutes_rank0 = 0.0;
        h->local_stop = h->stop_any = false;
        // the reference compares whole minutes (Timer<..., std::chrono::minutes>, src/main.cpp:36);
        // rank 0's clock decides for every rank
        auto minutes = [&] { return std::floor(h->elapsed_minutes_rank0); };
        // With the ladder on the device the loop runs in batches: up to kBatch steps are queued,
        // then one read-back brings their traces, rank 0's clock and every rank's stop word, the
        // callback is told about each of those steps (with its own posterior and hot chain), and
        // the stop / max_runtime decisions are taken -- on every rank from the same gathered words,
        // so all ranks leave after the same step.  On the host ladder a batch is one step.
        const int kBatch = h->dev_mode ? 32 : 1;
        auto phase_loop = [&](int phase, int total) {
            int step = 0;
            while (!h->stop_any && step < total && minutes() < limit)
            {
                const int b0 = step;
                const size_t s0 = h->swap_store.size();
                do
                {
                    if (phase == 0) h->burnin_step(step);
                    else h->sample_step(step);
                    ++step;
                } while (step < total && step - b0 < kBatch);
                h->flush();
                const auto &post = phase == 0 ? h->posterior_burnin : h->posterior_sample;
                for (int s = b0; s < step && progress; ++s)
                {
                    const size_t si = s0 + (size_t)(s - b0);
                    const int hot = si < h->swap_store.size() ? h->swap_store[si] : h->swap_indices[0];
                    if (progress(ctx, phase, s + 1, post[post.size() - (size_t)(step - s)], hot) != 0)
                    {
                        h->local_stop = true;
                        if (h->world == 1) h->stop_any = true;  // else: agreed through the next exchange
                        break;
                    }
                }
            }
            return step;
        };
        phase_loop(0, h->prm.burnin);
        const int sampled = phase_loop(1, h->prm.samples);
        if (max_runtime_reached) *max_runtime_reached = minutes() >= limit && sampled < h->prm.samples;
        h->finalize();
        if (runtime_minutes) *runtime_minutes = minutes();
        if (h->stop_any) throw DriverError{MOIRE_ESTATE, "interrupted by the progress callback"};
    });
}

int moire_mcmc_get_mode(moire_mcmc *h, int32_t *device_ladder, int32_t *step_graph)
{
    return guarded(h, [&] {
        if (device_ladder) *device_ladder = h->dev_mode ? 1 : 0;
        if (step_graph) *step_graph = h->dev_mode && h->graph_ok && (h->step_graph[0] || h->step_graph[1]) ? 1 : 0;
    });
}

int moire_mcmc_get_sizes(moire_mcmc *h, moire_mcmc_sizes *out)
{
    return guarded(h, [&] {
        h->flush();
        if (!out) throw DriverError{MOIRE_EINVAL, "null output"};
        out->num_chains = h->T;
        out->first_chain = h->first;
        out->local_chains = h->count;
        out->burnin_steps = (int32_t)h->llik_burnin.size();
        out->sample_steps = (int32_t)h->llik_sample.size();
        out->num_draws = (int32_t)h->draw_step.size();
        out->num_swap_store = (int32_t)h->swap_store.size();
        out->a_pad = h->AP;
    });
}

int moire_mcmc_get_traces(moire_mcmc *h, double *llik_burnin, double *prior_burnin,
                          double *posterior_burnin, double *llik_sample, double *prior_sample,
                          double *posterior_sample)
{
    return guarded(h, [&] {
        h->flush();
        copy_out(h->llik_burnin, llik_burnin);
        copy_out(h->prior_burnin, prior_burnin);
        copy_out(h->posterior_burnin, posterior_burnin);
        copy_out(h->llik_sample, llik_sample);
        copy_out(h->prior_sample, prior_sample);
        copy_out(h->posterior_sample, posterior_sample);
    });
}

int moire_mcmc_get_ladder(moire_mcmc *h, int32_t *swap_store, int64_t *swap_acceptances,
                          double *swap_barriers, double *temp_gradient, int32_t *swap_indices,
                          double *chain_temps)
{
    return guarded(h, [&] {
        h->flush();
        copy_out(h->swap_store, swap_store);
        copy_out(h->swap_acceptances, swap_acceptances);
        copy_out(h->swap_barriers, swap_barriers);
        copy_out(h->temp_gradient, temp_gradient);
        copy_out(h->swap_indices, swap_indices);
        copy_out(h->chain_temp, chain_temps);
    });
}

int moire_mcmc_get_hot(moire_mcmc *h, int32_t *hot)
{
    return guarded(h, [&] {
        h->flush();
        if (!hot) throw DriverError{MOIRE_EINVAL, "null output"};
        for (int c = 0; c < h->T; ++c) hot[c] = h->chain_hot[c];
    });
}

int moire_mcmc_get_draws(moire_mcmc *h, int32_t *draw_step, double *allele_freqs, int32_t *coi,
                         double *eps_neg, double *eps_pos, double *relatedness,
                         double *data_llik, double *lam_coi, uint64_t *latent)
{
    return guarded(h, [&] {
        h->flush();
        copy_out(h->draw_step, draw_step);
        copy_out(h->p_store, allele_freqs);
        copy_out(h->m_store, coi);
        copy_out(h->eps_neg_store, eps_neg);
        copy_out(h->eps_pos_store, eps_pos);
        copy_out(h->r_store, relatedness);
        copy_out(h->data_llik_store, data_llik);
        copy_out(h->mean_coi_store, lam_coi);
        copy_out(h->latent_store, latent);
    });
}

int moire_mcmc_swap_pass(moire_mcmc *h, const double *chain_llik, int32_t step, int32_t is_burnin)
{
    return guarded(h, [&] {
        h->flush();
        h->host_dirty = true;
        if (!chain_llik) throw DriverError{MOIRE_EINVAL, "null llik"};
        h->chain_llik.assign(chain_llik, chain_llik + h->T);
        if (h->swap_pass(step, is_burnin != 0)) h->push_local_temps();
    });
}

int moire_mcmc_set_swap_uniforms(moire_mcmc *h, const double *u, int32_t n)
{
    return guarded(h, [&] {
        h->forced_u.clear();
        if (u && n > 0) h->forced_u.assign(u, u + n);
    });
}

int moire_mcmc_set_ladder(moire_mcmc *h, const double *temp_gradient, const double *swap_barriers,
                          const int32_t *swap_indices, const double *chain_temps,
                          int64_t num_swaps, int32_t even_swap)
{
    return guarded(h, [&] {
        h->flush();
        h->host_dirty = true;
        const int T = h->T;
        if (temp_gradient) h->temp_gradient.assign(temp_gradient, temp_gradient + T);
        if (swap_barriers) h->swap_barriers.assign(swap_barriers, swap_barriers + (T - 1));
        if (swap_indices) h->swap_indices.assign(swap_indices, swap_indices + T);
        if (chain_temps)
        {
            h->chain_temp.assign(chain_temps, chain_temps + T);
            h->push_local_temps();
        }
        if (num_swaps >= 0) h->num_swaps = num_swaps;
        if (even_swap >= 0) h->even_swap = even_swap != 0;
    });
}

int moire_mcmc_set_local_scalars(moire_mcmc *h, const double *llik, const double *prior)
{
    return guarded(h, [&] {
        if (h->engine) throw DriverError{MOIRE_ESTATE, "only in ladder-only mode (device_id < 0)"};
        if (llik) h->injected_llik.assign(llik, llik + h->count);
        if (prior) h->injected_prior.assign(prior, prior + h->count);
    });
}

int moire_mcmc_adapt_temp(moire_mcmc *h)
{
    return guarded(h, [&] {
        h->flush();
        h->host_dirty = true;
        if (h->adapt_ladder()) h->push_local_temps();
    });
}

double moire_monotone_spline_solve(const double *x, const double *y, int32_t n, double y_target)
{
    if (!x || !y || n < 3) return std::numeric_limits<double>::quiet_NaN();
    MonotoneSpline s;
    s.fit(x, y, n);
    return s.solve_first(y_target);
}
}  // extern "C"
