// Host-side parallel-tempering driver (include/moire_mcmc.h).
//
// What the reference keeps on its master thread stays here, in double: the even/odd temperature
// swaps (MCMC::swap_chains, src/mcmc.cpp:190-240), the adaptive ladder (MCMC::adapt_temp,
// src/mcmc.cpp:244-304, with the monotone cubic spline of src/include/spline/src/spline.h), the
// traces and the stores of the temp = 1 chain (src/mcmc.cpp:175-177, 327-355) and the two loops
// of run_mcmc (src/main.cpp:52-93).  Everything per-chain goes through the engine's C ABI; no
// likelihood arithmetic runs here.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>

#include "../../include/moire_mcmc.h"
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
    std::vector<char> chain_hot;            // [T] by chain id (Chain::hot)
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

    ~moire_mcmc()
    {
        if (send_d) cudaFree(send_d);
        if (recv_d) cudaFree(recv_d);
        if (send_pin) cudaFreeHost(send_pin);
        if (recv_pin) cudaFreeHost(recv_pin);
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
        const int per = 2 * cmax + 1;
        if (world > 1 && nccl_comm && engine)
        {
            // device to device: the sums k_reduce left on the engine's stream, the collective on
            // the same stream, one read-back of all T pairs
            void *dl = nullptr, *dp = nullptr;
            ck(moire_engine_device_scalars(engine, &dl, &dp), "device_scalars");
            send_pin[0] = minutes;
            cu(cudaMemcpyAsync(send_d, dl, sizeof(double) * count, cudaMemcpyDeviceToDevice, stream), "d2d");
            cu(cudaMemcpyAsync(send_d + cmax, dp, sizeof(double) * count, cudaMemcpyDeviceToDevice, stream), "d2d");
            cu(cudaMemcpyAsync(send_d + 2 * cmax, send_pin, sizeof(double), cudaMemcpyHostToDevice, stream), "h2d");
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
        moire::Stream s;
        s.open(seed, swap_passes, moire::KIND_SWAP, 0xFFFFu, (uint32_t)pair, (uint32_t)(step & 0xFFFF));
        return s.next_uniform();
    }

    // MCMC::swap_chains (src/mcmc.cpp:190-240)
    bool swap_pass(int step, bool burnin)
    {
        bool moved = false;
        for (int i = even_swap ? 1 : 0; i < T - 1; i += 2)
        {
            const int a = swap_indices[i], b = swap_indices[i + 1];
            const double V_a = -chain_llik[a], temp_a = chain_temp[a];
            const double V_b = -chain_llik[b], temp_b = chain_temp[b];
            const double ratio = (temp_b - temp_a) * (V_b - V_a);
            const double rate = std::min(1.0, std::exp(ratio));
            if (burnin && step > prm.pre_adapt_steps) swap_barriers[i] += 1.0 - rate;
            const double u = std::log(uniform_for_pair(step, i));
            if ((ratio > 0 || u < ratio) && !std::isnan(ratio))
            {
                std::swap(swap_indices[i], swap_indices[i + 1]);
                chain_temp[a] = temp_b;
                chain_temp[b] = temp_a;
                if (i == 0)
                {
                    chain_hot[b] = 1;
                    chain_hot[a] = 0;
                }
                if (!burnin) swap_acceptances[i]++;
                moved = true;
            }
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
            latent_store.resize((D + 1) * NL);
            lat = latent_store.data() + D * NL;
        }
        if (!engine) return;
        ck(moire_engine_record_draw(engine, chain - first, p_store.data() + D * LA,
                                    m_store.data() + D * N, eps_neg_store.data() + D * N,
                                    eps_pos_store.data() + D * N, r_store.data() + D * N,
                                    data_llik_store.data() + D * N, mean_coi_store.data() + D, lat),
           "record_draw");
    }

    void burnin_step(int step)
    {
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
    h->send.assign(2 * h->cmax + 1, 0.0);
    h->recv.assign((size_t)world_size * (2 * h->cmax + 1), 0.0);
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
        const int per = 2 * h->cmax + 1;
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
            g_nccl_comm = comm;  // an earlier communicator stays alive: it may be in use
            g_nccl_world = h->world;
            g_nccl_rank = h->rank;
        }
        h->nccl_comm = g_nccl_comm;
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
            const int per = 2 * h->cmax + 1;
            h->host_allgather(h->send.data(), per, h->recv.data());
            any = 0.0;
            for (int r = 0; r < h->world; ++r) any += h->recv[(size_t)r * per];
        }
        if (failed) *failed = any > 0.0;
        h->initialised = any == 0.0;
        h->t_start = std::chrono::steady_clock::now();
        if (h->initialised) h->exchange_scalars();
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
        const double limit = (h->prm.max_runtime > 0 && std::isfinite(h->prm.max_runtime))
                                 ? h->prm.max_runtime
                                 : std::numeric_limits<double>::infinity();
        h->t_start = std::chrono::steady_clock::now();
        h->elapsed_minutes_rank0 = 0.0;
        // the reference compares whole minutes (Timer<..., std::chrono::minutes>, src/main.cpp:36);
        // rank 0's clock decides for every rank
        auto minutes = [&] { return std::floor(h->elapsed_minutes_rank0); };
        bool stop = false;
        int step = 0;
        while (!stop && step < h->prm.burnin && minutes() < limit)
        {
            h->burnin_step(step);
            ++step;
            if (progress)
                stop = progress(ctx, 0, step, h->posterior_burnin.back(), h->swap_indices[0]) != 0;
        }
        step = 0;
        while (!stop && step < h->prm.samples && minutes() < limit)
        {
            h->sample_step(step);
            ++step;
            if (progress)
                stop = progress(ctx, 1, step, h->posterior_sample.back(), h->swap_indices[0]) != 0;
        }
        if (max_runtime_reached) *max_runtime_reached = minutes() >= limit && step < h->prm.samples;
        h->finalize();
        if (runtime_minutes) *runtime_minutes = minutes();
        if (stop) throw DriverError{MOIRE_ESTATE, "interrupted by the progress callback"};
    });
}

int moire_mcmc_get_sizes(moire_mcmc *h, moire_mcmc_sizes *out)
{
    return guarded(h, [&] {
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
        if (!hot) throw DriverError{MOIRE_EINVAL, "null output"};
        for (int c = 0; c < h->T; ++c) hot[c] = h->chain_hot[c];
    });
}

int moire_mcmc_get_draws(moire_mcmc *h, int32_t *draw_step, double *allele_freqs, int32_t *coi,
                         double *eps_neg, double *eps_pos, double *relatedness,
                         double *data_llik, double *lam_coi, uint64_t *latent)
{
    return guarded(h, [&] {
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
