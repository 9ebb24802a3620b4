// Device-resident state of a block of tempered chains, as a POD view handed to kernels by value.
//
// Replaces the public members of the reference's Chain (src/chain.h:84-142) for ALL chains of one
// GPU at once.  Per-cell arrays are [T][N][L] (sample-major rows, like
// genotyping_llik_new[sample * L + locus], src/chain.cpp:1116), per-sample arrays [T][N], allele
// frequencies [T][L][AP].  Only the accepted ("old") state is stored: proposals live in
// registers / shared memory of the kernel that makes the accept decision, so the reference's
// new/old double buffering and its save_* / restore_* copies (src/chain.cpp:1225-1286) vanish.
// The cell log-likelihood is kept as its two addends (transmission, observation) because two of
// the eight updates change only one of them.
#pragma once
#include <cstdint>

#include "mask.cuh"

namespace moire
{
// Programmatic dependent launch (sm_90+): the kernels of a step form one chain on one stream, each a
// few tens of microseconds at most, so the launch gap between two of them is a visible share of a
// step on small shards (a ladder cut over 8 GPUs, Namibia-sized data).  Launched with the
// programmatic-stream-serialization attribute (engine.cu: launch_chain) a kernel may become resident
// while its predecessor is still running; pdl_enter() lets ITS successor do the same and then waits
// until every predecessor has completed and flushed its memory -- nothing written by an earlier
// kernel is read, and nothing an earlier kernel reads is written, before it.  Both instructions are
// no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_enter()
{
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// kernels take the iteration by value, or read it from View::iter when handed this
constexpr int kIterFromDevice = -1;

struct View
{
    int N, L, T, AP, nA;
    int allow_rel, burnin, max_coi, chain_offset;
    unsigned long long seed;

    // data (read-only)
    const mask_t *obs;              // [N][L] observed alleles (mask.cuh)
    const unsigned char *missing;   // [N][L]
    const int *nall;                // [L]
    const int *aclass;              // [L] index into avals
    const int *avals;               // [nA] distinct allele counts
    const int *obs_coi;             // [N]
    const double *logC;             // [kLogCDim][kLogCDim] log C(n, k)
    const double *lgam;             // [128] lgamma(i + 1)

    // per cell [T][N][L]
    mask_t *latent;
    double *lgadj;
    double *cellT;
    double *cellO;

    // per sample [T][N]
    int *m;
    double *r, *log_r, *log_1mr, *odds;  // odds = r / (1 - r), what the linear-domain mixture uses
    double *eps_pos, *eps_neg;
    double *sd_eps_neg, *sd_eps_pos, *sd_r, *sd_mr;
    int *acc_eps_neg, *acc_eps_pos, *acc_m, *acc_r, *acc_mr, *acc_samples;
    double *pr_eps_neg, *pr_eps_pos, *pr_coi, *pr_r;

    // per locus [T][L][AP]
    double *p, *p_sd;
    int *p_acc, *p_att;

    // per chain [T]
    double *mean_coi, *mean_coi_sd, *mean_coi_acc, *hyper;
    double *temp, *llik, *prior;
    int *active;  // init retry mask

    unsigned long long *evals;  // [16] non-missing cell evaluations, by update kind
    const int *iter;            // [1] the step's iteration for kernels launched with iteration < 0
                                // (the captured step graph: its launch parameters never change)
    unsigned long long *trace;  // tuning aid (MOIRE_B200_TRACE): per-warp clocks of update_p, or null

    // parameters (src/parameters.h:31-47) and host-precomputed constants
    double mean_coi_shape, mean_coi_scale, lg_shape, l_scale;
    double max_eps_pos, eps_pos_a, eps_pos_b, eps_pos_lbeta;
    double max_eps_neg, eps_neg_a, eps_neg_b, eps_neg_lbeta;
    double r_a, r_b, r_lbeta;
};
}  // namespace moire
