// TEST INFRASTRUCTURE ONLY (oracle/): drives bindings/r/main_b200.cpp -- the replacement for the
// reference's src/main.cpp -- without R.  Builds the `args` list run_mcmc_rcpp receives from
// R/mcmc.R (functional Rcpp stand-in, oracle/shim_r/Rcpp.h), calls run_mcmc(args) exactly as
// RcppExports would, and flattens the returned list for the Python test (tests/test_binding.py).
#include <cstring>
#include <string>
#include <vector>

#include <Rcpp.h>

Rcpp::List run_mcmc(Rcpp::List args);

namespace
{
Rcpp::List g_result;
std::string g_error, g_names;

SEXP by_path(const int *path, int depth)
{
    SEXP x = g_result.node;
    for (int k = 0; k < depth; ++k)
    {
        if (!x || x->type != VECSXP || path[k] < 0 || (size_t)path[k] >= x->l.size()) return nullptr;
        x = x->l[path[k]];
    }
    return x;
}
}  // namespace

extern "C"
{
// scalars: the 27 values R/mcmc.R puts next to data / is_missing, in the order of `names` below
int binding_run(int L, int N, const int *num_alleles, const unsigned char *obs /*per locus [N][A_j]*/,
                const unsigned char *is_missing /*[L][N]*/, const double *scalars, const double *pt_chains, int T)
{
    static const char *names[] = {"verbose", "simple_verbose", "use_message", "allow_relatedness", "chain_number",
                                  "thin", "burnin", "samples", "pt_num_threads", "adapt_temp", "pre_adapt_steps",
                                  "temp_adapt_steps", "max_initialization_tries", "record_latent_genotypes",
                                  "max_runtime", "max_coi", "mean_coi_shape", "mean_coi_scale", "max_eps_pos",
                                  "eps_pos_alpha", "eps_pos_beta", "max_eps_neg", "eps_neg_alpha", "eps_neg_beta",
                                  "r_alpha", "r_beta"};
    try
    {
        Rcpp::shim_clear();
        Rcpp::List data;
        size_t off = 0;
        for (int j = 0; j < L; ++j)
        {
            Rcpp::List locus;
            for (int i = 0; i < N; ++i)
            {
                std::vector<double> v(obs + off, obs + off + num_alleles[j]);
                off += num_alleles[j];
                locus.push_back(Rcpp::wrap(v));
            }
            data.push_back(locus);
        }
        SEXP miss = Rcpp::shim_new(LGLSXP);  // matrix(loci x samples), column-major like R
        miss->nrow = L;
        miss->ncol = N;
        miss->i.resize((size_t)L * N);
        for (int j = 0; j < L; ++j)
            for (int i = 0; i < N; ++i) miss->i[(size_t)i * L + j] = is_missing[(size_t)j * N + i];
        Rcpp::List args;
        Rcpp::StringVector arg_names;
        args.push_back(data);
        arg_names.push_back("data");
        args.push_back(miss);
        arg_names.push_back("is_missing");
        for (size_t k = 0; k < sizeof(names) / sizeof(names[0]); ++k)
        {
            args.push_back(Rcpp::wrap(scalars[k]));
            arg_names.push_back(names[k]);
        }
        args.push_back(Rcpp::wrap(std::vector<double>(pt_chains, pt_chains + T)));
        arg_names.push_back("pt_chains");
        args.names() = arg_names;
        g_result = run_mcmc(args);
        g_names.clear();
        for (const auto &n : g_result.node->names) g_names += n + ",";
        return 0;
    }
    catch (const std::exception &e)
    {
        g_error = e.what();
        return 1;
    }
}

const char *binding_error() { return g_error.c_str(); }
const char *binding_result_names() { return g_names.c_str(); }

// the node at result[[path[0]]][[path[1]]]...: its type (0 NULL, 10 lgl, 13 int, 14 real, 16 str, 19 list),
// its length, and -- for atomic vectors -- its values as doubles
int binding_node(const int *path, int depth, int *type, int *length, double *values, int max_values)
{
    SEXP x = by_path(path, depth);
    if (!x)
    {
        *type = 0;
        *length = 0;
        return depth == 0 ? 0 : 1;
    }
    *type = x->type;
    *length = (int)Rcpp::shim_length(x);
    if (values && (x->type == LGLSXP || x->type == INTSXP || x->type == REALSXP))
        for (int k = 0; k < *length && k < max_values; ++k) values[k] = Rcpp::shim_num(x, k);
    return 0;
}

const char *binding_node_names(const int *path, int depth)
{
    static std::string out;
    out.clear();
    SEXP x = by_path(path, depth);
    if (x)
        for (const auto &n : x->names) out += n + ",";
    return out.c_str();
}
}
