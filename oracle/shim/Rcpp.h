// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <Rcpp.h>.
//
// Purpose: let the reference's hot-path translation units
// (/root/reference/src/{chain,sampler,prob_any_missing,combination_indices_generator,
// genotyping_data,parameters,mcmc_utils,mcmc}.cpp) compile UNMODIFIED, in place, in a
// container that has no R and no Rcpp.  Nothing here marshals real R objects: every
// conversion throws if it is ever reached at run time.  The harness
// (oracle/ref_harness.cpp) fills GenotypingData's public statics and Parameters'
// public fields directly and never goes through an Rcpp::List.
#pragma once
#ifndef MOIRE_ORACLE_SHIM_RCPP_H
#define MOIRE_ORACLE_SHIM_RCPP_H

#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

struct ShimSexpRec
{
};
typedef ShimSexpRec *SEXP;

#ifndef NA_INTEGER
#define NA_INTEGER (std::numeric_limits<int>::min())
#endif
#define R_NilValue (static_cast<SEXP>(nullptr))

#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16

namespace Rcpp
{
[[noreturn]] inline void shim_unreachable(const char *what)
{
    throw std::logic_error(std::string("Rcpp shim: ") + what +
                           " is not available without R");
}

struct ShimAny
{
    ShimAny() = default;
    template <class T>
    ShimAny(const T &)
    {
    }
    operator SEXP() const { return nullptr; }
};

struct ShimNamed
{
    explicit ShimNamed(const char *) {}
    explicit ShimNamed(const std::string &) {}
    template <class T>
    ShimAny operator=(const T &) const
    {
        return ShimAny{};
    }
};
inline ShimNamed Named(const char *n) { return ShimNamed(n); }
inline ShimNamed Named(const std::string &n) { return ShimNamed(n); }

template <class Elem>
struct ShimVector
{
    std::vector<Elem> v;
    ShimVector() = default;
    ShimVector(SEXP) {}
    ShimVector(const ShimAny &) {}
    ShimVector(const char *s) { (void)s; }
    ShimVector(const std::string &s) { (void)s; }
    template <class T>
    void push_back(const T &x)
    {
        v.push_back(Elem(x));
    }
    std::size_t size() const { return v.size(); }
    ShimAny operator[](std::size_t) const { return ShimAny{}; }
    ShimAny operator[](const char *) const { return ShimAny{}; }
    ShimAny operator[](const std::string &) const { return ShimAny{}; }
    operator SEXP() const { return nullptr; }
    ShimAny names() { return ShimAny{}; }
};

struct List : ShimVector<ShimAny>
{
    using ShimVector<ShimAny>::ShimVector;
    List() = default;
    template <class... Args>
    static List create(const Args &...)
    {
        return List{};
    }
};

typedef ShimVector<double> NumericVector;
typedef ShimVector<int> IntegerVector;
typedef ShimVector<std::string> CharacterVector;
typedef ShimVector<std::string> StringVector;
typedef ShimVector<int> LogicalVector;

template <int RTYPE>
struct Matrix
{
    Matrix() = default;
    Matrix(SEXP) {}
    Matrix(const ShimAny &) {}
    std::size_t nrow() const { return 0; }
    std::size_t ncol() const { return 0; }
    int at(std::size_t, std::size_t) const { shim_unreachable("Matrix::at"); }
};

namespace traits
{
template <class T>
struct r_sexptype_traits
{
    enum
    {
        rtype = REALSXP
    };
};
template <>
struct r_sexptype_traits<bool>
{
    enum
    {
        rtype = LGLSXP
    };
};
template <>
struct r_sexptype_traits<int>
{
    enum
    {
        rtype = INTSXP
    };
};
}  // namespace traits

template <class T, class U>
T as(const U &)
{
    shim_unreachable("Rcpp::as");
}

template <class T>
SEXP wrap(const T &)
{
    return nullptr;
}

static std::ostream &Rcout = std::cout;

inline void message(SEXP) {}
inline void checkUserInterrupt() {}
}  // namespace Rcpp

#include "Rmath.h"

#endif
