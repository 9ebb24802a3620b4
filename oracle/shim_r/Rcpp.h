// TEST INFRASTRUCTURE ONLY (oracle/): a small FUNCTIONAL stand-in for <Rcpp.h>, used to compile and
// run the R binding of this repository (bindings/r/main_b200.cpp) together with the reference's own
// argument parsers (src/parameters.cpp, src/genotyping_data.cpp, src/mcmc_utils.cpp) in a container
// that has no R.  Unlike oracle/shim/Rcpp.h (stubs that let the reference's hot path compile), the
// objects here hold data: an Rcpp::List can be built, indexed by name, converted with Rcpp::as and
// inspected, which is what tests/test_binding.py does with the 23-field result.
// Only the surface those translation units use is provided.  SEXP nodes live in an arena that the
// harness clears; nothing here is thread-safe or fast.
#pragma once
#ifndef MOIRE_ORACLE_SHIM_R_RCPP_H
#define MOIRE_ORACLE_SHIM_R_RCPP_H
#define MOIRE_ORACLE_SHIM_RCPP_H  // keeps oracle/shim/Rcpp.h out if both directories are on the path

#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19

struct ShimNode
{
    int type = 0;  // 0 = NULL
    std::vector<int> i;          // LGLSXP / INTSXP
    std::vector<double> d;       // REALSXP
    std::vector<std::string> s;  // STRSXP
    std::vector<ShimNode *> l;   // VECSXP
    std::vector<std::string> names;
    int nrow = 0, ncol = 0;      // matrices (column-major, like R)
};
typedef ShimNode *SEXP;

#ifndef NA_INTEGER
#define NA_INTEGER (std::numeric_limits<int>::min())
#endif
#define R_NilValue (static_cast<SEXP>(nullptr))

namespace Rcpp
{
inline std::vector<std::unique_ptr<ShimNode>> &shim_arena()
{
    static std::vector<std::unique_ptr<ShimNode>> a;
    return a;
}
inline SEXP shim_new(int type)
{
    shim_arena().emplace_back(new ShimNode());
    shim_arena().back()->type = type;
    return shim_arena().back().get();
}
inline void shim_clear() { shim_arena().clear(); }

[[noreturn]] inline void stop(const std::string &msg) { throw std::runtime_error(msg); }
inline void checkUserInterrupt() {}
static std::ostream &Rcout = std::cout;
inline void message(SEXP) {}

inline std::size_t shim_length(SEXP x)
{
    if (!x) return 0;
    switch (x->type)
    {
        case LGLSXP: case INTSXP: return x->i.size();
        case REALSXP: return x->d.size();
        case STRSXP: return x->s.size();
        case VECSXP: return x->l.size();
        default: return 0;
    }
}
inline double shim_num(SEXP x, std::size_t k)
{
    if (!x || k >= shim_length(x)) stop("Rcpp shim: index out of range / NULL");
    if (x->type == REALSXP) return x->d[k];
    if (x->type == INTSXP || x->type == LGLSXP) return (double)x->i[k];
    stop("Rcpp shim: not a numeric vector");
}

// ---- wrap -------------------------------------------------------------------------------------
inline SEXP wrap(SEXP x) { return x; }
inline SEXP wrap(bool x) { SEXP n = shim_new(LGLSXP); n->i.push_back(x); return n; }
inline SEXP wrap(int x) { SEXP n = shim_new(INTSXP); n->i.push_back(x); return n; }
inline SEXP wrap(long x) { SEXP n = shim_new(REALSXP); n->d.push_back((double)x); return n; }
inline SEXP wrap(double x) { SEXP n = shim_new(REALSXP); n->d.push_back(x); return n; }
inline SEXP wrap(float x) { return wrap((double)x); }
inline SEXP wrap(const std::string &x) { SEXP n = shim_new(STRSXP); n->s.push_back(x); return n; }
inline SEXP wrap(const char *x) { return wrap(std::string(x)); }
template <class T>
SEXP wrap(const std::vector<T> &v)
{
    if constexpr (std::is_same<T, bool>::value)
    {
        SEXP n = shim_new(LGLSXP);
        for (bool b : v) n->i.push_back(b);
        return n;
    }
    else if constexpr (std::is_integral<T>::value && sizeof(T) <= 4)
    {
        SEXP n = shim_new(INTSXP);
        n->i.assign(v.begin(), v.end());
        return n;
    }
    else if constexpr (std::is_arithmetic<T>::value)
    {
        SEXP n = shim_new(REALSXP);
        n->d.assign(v.begin(), v.end());
        return n;
    }
    else if constexpr (std::is_same<T, std::string>::value)
    {
        SEXP n = shim_new(STRSXP);
        n->s = v;
        return n;
    }
    else
    {
        SEXP n = shim_new(VECSXP);
        for (const auto &e : v) n->l.push_back(wrap(e));
        return n;
    }
}

// ---- as ---------------------------------------------------------------------------------------
template <class T>
struct shim_as
{
    static T get(SEXP x)
    {
        if constexpr (std::is_same<T, std::string>::value)
        {
            if (!x || x->type != STRSXP || x->s.empty()) stop("Rcpp shim: not a string");
            return x->s[0];
        }
        else
            return static_cast<T>(shim_num(x, 0));
    }
};
template <class E>
struct shim_as<std::vector<E>>
{
    static std::vector<E> get(SEXP x)
    {
        std::vector<E> out;
        if constexpr (std::is_same<E, std::string>::value)
        {
            if (x) out = x->s;
        }
        else
            for (std::size_t k = 0; k < shim_length(x); ++k) out.push_back(static_cast<E>(shim_num(x, k)));
        return out;
    }
};
template <class T>
T as(SEXP x)
{
    return shim_as<T>::get(x);
}

// ---- vectors ----------------------------------------------------------------------------------
struct ShimNamed
{
    std::string name;
    struct Value
    {
        std::string name;
        SEXP value;
    };
    template <class T>
    Value operator=(const T &x) const
    {
        return Value{name, wrap(x)};
    }
};
inline ShimNamed Named(const std::string &n) { return ShimNamed{n}; }

template <int TYPE>
struct ShimVector
{
    SEXP node;
    ShimVector() : node(shim_new(TYPE)) {}
    ShimVector(SEXP x) : node(x ? x : shim_new(TYPE)) {}
    ShimVector(const char *s) : node(shim_new(TYPE)) { push_back(s); }
    ShimVector(const std::string &s) : node(shim_new(TYPE)) { push_back(s); }
    ShimVector(std::initializer_list<const char *> init) : node(shim_new(TYPE))
    {
        for (const char *s : init) push_back(s);
    }
    operator SEXP() const { return node; }
    std::size_t size() const { return shim_length(node); }
    void push_back(SEXP x)
    {
        if (TYPE != VECSXP) stop("Rcpp shim: push_back(SEXP) on an atomic vector");
        node->l.push_back(x);
    }
    template <class T>
    void push_back(const T &x)
    {
        if constexpr (TYPE == VECSXP) node->l.push_back(wrap(x));
        else if constexpr (TYPE == STRSXP) node->s.push_back(std::string(x));
        else if constexpr (TYPE == REALSXP) node->d.push_back((double)x);
        else node->i.push_back((int)x);
    }
    SEXP operator[](std::size_t k) const
    {
        if (TYPE != VECSXP) stop("Rcpp shim: operator[] on an atomic vector");
        if (k >= node->l.size()) stop("Rcpp shim: list index out of range");
        return node->l[k];
    }
    SEXP operator[](const std::string &name) const
    {
        for (std::size_t k = 0; k < node->names.size() && k < node->l.size(); ++k)
            if (node->names[k] == name) return node->l[k];
        stop("Rcpp shim: no list element named '" + name + "'");
    }
    SEXP operator[](const char *name) const { return (*this)[std::string(name)]; }
    struct NamesProxy
    {
        SEXP node;
        void operator=(const ShimVector<STRSXP> &n) { node->names = n.node->s; }
    };
    NamesProxy names() { return NamesProxy{node}; }
};

struct List : ShimVector<VECSXP>
{
    using ShimVector<VECSXP>::ShimVector;
    List() = default;
    template <class... Args>
    static List create(const Args &...args)
    {
        List out;
        (void)std::initializer_list<int>{(out.node->l.push_back(args.value), out.node->names.push_back(args.name), 0)...};
        return out;
    }
};
inline SEXP wrap(const List &x) { return x.node; }

typedef ShimVector<REALSXP> NumericVector;
typedef ShimVector<INTSXP> IntegerVector;
typedef ShimVector<STRSXP> CharacterVector;
typedef ShimVector<STRSXP> StringVector;
typedef ShimVector<LGLSXP> LogicalVector;
template <class T, int TYPE>
T as(const ShimVector<TYPE> &x)
{
    return shim_as<T>::get(x.node);
}

template <int RTYPE>
struct Matrix
{
    SEXP node;
    Matrix(SEXP x) : node(x) {}
    std::size_t nrow() const { return node ? node->nrow : 0; }
    std::size_t ncol() const { return node ? node->ncol : 0; }
    double at(std::size_t r, std::size_t c) const { return shim_num(node, c * nrow() + r); }
    double operator()(std::size_t r, std::size_t c) const { return at(r, c); }
};

namespace traits
{
template <class T>
struct r_sexptype_traits
{
    enum { rtype = REALSXP };
};
template <>
struct r_sexptype_traits<bool>
{
    enum { rtype = LGLSXP };
};
template <>
struct r_sexptype_traits<int>
{
    enum { rtype = INTSXP };
};
}  // namespace traits
}  // namespace Rcpp

#include "Rmath.h"

#endif
