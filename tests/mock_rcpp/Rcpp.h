// A minimal stand-in for <Rcpp.h>: just enough of the API r-pkg/src/cusmc_glue.cpp uses for
// `g++ -fsyntax-only` to parse and type-check it in an image without R (tests/test_rpkg_cpu.py).
// Nothing here is linked or run.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <string>

typedef struct SEXPREC *SEXP;
typedef std::ptrdiff_t R_xlen_t;
bool Rf_isMatrix(SEXP);

namespace R {
double unif_rand();
}

namespace Rcpp {

struct RNGScope {
    RNGScope();
    ~RNGScope();
};

template <typename... A>
[[noreturn]] void stop(const char *fmt, A... args);

struct Dimension {
    Dimension(int, int);
    Dimension(int, int, int);
};

struct AttributeProxy {
    AttributeProxy &operator=(const Dimension &);
};

class NumericVector {
public:
    NumericVector();
    explicit NumericVector(int n);
    explicit NumericVector(R_xlen_t n);
    R_xlen_t size() const;
    double *begin();
    const double *begin() const;
    double *end();
    const double *end() const;
    double &operator[](R_xlen_t i);
    const double &operator[](R_xlen_t i) const;
    AttributeProxy attr(const char *name);
    operator SEXP() const;
};

class NumericMatrix {
public:
    NumericMatrix();
    NumericMatrix(int rows, int cols);
    int nrow() const;
    int ncol() const;
    double *begin();
    const double *begin() const;
    double *end();
    const double *end() const;
    operator SEXP() const;
};

template <typename T>
T as(SEXP);
SEXP wrap(double);

struct NamedValue {
    template <typename T>
    NamedValue operator=(const T &) const;
};
NamedValue Named(const char *);

class List {
public:
    template <typename... A>
    static List create(const A &...);
    operator SEXP() const;
};

}  // namespace Rcpp

// ---- what r-pkg/src/exports.cpp needs on top -----------------------------------------------------
#define BEGIN_RCPP try {
#define END_RCPP   \
    }              \
    catch (...) {} \
    return nullptr;
#define FALSE 0
typedef void *(*DL_FUNC)();
struct DllInfo;
struct R_CallMethodDef {
    const char *name;
    DL_FUNC fun;
    int numArgs;
};
int R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *, const void *, const void *);
int R_useDynamicSymbols(DllInfo *, int);
namespace Rcpp {
SEXP wrap(const NumericVector &);
SEXP wrap(const List &);
}
