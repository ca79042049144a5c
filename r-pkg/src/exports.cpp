// exports.cpp -- the .Call entry points `_CuSMC_*` behind R/cusmc.R and their registration
// (what Rcpp::compileAttributes() would generate into RcppExports.cpp; written out so that the
// package builds with R CMD INSTALL alone).  Symbol names and arities are the reference's
// (ref: src/RcppExports.cpp:105-113): 2, 3, 3, 4, 14, 3.
#include <Rcpp.h>

#include <string>

Rcpp::NumericVector MVN(Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma);
SEXP MVNPDF(SEXP x, Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma);
Rcpp::NumericVector MVT(Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma, float nu);
SEXP MVTPDF(SEXP x, Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma, float nu);
Rcpp::NumericVector metropolis_hastings(Rcpp::NumericVector w, int N, int B);
Rcpp::List run(unsigned N, unsigned d, unsigned timeSteps, Rcpp::NumericMatrix Y, Rcpp::NumericVector m0,
               Rcpp::NumericMatrix C0, Rcpp::NumericMatrix F, Rcpp::NumericMatrix G, Rcpp::NumericMatrix V,
               Rcpp::NumericMatrix W, float df, std::string resampler, std::string distribution, unsigned p);

// R object -> C++ argument, as Rcpp's generated stubs convert them
template <typename T>
static T arg(SEXP s) { return Rcpp::as<T>(s); }
using NV = Rcpp::NumericVector;
using NM = Rcpp::NumericMatrix;

extern "C" {

SEXP _CuSMC_MVN(SEXP mu, SEXP sigma)
{
    BEGIN_RCPP
    return Rcpp::wrap(MVN(arg<NV>(mu), arg<NM>(sigma)));
    END_RCPP
}

SEXP _CuSMC_MVNPDF(SEXP x, SEXP mu, SEXP sigma)
{
    BEGIN_RCPP
    return MVNPDF(x, arg<NV>(mu), arg<NM>(sigma));
    END_RCPP
}

SEXP _CuSMC_MVT(SEXP mu, SEXP sigma, SEXP nu)
{
    BEGIN_RCPP
    return Rcpp::wrap(MVT(arg<NV>(mu), arg<NM>(sigma), arg<float>(nu)));
    END_RCPP
}

SEXP _CuSMC_MVTPDF(SEXP x, SEXP mu, SEXP sigma, SEXP nu)
{
    BEGIN_RCPP
    return MVTPDF(x, arg<NV>(mu), arg<NM>(sigma), arg<float>(nu));
    END_RCPP
}

SEXP _CuSMC_metropolis_hastings(SEXP w, SEXP N, SEXP B)
{
    BEGIN_RCPP
    return Rcpp::wrap(metropolis_hastings(arg<NV>(w), arg<int>(N), arg<int>(B)));
    END_RCPP
}

SEXP _CuSMC_run(SEXP N, SEXP d, SEXP timeSteps, SEXP Y, SEXP m0, SEXP C0, SEXP F, SEXP G, SEXP V, SEXP W, SEXP df,
                SEXP resampler, SEXP distribution, SEXP p)
{
    BEGIN_RCPP
    return Rcpp::wrap(run(arg<unsigned>(N), arg<unsigned>(d), arg<unsigned>(timeSteps), arg<NM>(Y), arg<NV>(m0),
                          arg<NM>(C0), arg<NM>(F), arg<NM>(G), arg<NM>(V), arg<NM>(W), arg<float>(df),
                          arg<std::string>(resampler), arg<std::string>(distribution), arg<unsigned>(p)));
    END_RCPP
}

static const R_CallMethodDef kCalls[] = {
    {"_CuSMC_MVN", (DL_FUNC)&_CuSMC_MVN, 2},
    {"_CuSMC_MVNPDF", (DL_FUNC)&_CuSMC_MVNPDF, 3},
    {"_CuSMC_MVT", (DL_FUNC)&_CuSMC_MVT, 3},
    {"_CuSMC_MVTPDF", (DL_FUNC)&_CuSMC_MVTPDF, 4},
    {"_CuSMC_run", (DL_FUNC)&_CuSMC_run, 14},
    {"_CuSMC_metropolis_hastings", (DL_FUNC)&_CuSMC_metropolis_hastings, 3},
    {NULL, NULL, 0}};

void R_init_CuSMC(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, kCalls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

}  // extern "C"
