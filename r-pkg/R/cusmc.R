# The six entry points of the reference package (ref: R/RcppExports.R:17-103, NAMESPACE:3-8), same
# names, argument order and defaults.  (What Rcpp::compileAttributes() generates from
# src/cusmc_glue.cpp; written out so the package builds without that step.)

MVN <- function(mu, sigma) {
    .Call(`_CuSMC_MVN`, mu, sigma)
}

MVNPDF <- function(x, mu, sigma) {
    .Call(`_CuSMC_MVNPDF`, x, mu, sigma)
}

MVT <- function(mu, sigma, nu) {
    .Call(`_CuSMC_MVT`, mu, sigma, nu)
}

MVTPDF <- function(x, mu, sigma, nu) {
    .Call(`_CuSMC_MVTPDF`, x, mu, sigma, nu)
}

run <- function(N, d, timeSteps, Y, m0, C0, F, G, V, W, df, resampler, distribution, p = 0L) {
    .Call(`_CuSMC_run`, N, d, timeSteps, Y, m0, C0, F, G, V, W, df, resampler, distribution, p)
}

metropolis_hastings <- function(w, N, B) {
    .Call(`_CuSMC_metropolis_hastings`, w, N, B)
}
