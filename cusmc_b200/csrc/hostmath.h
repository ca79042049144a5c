// hostmath.h -- the O(d^3) set-up algebra that stays on the host (once per call / per
// run, d <= 32): Cholesky, triangular inverse, tiny products.  Column-major, ld = rows,
// like the Eigen matrices the reference hands over (src/mcmc.cpp:275-280 computes
// V.determinant(), V.inverse() and the eigen factor on the host too).
#pragma once

#include <cmath>
#include <vector>

namespace hostmath {

inline double &at(std::vector<double> &M, int r, int c, int ld) { return M[(size_t)c * ld + r]; }
inline double at(const double *M, int r, int c, int ld) { return M[(size_t)c * ld + r]; }

// Lower Cholesky factor of the symmetric matrix A (only the lower triangle is read).
// Returns 0, or k+1 when pivot k is not positive (A is not SPD).
inline int cholesky_lower(const double *A, int d, std::vector<double> &L)
{
    L.assign((size_t)d * d, 0.0);
    for (int c = 0; c < d; ++c) {
        double s = at(A, c, c, d);
        for (int k = 0; k < c; ++k) s -= at(L, c, k, d) * at(L, c, k, d);
        if (!(s > 0.0) || !std::isfinite(s)) return c + 1;
        double lcc = std::sqrt(s);
        at(L, c, c, d) = lcc;
        for (int r = c + 1; r < d; ++r) {
            double t = at(A, r, c, d);
            for (int k = 0; k < c; ++k) t -= at(L, r, k, d) * at(L, c, k, d);
            at(L, r, c, d) = t / lcc;
        }
    }
    return 0;
}

// W = L^-1 for lower-triangular L.
inline void tri_inverse_lower(const std::vector<double> &L, int d, std::vector<double> &W)
{
    W.assign((size_t)d * d, 0.0);
    for (int c = 0; c < d; ++c) {
        at(W, c, c, d) = 1.0 / L[(size_t)c * d + c];
        for (int r = c + 1; r < d; ++r) {
            double s = 0.0;
            for (int k = c; k < r; ++k) s -= L[(size_t)k * d + r] * W[(size_t)c * d + k];
            at(W, r, c, d) = s / L[(size_t)r * d + r];
        }
    }
}

inline double logdet_from_cholesky(const std::vector<double> &L, int d)
{
    double s = 0.0;
    for (int k = 0; k < d; ++k) s += std::log(L[(size_t)k * d + k]);
    return 2.0 * s;
}

inline bool is_identity(const double *F, int rows, int cols)
{
    if (rows != cols) return false;
    for (int c = 0; c < cols; ++c)
        for (int r = 0; r < rows; ++r)
            if (at(F, r, c, rows) != (r == c ? 1.0 : 0.0)) return false;
    return true;
}

// Log normalising constants.  nu is a float and (nu + d) is a float sum, as in
// the reference (src/statistics.cc.cpp:300-302).
inline double mvn_lognorm(double logdet, int d)
{
    return -(0.5 * d * std::log(2.0 * M_PI) + 0.5 * logdet);
}
inline double mvt_half_nu_plus_d(float nu, int d)
{
    float s = nu + (float)(unsigned)d;
    return 0.5 * s;
}
inline double mvt_lognorm(double logdet, int d, float nu)
{
    return -0.5 * d * std::log(M_PI * nu) - 0.5 * logdet + std::lgamma(mvt_half_nu_plus_d(nu, d)) -
           std::lgamma(0.5 * nu);
}

}  // namespace hostmath
