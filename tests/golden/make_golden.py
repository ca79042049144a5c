"""Generates tests/golden/golden.json.

The reference (tkamucheka/CuSMC) cannot be built or imported in this image (it
needs R + Rcpp + RcppEigen), so the golden set has two parts:

1. "reference": the only known answers the reference records
     MVNPDF(c(0,0), c(0,0), diag(2))      = 0.1591549   CuSMC/CuSMC.tex:95-105
     MVTPDF(0_3, 0_3, diag(3), nu = 3.0)  = 0.07799708  CuSMC/CuSMC.tex:131-142
     metropolis_hastings(c(0,0), 2, 10)   = (0, 1)      man/metropolis_hastings.Rd:22-27
   typed in by hand from those files (7 significant digits as printed by R).

2. "independent": the same closed-form densities (src/statistics.cc.cpp:171-211,
   295-340) evaluated with mpmath at 60 digits on seeded inputs, plus scipy's
   multivariate_normal / multivariate_t as a second opinion.  These pin the
   oracle's arithmetic far below the 1e-10 acceptance tolerance.

Run:  python tests/golden/make_golden.py     (deterministic; commit the output)
"""
import json
import os

import mpmath as mp
import numpy as np
from scipy import stats

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))


def mp_logpdf(kind, x, mu, sigma, nu):
    d = len(x)
    S = mp.matrix(sigma.tolist())
    r = mp.matrix([mp.mpf(float(a)) - mp.mpf(float(b)) for a, b in zip(x, mu)])
    q = (r.T * mp.inverse(S) * r)[0]
    logdet = mp.log(mp.det(S))
    if kind == "mvn":
        return -mp.mpf(d) / 2 * mp.log(2 * mp.pi) - logdet / 2 - q / 2
    nu_f = mp.mpf(float(np.float32(nu)))                    # nu is a float in the reference (Q9)
    nu_n = mp.mpf(float(np.float32(nu) + np.float32(d)))    # float sum
    return (-mp.mpf(d) / 2 * mp.log(mp.pi * nu_f) - logdet / 2
            + mp.loggamma(nu_n / 2) - mp.loggamma(nu_f / 2)
            - nu_n / 2 * mp.log(1 + q / nu_f))


def make_cases():
    cases = []
    seed = 20201
    for d in (1, 2, 3, 5, 8, 16, 32):
        for kind, nu in (("mvn", 0.0), ("mvt", 5.0), ("mvt", 2.5)):
            if kind == "mvn" and nu != 0.0:
                continue
            rng = np.random.default_rng(seed)
            seed += 1
            A = rng.standard_normal((d, d))
            sigma = A @ A.T / d + np.eye(d)
            mu = rng.standard_normal(d)
            xs = rng.standard_normal((4, d)) * 1.5 + mu
            logs = [mp_logpdf(kind, x, mu, sigma, nu) for x in xs]
            if kind == "mvn":
                sp = stats.multivariate_normal(mean=mu, cov=sigma).logpdf(xs)
            else:
                sp = stats.multivariate_t(loc=mu, shape=sigma, df=float(np.float32(nu))).logpdf(xs)
            sp = np.atleast_1d(sp)
            for lg, s in zip(logs, sp):
                assert abs(float(lg) - s) <= 1e-9 * max(1.0, abs(s)), (kind, d, float(lg), s)
            cases.append(dict(kind=kind, d=d, nu=nu, mu=mu.tolist(), sigma=sigma.tolist(),
                              x=xs.tolist(),
                              logpdf=[float(v) for v in logs],
                              pdf=[float(mp.e ** v) for v in logs]))
    return cases


def main():
    golden = {
        "reference": {
            "MVNPDF": {"x": [0, 0], "mu": [0, 0], "sigma": [[1, 0], [0, 1]], "value": 0.1591549,
                       "source": "CuSMC/CuSMC.tex:95-105"},
            "MVTPDF": {"x": [0, 0, 0], "mu": [0, 0, 0],
                       "sigma": [[1, 0, 0], [0, 1, 0], [0, 0, 1]], "nu": 3.0, "value": 0.07799708,
                       "source": "CuSMC/CuSMC.tex:131-142"},
            "metropolis_hastings": {"w": [0, 0], "N": 2, "B": 10, "value": [0, 1],
                                    "source": "man/metropolis_hastings.Rd:22-27 + src/samplers.cpp:30"},
        },
        # Random123 known-answer vectors for philox4x32-10 (kat_vectors of the Random123 1.09
        # distribution), used to pin the counter-based generator.
        "philox4x32_10": [
            {"ctr": [0, 0, 0, 0], "key": [0, 0],
             "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
            {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2,
             "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
            {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
             "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
        ],
        "independent": make_cases(),
    }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote", len(golden["independent"]), "independent cases")


if __name__ == "__main__":
    main()
