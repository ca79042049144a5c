"""The Rcpp veneer (r-pkg/) cannot be built here -- the image has no R -- but it
must at least parse and type-check against the current C ABI: g++ -fsyntax-only with a minimal stand-in
for <Rcpp.h> (tests/mock_rcpp)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "r-pkg", "src", "cusmc_glue.cpp")


def test_glue_type_checks_against_the_c_abi():
    for src in (GLUE, os.path.join(ROOT, "r-pkg", "src", "exports.cpp")):
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra",
                            "-I", os.path.join(ROOT, "tests", "mock_rcpp"), "-I", os.path.join(ROOT, "include"), src],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_registered_calls_match_the_stubs():
    """Every `_CuSMC_*` symbol R/cusmc.R calls is defined and registered with the same arity."""
    exp = open(os.path.join(ROOT, "r-pkg", "src", "exports.cpp")).read()
    rsrc = open(os.path.join(ROOT, "r-pkg", "R", "cusmc.R")).read()
    for m in re.finditer(r"\.Call\(`(_CuSMC_\w+)`((?:, \w+)*)\)", rsrc):
        name, n = m.group(1), m.group(2).count(",")
        assert re.search(r'\{"%s", \(DL_FUNC\)&%s, %d\}' % (name, name, n), exp), name
        assert re.search(r"SEXP %s\(" % name, exp), name


def test_exports_match_the_reference_namespace():
    """Same six names and arities as the reference's NAMESPACE / RcppExports.R (2, 3, 3, 4, 14, 3)."""
    want = {"MVN": 2, "MVNPDF": 3, "MVT": 3, "MVTPDF": 4, "run": 14, "metropolis_hastings": 3}
    src = open(GLUE).read()
    got = {}
    for m in re.finditer(r"// \[\[Rcpp::export\]\]\n[\w:<> ]+?\b(\w+)\(([^)]*)\)", src):
        got[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])
    assert got == want
    ns = open(os.path.join(ROOT, "r-pkg", "NAMESPACE")).read()
    rsrc = open(os.path.join(ROOT, "r-pkg", "R", "cusmc.R")).read()
    for name, arity in want.items():
        assert re.search(r"export\([^)]*\b%s\b" % name, ns), name
        m = re.search(r"^%s <- function\(([^)]*)\)" % name, rsrc, re.M)
        assert m and len(m.group(1).split(",")) == arity, name
        assert "`_CuSMC_%s`" % name in rsrc
