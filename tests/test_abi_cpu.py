"""CPU tests of the drop-in boundary: the shared library loads, exports every entry point
include/cusmc_b200.h declares, agrees with the Python binding table, and fails loudly (never
falls back to a CPU path) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cusmc_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cusmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import cusmc_b200._lib as L
    lib = L.load()
    names = declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libcusmc_b200.so does not export %s" % n
    assert sorted(L.PROTOTYPES) == names          # binding table and header agree, both ways
    assert lib.cusmc_version() == 100


def test_library_is_sm100a_only():
    so = os.path.join(ROOT, "cusmc_b200", "libcusmc_b200.so")
    out = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_headers_compile_as_plain_c(tmp_path):
    """The boundary is a C ABI: the public headers must be consumable by a C compiler."""
    src = tmp_path / "abi.c"
    src.write_text('#include "cusmc_b200.h"\n#include "cusmc_philox.h"\n'
                   "int main(void){ cusmc_filter_config c; (void)c; double s, k; cusmc_det_sincospi(0.25, &s, &k);"
                   " return cusmc_fixed_shift(1024) == 51 && cusmc_det_exp(0.0) == 1.0 ? 0 : 1; }\n")
    exe = tmp_path / "abi"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-lm"])
    assert subprocess.call([str(exe)]) == 0


def test_public_header_math_matches_oracle(tmp_path, orc):
    """include/cusmc_detmath.h / cusmc_philox.h compiled for the HOST give the oracle's bits: the
    reproducibility contract of the fixed-point resamplers."""
    src = tmp_path / "m.c"
    src.write_text(r'''
#include <stdio.h>
#include "cusmc_philox.h"
int main(void) {
    for (int i = 0; i < 2000; ++i) {
        double x = -745.0 + 0.7451 * i;
        printf("%a %a\n", cusmc_det_exp(x), cusmc_det_log(1e-300 * (i + 1) * (i + 1) * 1e297));
    }
    for (int i = 0; i < 64; ++i) {
        double z0, z1;
        cusmc_normal_pair(cusmc_rng(99, CUSMC_STREAM_NORMAL, 3, (uint64_t)i, 1), &z0, &z1);
        cusmc_u32x4 r = cusmc_rng(99, CUSMC_STREAM_METROPOLIS, 3, (uint64_t)i, 2);
        printf("%a %a %a %llu\n", z0, z1, cusmc_u01(r.v[0], r.v[1]), (unsigned long long)cusmc_uint_below(r.v[2], r.v[3], 1000));
    }
    return 0;
}''')
    exe = tmp_path / "m"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-lm"])
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    for i in range(2000):
        e, l = lines[i].split()
        x = -745.0 + 0.7451 * i
        assert float.fromhex(e) == orc.det_exp([x])[0]
        assert float.fromhex(l) == orc.det_log([1e-300 * (i + 1) * (i + 1) * 1e297])[0]
    zz = orc.lib
    for i in range(64):
        z0, z1, u, j = lines[2000 + i].split()
        buf = (C.c_double * 2)()
        zz.orc_rng_normal_pair(99, 1, 3, i, 1, buf)
        assert float.fromhex(z0) == buf[0] and float.fromhex(z1) == buf[1]
        uu, jj = C.c_double(), C.c_uint32()
        zz.orc_rng_metropolis(99, 3, i, 2, 1000, C.byref(uu), C.byref(jj))
        assert float.fromhex(u) == uu.value and int(j) == jj.value


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product must refuse to work -- not route through a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import cusmc_b200
    h = C.c_void_p()
    assert cusmc_b200.load().cusmc_ctx_create(C.byref(h), 0) == 2        # CUSMC_ERR_CUDA
    with pytest.raises(cusmc_b200.CusmcError):
        cusmc_b200.Context(0)
    with pytest.raises(cusmc_b200.CusmcError):
        cusmc_b200.MVNPDF([0.0, 0.0], [0.0, 0.0], np.eye(2))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under cusmc_b200/ may reference it."""
    pkg = os.path.join(ROOT, "cusmc_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "libcusmc_oracle" not in txt and "cusmc_oracle.h" not in txt, f


def test_argument_validation_without_device():
    """Unknown registry keys are immediate errors in the host mirror (the reference dies with
    std::bad_function_call)."""
    from cusmc_b200 import api
    with pytest.raises(ValueError):
        api._kind("normal")
    with pytest.raises(ValueError):
        api._resampler("gibbs")
    assert api._kind("mvt") == 1 and api._resampler("systematic") == 1
    assert np.array_equal(api._colmajor(np.array([[1.0, 2.0], [3.0, 4.0]])), [1.0, 3.0, 2.0, 4.0])


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    import json
    line = json.loads(out)
    assert line["impl"] == "reference" and line["metric"] == "mvn_logpdf_evals_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_struct_layouts_match_the_binding(tmp_path):
    """cusmc_filter_config / cusmc_filter_draws cross the ABI by pointer: the ctypes mirrors must have the
    C compiler's size and field offsets (a drifted field would silently shift every later argument)."""
    import cusmc_b200._lib as L
    structs = {"cusmc_filter_config": L.FilterConfig, "cusmc_filter_draws": L.FilterDraws}
    lines = []
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cusmc_b200.h"\nint main(void){\n'
                   + "\n".join(lines) + "\nreturn 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, f)]) == getattr(cls, f).offset, "%s.%s" % (cname, f)
    # and the header has no field the binding lacks
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for cname, cls in structs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names += [re.sub(r"[\s\*]", "", n).split(" ")[-1] for n in re.sub(r"^(const\s+)?\w+\s+", "", decl).split(",")]
        assert names == [f for f, _ in cls._fields_], (cname, names)
