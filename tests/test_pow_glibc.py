"""csrc/grrt_pow.cuh restates glibc's x86-64 FMA pow() operation by operation so that the RKF45
step-size controller (reference src/rendering/runge_kutta.rs:155-157, f64::powf -> libm pow) produces
the same bits on the device as on a linux-gnu host.  The header is plain C++ when compiled without nvcc:
compile it with g++ and compare against libm on random inputs."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include "%(hdr)s"
#include <cstdio>
#include <cstdlib>
int main() {
    srand48(7);
    long bad = 0, n = 3000000, counted = 0;
    grrt::PowTables T = grrt::pow_tables_global();
    for (long i = 0; i < n; ++i) {
        double x = std::exp((drand48() * 60 - 30) * 2.302585) * (1 + drand48());
        double y = 0.2;
        if ((i & 3) == 1) y = drand48() * 4 - 2;
        if ((i & 3) == 2) x = 1.0 + (drand48() - 0.5) * 1e-3;
        if ((i & 7) == 7) x = std::exp(drand48() * 1400 - 700);
        double ref = std::pow(x, y);
        if (!(ref > 1e-290 && ref < 1e290)) continue;
        ++counted;
        if (grrt::pow_glibc(T, x, y) != ref) ++bad;
    }
    std::printf("%%ld %%ld\n", bad, counted);
    return 0;
}
'''


def _fma_cpu():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(not _fma_cpu(), reason="glibc selects __pow_fma only on CPUs with FMA")
def test_pow_matches_libm_bit_for_bit():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "h.cpp"), os.path.join(d, "h")
        open(src, "w").write(HARNESS % {"hdr": os.path.join(ROOT, "gr_raytracer_b200", "csrc", "grrt_pow.cuh")})
        subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"])
        bad, counted = map(int, subprocess.check_output([exe]).split())
    assert counted > 2000000
    assert bad == 0, f"{bad} of {counted} results differ from libm pow()"
