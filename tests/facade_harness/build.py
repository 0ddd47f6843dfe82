"""Builds the facade test driver (tests/facade_harness/facade_driver.cpp) — TEST INFRASTRUCTURE.

The driver includes the reference's UNMODIFIED module headers from /root/reference/include, so it can only
be compiled where the reference is mounted; its outputs go to oracle/_ref/ (git-ignored, not
gpurun-ignored: they travel to the GPU box like libsealref.so):

  oracle/_ref/libmoai_b200_mock.so      mock_cabi.c, the CPU test double of the C ABI (over the C oracle)
  oracle/_ref/libfacade_driver_mock.so  driver + facade bound to the test double   (`-m "not gpu"` tests)
  oracle/_ref/libfacade_driver.so       driver + facade bound to libmoai_b200.so   (`-m gpu` tests)
  oracle/_ref/libfacade_driver_fused.so the same driver with include/facade_fused first on the include path: the
                                        module functions resolve to the fused device pipelines (`-m gpu` tests)
  oracle/_ref/libfacade_driver_fused_mock.so  the fused variant on the test double (ct-pt matmuls only on the CPU)
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")
PKG = os.path.join(ROOT, "moai-fhe-transformerinference-public_b200")
MOCK_SO = os.path.join(OUT, "libmoai_b200_mock.so")
DRIVER_MOCK_SO = os.path.join(OUT, "libfacade_driver_mock.so")
DRIVER_SO = os.path.join(OUT, "libfacade_driver.so")
DRIVER_FUSED_SO = os.path.join(OUT, "libfacade_driver_fused.so")
DRIVER_FUSED_MOCK_SO = os.path.join(OUT, "libfacade_driver_fused_mock.so")


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps if os.path.exists(d))


def build(force=False):
    """Returns True when the driver libraries exist afterwards."""
    if not os.path.isdir(REF):
        return os.path.exists(DRIVER_SO)
    os.makedirs(OUT, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    hdrs = [os.path.join(inc, f) for f in ("moai_b200.h", "moai_b200_modules.h", "moai_b200_seal.hpp")] + \
           [os.path.join(inc, "moai_b200_seal_prng.hpp")] + \
           [os.path.join(inc, "facade", f) for f in ("Bootstrapper.h", "ckks_evaluator.h", "seal/seal.h")]
    mock_src = os.path.join(HERE, "mock_cabi.c")
    oracle_so = os.path.join(ROOT, "oracle", "libckks_oracle.so")
    if force or _stale(MOCK_SO, [mock_src, oracle_so] + hdrs[:2]):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-I" + inc, mock_src, "-o", MOCK_SO,
                               "-L" + os.path.dirname(oracle_so), "-lckks_oracle", "-Wl,-rpath,$ORIGIN/.."])
    drv_src = os.path.join(HERE, "facade_driver.cpp")
    common = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-w"]
    plain_inc = ["-I" + os.path.join(inc, "facade"), "-I" + inc, "-I" + os.path.join(REF, "include"), drv_src]
    fused_inc = ["-I" + os.path.join(inc, "facade_fused")] + plain_inc
    fused_hdrs = [os.path.join(inc, "moai_b200_fused_modules.hpp")]
    mock_link = ["-L" + OUT, "-lmoai_b200_mock", "-Wl,-rpath,$ORIGIN"]
    lib = os.path.join(PKG, "libmoai_b200.so")
    real_link = ["-L" + PKG, "-lmoai_b200", "-Wl,-rpath,$ORIGIN/../../" + os.path.basename(PKG)]
    jobs = []   # the driver includes the reference's whole include.hpp: ~1 min per variant, so build them side by side
    if force or _stale(DRIVER_MOCK_SO, [drv_src, MOCK_SO] + hdrs):
        jobs.append(common + plain_inc + ["-o", DRIVER_MOCK_SO] + mock_link)
    if force or _stale(DRIVER_FUSED_MOCK_SO, [drv_src, MOCK_SO] + hdrs + fused_hdrs):
        jobs.append(common + fused_inc + ["-o", DRIVER_FUSED_MOCK_SO] + mock_link)
    if os.path.exists(lib):
        if force or _stale(DRIVER_SO, [drv_src, lib] + hdrs):
            jobs.append(common + plain_inc + ["-o", DRIVER_SO] + real_link)
        if force or _stale(DRIVER_FUSED_SO, [drv_src, lib] + hdrs + fused_hdrs):
            jobs.append(common + fused_inc + ["-o", DRIVER_FUSED_SO] + real_link)
    procs = [subprocess.Popen(j) for j in jobs]
    failed = [j for j, p_ in zip(jobs, procs) if p_.wait() != 0]
    if failed:
        raise subprocess.CalledProcessError(1, failed[0])
    return os.path.exists(DRIVER_SO)


if __name__ == "__main__":
    print(build(force=True))
