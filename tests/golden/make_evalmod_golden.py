#!/usr/bin/env python
"""Writes tests/golden/evalmod_remez_K25_w10_d59_r2.json: the EvalMod polynomial the REFERENCE generates.

The reference's own multi-interval Remez (M/source/bootstrapping/common/Remez.cpp:557-586, driven by
ModularReducer.cpp:34-48 through Bootstrapper::prepare_mod_polynomial, Bootstrapper.cpp:1973-1977) is run
unmodified inside oracle/_ref (NTL::RR provided over libmpfr by oracle/refbuild/ntl_shim) with the driver's
parameters (M/test/test_full_scheme.hpp:345-350: K = 25, degree 59, 2 double angles, inverse degree 1, loge = 10).
Needs /root/reference (to build oracle/_ref); the JSON travels with the repo.  Run: python tests/golden/make_evalmod_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

oracle.build_ref()
r = oracle.SealRef(13, [51, 46, 46] + [51] * 14 + [58], hamming_weight=64, seed=5)
r.make_relin_key()
r.boot_create()
cheb, sic = r.boot_polynomial()
out = {"source": "reference Remez (common/Remez.cpp) run in oracle/_ref; cosine cos(2*pi*(x - 1/4)/4) on the union of "
                 "[i - 2^-10, i + 2^-10], |i| < 25, Chebyshev basis in x/25, multiplied by scale_inverse_coeff "
                 "(ModularReducer.cpp:42-47)",
       "boundary_K": 25, "deg": 59, "double_angles": 2, "log_width": 10,
       "scale_inverse_coeff": float(sic).hex(), "cheb_times_sic": [float(c).hex() for c in cheb]}
with open(os.path.join(HERE, "evalmod_remez_K25_w10_d59_r2.json"), "w") as f:
    json.dump(out, f, indent=1)
print("scale_inverse_coeff", sic, "cheb[0..3]", cheb[:4])
