"""oracle — CPU checkers for the B200 CKKS backend.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package.  The product package never does.

Two checkers live here:

* :class:`Oracle` — ``ckks_oracle.c``, a plain-C restatement of the reference's algorithms
  (each function cites the reference file:line it follows).  Parity status: *pinned* against
  the reference's own KATs and against :class:`SealRef` (``tests/test_oracle_pinned.py``).
* :class:`SealRef` — the reference's vendored SEAL-4.1-bs itself, compiled from the sources
  where they lie under ``/root/reference`` into ``oracle/_ref/libsealref.so`` by
  ``oracle/refbuild/Makefile`` (git-ignored, travels to the GPU box with the snapshot).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libckks_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libsealref.so")
REFERENCE_ROOT = "/root/reference"

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int)
f64p = C.POINTER(C.c_double)


def _p(a, t=u64p):
    return a.ctypes.data_as(t)


def build_oracle(force=False):
    """Compile ckks_oracle.c (gcc) -> oracle/libckks_oracle.so."""
    src = os.path.join(_HERE, "ckks_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["/usr/bin/gcc", "-O3", "-fopenmp", "-shared", "-fPIC", src, "-o", _ORACLE_SO, "-lm"])
    return _ORACLE_SO


def build_ref(force=False):
    """Compile the reference's SEAL-4.1-bs into oracle/_ref (only where /root/reference exists)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return _REF_SO if os.path.exists(_REF_SO) else None
    rb = os.path.join(_HERE, "refbuild")
    deps = [os.path.join(rb, "ref_wrap.cpp"), os.path.join(rb, "Makefile"), os.path.join(rb, "ntl_shim", "NTL", "RR.h"),
            os.path.join(rb, "ntl_shim", "NTL", "mat_RR.h")]
    if force or not os.path.exists(_REF_SO) or os.path.getmtime(_REF_SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "refbuild")])
    return _REF_SO


def have_ref():
    return os.path.exists(_REF_SO)


MOAI_BITS = [51] + [46] * 20 + [51] * 14 + [58]  # M/test/test_full_scheme.hpp:356-378


class Oracle:
    """ctypes view of ckks_oracle.c.  Arrays are numpy uint64, SEAL layout [poly][limb][n]."""

    def __init__(self, log_n, bits=None, primes=None):
        build_oracle()
        self.lib = C.CDLL(_ORACLE_SO)
        L = self.lib
        L.orc_create.restype = C.c_void_p
        for name in ("orc_mulmod", "orc_powmod", "orc_invmod", "orc_shoup_quotient", "orc_mul_lazy",
                     "orc_barrett_reduce_128", "orc_minimal_primitive_root"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = None
        L.orc_elt_from_step.restype = C.c_uint32
        for name in ('orc_mulmod','orc_powmod','orc_invmod','orc_shoup_quotient','orc_minimal_primitive_root'):
            getattr(L, name).argtypes = [C.c_uint64] * (3 if name in ('orc_mulmod','orc_powmod') else 2)
        L.orc_mul_lazy.argtypes = [C.c_uint64] * 4
        L.orc_barrett_reduce_128.argtypes = [C.c_uint64] * 3
        self.log_n = log_n
        self.n = 1 << log_n
        self.bits = list(bits) if bits is not None else None
        if primes is not None:
            L.orc_create_from_primes.restype = C.c_void_p
            parr = (C.c_uint64 * len(primes))(*primes)
            self.h = C.c_void_p(L.orc_create_from_primes(C.c_int(log_n), parr, C.c_int(len(primes))))
            bits = [int(p).bit_length() for p in primes]
            self.bits = bits
        else:
            arr = (C.c_int * len(bits))(*bits)
            self.h = C.c_void_p(L.orc_create(C.c_int(log_n), arr, C.c_int(len(bits))))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.kl = len(bits)
        q = np.zeros(self.kl, dtype=np.uint64)
        L.orc_primes(self.h, _p(q))
        self.q = q

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    # ---- tables
    def ntt_tables(self, limb):
        n = self.n
        out = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
        invn = np.zeros(2, dtype=np.uint64)
        self.lib.orc_ntt_tables(self.h, C.c_int(limb), *[_p(a) for a in out], _p(invn))
        return out + [invn]

    def fft_tables(self):
        n = self.n
        roots = np.zeros(2 * n, dtype=np.float64)
        inv = np.zeros(2 * n, dtype=np.float64)
        imap = np.zeros(n, dtype=np.uint64)
        self.lib.orc_fft_tables(self.h, _p(roots, f64p), _p(inv, f64p), _p(imap))
        return roots, inv, imap

    # ---- transforms (in place on a copy)
    def ntt(self, limb, v):
        v = np.ascontiguousarray(v, dtype=np.uint64).copy()
        self.lib.orc_ntt(self.h, C.c_int(limb), _p(v))
        return v

    def intt(self, limb, v):
        v = np.ascontiguousarray(v, dtype=np.uint64).copy()
        self.lib.orc_intt(self.h, C.c_int(limb), _p(v))
        return v

    # ---- element-wise
    def _binop(self, op, a, b, polys, limbs):
        out = np.empty_like(a)
        self.lib.orc_addsub(self.h, C.c_int(op), _p(a), _p(b if b is not None else a), C.c_int(polys), C.c_int(limbs),
                            _p(out))
        return out

    def add(self, a, b, polys, limbs):
        return self._binop(0, a, b, polys, limbs)

    def sub(self, a, b, polys, limbs):
        return self._binop(1, a, b, polys, limbs)

    def negate(self, a, polys, limbs):
        return self._binop(2, a, None, polys, limbs)

    def addsub_plain(self, op, ct, pt, polys, limbs):
        out = np.empty_like(ct)
        self.lib.orc_addsub_plain(self.h, C.c_int(op), _p(ct), _p(pt), C.c_int(polys), C.c_int(limbs), _p(out))
        return out

    def multiply_plain(self, ct, pt, polys, limbs):
        out = np.empty_like(ct)
        self.lib.orc_multiply_plain(self.h, _p(ct), _p(pt), C.c_int(polys), C.c_int(limbs), _p(out))
        return out

    def multiply(self, a, b, limbs):
        out = np.empty(3 * limbs * self.n, dtype=np.uint64)
        self.lib.orc_multiply(self.h, _p(a), _p(b), C.c_int(limbs), _p(out))
        return out

    def square(self, a, limbs):
        return self.multiply(a, a, limbs)

    def rescale(self, ct, polys, limbs):
        out = np.empty(polys * (limbs - 1) * self.n, dtype=np.uint64)
        self.lib.orc_rescale(self.h, _p(ct), C.c_int(polys), C.c_int(limbs), _p(out))
        return out

    def mod_switch(self, ct, polys, limbs):
        out = np.empty(polys * (limbs - 1) * self.n, dtype=np.uint64)
        self.lib.orc_mod_switch(self.h, _p(ct), C.c_int(polys), C.c_int(limbs), _p(out))
        return out

    # ---- Galois / key switching
    def elt_from_step(self, step):
        return int(self.lib.orc_elt_from_step(self.h, C.c_int(step)))

    def galois_table(self, elt):
        t = np.zeros(self.n, dtype=np.uint32)
        self.lib.orc_galois_table(self.h, C.c_uint32(elt), _p(t, u32p))
        return t

    def switch_key(self, ct, target, limbs, key):
        ct = ct.copy()
        self.lib.orc_switch_key(self.h, _p(ct), _p(target), C.c_int(limbs), _p(key))
        return ct

    def relinearize(self, ct3, limbs, key):
        out = np.empty(2 * limbs * self.n, dtype=np.uint64)
        self.lib.orc_relinearize(self.h, _p(ct3), C.c_int(limbs), _p(key), _p(out))
        return out

    def apply_galois(self, ct, limbs, elt, key):
        out = np.empty(2 * limbs * self.n, dtype=np.uint64)
        self.lib.orc_apply_galois(self.h, _p(ct), C.c_int(limbs), C.c_uint32(elt), _p(key), _p(out))
        return out

    def naf_steps(self, steps):
        out = (C.c_int * 64)()
        k = self.lib.orc_naf_steps(self.h, C.c_int(steps), out)
        return [out[i] for i in range(k)]

    # ---- encoder
    def encode_scalar_consts(self, value, scale, limbs):
        out = np.zeros(limbs, dtype=np.uint64)
        self.lib.orc_encode_scalar_consts(self.h, C.c_double(value), C.c_double(scale), C.c_int(limbs), _p(out))
        return out

    def encode_scalar(self, value, scale, limbs):
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self.lib.orc_encode_scalar(self.h, C.c_double(value), C.c_double(scale), C.c_int(limbs), _p(out))
        return out

    def encode(self, values, scale, limbs):
        """values: complex or real array of length <= n/2."""
        v = np.asarray(values, dtype=np.complex128)
        ri = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(-1))
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        rc = self.lib.orc_encode_vector(self.h, _p(ri, f64p), C.c_int(len(v)), C.c_double(scale), C.c_int(limbs),
                                        _p(out))
        if rc:
            raise ValueError("encode overflow")
        return out

    def decode(self, pt, limbs, scale):
        out = np.zeros(self.n, dtype=np.float64)
        self.lib.orc_decode(self.h, _p(pt), C.c_int(limbs), C.c_double(scale), _p(out, f64p))
        return out[0::2] + 1j * out[1::2]

    def modraise(self, ct, polys, limbs):
        out = np.empty(polys * limbs * self.n, dtype=np.uint64)
        self.lib.orc_modraise(self.h, _p(ct), C.c_int(polys), C.c_int(limbs), _p(out))
        return out

    # ---- client side helpers (own PRNG)
    def gen_secret(self, seed, hamming_weight=0):
        sk = np.zeros(self.kl * self.n, dtype=np.uint64)
        self.lib.orc_gen_secret(self.h, C.c_uint64(seed), C.c_int(hamming_weight), _p(sk))
        return sk

    def encrypt_sym(self, sk, seed, pt, limbs):
        ct = np.zeros(2 * limbs * self.n, dtype=np.uint64)
        self.lib.orc_encrypt_sym(self.h, _p(sk), C.c_uint64(seed), _p(pt), C.c_int(limbs), _p(ct))
        return ct

    def decrypt(self, sk, ct, polys, limbs):
        pt = np.zeros(limbs * self.n, dtype=np.uint64)
        self.lib.orc_decrypt(self.h, _p(sk), _p(ct), C.c_int(polys), C.c_int(limbs), _p(pt))
        return pt

    def gen_relin_key(self, sk, seed):
        s2 = np.zeros(self.kl * self.n, dtype=np.uint64)
        self.lib.orc_secret_squared(self.h, _p(sk), _p(s2))
        return self._gen_ksk(sk, seed, s2)

    def gen_galois_key(self, sk, seed, elt):
        sg = np.zeros(self.kl * self.n, dtype=np.uint64)
        self.lib.orc_secret_galois(self.h, _p(sk), C.c_uint32(elt), _p(sg))
        return self._gen_ksk(sk, seed, sg)

    def _gen_ksk(self, sk, seed, new_key):
        out = np.zeros((self.kl - 1) * 2 * self.kl * self.n, dtype=np.uint64)
        self.lib.orc_gen_kswitch_key(self.h, _p(sk), C.c_uint64(seed), _p(new_key), _p(out))
        return out

    # ---- module level
    def ct_pt_matmul_scalar(self, X, W, K, Cc, limbs, scale, c_begin=0, c_end=None):
        c_end = Cc if c_end is None else c_end
        W = np.ascontiguousarray(W, dtype=np.float64)
        out = np.zeros((c_end - c_begin) * 2 * (limbs - 1) * self.n, dtype=np.uint64)
        self.lib.orc_ct_pt_matmul_scalar(self.h, _p(X), _p(W, f64p), C.c_int(K), C.c_int(Cc), C.c_int(limbs),
                                         C.c_double(scale), C.c_int(c_begin), C.c_int(c_end), _p(out))
        return out

    def ct_pt_matmul_masked(self, X, W, mask, K, Cc, limbs, scale, c_begin=0, c_end=None):
        c_end = Cc if c_end is None else c_end
        W = np.ascontiguousarray(W, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        out = np.zeros((c_end - c_begin) * 2 * (limbs - 1) * self.n, dtype=np.uint64)
        self.lib.orc_ct_pt_matmul_masked(self.h, _p(X), _p(W, f64p), _p(mask, i32p), C.c_int(K), C.c_int(Cc),
                                         C.c_int(limbs), C.c_double(scale), C.c_int(c_begin), C.c_int(c_end), _p(out))
        return out


# op codes of ref_eval (oracle/refbuild/ref_wrap.cpp)
(OP_ADD, OP_SUB, OP_MULTIPLY, OP_SQUARE, OP_RELINEARIZE, OP_RESCALE, OP_MOD_SWITCH, OP_ROTATE, OP_CONJUGATE,
 OP_MULTIPLY_PLAIN, OP_ADD_PLAIN, OP_SUB_PLAIN, OP_NEGATE, OP_MULTIPLY_CONST, OP_ADD_CONST, OP_DOUBLE,
 OP_ADD_REDUCED_ERROR, OP_SUB_REDUCED_ERROR, OP_MULTIPLY_REDUCED_ERROR, OP_MULTIPLY_VECTOR_REDUCED_ERROR) = range(20)


class SealRef:
    """ctypes view of oracle/_ref/libsealref.so — the reference's real SEAL-4.1-bs."""

    def __init__(self, log_n, bits, hamming_weight=0, seed=1):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libsealref.so not built (run oracle.build_ref() where /root/reference exists)")
        self.lib = C.CDLL(_REF_SO)
        L = self.lib
        L.ref_create.restype = C.c_void_p
        L.ref_error.restype = C.c_char_p
        L.ref_galois_elt_from_step.restype = C.c_uint32
        self.log_n, self.n, self.bits = log_n, 1 << log_n, list(bits)
        arr = (C.c_int * len(bits))(*bits)
        self.h = C.c_void_p(L.ref_create(C.c_int(log_n), arr, C.c_int(len(bits)), C.c_int(hamming_weight),
                                         C.c_uint64(seed)))
        if not L.ref_ok(self.h):
            raise RuntimeError("ref_create: " + L.ref_error(self.h).decode())
        self.kl = L.ref_n_key_limbs(self.h)
        q = np.zeros(self.kl, dtype=np.uint64)
        L.ref_primes(self.h, _p(q))
        self.q = q

    def __del__(self):
        try:
            self.lib.ref_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.lib.ref_error(self.h).decode())

    def ntt_tables(self, limb):
        n = self.n
        out = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
        invn = np.zeros(2, dtype=np.uint64)
        self.lib.ref_ntt_tables(self.h, C.c_int(limb), *[_p(a) for a in out], _p(invn))
        return out + [invn]

    def ntt(self, limb, v):
        v = np.ascontiguousarray(v, dtype=np.uint64).copy()
        self.lib.ref_ntt(self.h, C.c_int(limb), _p(v), C.c_int(v.size // self.n))
        return v

    def intt(self, limb, v):
        v = np.ascontiguousarray(v, dtype=np.uint64).copy()
        self.lib.ref_intt(self.h, C.c_int(limb), _p(v), C.c_int(v.size // self.n))
        return v

    def secret_key(self):
        sk = np.zeros(self.kl * self.n, dtype=np.uint64)
        self.lib.ref_secret_key(self.h, _p(sk))
        return sk

    def make_relin_key(self):
        self._chk(self.lib.ref_make_relin_key(self.h))

    def make_galois_keys(self, steps, conjugate=False):
        arr = (C.c_int * max(1, len(steps)))(*steps)
        self._chk(self.lib.ref_make_galois_keys(self.h, arr, C.c_int(len(steps)), C.c_int(int(conjugate))))

    def elt_from_step(self, step):
        return int(self.lib.ref_galois_elt_from_step(self.h, C.c_int(step)))

    def has_galois_key(self, elt):
        return bool(self.lib.ref_has_galois_key(self.h, C.c_uint32(elt)))

    def export_relin_key(self):
        out = np.zeros((self.kl - 1) * 2 * self.kl * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_export_kswitch_key(self.h, C.c_int(0), C.c_uint32(0), _p(out)))
        return out

    def export_galois_key(self, elt):
        out = np.zeros((self.kl - 1) * 2 * self.kl * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_export_kswitch_key(self.h, C.c_int(1), C.c_uint32(elt), _p(out)))
        return out

    def encode(self, values, scale, limbs):
        v = np.asarray(values, dtype=np.complex128)
        ri = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(-1))
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_encode_complex(self.h, _p(ri, f64p), C.c_int(len(v)), C.c_int(limbs), C.c_double(scale),
                                              _p(out)))
        return out

    def encode_real(self, values, scale, limbs):
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_encode_real(self.h, _p(v, f64p), C.c_int(len(v)), C.c_int(limbs), C.c_double(scale),
                                           _p(out)))
        return out

    def encode_scalar(self, value, scale, limbs):
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_encode_scalar(self.h, C.c_double(value), C.c_int(limbs), C.c_double(scale), _p(out)))
        return out

    def decode(self, pt, limbs, scale):
        out = np.zeros(self.n, dtype=np.float64)
        self._chk(self.lib.ref_decode(self.h, _p(pt), C.c_int(limbs), C.c_double(scale), _p(out, f64p)))
        return out[0::2] + 1j * out[1::2]

    def encrypt(self, pt, limbs, scale):
        out = np.zeros(2 * limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_encrypt(self.h, _p(pt), C.c_int(limbs), C.c_double(scale), _p(out)))
        return out

    def decrypt(self, ct, size, limbs, scale):
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_decrypt(self.h, _p(ct), C.c_int(size), C.c_int(limbs), C.c_double(scale), _p(out)))
        return out

    def eval(self, op, a, size_a, limbs_a, scale_a, b=None, size_b=0, limbs_b=0, scale_b=1.0, iarg=0, darg=0.0,
             varg=None):
        """Returns (raw, size, limbs, scale)."""
        out = np.zeros(3 * limbs_a * self.n, dtype=np.uint64)
        osz, olm, osc = C.c_int(0), C.c_int(0), C.c_double(0)
        bb = _p(b) if b is not None else None
        vv = None
        if varg is not None:
            v = np.asarray(varg, dtype=np.complex128)
            ri = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(-1))
            vv = _p(ri, f64p)
        self._chk(self.lib.ref_eval(self.h, C.c_int(op), _p(a), C.c_int(size_a), C.c_int(limbs_a), C.c_double(scale_a),
                                    bb, C.c_int(size_b), C.c_int(limbs_b), C.c_double(scale_b), C.c_int(iarg),
                                    C.c_double(darg), vv, _p(out), C.byref(osz), C.byref(olm), C.byref(osc)))
        return out[: osz.value * olm.value * self.n].copy(), osz.value, olm.value, osc.value

    def set_threads(self, n=None):
        """Use n (default: every host core) OpenMP threads whatever OMP_NUM_THREADS says; returns the count."""
        n = n or os.cpu_count() or 1
        self.lib.ref_set_omp_threads(C.c_int(n))
        return self.omp_threads()

    def omp_threads(self):
        return int(self.lib.ref_omp_threads())

    def ct_pt_matmul(self, variant, X, W, mask, K, Cc, limbs, scale):
        """The reference's ct_pt_matrix_mul_* (M/source/matrix_mul/Ct_pt_matrix_mul.hpp) on raw buffers.
        Returns (out, seconds)."""
        W = np.ascontiguousarray(W, dtype=np.float64)
        out = np.zeros(Cc * 2 * (limbs - 1) * self.n, dtype=np.uint64)
        sec = C.c_double(0)
        m = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
            m = _p(mask, i32p)
        self._chk(self.lib.ref_ct_pt_matmul(self.h, C.c_int(variant), _p(X), _p(W, f64p), m, C.c_int(K), C.c_int(Cc),
                                            C.c_int(limbs), C.c_double(scale), _p(out), C.byref(sec)))
        return out, sec.value

    def galois_elts(self):
        buf = (C.c_uint32 * 4096)()
        k = self.lib.ref_galois_elts(self.h, buf, C.c_int(4096))
        return [int(buf[i]) for i in range(k)]

    def gelu_v2(self, x, count, limbs, scale):
        out = np.zeros(count * 2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.ref_gelu_v2(self.h, _p(x), C.c_int(count), C.c_int(limbs), C.c_double(scale), _p(out),
                                       C.byref(ol), C.byref(osc)))
        return out[: count * 2 * ol.value * self.n].copy(), ol.value, osc.value

    def layernorm(self, variant, x, num_ct, limbs, scale, gamma, beta, bias_vec):
        out = np.zeros(num_ct * 2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        g = np.ascontiguousarray(gamma, dtype=np.float64)
        b = np.ascontiguousarray(beta, dtype=np.float64)
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        self._chk(self.lib.ref_layernorm(self.h, C.c_int(variant), _p(x), C.c_int(num_ct), C.c_int(limbs),
                                         C.c_double(scale), _p(g, f64p), _p(b, f64p), _p(bv, i32p), _p(out),
                                         C.byref(ol), C.byref(osc)))
        return out[: num_ct * 2 * ol.value * self.n].copy(), ol.value, osc.value

    def save_ciphertext(self, ct, size, limbs, scale):
        """Ciphertext::save(compr_mode_type::none) of the real library -> bytes."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        cap = C.c_int64(ct.nbytes + 4096)
        buf = np.zeros(cap.value, dtype=np.uint8)
        self._chk(self.lib.ref_save_ciphertext(self.h, _p(ct), C.c_int(size), C.c_int(limbs), C.c_double(scale),
                                               buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(cap)))
        return buf[: cap.value].tobytes()

    def load_ciphertext(self, blob):
        """Ciphertext::load (with SEAL's validity checks) -> (residues, size, limbs, scale)."""
        b = np.frombuffer(blob, dtype=np.uint8)
        out = np.zeros(len(blob) // 4 + 8, dtype=np.uint64)      # a seeded stream expands to twice its size
        size, limbs, scale = C.c_int(0), C.c_int(0), C.c_double(0)
        self._chk(self.lib.ref_load_ciphertext(self.h, b.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(len(blob)),
                                               _p(out), C.c_int64(out.size), C.byref(size), C.byref(limbs),
                                               C.byref(scale)))
        return out[: size.value * limbs.value * self.n].copy(), size.value, limbs.value, scale.value

    # ---- randomness and key wire format (client-side pieces of the facade) ----
    def prng_bytes(self, seed8, n):
        out = np.zeros(n, dtype=np.uint8)
        s = np.ascontiguousarray(seed8, dtype=np.uint64)
        self.lib.ref_prng_bytes(_p(s), C.c_int64(n), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def sample_uniform(self, seed8):
        out = np.zeros(self.kl * self.n, dtype=np.uint64)
        s = np.ascontiguousarray(seed8, dtype=np.uint64)
        self._chk(self.lib.ref_sample_uniform(self.h, _p(s), _p(out)))
        return out

    def public_key(self):
        out = np.zeros(2 * self.kl * self.n, dtype=np.uint64)
        self.lib.ref_public_key(self.h, _p(out))
        return out

    def save_keys(self, kind, seeded, steps=(), conjugate=False):
        """kind 0 RelinKeys, 1 GaloisKeys, 2 PublicKey -> bytes of save(compr_mode_type::none); seeded: the
        Serializable<> form of freshly generated keys."""
        cap = C.c_int64(((self.kl - 1) * 2 * self.kl * self.n * 8 + 4096) * max(1, len(steps) + int(conjugate)) + 65536)
        buf = np.zeros(cap.value, dtype=np.uint8)
        arr = (C.c_int * max(1, len(steps)))(*steps)
        self._chk(self.lib.ref_save_keys(self.h, C.c_int(kind), C.c_int(int(seeded)), arr, C.c_int(len(steps)),
                                         C.c_int(int(conjugate)), buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(cap)))
        return buf[: cap.value].tobytes()

    def load_keys(self, kind, blob):
        b = np.frombuffer(blob, dtype=np.uint8)
        self._chk(self.lib.ref_load_keys(self.h, C.c_int(kind), b.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(len(blob))))

    def save_ciphertext_seeded(self, pt, limbs, scale):
        cap = C.c_int64(2 * limbs * self.n * 8 + 4096)
        buf = np.zeros(cap.value, dtype=np.uint8)
        self._chk(self.lib.ref_save_ciphertext_seeded(self.h, _p(pt), C.c_int(limbs), C.c_double(scale),
                                                      buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(cap)))
        return buf[: cap.value].tobytes()

    def batch_input(self, X, scale):
        X = np.ascontiguousarray(X, dtype=np.float64)
        num_X, num_row, num_col = X.shape
        out = np.zeros(num_col * 2 * (self.kl - 1) * self.n, dtype=np.uint64)
        self._chk(self.lib.ref_batch_input(self.h, _p(X, f64p), C.c_int(num_X), C.c_int(num_row), C.c_int(num_col),
                                           C.c_double(scale), _p(out)))
        return out

    def ct_ct_matmul(self, which, X, nX, W, nW, limbs, scale_X, scale_W, col_X, row_X, col_W, row_W, num_batch):
        out = np.zeros(max(row_X, col_W) * 2 * limbs * self.n, dtype=np.uint64)
        oc, ol, osc = C.c_int(0), C.c_int(0), C.c_double(0)
        self._chk(self.lib.ref_ct_ct_matmul(self.h, C.c_int(which), _p(X), C.c_int(nX), _p(W), C.c_int(nW),
                                            C.c_int(limbs), C.c_double(scale_X), C.c_double(scale_W), C.c_int(col_X),
                                            C.c_int(row_X), C.c_int(col_W), C.c_int(row_W), C.c_int(num_batch),
                                            _p(out), C.byref(oc), C.byref(ol), C.byref(osc)))
        return out[: oc.value * 2 * ol.value * self.n].copy(), oc.value, ol.value, osc.value

    # ---- the reference's Bootstrapper (M/source/bootstrapping/Bootstrapper.cpp), compiled unmodified ----
    def boot_create(self, logn=None, total_level=None, loge=10, final_scale=2.0 ** 46, boundary_K=25, deg=59,
                    scale_factor=2, inverse_deg=1):
        """Bootstrapper(...) + prepare_mod_polynomial() as M/test/test_full_scheme.hpp:413-433 does (the reference's own
        Remez runs here: ~15 s).  Needs make_relin_key() first.  Returns the rotation steps the driver generates keys
        for (test_full_scheme.hpp:436-443)."""
        logn = self.log_n - 1 if logn is None else logn
        total_level = self.kl - 2 if total_level is None else total_level
        self._chk(self.lib.ref_boot_create(self.h, C.c_int(loge), C.c_int(logn), C.c_int(self.log_n - 1),
                                           C.c_int(total_level), C.c_double(final_scale), C.c_int(boundary_K),
                                           C.c_int(deg), C.c_int(scale_factor), C.c_int(inverse_deg)))
        buf = (C.c_int * 4096)()
        k = self.lib.ref_boot_steps(self.h, buf, C.c_int(4096))
        return [int(buf[i]) for i in range(k)]

    def boot_prepare(self):
        """generate_LT_coefficient_3() (test_full_scheme.hpp:445-448); Galois keys must have been made."""
        self._chk(self.lib.ref_boot_prepare(self.h))

    def boot_polynomial(self):
        """(Chebyshev coefficients of the EvalMod cosine in x / K with the inverse-sine constant folded in,
        scale_inverse_coeff) as the reference generated them."""
        buf = np.zeros(256, dtype=np.float64)
        sic = C.c_double(0)
        deg = self.lib.ref_boot_polynomial(self.h, _p(buf, f64p), C.c_int(256), C.byref(sic))
        if deg < 0:
            raise RuntimeError("no bootstrapper")
        return buf[: deg + 1].copy(), sic.value

    def bootstrap_3(self, ct, scale):
        """Bootstrapper::bootstrap_3 on one ciphertext [2][1][n] -> (residues [2][limbs][n], limbs, scale)."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1)
        out = np.zeros(2 * (self.kl - 1) * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.ref_boot_bootstrap_3(self.h, _p(ct), C.c_double(scale), _p(out), C.byref(ol), C.byref(osc)))
        return out[: 2 * ol.value * self.n].copy(), ol.value, osc.value

    def boot_phase(self, phase, cts, limbs, scale):
        """One phase of bootstrap_full_3 (0 modraise, 1 coefftoslot_full_3 -> 2 cts, 2 modular_reduction,
        3 slottocoeff_full_3 <- 2 cts) -> (residues, limbs, scale)."""
        cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1)
        out = np.zeros(2 * 2 * (self.kl - 1) * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.ref_boot_phase(self.h, C.c_int(phase), _p(cts), C.c_int(limbs), C.c_double(scale), _p(out),
                                          C.byref(ol), C.byref(osc)))
        cnt = 2 if phase == 1 else 1
        return out[: cnt * 2 * ol.value * self.n].copy(), ol.value, osc.value

    def softmax_boot(self, x, num, limbs, scale, bias_vec, input_num, iters, layer_id):
        """softmax_boot (M/source/non_linear_func/softmax.hpp:308-581) -> (residues, limbs, scale)."""
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
        out = np.zeros(num * 2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        self._chk(self.lib.ref_softmax_boot(self.h, _p(x), C.c_int(num), C.c_int(limbs), C.c_double(scale), _p(bv, i32p),
                                            C.c_int(input_num), C.c_int(iters), C.c_int(layer_id), _p(out),
                                            C.byref(ol), C.byref(osc)))
        return out[: num * 2 * ol.value * self.n].copy(), ol.value, osc.value

    def single_att_block(self, X, num_col, limbs, scale, WQ, WK, WV, bQ, bK, bV, bias_vec, input_num, num_batch, iters,
                         layer_id):
        """single_att_block (M/source/att_block/single_att_block.hpp:10-207) -> (residues, count, limbs, scale)."""
        X = np.ascontiguousarray(X, dtype=np.uint64).reshape(-1)
        col_W = WQ.shape[1]
        ws = [np.ascontiguousarray(w, dtype=np.float64) for w in (WQ, WK, WV)]
        bs = [np.ascontiguousarray(b, dtype=np.float64) for b in (bQ, bK, bV)]
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        out = np.zeros(col_W * 2 * limbs * self.n, dtype=np.uint64)
        oc, ol, osc = C.c_int(0), C.c_int(0), C.c_double(0)
        self._chk(self.lib.ref_single_att_block(self.h, _p(X), C.c_int(num_col), C.c_int(limbs), C.c_double(scale),
                                                _p(ws[0], f64p), _p(ws[1], f64p), _p(ws[2], f64p), C.c_int(col_W),
                                                _p(bs[0], f64p), _p(bs[1], f64p), _p(bs[2], f64p), _p(bv, i32p),
                                                C.c_int(input_num), C.c_int(num_batch), C.c_int(iters), C.c_int(layer_id),
                                                _p(out), C.byref(oc), C.byref(ol), C.byref(osc)))
        return out[: oc.value * 2 * ol.value * self.n].copy(), oc.value, ol.value, osc.value

    def exp_inverse(self, which, x, count, limbs, scale, iters=16):
        """The reference header's exp (which=0, softmax.hpp:9-47) / inverse (which=1, :49-82) -> (residues, limbs, scale)."""
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
        out = np.zeros(count * 2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.ref_exp_inverse(self.h, C.c_int(which), _p(x), C.c_int(count), C.c_int(limbs), C.c_double(scale),
                                           C.c_int(iters), _p(out), C.byref(ol), C.byref(osc)))
        return out[: count * 2 * ol.value * self.n].copy(), ol.value, osc.value

