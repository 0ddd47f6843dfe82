"""ctypes view of the facade test driver (tests/facade_harness/facade_driver.cpp) — TEST INFRASTRUCTURE.

`FacadeDriver(mock=True)` binds the header-only facade (include/moai_b200_seal.hpp) to the CPU test double
of the C ABI (host-logic tests without a GPU); `mock=False` binds it to libmoai_b200.so (the `-m gpu`
tests).  Method names and argument meaning mirror oracle.SealRef, so a test drives both with the same
lines and compares residues bit for bit.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int)
f64p = C.POINTER(C.c_double)


def _p(a, t=u64p):
    return a.ctypes.data_as(t)


def _path(mock, fused):
    if fused:
        return _build.DRIVER_FUSED_MOCK_SO if mock else _build.DRIVER_FUSED_SO
    return _build.DRIVER_MOCK_SO if mock else _build.DRIVER_SO


def available(mock, fused=False):
    _build.build()
    return os.path.exists(_path(mock, fused))


class FacadeError(RuntimeError):
    pass


class FacadeDriver:
    def __init__(self, log_n, bits=None, primes=None, mock=True, device=0, fused=False, prng_seed=None,
                 hamming_weight=0):
        path = _path(mock, fused)
        self.lib = C.CDLL(path)
        L = self.lib
        L.fd_create.restype = C.c_void_p
        L.fd_error.restype = C.c_char_p
        self.log_n, self.n = log_n, 1 << log_n
        if bits is not None:
            arr = (C.c_int * len(bits))(*bits)
            self.h = C.c_void_p(L.fd_create(C.c_int(log_n), arr, None, C.c_int(len(bits)), C.c_int(device)))
        else:
            pr = np.ascontiguousarray(primes, dtype=np.uint64)
            self.h = C.c_void_p(L.fd_create(C.c_int(log_n), None, _p(pr), C.c_int(len(pr)), C.c_int(device)))
        if not L.fd_ok(self.h):
            raise FacadeError("fd_create: " + L.fd_error(self.h).decode())
        if prng_seed is not None:
            self.set_prng_seed(prng_seed, hamming_weight)       # same seeding rule as oracle.SealRef(seed=...)
        self.kl = L.fd_n_key_limbs(self.h)
        q = np.zeros(self.kl, dtype=np.uint64)
        L.fd_primes(self.h, _p(q))
        self.q = q

    def __del__(self):
        try:
            self.lib.fd_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise FacadeError(self.lib.fd_error(self.h).decode())

    # ---- keys from the client (raw residues exported by stock SEAL) ----
    def set_relin(self, key):
        self._chk(self.lib.fd_set_relin(self.h, _p(np.ascontiguousarray(key, dtype=np.uint64))))

    def add_galois(self, elt, key):
        self._chk(self.lib.fd_add_galois(self.h, C.c_uint32(elt), _p(np.ascontiguousarray(key, dtype=np.uint64))))

    def add_galois_fast(self, elt, key, max_limbs):
        self._chk(self.lib.fd_add_galois_fast(self.h, C.c_uint32(elt), _p(np.ascontiguousarray(key, dtype=np.uint64)),
                                              C.c_int(max_limbs)))

    def set_secret(self, sk):
        self._chk(self.lib.fd_set_secret(self.h, _p(np.ascontiguousarray(sk, dtype=np.uint64))))

    def take_keys_from(self, ref, galois=True, secret=True):
        """Upload everything a SealRef holds: relin key, all Galois keys, the secret key."""
        try:
            self.set_relin(ref.export_relin_key())
        except RuntimeError:
            pass
        if galois:
            for elt in ref.galois_elts():
                self.add_galois(elt, ref.export_galois_key(elt))
        if secret:
            self.set_secret(ref.secret_key())

    def chain_index(self, limbs):
        return int(self.lib.fd_chain_index(self.h, C.c_int(limbs)))

    # ---- client-side pieces ----
    def prng_bytes(self, seed8, n):
        out = np.zeros(n, dtype=np.uint8)
        s = np.ascontiguousarray(seed8, dtype=np.uint64)
        self.lib.fd_prng_bytes(_p(s), C.c_int64(n), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def sample_uniform(self, seed8):
        out = np.zeros(self.kl * self.n, dtype=np.uint64)
        s = np.ascontiguousarray(seed8, dtype=np.uint64)
        self._chk(self.lib.fd_sample_uniform(self.h, _p(s), _p(out)))
        return out

    def set_prng_seed(self, seed, hamming_weight=0):
        self._chk(self.lib.fd_set_prng_seed(self.h, C.c_uint64(seed), C.c_int(hamming_weight)))

    def set_public_key(self, pk):
        self._chk(self.lib.fd_set_public_key(self.h, _p(np.ascontiguousarray(pk, dtype=np.uint64))))

    def load_keys(self, kind, blob):
        b = np.frombuffer(blob, dtype=np.uint8)
        seeded = C.c_int(0)
        self._chk(self.lib.fd_load_keys(self.h, C.c_int(kind), b.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(len(blob)),
                                        C.byref(seeded)))
        return seeded.value

    def export_key(self, kind, elt=0):
        words = (1 if kind == 2 else self.kl - 1) * 2 * self.kl * self.n
        out = np.zeros(words, dtype=np.uint64)
        self._chk(self.lib.fd_export_key(self.h, C.c_int(kind), C.c_uint32(elt), _p(out)))
        return out

    def keygen(self, steps=None, conjugate=False):
        if steps is None:
            self._chk(self.lib.fd_keygen(self.h, None, C.c_int(0), C.c_int(0)))
        else:
            arr = (C.c_int * max(1, len(steps)))(*steps)
            self._chk(self.lib.fd_keygen(self.h, arr, C.c_int(len(steps)), C.c_int(int(conjugate))))

    def export_secret(self):
        out = np.zeros(self.kl * self.n, dtype=np.uint64)
        self._chk(self.lib.fd_export_secret(self.h, _p(out)))
        return out

    def has_galois(self, elt):
        return bool(self.lib.fd_has_galois(self.h, C.c_uint32(elt)))

    def encrypt(self, pt, limbs, scale):
        out = np.zeros(2 * limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.fd_encrypt(self.h, _p(pt), C.c_int(limbs), C.c_double(scale), _p(out)))
        return out

    def save_ciphertext(self, ct, size, limbs, scale):
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        cap = C.c_int64(ct.nbytes + 4096)
        buf = np.zeros(cap.value, dtype=np.uint8)
        self._chk(self.lib.fd_save_ciphertext(self.h, _p(ct), C.c_int(size), C.c_int(limbs), C.c_double(scale),
                                              buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(cap)))
        return buf[: cap.value].tobytes()

    def load_ciphertext(self, blob, max_words):
        b = np.frombuffer(blob, dtype=np.uint8)
        out = np.zeros(max_words, dtype=np.uint64)
        size, limbs, scale = C.c_int(0), C.c_int(0), C.c_double(0)
        self._chk(self.lib.fd_load_ciphertext(self.h, b.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(len(blob)),
                                              _p(out), C.c_int64(out.size), C.byref(size), C.byref(limbs), C.byref(scale)))
        return out[: size.value * limbs.value * self.n].copy(), size.value, limbs.value, scale.value

    def batch_input(self, X, scale):
        X = np.ascontiguousarray(X, dtype=np.float64)
        num_X, num_row, num_col = X.shape
        out = np.zeros(num_col * 2 * (self.kl - 1) * self.n, dtype=np.uint64)
        self._chk(self.lib.fd_batch_input(self.h, _p(X, f64p), C.c_int(num_X), C.c_int(num_row), C.c_int(num_col),
                                          C.c_double(scale), _p(out)))
        return out

    def alias_checks(self, a, b, limbs, scale):
        """Value-semantics / aliasing self-checks inside the driver; returns the number of failed checks."""
        failed = C.c_int(-1)
        self._chk(self.lib.fd_alias_checks(self.h, _p(a), _p(b), C.c_int(limbs), C.c_double(scale), C.byref(failed)))
        return failed.value

    # ---- same calls as oracle.SealRef ----
    def eval(self, op, a, size_a, limbs_a, scale_a, b=None, size_b=0, limbs_b=0, scale_b=1.0, iarg=0, darg=0.0,
             varg=None):
        out = np.zeros(3 * max(limbs_a, limbs_b) * self.n, dtype=np.uint64)
        osz, olm, osc = C.c_int(0), C.c_int(0), C.c_double(0)
        bb = _p(b) if b is not None else None
        vv = None
        if varg is not None:
            v = np.asarray(varg, dtype=np.complex128)
            ri = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(-1))
            vv = _p(ri, f64p)
        self._chk(self.lib.fd_eval(self.h, C.c_int(op), _p(a), C.c_int(size_a), C.c_int(limbs_a), C.c_double(scale_a),
                                   bb, C.c_int(size_b), C.c_int(limbs_b), C.c_double(scale_b), C.c_int(iarg),
                                   C.c_double(darg), vv, _p(out), C.byref(osz), C.byref(olm), C.byref(osc)))
        return out[: osz.value * olm.value * self.n].copy(), osz.value, olm.value, osc.value

    def encode(self, values, scale, limbs):
        v = np.asarray(values, dtype=np.complex128)
        ri = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(-1))
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.fd_encode_complex(self.h, _p(ri, f64p), C.c_int(len(v)), C.c_int(limbs), C.c_double(scale),
                                             _p(out)))
        return out

    def encode_real(self, values, scale, limbs):
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.fd_encode_real(self.h, _p(v, f64p), C.c_int(len(v)), C.c_int(limbs), C.c_double(scale),
                                          _p(out)))
        return out

    def decode(self, pt, limbs, scale):
        out = np.zeros(self.n, dtype=np.float64)
        self._chk(self.lib.fd_decode(self.h, _p(pt), C.c_int(limbs), C.c_double(scale), _p(out, f64p)))
        return out[0::2] + 1j * out[1::2]

    def decrypt(self, ct, size, limbs, scale):
        out = np.zeros(limbs * self.n, dtype=np.uint64)
        self._chk(self.lib.fd_decrypt(self.h, _p(ct), C.c_int(size), C.c_int(limbs), C.c_double(scale), _p(out)))
        return out

    def ct_pt_matmul(self, variant, X, W, mask, K, Cc, limbs, scale):
        W = np.ascontiguousarray(W, dtype=np.float64)
        out = np.zeros(Cc * 2 * (limbs - 1) * self.n, dtype=np.uint64)
        sec = C.c_double(0)
        m = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
            m = _p(mask, i32p)
        self._chk(self.lib.fd_ct_pt_matmul(self.h, C.c_int(variant), _p(X), _p(W, f64p), m, C.c_int(K), C.c_int(Cc),
                                           C.c_int(limbs), C.c_double(scale), _p(out), C.byref(sec)))
        return out, sec.value

    def gelu_v2(self, x, count, limbs, scale):
        out = np.zeros(count * 2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.fd_gelu_v2(self.h, _p(x), C.c_int(count), C.c_int(limbs), C.c_double(scale), _p(out),
                                      C.byref(ol), C.byref(osc)))
        return out[: count * 2 * ol.value * self.n].copy(), ol.value, osc.value

    def layernorm(self, variant, x, num_ct, limbs, scale, gamma, beta, bias_vec, want_printed=False):
        out = np.zeros(num_ct * 2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        g = np.ascontiguousarray(gamma, dtype=np.float64)
        b = np.ascontiguousarray(beta, dtype=np.float64)
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        cap = 1 << 20
        buf = C.create_string_buffer(cap)
        self._chk(self.lib.fd_layernorm(self.h, C.c_int(variant), _p(x), C.c_int(num_ct), C.c_int(limbs),
                                        C.c_double(scale), _p(g, f64p), _p(b, f64p), _p(bv, i32p), _p(out),
                                        C.byref(ol), C.byref(osc), buf, C.c_int(cap)))
        res = (out[: num_ct * 2 * ol.value * self.n].copy(), ol.value, osc.value)
        return res + (buf.value.decode(),) if want_printed else res

    def ct_ct_matmul(self, which, X, nX, W, nW, limbs, scale_X, scale_W, col_X, row_X, col_W, row_W, num_batch):
        out = np.zeros(max(row_X, col_W) * 2 * limbs * self.n, dtype=np.uint64)
        oc, ol, osc = C.c_int(0), C.c_int(0), C.c_double(0)
        self._chk(self.lib.fd_ct_ct_matmul(self.h, C.c_int(which), _p(X), C.c_int(nX), _p(W), C.c_int(nW),
                                           C.c_int(limbs), C.c_double(scale_X), C.c_double(scale_W), C.c_int(col_X),
                                           C.c_int(row_X), C.c_int(col_W), C.c_int(row_W), C.c_int(num_batch),
                                           _p(out), C.byref(oc), C.byref(ol), C.byref(osc)))
        return out[: oc.value * 2 * ol.value * self.n].copy(), oc.value, ol.value, osc.value

    def exp(self, x, limbs, scale):
        out = np.zeros(2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.fd_exp(self.h, _p(x), C.c_int(limbs), C.c_double(scale), _p(out), C.byref(ol), C.byref(osc)))
        return out[: 2 * ol.value * self.n].copy(), ol.value, osc.value

    def inverse(self, x, limbs, scale, iters):
        out = np.zeros(2 * limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.fd_inverse(self.h, _p(x), C.c_int(limbs), C.c_double(scale), C.c_int(iters), _p(out),
                                      C.byref(ol), C.byref(osc)))
        return out[: 2 * ol.value * self.n].copy(), ol.value, osc.value

    # ---- bootstrapping (GPU only) ----
    def boot_create(self, loge, logn, total_level, final_scale, boundary_K=25, deg=59, scale_factor=2, hoisting=False):
        steps = (C.c_int * 4096)()
        cnt = C.c_int(0)
        self._chk(self.lib.fd_boot_create(self.h, C.c_int(loge), C.c_int(logn), C.c_int(total_level),
                                          C.c_double(final_scale), C.c_int(boundary_K), C.c_int(deg),
                                          C.c_int(scale_factor), C.c_int(int(hoisting)), steps, C.c_int(4096),
                                          C.byref(cnt)))
        return [int(steps[i]) for i in range(cnt.value)]

    def bootstrap_3(self, x, scale, max_limbs):
        out = np.zeros(2 * max_limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        self._chk(self.lib.fd_bootstrap_3(self.h, _p(x), C.c_double(scale), _p(out), C.byref(ol), C.byref(osc)))
        return out[: 2 * ol.value * self.n].copy(), ol.value, osc.value

    def boot_combined(self, x, n_cts, scale, max_limbs, real_slots=True, max_batch=64, linger_us=300):
        out = np.zeros(n_cts * 2 * max_limbs * self.n, dtype=np.uint64)
        ol, calls = C.c_int(0), C.c_int(0)
        self._chk(self.lib.fd_boot_combined(self.h, _p(x), C.c_int(n_cts), C.c_double(scale), C.c_int(int(real_slots)),
                                            C.c_int(max_batch), C.c_int(linger_us), _p(out), C.byref(ol), C.byref(calls)))
        return out[: n_cts * 2 * ol.value * self.n].reshape(n_cts, 2, ol.value, self.n).copy(), calls.value

    def parallel_chain(self, x, n_cts, limbs, scale, lanes, with_rotation=True):
        """square / multiply / relinearize / add / rescale (/ rotate) on n ciphertexts from an OpenMP loop, with one lane
        per thread (lanes=True) or the context mutex (lanes=False); returns ([n][2][limbs-1][N], threads used)."""
        out = np.zeros(n_cts * 2 * (limbs - 1) * self.n, dtype=np.uint64)
        th, ms = C.c_int(0), C.c_double(0)
        self._chk(self.lib.fd_parallel_chain(self.h, _p(x), C.c_int(n_cts), C.c_int(limbs), C.c_double(scale),
                                             C.c_int(int(lanes)), C.c_int(int(with_rotation)), _p(out), C.byref(th),
                                             C.byref(ms)))
        self.last_loop_ms = ms.value
        return out.reshape(n_cts, 2, limbs - 1, self.n), th.value

    def softmax_boot(self, x, num, limbs, scale, bias_vec, input_num, iters, layer_id, max_limbs):
        out = np.zeros(num * 2 * max_limbs * self.n, dtype=np.uint64)
        ol, osc = C.c_int(0), C.c_double(0)
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        self._chk(self.lib.fd_softmax_boot(self.h, _p(x), C.c_int(num), C.c_int(limbs), C.c_double(scale), _p(bv, i32p),
                                           C.c_int(input_num), C.c_int(iters), C.c_int(layer_id), _p(out), C.byref(ol),
                                           C.byref(osc)))
        return out[: num * 2 * ol.value * self.n].copy(), ol.value, osc.value


def reference_seal_ckks_test(mock=True, fused=False):
    """Runs the reference's own SEAL_ckks_test() through the facade; returns (ok, printed text)."""
    lib = C.CDLL(_path(mock, fused))
    buf = C.create_string_buffer(1 << 16)
    rc = lib.fd_reference_seal_ckks_test(buf, C.c_int(1 << 16))
    return rc == 0, buf.value.decode(errors="replace")
