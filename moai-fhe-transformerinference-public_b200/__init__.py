"""moai-fhe-transformerinference-public_b200 — B200-native CKKS evaluation backend (host-side Python mirror).

The product is ``libmoai_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/moai_b200.h``).  This module is the thin Python host side used by the tests and
``bench.py``: it loads the library with ctypes and mirrors the reference's ``seal::Evaluator``
method names (S/evaluator.h:93-1386) on *batches* of ciphertexts held in torch CUDA tensors
(torch only provides device memory, streams and ``torch.distributed`` plumbing).

There is NO CPU fallback: constructing :class:`Backend` without a CUDA device, or without the
built extension, raises.  Nothing here imports ``oracle/``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmoai_b200.so")

MOAI_BITS = [51] + [46] * 20 + [51] * 14 + [58]  # M/test/test_full_scheme.hpp:356-378
MOAI_LOG_N = 16

_u64p = C.POINTER(C.c_uint64)
_lib = None


class MoaiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmoai_b200 status %d: %s" % (code, msg))
        self.code = code


def load_library():
    """dlopen the in-tree CUDA extension; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libmoai_b200.so is missing: run `python __graft_entry__.py` (build()) first; "
                              "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.moai_last_error.restype = C.c_char_p
    return _lib


def exported_symbols():
    """Every entry point include/moai_b200.h declares (checked by the CPU test-suite)."""
    import re
    hdr = open(os.path.join(_HERE, "..", "include", "moai_b200.h")).read()
    hdr += open(os.path.join(_HERE, "..", "include", "moai_b200_modules.h")).read() \
        if os.path.exists(os.path.join(_HERE, "..", "include", "moai_b200_modules.h")) else ""
    return sorted(set(re.findall(r"\b(moai_[a-z0-9_]+)\s*\(", hdr)))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def to_device(a, device="cuda"):
    """numpy uint64 array -> torch int64 CUDA tensor holding the same bits."""
    import torch
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return torch.from_numpy(a.view(np.int64)).to(device)


def to_host(t):
    return t.detach().cpu().numpy().view(np.uint64)


class Backend:
    """One CKKS context on one GPU.  Tensors are int64 CUDA tensors carrying uint64 residues,
    shaped [batch, size, limbs, n] (ciphertexts) or [limbs, n] (plaintexts)."""

    def __init__(self, log_n, primes, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("moai_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.lib = load_library()
        self.log_n, self.n = log_n, 1 << log_n
        self.primes = [int(p) for p in primes]
        self.kl = len(self.primes)
        self.device = torch.device("cuda", device)
        arr = (C.c_uint64 * self.kl)(*self.primes)
        h = C.c_void_p()
        self._chk(self.lib.moai_context_create(C.c_int32(log_n), arr, C.c_int32(self.kl), C.c_int32(device),
                                               C.byref(h)))
        self.h = h
        self.use_torch_stream()

    def fork(self):
        """A lane of this context (moai_context_fork): the same tables, its own CUDA stream and arena — one per host thread.
        Close lanes before their parent; results of a lane are complete after lane.synchronize()."""
        lane = object.__new__(Backend)
        lane.torch, lane.lib = self.torch, self.lib
        lane.log_n, lane.n, lane.primes, lane.kl, lane.device = self.log_n, self.n, self.primes, self.kl, self.device
        h = C.c_void_p()
        self._chk(self.lib.moai_context_fork(self.h, C.byref(h)))
        lane.h = h
        return lane

    def close(self):
        if getattr(self, "h", None):
            self.lib.moai_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise MoaiError(rc, self.lib.moai_last_error().decode())

    def use_torch_stream(self):
        """Launch on torch's current stream so torch.cuda.Event timing sees the kernels."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream
        self._chk(self.lib.moai_set_stream(self.h, C.c_void_p(s)))

    # ---- one packed batch over several GPUs (moai_comm_*): call after torch.distributed.init_process_group
    def comm_init(self, dist):
        """Join this context to an NCCL communicator spanning the ranks of torch.distributed process group `dist`
        (rank 0 creates the id, torch.distributed ships the 128 bytes)."""
        rank, world = dist.get_rank(), dist.get_world_size()
        idb = (C.c_uint8 * 128)()
        if rank == 0:
            self._chk(self.lib.moai_comm_unique_id(idb))
        t = self.torch.tensor(list(idb), dtype=self.torch.uint8, device=self.device)
        dist.broadcast(t, src=0)
        idb = (C.c_uint8 * 128)(*[int(v) for v in t.cpu().tolist()])
        self._chk(self.lib.moai_comm_init(self.h, idb, C.c_int32(rank), C.c_int32(world)))
        self.comm_rank, self.comm_world = rank, world

    def comm_stats(self):
        g, b = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.moai_comm_stats(self.h, C.byref(g), C.byref(b)))
        return g.value, b.value

    def profile(self, on):
        self._chk(self.lib.moai_profile_enable(self.h, C.c_int32(int(on))))

    def profile_get(self, name):
        ms, cnt = C.c_double(), C.c_int64()
        self._chk(self.lib.moai_profile_get(self.h, name.encode(), C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def profile_dump(self):
        buf = C.create_string_buffer(1 << 16)
        self._chk(self.lib.moai_profile_dump(self.h, buf, C.c_int32(1 << 16)))
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                name, ms, cnt = item.rsplit(":", 2)
                out[name] = (float(ms), int(cnt))
        return out

    def launch_count(self):
        v = C.c_uint64()
        self._chk(self.lib.moai_launch_count(self.h, C.byref(v)))
        return v.value

    def synchronize(self):
        self._chk(self.lib.moai_synchronize(self.h))

    def empty(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.int64, device=self.device)

    # ---- NTT (A1/A2)
    def ntt_forward_(self, x):
        b, p, l, n = x.shape
        self._chk(self.lib.moai_ntt_forward(self.h, _ptr(x), C.c_int64(b), C.c_int32(p), C.c_int32(l)))
        return x

    def ntt_inverse_(self, x):
        b, p, l, n = x.shape
        self._chk(self.lib.moai_ntt_inverse(self.h, _ptr(x), C.c_int64(b), C.c_int32(p), C.c_int32(l)))
        return x

    def ntt_forward_limb_(self, x, limb):
        self._chk(self.lib.moai_ntt_forward_limb(self.h, _ptr(x), C.c_int64(x.numel() // self.n), C.c_int32(limb)))
        return x

    def ntt_inverse_limb_(self, x, limb):
        self._chk(self.lib.moai_ntt_inverse_limb(self.h, _ptr(x), C.c_int64(x.numel() // self.n), C.c_int32(limb)))
        return x

    # ---- element-wise (A3-A5, A10)
    def _bin(self, fn, a, b, out=None):
        bt, p, l, n = a.shape
        out = self.torch.empty_like(a) if out is None else out
        self._chk(fn(self.h, _ptr(a), _ptr(b), _ptr(out), C.c_int64(bt), C.c_int32(p), C.c_int32(l)))
        return out

    def add(self, a, b, out=None):
        return self._bin(self.lib.moai_add, a, b, out)

    def sub(self, a, b, out=None):
        return self._bin(self.lib.moai_sub, a, b, out)

    def negate(self, a, out=None):
        bt, p, l, n = a.shape
        out = self.torch.empty_like(a) if out is None else out
        self._chk(self.lib.moai_negate(self.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(p), C.c_int32(l)))
        return out

    def _plain(self, fn, ct, pt, out=None):
        bt, p, l, n = ct.shape
        stride = 0 if pt.dim() == 2 else l * n
        out = self.torch.empty_like(ct) if out is None else out
        self._chk(fn(self.h, _ptr(ct), _ptr(pt), _ptr(out), C.c_int64(bt), C.c_int32(p), C.c_int32(l),
                     C.c_int64(stride)))
        return out

    def add_plain(self, ct, pt, out=None):
        return self._plain(self.lib.moai_add_plain, ct, pt, out)

    def sub_plain(self, ct, pt, out=None):
        return self._plain(self.lib.moai_sub_plain, ct, pt, out)

    def multiply_plain(self, ct, pt, out=None):
        return self._plain(self.lib.moai_multiply_plain, ct, pt, out)

    def multiply(self, a, b, out=None, accumulate=False):
        bt, p, l, n = a.shape
        out = self.empty(bt, 3, l, n) if out is None else out
        self._chk(self.lib.moai_multiply(self.h, _ptr(a), _ptr(b), _ptr(out), C.c_int64(bt), C.c_int32(l),
                                         C.c_int32(int(accumulate))))
        return out

    def square(self, a, out=None):
        bt, p, l, n = a.shape
        out = self.empty(bt, 3, l, n) if out is None else out
        self._chk(self.lib.moai_square(self.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(l)))
        return out

    # ---- level changes (A8, A9, C1)
    def rescale_to_next(self, a, out=None):
        bt, p, l, n = a.shape
        out = self.empty(bt, p, l - 1, n) if out is None else out
        self._chk(self.lib.moai_rescale_to_next(self.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(p), C.c_int32(l)))
        return out

    def mod_switch_to(self, a, limbs_out, out=None):
        bt, p, l, n = a.shape
        out = self.empty(bt, p, limbs_out, n) if out is None else out
        self._chk(self.lib.moai_mod_switch_to(self.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(p), C.c_int32(l),
                                              C.c_int32(limbs_out)))
        return out

    def mod_switch_to_next(self, a):
        return self.mod_switch_to(a, a.shape[2] - 1)

    def mod_raise(self, a, limbs_out):
        bt, p, l, n = a.shape
        assert l == 1
        out = self.empty(bt, p, limbs_out, n)
        self._chk(self.lib.moai_mod_raise(self.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(p),
                                          C.c_int32(limbs_out)))
        return out

    # ---- key switching (A6, A7)
    def galois_elt_from_step(self, step):
        e = C.c_uint32()
        self._chk(self.lib.moai_galois_elt_from_step(self.h, C.c_int32(step), C.byref(e)))
        return e.value

    def rotate_naf_steps(self, steps):
        out = (C.c_int32 * 64)()
        cnt = C.c_int32()
        self._chk(self.lib.moai_rotate_naf_steps(self.h, C.c_int32(steps), out, C.byref(cnt)))
        return [out[i] for i in range(cnt.value)]

    def apply_galois(self, a, elt, ksk, out=None):
        bt, p, l, n = a.shape
        out = self.torch.empty_like(a) if out is None else out
        self._chk(self.lib.moai_apply_galois(self.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(l), C.c_uint32(elt),
                                             _ptr(ksk)))
        return out

    def rotate_vector(self, a, steps, galois_keys):
        """Evaluator::rotate_vector incl. the NAF fallback for missing keys
        (S/evaluator.cpp:2667-2722).  galois_keys: dict galois_elt -> device key tensor."""
        if steps == 0:
            return a
        elt = self.galois_elt_from_step(steps)
        if elt in galois_keys:
            return self.apply_galois(a, elt, galois_keys[elt])
        if abs(steps) & (abs(steps) - 1) == 0:  # NAF has a single term: nothing to decompose into
            raise MoaiError(1, "Galois key not present")
        terms = self.rotate_naf_steps(steps)
        for s in terms:
            a = self.rotate_vector(a, s, galois_keys)
        return a

    def complex_conjugate(self, a, galois_keys):
        elt = self.galois_elt_from_step(0)
        if elt not in galois_keys:
            raise MoaiError(1, "Galois key not present")
        return self.apply_galois(a, elt, galois_keys[elt])

    def relinearize(self, a3, relin_key, out=None):
        bt, p, l, n = a3.shape
        out = self.empty(bt, 2, l, n) if out is None else out
        self._chk(self.lib.moai_relinearize(self.h, _ptr(a3), _ptr(out), C.c_int64(bt), C.c_int32(l), _ptr(relin_key)))
        return out

    def switch_key_(self, ct, target, ksk):
        bt, p, l, n = ct.shape
        self._chk(self.lib.moai_switch_key(self.h, _ptr(ct), _ptr(target), C.c_int64(bt), C.c_int32(l), _ptr(ksk)))
        return ct

    # ---- scalar plaintexts (A11, A13)
    def encode_scalar_consts(self, value, scale, limbs):
        out = (C.c_uint64 * limbs)()
        self._chk(self.lib.moai_encode_scalar_consts(self.h, C.c_double(value), C.c_double(scale), C.c_int32(limbs),
                                                     out))
        return np.array(list(out), dtype=np.uint64)

    def multiply_const(self, a, value, scale, out=None):
        bt, p, l, n = a.shape
        k = self.encode_scalar_consts(value, scale, l)
        out = self.torch.empty_like(a) if out is None else out
        self._chk(self.lib.moai_multiply_scalar(self.h, _ptr(a), k.ctypes.data_as(_u64p), _ptr(out), C.c_int64(bt),
                                                C.c_int32(p), C.c_int32(l)))
        return out

    def add_const(self, a, value, scale, out=None):
        bt, p, l, n = a.shape
        k = self.encode_scalar_consts(value, scale, l)
        out = self.torch.empty_like(a) if out is None else out
        self._chk(self.lib.moai_add_scalar(self.h, _ptr(a), k.ctypes.data_as(_u64p), _ptr(out), C.c_int64(bt),
                                           C.c_int32(p), C.c_int32(l)))
        return out

    # ---- fused modules (B1/B2)
    def ct_pt_matrix_mul_wo_pre(self, enc_X, W, scale, out=None):
        """ct_pt_matrix_mul_wo_pre / _wo_pre_large (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:4-101).
        enc_X: [K, 2, limbs, n] device tensor; W: K x C float64 (host).  Returns [C, 2, limbs-1, n]."""
        K, p, l, n = enc_X.shape
        W = np.ascontiguousarray(W, dtype=np.float64)
        if W.shape[0] != K:
            raise MoaiError(1, "bad dimensions of X or W")
        Cc = W.shape[1]
        out = self.empty(Cc, 2, l - 1, n) if out is None else out
        self._chk(self.lib.moai_ct_pt_matrix_mul_wo_pre(
            self.h, _ptr(enc_X), W.ctypes.data_as(C.POINTER(C.c_double)), C.c_int32(K), C.c_int32(Cc), C.c_int32(K),
            C.c_int32(l), C.c_double(scale), _ptr(out)))
        return out

    ct_pt_matrix_mul_wo_pre_large = ct_pt_matrix_mul_wo_pre

    def ct_pt_matrix_mul_wo_pre_host(self, host_X, W, scale, out=None):
        """The same module on HOST tensors (pinned CPU int64 [K, 2, limbs, n] -> [C, 2, limbs-1, n]):
        upload, tensor-core GEMM and download pipelined inside the library
        (moai_ct_pt_matrix_mul_wo_pre_host)."""
        K, p, l, n = host_X.shape
        assert host_X.device.type == "cpu" and host_X.is_contiguous()
        W = np.ascontiguousarray(W, dtype=np.float64)
        if W.shape[0] != K:
            raise MoaiError(1, "bad dimensions of X or W")
        Cc = W.shape[1]
        if out is None:
            out = self.torch.empty((Cc, 2, l - 1, n), dtype=self.torch.int64, pin_memory=True)
        self._chk(self.lib.moai_ct_pt_matrix_mul_wo_pre_host(
            self.h, C.c_void_p(host_X.data_ptr()), W.ctypes.data_as(C.POINTER(C.c_double)), C.c_int32(K),
            C.c_int32(Cc), C.c_int32(K), C.c_int32(l), C.c_double(scale), C.c_void_p(out.data_ptr())))
        return out

    def ct_pt_matrix_mul_wo_pre_w_mask(self, enc_X, W, bias_vec, scale, out=None):
        """ct_pt_matrix_mul_wo_pre_w_mask (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:103-170)."""
        K, p, l, n = enc_X.shape
        W = np.ascontiguousarray(W, dtype=np.float64)
        mask = np.ascontiguousarray(bias_vec, dtype=np.int32)
        if W.shape[0] != K or mask.size != n // 2:
            raise MoaiError(1, "bad dimensions of X or W")
        Cc = W.shape[1]
        out = self.empty(Cc, 2, l - 1, n) if out is None else out
        self._chk(self.lib.moai_ct_pt_matrix_mul_wo_pre_w_mask(
            self.h, _ptr(enc_X), W.ctypes.data_as(C.POINTER(C.c_double)), mask.ctypes.data_as(C.POINTER(C.c_int32)),
            C.c_int32(K), C.c_int32(Cc), C.c_int32(K), C.c_int32(l), C.c_double(scale), _ptr(out)))
        return out

    def ct_pt_matrix_mul_wo_pre_w_mask_fast(self, enc_X, W, bias_vec, scale, out=None):
        """Fast-mode masked matmul: one tensor-core GEMM times ONE mask plaintext (tolerance, not bit-exact)."""
        K, p, l, n = enc_X.shape
        W = np.ascontiguousarray(W, dtype=np.float64)
        mask = np.ascontiguousarray(bias_vec, dtype=np.int32)
        if W.shape[0] != K or mask.size != n // 2:
            raise MoaiError(1, "bad dimensions of X or W")
        Cc = W.shape[1]
        out = self.empty(Cc, 2, l - 1, n) if out is None else out
        self._chk(self.lib.moai_ct_pt_matrix_mul_wo_pre_w_mask_fast(
            self.h, _ptr(enc_X), W.ctypes.data_as(C.POINTER(C.c_double)), mask.ctypes.data_as(C.POINTER(C.c_int32)),
            C.c_int32(K), C.c_int32(Cc), C.c_int32(K), C.c_int32(l), C.c_double(scale), _ptr(out)))
        return out

    # ---- vector encoder (A12)
    def encode(self, values, scale, limbs):
        """CKKSEncoder::encode for a batch: values [count, n_vals] (or [n_vals]) real/complex ->
        device plaintexts [count, limbs, n] (or [limbs, n])."""
        v = np.asarray(values)
        single = v.ndim == 1
        v = np.atleast_2d(v).astype(np.complex128)
        count, n_vals = v.shape
        ri = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1))
        d = self.torch.from_numpy(ri).to(self.device)
        out = self.empty(count, limbs, self.n)
        self._chk(self.lib.moai_encode_vector(self.h, _ptr(d), C.c_int64(count), C.c_int32(n_vals), C.c_double(scale),
                                              C.c_int32(limbs), _ptr(out)))
        return out[0] if single else out

    # ---- evaluation keys + MOAI modules (B4-B9)
    def make_keys(self, relin=None, galois=None, galois_fast=None, grouped=None, single=None):
        """relin: device key tensor; galois: dict galois_elt -> device key tensor (SEAL layout,
        bit-exact rotations); galois_fast: dict galois_elt -> tensor or list of tensors produced by
        key_prepare(pre_permute=True), shaped [L, 2, L + 1, n] (hoisted fast-mode rotations);
        grouped: dict galois_elt (0 = relinearisation key) -> list of GroupedKey from key_prepare_grouped;
        single: dict galois_elt -> tensor [1, 2, kl, n] from key_prepare_single (rotations of mod-raised ciphertexts).
        The returned handle keeps the tensors alive."""
        h = C.c_void_p()
        self._chk(self.lib.moai_keys_create(self.h, C.byref(h)))
        keep = []
        if relin is not None:
            self._chk(self.lib.moai_keys_set_relin(h, _ptr(relin)))
            keep.append(relin)
        for elt, t in (galois or {}).items():
            if t.dim() != 4 or t.shape[2] == self.kl:   # flat or [kl - 1, 2, kl, n]: a full SEAL-layout key
                self._chk(self.lib.moai_keys_add_galois(h, C.c_uint32(elt), _ptr(t)))
            else:   # level-truncated SEAL-exact key [L, 2, L + 1, n] (key_prepare(pre_permute=False))
                self._chk(self.lib.moai_keys_add_galois_truncated(h, C.c_uint32(elt), _ptr(t), C.c_int32(t.shape[2])))
            keep.append(t)
        for elt, ts in (galois_fast or {}).items():
            for t in (ts if isinstance(ts, (list, tuple)) else [ts]):
                self._chk(self.lib.moai_keys_add_galois_fast(h, C.c_uint32(elt), _ptr(t), C.c_int32(t.shape[2])))
                keep.append(t)
        for elt, t in (single or {}).items():
            self._chk(self.lib.moai_keys_add_single(h, C.c_uint32(elt), _ptr(t)))
            keep.append(t)
        for elt, gks in (grouped or {}).items():
            for gk in (gks if isinstance(gks, (list, tuple)) else [gks]):
                self._chk(self.lib.moai_keys_add_grouped(h, C.c_uint32(elt), _ptr(gk.t), C.c_int32(gk.k_extra),
                                                         C.c_int32(gk.max_limbs)))
                keep.append(gk.t)
        return _KeysHandle(self.lib, h, keep)

    def ksg_best_extra(self, limbs):
        """Extra primes the library's cost model prefers for a key switch at `limbs` (0 = SEAL's digits)."""
        k = C.c_int32()
        self._chk(self.lib.moai_ksg_best_extra(self.h, C.c_int32(limbs), C.byref(k)))
        return k.value

    def ksg_key_shape(self, k_extra, max_limbs):
        d, kl = C.c_int32(), C.c_int32()
        self._chk(self.lib.moai_ksg_key_shape(self.h, C.c_int32(k_extra), C.c_int32(max_limbs), C.byref(d), C.byref(kl)))
        return d.value, kl.value

    def expand_seeds(self, seeds, limbs):
        """seeds: [count, 8] uint64 (numpy) -> device [count, limbs, n]: SEAL's sample_poly_uniform(Blake2xbPRNG(seed))
        regenerated on the device (the uniform half of a seeded key digit / ciphertext)."""
        s = np.ascontiguousarray(seeds, dtype=np.uint64).reshape(-1, 8)
        out = self.empty(s.shape[0], limbs, self.n)
        self._chk(self.lib.moai_expand_seeds(self.h, s.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(s.shape[0]),
                                             C.c_int32(limbs), _ptr(out), C.c_int64(limbs * self.n)))
        return out

    def ksg_plan(self, levels):
        """Grouped-key variants that give every level in `levels` its preferred digit layout: dict k_extra ->
        max_limbs (one key per distinct k, truncated to the highest level that wants it; levels that prefer SEAL's
        digits are left to the SEAL-layout / pre-permuted key)."""
        plan = {}
        for lv in levels:
            k = self.ksg_best_extra(lv)
            if k:
                plan[k] = max(plan.get(k, 0), lv)
        return plan

    def random_grouped_key(self, k_extra, max_limbs, generator=None):
        """Uniformly random residues in the layout of a grouped key (timing runs: a key switch's time does not depend
        on key values)."""
        d, gkl = self.ksg_key_shape(k_extra, max_limbs)
        t = self.torch.empty((d, 2, gkl, self.n), dtype=self.torch.int64, device=self.device)
        ids = list(range(max_limbs)) + list(range(self.kl - 1 - k_extra, self.kl - 1)) + [self.kl - 1]
        for j, pid in enumerate(ids):
            t[:, :, j, :] = self.torch.randint(0, int(self.primes[pid]), (d, 2, self.n), generator=generator,
                                               device=self.device, dtype=self.torch.int64)
        return GroupedKey(t, k_extra, max_limbs)

    def key_prepare_grouped(self, ksk, elt, max_limbs, k_extra=None, pre_permute=True):
        """SEAL-layout key [kl-1, 2, kl, n] -> grouped-digit key [digits, 2, max_limbs + k + 1, n] for the fast-mode
        key switch at levels <= max_limbs (include/moai_b200_modules.h); elt = 0 with pre_permute=False for the
        relinearisation key.  Returns None when the cost model prefers SEAL's digits (k = 0) at this level."""
        k = self.ksg_best_extra(max_limbs) if k_extra is None else k_extra
        if k == 0:
            return None
        d, kl = self.ksg_key_shape(k, max_limbs)
        out = self.empty(d, 2, kl, self.n)
        self._chk(self.lib.moai_key_prepare_grouped(self.h, _ptr(ksk), C.c_uint32(elt), C.c_int32(k), C.c_int32(max_limbs),
                                                    C.c_int32(int(pre_permute)), _ptr(out)))
        return GroupedKey(out, k, max_limbs)

    def key_prepare_single(self, ksk, elt, pre_permute=False):
        """SEAL-layout Galois key -> single-digit key [1, 2, kl, n] (sum of all digits, NATURAL order: the fused
        first-stage kernel gathers the digit instead of permuting the key) for rotations of a mod-raised ciphertext
        (first CoeffToSlot stage, hoisting mode 2)."""
        out = self.empty(1, 2, self.kl, self.n)
        self._chk(self.lib.moai_key_prepare_single(self.h, _ptr(ksk), C.c_uint32(elt), C.c_int32(int(pre_permute)), _ptr(out)))
        return out

    def key_prepare(self, ksk, elt, max_limbs=None, pre_permute=True):
        """SEAL-layout Galois key [kl-1, 2, kl, n] -> level-truncated (and pre-permuted) key
        [L, 2, L + 1, n] for the hoisted fast-mode rotations (include/moai_b200_modules.h)."""
        L = self.kl - 1 if max_limbs is None else max_limbs
        out = self.empty(L, 2, L + 1, self.n)
        self._chk(self.lib.moai_key_prepare(self.h, _ptr(ksk), C.c_uint32(elt), C.c_int32(L),
                                            C.c_int32(int(pre_permute)), _ptr(out)))
        return out

    def rotate_vector_keys(self, keys, a, steps):
        """Evaluator::rotate_vector through a key handle (SEAL-layout key, else pre-permuted key,
        else SEAL's NAF fallback)."""
        bt, p, l, n = a.shape
        out = self.torch.empty_like(a)
        self._chk(self.lib.moai_rotate_vector(self.h, keys.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(l),
                                              C.c_int32(steps)))
        return out

    def relinearize_keys(self, keys, a3):
        bt, p, l, n = a3.shape
        out = self.empty(bt, 2, l, n)
        self._chk(self.lib.moai_relinearize_keys(self.h, keys.h, _ptr(a3), _ptr(out), C.c_int64(bt), C.c_int32(l)))
        return out

    def relin_rescale_keys(self, keys, a3):
        """rescale_to_next(relinearize(a3)); one merged division with a grouped-digit key (fast mode)."""
        bt, p, l, n = a3.shape
        out = self.empty(bt, 2, l - 1, n)
        self._chk(self.lib.moai_relin_rescale_keys(self.h, keys.h, _ptr(a3), _ptr(out), C.c_int64(bt), C.c_int32(l)))
        return out

    def complex_conjugate_keys(self, keys, a):
        bt, p, l, n = a.shape
        out = self.torch.empty_like(a)
        self._chk(self.lib.moai_complex_conjugate_keys(self.h, keys.h, _ptr(a), _ptr(out), C.c_int64(bt), C.c_int32(l)))
        return out

    def rotate_many(self, keys, a, steps):
        """[len(steps), batch, 2, limbs, n]: every rotation of the batch, one shared digit
        decomposition when all steps have pre-permuted keys."""
        bt, p, l, n = a.shape
        st = (C.c_int32 * len(steps))(*steps)
        out = self.empty(len(steps), bt, 2, l, n)
        self._chk(self.lib.moai_rotate_many(self.h, keys.h, _ptr(a), C.c_int64(bt), C.c_int32(l), st,
                                            C.c_int32(len(steps)), _ptr(out)))
        return out

    def _module(self, fn, x, scale, *extra, out_count=None, out=None):
        bt, p, l, n = x.shape
        cnt = bt if out_count is None else out_count
        buf = self.empty(cnt, 2, l, n) if out is None else out
        ol, osc = C.c_int32(), C.c_double()
        self._chk(fn(*extra, _ptr(buf), C.byref(ol), C.byref(osc)))
        flat = buf.reshape(-1)[: cnt * 2 * ol.value * n]
        return flat.reshape(cnt, 2, ol.value, n), osc.value

    def gelu_v2(self, keys, x, scale):
        bt, p, l, n = x.shape
        return self._module(self.lib.moai_gelu_v2, x, scale, self.h, keys.h, _ptr(x), C.c_int64(bt), C.c_int32(l),
                            C.c_double(scale))

    def layernorm(self, keys, x, scale, gamma, beta, bias_vec, variant=1):
        bt, p, l, n = x.shape
        g = np.ascontiguousarray(gamma, dtype=np.float64)
        b = np.ascontiguousarray(beta, dtype=np.float64)
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        return self._module(self.lib.moai_layernorm, x, scale, self.h, keys.h, _ptr(x), C.c_int32(bt), C.c_int32(l),
                            C.c_double(scale), g.ctypes.data_as(dp), b.ctypes.data_as(dp),
                            bv.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(variant))

    def exp(self, keys, x, scale):
        bt, p, l, n = x.shape
        return self._module(self.lib.moai_exp, x, scale, self.h, keys.h, _ptr(x), C.c_int64(bt), C.c_int32(l),
                            C.c_double(scale))

    def inverse(self, keys, x, scale, iters):
        bt, p, l, n = x.shape
        return self._module(self.lib.moai_inverse, x, scale, self.h, keys.h, _ptr(x), C.c_int64(bt), C.c_int32(l),
                            C.c_double(scale), C.c_int32(iters))

    def ct_ct_matrix_mul_colpacking(self, keys, X, W, scale_X, scale_W, col_X, row_X, col_W, row_W, num_batch):
        l = X.shape[2]
        return self._module(self.lib.moai_ct_ct_matrix_mul_colpacking, X, scale_X, self.h, keys.h, _ptr(X), _ptr(W),
                            C.c_int32(l), C.c_double(scale_X), C.c_double(scale_W), C.c_int32(col_X), C.c_int32(row_X),
                            C.c_int32(col_W), C.c_int32(row_W), C.c_int32(num_batch), out_count=row_X)

    def ct_ct_matrix_mul_diagpacking(self, keys, X, W, scale_X, scale_W, col_X, row_X, col_W, row_W, num_batch):
        l = X.shape[2]
        return self._module(self.lib.moai_ct_ct_matrix_mul_diagpacking, X, scale_X, self.h, keys.h, _ptr(X), _ptr(W),
                            C.c_int32(l), C.c_double(scale_X), C.c_double(scale_W), C.c_int32(col_X), C.c_int32(row_X),
                            C.c_int32(col_W), C.c_int32(row_W), C.c_int32(num_batch), out_count=col_W)


def attention_rotation_steps(num_batch, tokens=128):
    """Rotation steps (left, in slots) one attention head takes in fast mode, so that no rotation
    falls back to SEAL's NAF chain: {level tag: steps}.  "qk": ct_ct_matrix_mul_colpacking at the
    Q/K level (giant 16 a, hoisted baby b; csrc/modules.cu); "sv": ct_ct_matrix_mul_diagpacking at
    the softmax-output level (Ct_ct_matrix_mul.hpp:70-151: g = ceil(sqrt(tokens)) baby steps of V,
    group rotations of the diagonals, giant rotations of the partial sums)."""
    import math
    qk = {b * num_batch for b in range(1, 16)} | {a * num_batch for a in range(16, tokens, 16)}
    g = int(math.sqrt(tokens))
    g += g * g < tokens
    b = -(-tokens // g)
    sv = {k * num_batch for k in range(1, g)}
    sv |= {(tokens - i * g) * num_batch for i in range(1, b) if i * g < tokens}
    sv |= {j * g * num_batch for j in range(1, b)}
    return {"qk": sorted(qk), "sv": sorted(sv)}


class LayerWeightsC(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32), ("inter", C.c_int32)] + \
        [(k, C.POINTER(C.c_double)) for k in ("WQ", "WK", "WV", "bQ", "bK", "bV", "selfoutput", "selfoutput_bias",
                                              "ln1_gamma", "ln1_beta", "inter_weight", "inter_bias", "final_weight",
                                              "final_bias", "ln2_gamma", "ln2_beta")]


class Bootstrapper:
    """Bootstrapper (M/source/bootstrapping/Bootstrapper.h:15-221) bound to one Backend."""

    def __init__(self, be, total_limbs, final_scale=2.0 ** 46, boundary_K=25, deg=59, double_angles=2, log_width=10):
        self.be = be
        h = C.c_void_p()
        be._chk(be.lib.moai_bootstrapper_create(be.h, C.c_int32(total_limbs), C.c_double(final_scale),
                                                C.c_int32(boundary_K), C.c_int32(deg), C.c_int32(double_angles),
                                                C.c_int32(log_width), C.byref(h)))
        self.h = h
        self.total_limbs = total_limbs

    def set_hoisting(self, on=True):
        """Plan the linear stages for hoisted baby steps; call before required_steps()."""
        self.be._chk(self.be.lib.moai_bootstrapper_set_hoisting(self.h, C.c_int32(int(on))))

    def required_steps(self):
        buf = (C.c_int32 * 1024)()
        cnt = C.c_int32()
        self.be._chk(self.be.lib.moai_bootstrapper_required_steps(self.h, buf, C.c_int32(1024), C.byref(cnt)))
        return [int(buf[i]) for i in range(cnt.value)]

    def required_step_levels(self):
        """dict step -> sorted list of levels (limb counts) the rotation key is used at; step 0 = conjugation."""
        cap = 1024
        st, lv = (C.c_int32 * cap)(), (C.c_int32 * cap)()
        cnt = C.c_int32()
        self.be._chk(self.be.lib.moai_bootstrapper_required_step_levels(self.h, st, lv, C.c_int32(cap), C.byref(cnt)))
        out = {}
        for i in range(cnt.value):
            out.setdefault(st[i], []).append(lv[i])
        return {k: sorted(set(v)) for k, v in out.items()}

    def bootstrap_3(self, keys, x, scale):
        """x: [batch, 2, 1, n] at chain_index 0 -> ([batch, 2, total_limbs - 14, n], final_scale)."""
        be = self.be
        bt, p, l, n = x.shape
        out = be.empty(bt, 2, self.total_limbs - 14, n)
        ol, osc = C.c_int32(), C.c_double()
        be._chk(be.lib.moai_bootstrap(be.h, self.h, keys.h, _ptr(x), C.c_int64(bt), C.c_double(scale), _ptr(out),
                                      C.byref(ol), C.byref(osc)))
        assert ol.value == self.total_limbs - 14
        return out, osc.value

    def bootstrap_phase_debug(self, keys, x, scale, stop_after):
        """The pipeline stopped after ModRaise (1), CoeffToSlot (2) or EvalMod (3): (ciphertexts
        [count, 2, limbs, n], scale); count = batch (1) or 2 * batch (2, 3: real halves then imaginary halves)."""
        be = self.be
        bt, p, l, n = x.shape
        buf = be.empty(2 * bt, 2, self.total_limbs, n)
        oc, ol, osc = C.c_int64(), C.c_int32(), C.c_double()
        be._chk(be.lib.moai_bootstrap_phase_debug(be.h, self.h, keys.h, _ptr(x), C.c_int64(bt), C.c_double(scale),
                                                  C.c_int32(stop_after), _ptr(buf), C.byref(oc), C.byref(ol),
                                                  C.byref(osc)))
        flat = buf.reshape(-1)[: oc.value * 2 * ol.value * n]
        return flat.reshape(oc.value, 2, ol.value, n), osc.value

    def bootstrap_real(self, keys, x, scale, chunk_pairs=32):
        """Real-slot messages, two per bootstrapping (moai_bootstrap_real): x [batch, 2, 1, n] ->
        ([batch, 2, total_limbs - 14, n], final_scale) with ceil(batch / 2) bootstrappings."""
        be = self.be
        bt, p, l, n = x.shape
        out = be.empty(bt, 2, self.total_limbs - 14, n)
        ol, osc = C.c_int32(), C.c_double()
        be._chk(be.lib.moai_bootstrap_real(be.h, self.h, keys.h, _ptr(x), C.c_int64(bt), C.c_double(scale),
                                           C.c_int64(chunk_pairs), _ptr(out), C.byref(ol), C.byref(osc)))
        assert ol.value == self.total_limbs - 14
        return out, osc.value

    def softmax_boot(self, keys, x, scale, bias_vec, input_num, iters=16, layer_id=0):
        be = self.be
        bt, p, l, n = x.shape
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        return be._module(be.lib.moai_softmax_boot, x, scale, be.h, keys.h, self.h, _ptr(x), C.c_int32(bt), C.c_int32(l),
                          C.c_double(scale), bv.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(input_num),
                          C.c_int32(iters), C.c_int32(layer_id))

    def single_att_block(self, keys, x, scale, WQ, WK, WV, bQ, bK, bV, bias_vec, input_num, num_batch, iters=16,
                         layer_id=0):
        be = self.be
        bt, p, l, n = x.shape
        dp = C.POINTER(C.c_double)
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (WQ, WK, WV, bQ, bK, bV)]
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        col_W = arrs[3].size
        return be._module(be.lib.moai_single_att_block, x, scale, be.h, keys.h, self.h, _ptr(x), C.c_int32(bt),
                          C.c_int32(l), C.c_double(scale), *[a.ctypes.data_as(dp) for a in arrs], C.c_int32(col_W),
                          bv.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(input_num), C.c_int32(num_batch),
                          C.c_int32(iters), C.c_int32(layer_id), out_count=col_W)

    def encoder_layer(self, keys, x, scale, weights, bias_vec, input_num, num_batch, layer_id=0, boot_chunk=32,
                      inplace=False):
        """weights: dict with the fields of moai_layer_weights (numpy float64 arrays).  inplace=True
        writes the layer output over x (same shape; saves one 15.75 GiB buffer at the repo's size)."""
        be = self.be
        bt, p, l, n = x.shape
        dp = C.POINTER(C.c_double)
        keep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in weights.items()
                if k not in ("hidden", "heads", "head_dim", "inter")}
        w = LayerWeightsC(hidden=weights["hidden"], heads=weights["heads"], head_dim=weights["head_dim"],
                          inter=weights["inter"], **{k: v.ctypes.data_as(dp) for k, v in keep.items()})
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        return be._module(be.lib.moai_encoder_layer, x, scale, be.h, keys.h, self.h, _ptr(x), C.c_int32(l),
                          C.c_double(scale), C.byref(w), bv.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(input_num),
                          C.c_int32(num_batch), C.c_int32(layer_id), C.c_int64(boot_chunk),
                          out=x if inplace else None)

    def layer_weights(self, weights):
        """dict with the fields of moai_layer_weights -> (C struct, keep-alive list)."""
        dp = C.POINTER(C.c_double)
        keep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in weights.items()
                if k not in ("hidden", "heads", "head_dim", "inter")}
        w = LayerWeightsC(hidden=weights["hidden"], heads=weights["heads"], head_dim=weights["head_dim"],
                          inter=weights["inter"], **{k: v.ctypes.data_as(dp) for k, v in keep.items()})
        return w, keep

    def encoder_layer_stage(self, keys, stage, x, aux, scale, cweights, bias_vec, input_num, num_batch, layer_id=0,
                            boot_chunk=32):
        """One bootstrap-delimited quarter of the encoder layer, in place on the two persistent buffers x and aux
        (moai_encoder_layer_stage).  cweights: the struct returned by layer_weights()."""
        be = self.be
        bt, p, l, n = x.shape
        bv = np.ascontiguousarray(bias_vec, dtype=np.int32)
        be._chk(be.lib.moai_encoder_layer_stage(be.h, keys.h, self.h, C.c_int32(stage), _ptr(x), _ptr(aux), C.c_int32(l),
                                                C.c_double(scale), C.byref(cweights),
                                                bv.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(input_num),
                                                C.c_int32(num_batch), C.c_int32(layer_id), C.c_int64(boot_chunk)))

    def __del__(self):
        try:
            self.be.lib.moai_bootstrapper_destroy(self.h)
        except Exception:
            pass


class GroupedKey:
    """A grouped-digit key-switching key on the device (csrc/ksgroup.hpp)."""

    def __init__(self, t, k_extra, max_limbs):
        self.t, self.k_extra, self.max_limbs = t, k_extra, max_limbs


class _KeysHandle:
    def __init__(self, lib, h, keep):
        self.lib, self.h, self.keep = lib, h, keep

    def __del__(self):
        try:
            self.lib.moai_keys_destroy(self.h)
        except Exception:
            pass
