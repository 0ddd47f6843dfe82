"""Test infrastructure: CKKS secret / evaluation keys, symmetric encryption and decryption computed ON THE DEVICE
through the product's own C ABI (NTT, multiply_plain) plus torch for sampling and the additions.

Why: the full-size checks (N = 65536, 36 primes, Hamming-weight-192 secret: M/test/test_full_scheme.hpp:345-368) need
~50 Galois keys of 1.23 GiB each.  Generating them with the CPU oracle / the reference's KeyGenerator and shipping
them through numpy takes tens of minutes and > 100 GB of host memory; on the device it takes seconds.  The keys are
genuine RLWE key-switching keys in SEAL's layout ([digit][2][key limb][N], S/kswitchkeys.h:335-340; digit J encrypts
p * new_key in limb J only, S/keygenerator.cpp:303-336) — only the randomness differs from SEAL's PRNG stream, so
results are checked by DECRYPTION, not against SEAL's residues (bit-exactness of every op on SEAL-generated keys is
covered at smaller sizes by the other test modules)."""
import numpy as np


class DeviceKeyGen:
    def __init__(self, pkg, be, hamming_weight=192, seed=1, sigma=3.2):
        import torch
        self.torch, self.pkg, self.be = torch, pkg, be
        self.n, self.kl = be.n, be.kl
        self.dev = be.device
        self.q_host = [int(x) for x in be.primes]
        self.q = torch.tensor(self.q_host, dtype=torch.int64, device=self.dev)
        self.sigma = sigma
        rng = np.random.default_rng(seed)
        s = np.zeros(self.n, dtype=np.int64)
        if hamming_weight:
            idx = rng.choice(self.n, hamming_weight, replace=False)
            s[idx] = rng.choice([-1, 1], hamming_weight)
        else:
            s[:] = rng.integers(-1, 2, self.n)
        self.s_coef = s
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed)
        self.s_ntt = self.small_to_ntt(torch.from_numpy(s).to(self.dev)[None], self.kl)[0]        # [kl, n]
        p = self.q_host[-1]
        self.p_const = torch.stack([torch.full((self.n,), p % ql, dtype=torch.int64, device=self.dev)
                                    for ql in self.q_host])                                        # [kl, n]

    # -- helpers ------------------------------------------------------------------------------------------------
    def small_to_ntt(self, v, limbs):
        """v: [B, n] small signed integers -> [B, limbs, n] residues in NTT form (limb l <-> prime l; with
        limbs == kl the last limb is the special prime)."""
        t = self.torch
        q = self.q[:limbs] if limbs < self.kl else self.q
        r = t.remainder(v[:, None, :], q[None, :, None]).contiguous()
        self.be.ntt_forward_(r.view(v.shape[0], 1, limbs, self.n))
        return r

    def uniform(self, batch, limbs):
        t = self.torch
        out = t.empty((batch, limbs, self.n), dtype=t.int64, device=self.dev)
        for l in range(limbs):
            ql = self.q_host[l] if limbs < self.kl or l < self.kl - 1 else self.q_host[-1]
            out[:, l, :] = t.randint(0, ql, (batch, self.n), generator=self.gen, device=self.dev, dtype=t.int64)
        return out

    def noise(self, batch, limbs):
        t = self.torch
        e = t.round(t.randn((batch, self.n), generator=self.gen, device=self.dev, dtype=t.float64) * self.sigma)
        return self.small_to_ntt(e.to(t.int64), limbs)

    def _mul_s(self, a, s_ntt):
        """a: [B, limbs, n] (.) s_ntt [limbs, n] -> [B, limbs, n]"""
        B, l, n = a.shape
        return self.be.multiply_plain(a.view(B, 1, l, n), s_ntt.contiguous()).view(B, l, n)

    # -- key-switching keys -------------------------------------------------------------------------------------
    def _ksk(self, new_key_ntt):
        """new_key_ntt: [kl, n] (NTT form at every key limb) -> key [kl - 1, 2, kl, n] in SEAL's layout."""
        t = self.torch
        D, kl, n = self.kl - 1, self.kl, self.n
        a = self.uniform(D, kl)
        c0 = self.noise(D, kl)
        c0 = t.remainder(c0 - self._mul_s(a, self.s_ntt), self.q[None, :, None])
        ps = self._mul_s(new_key_ntt[None], self.p_const)[0]                                      # p * new_key mod q_l
        for j in range(D):
            c0[j, j] = t.remainder(c0[j, j] + ps[j], self.q[j])
        return t.stack([c0, a], dim=1).contiguous()

    def relin_key(self):
        s2 = self._mul_s(self.s_ntt[None], self.s_ntt)[0]
        return self._ksk(s2)

    def galois_key(self, elt):
        """Key for the automorphism X -> X^elt: encrypts p * s(X^elt) under s (S/keygenerator.cpp:195-232)."""
        n = self.n
        i = np.arange(n, dtype=np.int64)
        j = (i * int(elt)) % (2 * n)
        sp = np.zeros(n, dtype=np.int64)
        lo = j < n
        np.add.at(sp, j[lo], self.s_coef[lo])
        np.add.at(sp, j[~lo] - n, -self.s_coef[~lo])
        sp_ntt = self.small_to_ntt(self.torch.from_numpy(sp).to(self.dev)[None], self.kl)[0]
        return self._ksk(sp_ntt)

    # -- symmetric encryption / decryption at `limbs` data limbs -------------------------------------------------
    def encrypt(self, pt):
        """pt: [B, limbs, n] NTT-form plaintext residues -> ciphertexts [B, 2, limbs, n] = (pt + e - a s, a)."""
        t = self.torch
        B, l, n = pt.shape
        a = self.uniform(B, l)
        c0 = t.remainder(pt + self.noise(B, l) - self._mul_s(a, self.s_ntt[:l]), self.q[None, :l, None])
        return t.stack([c0, a], dim=1).contiguous()

    def decrypt(self, ct):
        """[B, 2 or 3, limbs, n] -> plaintext residues [B, limbs, n] = c0 + c1 s (+ c2 s^2)."""
        t = self.torch
        B, size, l, n = ct.shape
        s = self.s_ntt[:l].contiguous()
        m = ct[:, 0] + self._mul_s(ct[:, 1].contiguous(), s)
        if size == 3:
            m = m + self._mul_s(self._mul_s(ct[:, 2].contiguous(), s), s)
        return t.remainder(m, self.q[None, :l, None])

    def decrypt_decode(self, ct, scale, oracle, limbs_keep=2):
        """Decrypt on the device, drop to `limbs_keep` limbs (exact for CKKS: S/evaluator.cpp:1513-1545), decode on
        the host with the oracle -> complex [B, n / 2]."""
        B, size, l, n = ct.shape
        keep = min(limbs_keep, l)
        m = self.decrypt(ct[:, :, :keep].contiguous())
        host = self.pkg.to_host(m)
        return np.stack([oracle.decode(host[b].reshape(-1), keep, scale) for b in range(B)])
