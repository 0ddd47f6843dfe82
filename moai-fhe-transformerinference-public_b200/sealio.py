"""SEAL wire format for ciphertexts: the bytes `seal::Ciphertext::save` writes with
`compr_mode_type::none` and `Ciphertext::load` accepts, so that encrypted inputs, outputs and per-layer
checkpoints of the device pipeline interoperate with a stock SEAL 4.1 client.

Layout (little endian), restated from the reference:
  SEALHeader            S/serialization.h:76-93      magic 0xA15E, header size 0x10, version 4.1,
                                                     compr_mode 0, reserved, total size in bytes
  Ciphertext members    S/ciphertext.cpp:190-234     parms_id (4 x u64), is_ntt_form (u8), size, poly_modulus_degree,
                                                     coeff_modulus_size (u64 each), scale (f64), correction_factor (u64)
  DynArray<u64>         S/dynarray.h:560-581         its own SEALHeader, element count (u64), the residues
                                                     [size][coeff_modulus_size][N]
parms_id = BLAKE2b-256 over (scheme, N, primes of the level..., plain_modulus) as u64 words
(S/encryptionparams.cpp:124-158, S/util/hash.h:30-37).  Pure host code (numpy + hashlib), no device work:
the residues are exactly what `moai_memcpy_d2h` returns.
"""
import hashlib
import struct

import numpy as np

SEAL_MAGIC = 0xA15E
SEAL_HEADER_SIZE = 0x10
VERSION_MAJOR, VERSION_MINOR = 4, 1
SCHEME_CKKS = 2
_HDR = struct.Struct("<HBBBBHQ")


def parms_id(poly_modulus_degree, level_primes, scheme=SCHEME_CKKS, plain_modulus=0):
    """parms_id of the parameter set whose coeff_modulus is `level_primes` (data level k: q_0..q_k; key level:
    all primes including the special one)."""
    words = [scheme, poly_modulus_degree] + [int(p) for p in level_primes] + [plain_modulus]
    h = hashlib.blake2b(struct.pack("<%dQ" % len(words), *words), digest_size=32).digest()
    return struct.unpack("<4Q", h)


def _header(total_size):
    return _HDR.pack(SEAL_MAGIC, SEAL_HEADER_SIZE, VERSION_MAJOR, VERSION_MINOR, 0, 0, total_size)


def save_ciphertext(residues, level_primes, scale, is_ntt_form=True, correction_factor=1):
    """residues: uint64 array [size, limbs, N] (canonical, SEAL's layout) -> bytes of Ciphertext::save."""
    r = np.ascontiguousarray(residues, dtype=np.uint64)
    size, limbs, n = r.shape
    if limbs != len(level_primes):
        raise ValueError("one prime per limb expected")
    data = r.tobytes()
    dyn = _header(SEAL_HEADER_SIZE + 8 + len(data)) + struct.pack("<Q", r.size) + data
    members = struct.pack("<4Q", *parms_id(n, level_primes)) + struct.pack("<B", 1 if is_ntt_form else 0) + \
        struct.pack("<QQQ", size, n, limbs) + struct.pack("<d", scale) + struct.pack("<Q", correction_factor) + dyn
    return _header(SEAL_HEADER_SIZE + len(members)) + members


def load_ciphertext(blob, chain_primes=None):
    """bytes of Ciphertext::save (uncompressed) -> dict(residues [size, limbs, N], scale, is_ntt_form,
    parms_id, correction_factor).  With `chain_primes` (the data primes q_0.. in chain order) the parms_id is
    checked against the level the limb count implies, as SEAL's is_metadata_valid_for does."""
    def header(off):
        magic, hsize, vmaj, vmin, compr, _, total = _HDR.unpack_from(blob, off)
        if magic != SEAL_MAGIC or hsize != SEAL_HEADER_SIZE:
            raise ValueError("not a SEAL header")
        if vmaj != VERSION_MAJOR:
            raise ValueError("unsupported SEAL version %d.%d" % (vmaj, vmin))
        if compr != 0:
            raise ValueError("compressed streams are not supported (compr_mode %d)" % compr)
        return total

    total = header(0)
    if total != len(blob):
        raise ValueError("size field does not match the buffer")
    off = SEAL_HEADER_SIZE
    pid = struct.unpack_from("<4Q", blob, off)
    off += 32
    is_ntt = blob[off] != 0
    off += 1
    size, n, limbs = struct.unpack_from("<QQQ", blob, off)
    off += 24
    (scale,) = struct.unpack_from("<d", blob, off)
    off += 8
    (corr,) = struct.unpack_from("<Q", blob, off)
    off += 8
    dyn_total = header(off)
    (count,) = struct.unpack_from("<Q", blob, off + SEAL_HEADER_SIZE)
    if count != size * n * limbs or dyn_total != SEAL_HEADER_SIZE + 8 + 8 * count:
        raise ValueError("ciphertext data is invalid")       # seeded (half-size) ciphertexts are a client-side form
    data = np.frombuffer(blob, dtype=np.uint64, count=count, offset=off + SEAL_HEADER_SIZE + 8)
    if chain_primes is not None:
        if limbs > len(chain_primes) or pid != parms_id(n, chain_primes[:limbs]):
            raise ValueError("ciphertext data is invalid")   # parms_id unknown to this context
    return {"residues": data.reshape(size, limbs, n).copy(), "scale": scale, "is_ntt_form": is_ntt,
            "parms_id": pid, "correction_factor": corr}
