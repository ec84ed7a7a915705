"""TEST INFRASTRUCTURE ONLY -- ctypes front-end to the CPU oracle.

``oracle.lib``  = liboracle.so, the plain-C restatement (oracle/coati_oracle.c)
``oracle.ref``  = _ref/libcoati_ref.so, the UNMODIFIED reference hot path behind our shim
                  (oracle/ref_shim.cc); ``None`` when it has not been built.

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
bench.py may import this package; nothing under coati_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_u8p = C.POINTER(C.c_uint8)
_c_fp = C.POINTER(C.c_float)


def build(quiet: bool = True) -> None:
    """Build liboracle.so and (when /root/reference is present) _ref/libcoati_ref.so."""
    subprocess.run(["make", "-C", _HERE, "-j8", "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


if not os.path.exists(os.path.join(_HERE, "liboracle.so")):
    build()
lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
ref = _load(os.path.join(_HERE, "_ref", "libcoati_ref.so"))

lib.orc_log1p_exp.restype = C.c_float
lib.orc_log1p_exp.argtypes = [C.c_float]
lib.orc_log_sum_exp.restype = C.c_float
lib.orc_log_sum_exp.argtypes = [C.c_float, C.c_float]
lib.orc_rng_bits.restype = C.c_uint64
lib.orc_rng_f24.restype = C.c_float
lib.orc_fnv1.restype = C.c_uint32
lib.orc_end_stop_gap_score.restype = C.c_float
lib.orc_end_stop_gap_score.argtypes = [C.c_float, C.c_float]
if ref is not None:
    ref.coati_ref_rng_bits.restype = C.c_uint64
    ref.coati_ref_rng_f24.restype = C.c_float
    ref.coati_ref_viterbi_batch.restype = C.c_double

DEFAULT_G = np.float32(0.001)
DEFAULT_E = np.float32(1.0) - np.float32(1.0) / np.float32(6.0)  # structs.hpp:41


def _u8(x):
    return np.ascontiguousarray(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray))
                                else np.asarray(x, dtype=np.uint8))


def _ptr(arr, ty):
    return arr.ctypes.data_as(ty)


def encode_pair(anc: str, des: str):
    """marginal_seq_encoding (utils.cc:496-528).  Raises ValueError like the reference throws."""
    ab = anc.encode()
    out_a = np.zeros(len(ab), dtype=np.uint8)
    rc = lib.orc_encode_anc(ab, C.c_size_t(len(ab)), _ptr(out_a, _c_u8p))
    if rc == -1:
        raise ValueError("Ambiguous nucleotides in ancestor/reference.")
    if rc == -2:
        raise ValueError("Early stop codon in ancestor/reference.")
    if rc != 0:
        raise ValueError("Length of reference sequence must be multiple of 3.")
    db = des.encode()
    out_b = np.zeros(len(db), dtype=np.uint8)
    lib.orc_encode_des(db, C.c_size_t(len(db)), _ptr(out_b, _c_u8p))
    return out_a, out_b


def trim_end_stop(seq: str):
    """trim_end_stops for one sequence (utils.cc:945-967): returns (trimmed, stop-codon-or-'')."""
    b = seq.encode()
    if lib.orc_has_end_stop(b, C.c_size_t(len(b))):
        return seq[:-3], seq[-3:]
    return seq, ""


def restore_end_stops(sa: str, sb: str, score: float, stops, g=DEFAULT_G, e=DEFAULT_E):
    """restore_end_stops (utils.cc:1044-1063)."""
    s0, s1 = stops
    if len(s0) == len(s1):
        return sa + s0, sb + s1, score
    gs = np.float32(lib.orc_end_stop_gap_score(C.c_float(g), C.c_float(e)))
    if not s0:
        return sa + "---", sb + s1, np.float32(np.float32(score) + gs)
    return sa + s0, sb + "---", np.float32(np.float32(score) + gs)


def _table(table):
    t = np.ascontiguousarray(table, dtype=np.float32)
    assert t.shape == (183, 15)
    return t


def fill(semiring: int, a, b, table, g=DEFAULT_G, e=DEFAULT_E, k=1, with_trans=False, impl="oracle"):
    """forward_impl (align_pair.cc:62-139).  Returns (mch, del, ins[, trans(8,...)])."""
    a, b, t = _u8(a), _u8(b), _table(table)
    shape = (len(a) + k, len(b) + k)
    mats = [np.empty(shape, dtype=np.float32) for _ in range(3)]
    trans = np.empty((8,) + shape, dtype=np.float32) if with_trans else None
    if impl == "oracle":
        tp = (_c_fp * 8)(*[_ptr(trans[i], _c_fp) for i in range(8)]) if with_trans else None
        rc = lib.orc_fill(semiring, _ptr(a, _c_u8p), C.c_size_t(len(a)), _ptr(b, _c_u8p),
                          C.c_size_t(len(b)), _ptr(t, _c_fp), C.c_float(g), C.c_float(e),
                          C.c_size_t(k), *[_ptr(m, _c_fp) for m in mats], tp)
    else:
        assert ref is not None
        if semiring == 0:
            assert not with_trans
            n = len(a) + len(b) + 1
            rc = ref.coati_ref_viterbi(_ptr(a, _c_u8p), C.c_size_t(len(a)), _ptr(b, _c_u8p),
                                       C.c_size_t(len(b)), b"A" * len(a), b"A" * len(b),
                                       _ptr(t, _c_fp), C.c_float(g), C.c_float(e), C.c_size_t(k),
                                       C.create_string_buffer(n), C.create_string_buffer(n), None,
                                       None, *[_ptr(m, _c_fp) for m in mats])
        else:
            # member order of align_pair_work_t -> (mch del ins) + our trans order
            full = np.empty((11,) + shape, dtype=np.float32)
            mp = (_c_fp * 11)(*[_ptr(full[i], _c_fp) for i in range(11)])
            rc = ref.coati_ref_forward(_ptr(a, _c_u8p), C.c_size_t(len(a)), _ptr(b, _c_u8p),
                                       C.c_size_t(len(b)), _ptr(t, _c_fp), C.c_float(g),
                                       C.c_float(e), C.c_size_t(k), mp)
            mats = [full[0], full[1], full[2]]
            trans = full[3:]  # mch_mch mch_del mch_ins del_mch del_del ins_mch ins_del ins_ins
    if rc != 0:
        raise RuntimeError(f"fill failed rc={rc}")
    return (*mats, trans) if with_trans else tuple(mats)


def viterbi(anc: str, des: str, table, g=DEFAULT_G, e=DEFAULT_E, k=1, impl="oracle", enc=None):
    """viterbi_mem + traceback_viterbi.  Returns (aligned_anc, aligned_des, float32 score)."""
    a, b = enc if enc is not None else encode_pair(anc, des)
    t = _table(table)
    n = len(a) + len(b) + 1
    oa, ob = C.create_string_buffer(n), C.create_string_buffer(n)
    ol, sc = C.c_size_t(0), C.c_float(0)
    args = [_ptr(a, _c_u8p), C.c_size_t(len(a)), _ptr(b, _c_u8p), C.c_size_t(len(b)), anc.encode(),
            des.encode(), _ptr(t, _c_fp), C.c_float(g), C.c_float(e), C.c_size_t(k), oa, ob,
            C.byref(ol), C.byref(sc)]
    if impl == "oracle":
        rc = lib.orc_viterbi(*args)
    else:
        assert ref is not None
        rc = ref.coati_ref_viterbi(*args, None, None, None)
    if rc != 0:
        raise RuntimeError(f"viterbi failed rc={rc}")
    return oa.raw[:ol.value].decode(), ob.raw[:ol.value].decode(), np.float32(sc.value)


def directions(mch, dele, ins, la, lb, g=DEFAULT_G, e=DEFAULT_E, k=1):
    d = np.empty((la + k, lb + k), dtype=np.uint8)
    rc = lib.orc_directions(_ptr(mch, _c_fp), _ptr(dele, _c_fp), _ptr(ins, _c_fp), C.c_size_t(la),
                            C.c_size_t(lb), C.c_float(g), C.c_float(e), C.c_size_t(k),
                            _ptr(d, _c_u8p))
    assert rc == 0
    return d


def seed_state(seeds):
    """string_seed_seq + Random::Seed -> raw 128-bit Lehmer state as uint64[2] = {lo, hi}."""
    st = (C.c_uint64 * 2)()
    arr = (C.c_char_p * len(seeds))(*[s.encode() for s in seeds])
    lib.orc_rng_seed_strings(arr, C.c_size_t(len(seeds)), st)
    return np.array([st[0], st[1]], dtype=np.uint64)


def ref_seed_state(seeds):
    assert ref is not None
    st = (C.c_uint64 * 2)()
    arr = (C.c_char_p * len(seeds))(*[s.encode() for s in seeds])
    ref.coati_ref_seed(arr, C.c_size_t(len(seeds)), st)
    return np.array([st[0], st[1]], dtype=np.uint64)


def sample(anc: str, des: str, table, state, n, g=DEFAULT_G, e=DEFAULT_E, k=1, impl="oracle",
           timings=None):
    """forward + n x sampleback.  Returns (list[(a, b)], float32 scores, new_state, loglik|None)."""
    a, b = encode_pair(anc, des)
    t = _table(table)
    stride = len(a) + len(b) + 1
    oa, ob = C.create_string_buffer(n * stride), C.create_string_buffer(n * stride)
    ol = (C.c_size_t * n)()
    sc = np.zeros(n, dtype=np.float32)
    st = (C.c_uint64 * 2)(int(state[0]), int(state[1]))
    ll = C.c_float(0)
    args = [_ptr(a, _c_u8p), C.c_size_t(len(a)), _ptr(b, _c_u8p), C.c_size_t(len(b)), anc.encode(),
            des.encode(), _ptr(t, _c_fp), C.c_float(g), C.c_float(e), C.c_size_t(k), st,
            C.c_size_t(n), oa, ob, ol, _ptr(sc, _c_fp)]
    if impl == "oracle":
        rc = lib.orc_sample(*args, C.byref(ll))
        loglik = np.float32(ll.value)
    else:
        assert ref is not None
        tf, ts = C.c_double(0), C.c_double(0)
        rc = ref.coati_ref_sample(*args, C.byref(tf), C.byref(ts))
        loglik = None
        if timings is not None:
            timings["fill_s"], timings["sample_s"] = tf.value, ts.value
    if rc != 0:
        raise RuntimeError(f"sample failed rc={rc}")
    out = []
    for s in range(n):
        out.append((oa.raw[s * stride:s * stride + ol[s]].decode(),
                    ob.raw[s * stride:s * stride + ol[s]].decode()))
    return out, sc, np.array([st[0], st[1]], dtype=np.uint64), loglik


def alignment_score(aln_a: str, aln_b: str, table, g=DEFAULT_G, e=DEFAULT_E, k=1):
    """alignment_score (align_marginal.cc:373-473)."""
    if len(aln_a) != len(aln_b):
        raise ValueError("For alignment scoring both sequences must have equal length.")
    t = _table(table)
    sc = C.c_float(0)
    rc = lib.orc_alignment_score(aln_a.encode(), aln_b.encode(), C.c_size_t(len(aln_a)),
                                 _ptr(t, _c_fp), C.c_float(g), C.c_float(e), C.c_size_t(k),
                                 C.byref(sc))
    if rc != 0:
        raise ValueError(f"alignment_score rc={rc}")
    return np.float32(sc.value)


def path_score(aln_a: str, aln_b: str, a, b, table, g=DEFAULT_G, e=DEFAULT_E, k=1):
    """The alignment rows re-scored through forward_impl's own transition terms and association
    (align_pair.cc:81-138, oracle/long_pair.c).  For k = 1 this must equal the Viterbi score bit for bit."""
    a, b, t = _u8(a), _u8(b), _table(table)
    if len(aln_a) != len(aln_b):
        raise ValueError("rows differ in length")
    sc = C.c_float(0)
    rc = lib.orc_path_score(aln_a.encode(), aln_b.encode(), C.c_size_t(len(aln_a)), _ptr(a, _c_u8p),
                            C.c_size_t(len(a)), _ptr(b, _c_u8p), C.c_size_t(len(b)), _ptr(t, _c_fp),
                            C.c_float(g), C.c_float(e), C.c_size_t(k), C.byref(sc))
    if rc != 0:
        raise ValueError(f"path_score: rows are not an alignment of the pair (rc={rc})")
    return np.float32(sc.value)


def viterbi_score(a, b, table, g=DEFAULT_G, e=DEFAULT_E, k=1, threads=None):
    """Score-only rolling-row Viterbi, O(Lb) memory per strip, one strip of rows per thread
    (oracle/long_pair.c): max of the adjusted terminal M, D, I (align_pair.cc:130-138, :265)."""
    a, b, t = _u8(a), _u8(b), _table(table)
    sc = C.c_float(0)
    threads = int(threads or os.cpu_count() or 1)
    rc = lib.orc_viterbi_score(_ptr(a, _c_u8p), C.c_size_t(len(a)), _ptr(b, _c_u8p), C.c_size_t(len(b)),
                               _ptr(t, _c_fp), C.c_float(g), C.c_float(e), C.c_size_t(k), C.c_int(threads),
                               C.byref(sc))
    if rc != 0:
        raise RuntimeError(f"viterbi_score failed rc={rc}")
    return np.float32(sc.value)
