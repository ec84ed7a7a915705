"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's marginal table builder.

Follows (paths relative to /root/reference):
  mg94_p      src/lib/mutation_coati.cc:49-125   (Yang-94 nucleotide rates or GTR, MG94 codon Q)
  gtr_q       src/lib/mutation_coati.cc:317-354
  ecm_p       src/lib/mutation_ecm.cc:151-184    (data: src/include/coati/ecm_unrest.tcc)
  marginal_p  src/lib/mutation_coati.cc:164-202
  ambiguous_sum_p / ambiguous_best_p             src/lib/mutation_coati.cc:234-306

The matrix exponential lives in an un-vendored third-party dependency of the reference:
Eigen 3.4.0 ``unsupported/Eigen/MatrixFunctions`` (``MatrixBase::exp()``, float Pade +
scaling-and-squaring; call sites mutation_coati.cc:122, mutation_ecm.cc:181, io.cc:85).
Eigen is absent from this image, so P = expm(Q t) is evaluated here with scipy in float64 and
rounded to float32.  PARITY OF THE TABLE VALUES IS THEREFORE PINNED ONLY TO ~1e-5 relative
(the reference's own ``mg94_p`` test tolerance against the golden ``mg94P``); the dynamic
program is pinned bit-for-bit *given* table bytes, which is why every DP test feeds the same
bytes to oracle, reference and GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_libm = C.CDLL("libm.so.6")
_libm.logf.restype = C.c_float
_libm.logf.argtypes = [C.c_float]

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# utils.hpp:66-70 amino_group: ASCII amino-acid letter of each of the 61 sense codons
AMINO_GROUP = np.frombuffer(
    b"KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVYYSSSSCWCLFLF", dtype=np.uint8)

YANG94_Q = np.array([[-0.818, 0.132, 0.586, 0.1],      # mutation_coati.cc:66-69
                     [0.221, -1.349, 0.231, 0.897],
                     [0.909, 0.215, -1.322, 0.198],
                     [0.1, 0.537, 0.128, -0.765]], dtype=np.float32)

DEFAULT_PI = (0.308, 0.185, 0.199, 0.308)               # structs.hpp:74-75


def cod61_to_64(c: int) -> int:                          # utils.cc:1195-1211
    if c < 48:
        return c
    if c == 48:
        return 49
    if c < 54:
        return c + 2
    return c + 3


def get_nuc(cod61: int, pos: int) -> int:                # utils.cc:738-749
    return (cod61_to_64(cod61) >> (4 - 2 * pos)) & 3


def gtr_q(pi, sigma):                                    # mutation_coati.cc:317-354
    if any(s < 0 or s > 1 for s in sigma):
        raise ValueError("Sigma values must be in range [0,1].")
    q = np.zeros((4, 4), dtype=np.float32)
    idx = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for (i, j), s in zip(idx, sigma):
        q[i, j] = q[j, i] = np.float32(s)
    q *= np.asarray(pi, dtype=np.float32)[None, :]
    for i in range(4):
        q[i, i] = 0
        q[i, i] = -q[i].sum(dtype=np.float32)
    return q


def _expm_f32(Q, scale):
    from scipy.linalg import expm
    return expm(Q.astype(np.float64) * float(scale)).astype(np.float32)


def mg94_q(omega, pi, sigma=(0,) * 6):
    """Un-normalised MG94 Q and the normaliser d (mutation_coati.cc:58-118)."""
    pi = np.asarray(pi, dtype=np.float32)
    nuc_q = gtr_q(pi, sigma) if any(s > 0 for s in sigma) else YANG94_Q
    Q = np.zeros((61, 61), dtype=np.float32)
    d = np.float32(0)
    for i in range(61):
        ni = [get_nuc(i, p) for p in range(3)]
        Pi_i = np.float32(pi[ni[0]] * pi[ni[1]]) * pi[ni[2]]
        row = np.float32(0)
        for j in range(61):
            nj = [get_nuc(j, p) for p in range(3)]
            diff = [p for p in range(3) if ni[p] != nj[p]]
            if len(diff) == 1:
                w = np.float32(1.0) if AMINO_GROUP[i] == AMINO_GROUP[j] else np.float32(omega)
                p = diff[0]
                Q[i, j] = w * nuc_q[ni[p], nj[p]]
            row = np.float32(row + Q[i, j])
        Q[i, i] = -row
        d = np.float32(d + np.float32(Pi_i * row))
    return Q, d


def mg94_p(br_len=0.0133, omega=0.2, pi=DEFAULT_PI, sigma=(0,) * 6):
    if br_len <= 0:
        raise ValueError("Branch length must be positive.")
    Q, d = mg94_q(omega, pi, sigma)
    return _expm_f32(Q, np.float32(br_len) / d)


def load_ecm():
    z = np.load(os.path.join(_GOLDEN, "ecm_unrest.npz"))
    return z["exchang"], z["ecm_pi"]


def ecm_p(br_len=0.0133, omega=0.2):                     # mutation_ecm.cc:151-184
    if br_len <= 0:
        raise ValueError("Branch length must be positive.")
    exchang, ecm_pi = load_ecm()
    Q = np.zeros((61, 61), dtype=np.float32)
    d = np.float32(0)
    for i in range(61):
        row = np.float32(0)
        for j in range(61):
            if i == j:
                continue
            v = np.float32(exchang[i, j] * ecm_pi[j])      # k(i, j, 0) == 1
            if AMINO_GROUP[i] != AMINO_GROUP[j]:
                v = np.float32(v * np.float32(omega))
            Q[i, j] = v
            row = np.float32(row + v)
        Q[i, i] = -row
        d = np.float32(d + np.float32(ecm_pi[i] * row))
    return _expm_f32(Q, np.float32(br_len) / d)


def _lse(a, b):
    from . import lib
    return np.float32(lib.orc_log_sum_exp(C.c_float(a), C.c_float(b)))


def marginal_p(P, pi=DEFAULT_PI, amb="SUM", msub="SUM"):
    """183x15 float32 log-odds table (mutation_coati.cc:164-306)."""
    P = np.asarray(P, dtype=np.float32)
    pi = np.asarray(pi, dtype=np.float32)
    p = np.zeros((183, 15), dtype=np.float32)
    nuc_of = np.array([[get_nuc(i, pos) for pos in range(3)] for i in range(61)])
    for cod in range(61):
        for nuc in range(4):
            for pos in range(3):
                marg = np.float32(0)
                for i in range(61):
                    v = P[cod, i] if nuc_of[i, pos] == nuc else np.float32(0)
                    if msub == "SUM":
                        marg = np.float32(marg + v)
                    elif v > marg:
                        marg = v
                p[cod * 3 + pos, nuc] = _libm.logf(np.float32(marg / pi[nuc]))
    groups = [(0, 2), (1, 3), (0, 1), (2, 3), (1, 2), (0, 3),          # R Y M K S W
              (1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2), (0, 1, 2, 3)]  # B D H V N
    for row in range(183):
        for col, grp in enumerate(groups, start=4):
            acc = p[row, grp[0]]
            for x in grp[1:]:
                acc = _lse(acc, p[row, x]) if amb == "SUM" else max(acc, p[row, x])
            p[row, col] = acc
    return p


def build_table(model="mar-mg", br_len=0.0133, omega=0.2, pi=DEFAULT_PI, amb="SUM", msub="SUM"):
    """set_subst (utils.cc:595-618), marginal models only.  NB mar-ecm marginalises with the
    caller's pi (MG94 default), not ecm_pi, and sigma never reaches this path (utils.cc:603-606)."""
    if model == "mar-mg":
        P = mg94_p(br_len, omega, pi)
    elif model == "mar-ecm":
        P = ecm_p(br_len, omega)
    else:
        raise ValueError("Mutation model unknown.")
    return marginal_p(P, pi, amb, msub)
