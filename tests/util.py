"""Shared helpers for the parity tests: limb packing and oracle-side expected values."""
import random

import numpy as np

from oracle import zkt_oracle as O


def int_to_limbs(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def limbs_to_int(a):
    return sum(int(x) << (32 * i) for i, x in enumerate(a))


def scalars_to_array(scalars):
    """list of ints (< 2^256) -> (n, 8) uint32, little-endian limbs"""
    out = np.zeros((len(scalars), 8), dtype=np.uint32)
    for i, s in enumerate(scalars):
        out[i] = int_to_limbs(int(s), 8)
    return out


def g1_points_to_array(points):
    """oracle G1 points -> ((n, 24) uint32 canonical limbs, (n,) uint8 infinity flags)"""
    xy = np.zeros((len(points), 24), dtype=np.uint32)
    inf = np.zeros(len(points), dtype=np.uint8)
    for i, p in enumerate(points):
        if p is O.INF:
            inf[i] = 1
        else:
            xy[i] = O.g1_to_limbs(p)
    return xy, inf


def g2_points_to_array(points):
    xy = np.zeros((len(points), 48), dtype=np.uint32)
    inf = np.zeros(len(points), dtype=np.uint8)
    for i, p in enumerate(points):
        if p is O.INF:
            inf[i] = 1
        else:
            xy[i] = O.g2_to_limbs(p)
    return xy, inf


def g1_from_array(xy, is_inf):
    if is_inf:
        return O.INF
    return O.g1(limbs_to_int(xy[:12]), limbs_to_int(xy[12:24]))


def g2_from_array(xy, is_inf):
    if is_inf:
        return O.INF
    v = [limbs_to_int(xy[12 * k:12 * k + 12]) for k in range(4)]
    return O.g2(v[1], v[0], v[3], v[2])


def fast_mul(gen, k):
    """k * gen for the oracle's generators through a left-to-right ladder on the oracle's own
    affine law (same group element as O.scalar_mul; fewer Python-level additions)."""
    return O.scalar_mul(gen, k % O.R) if k % O.R else O.INF


def expected_from_dlogs(gen, dlogs, scalars):
    """sum_i s_i * (k_i * gen) == (sum_i s_i k_i mod r) * gen  -- one oracle scalar mul."""
    acc = 0
    for k, s in zip(dlogs, scalars):
        acc = (acc + int(k) * int(s)) % O.R
    return fast_mul(gen, acc)


def rand_scalars(rnd: random.Random, n, bound=None):
    bound = bound or O.R
    return [rnd.randrange(bound) for _ in range(n)]
