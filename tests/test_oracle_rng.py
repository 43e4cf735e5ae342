"""mod_random restatement (oracle) against known answers and an independent pure-Python
xorshift128 written from src/mod_random.f90:39-112; Philox against the Random123 vectors."""
import math
import os

import numpy as np

from oracle import pyoracle as po

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "forward_golden.npz"))
M32 = 0xFFFFFFFF


def py_seeds(rank):
    # init_random, src/mod_random.f90:49-52, in wrapping int32 arithmetic
    j = rank + 1
    out = []
    for i in (5551111, 453222, 4444431, 6765):
        v = (i * j ** 4 + 1000 * i * j ** 2 + i) & M32
        out.append(v - (1 << 32) if v & 0x80000000 else v)
    return out


class PyXorshift:
    def __init__(self, rank):
        self.x, self.y, self.z, self.w = [v & M32 for v in py_seeds(rank)]

    def raw(self):  # src/mod_random.f90:63-71 (ishft is a logical shift on 32 bits)
        t = (self.x ^ ((self.x << 11) & M32)) & M32
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = ((self.w ^ (self.w >> 19)) ^ (t ^ (t >> 8))) & M32
        return self.w - (1 << 32) if self.w & 0x80000000 else self.w

    def rand_u(self):
        return (float(self.raw()) + 2.0 ** 31) / 2.0 ** 32

    def rand_u2(self):
        return (float(self.raw()) + 2.0 ** 31 + 0.5) / 2.0 ** 32


def test_seeds_match_hand_derived_values():
    for r in range(4):
        assert po.rng_seeds(r) == list(GOLD["rng_seeds"][r])
        assert py_seeds(r) == list(GOLD["rng_seeds"][r])


def test_first_draws_rank0():
    assert np.array_equal(po.rng_draw(0, 0, 4), GOLD["rng_rank0_rand_u"])


def test_streams_match_python_restatement():
    for rank in (0, 1, 2, 3, 7, 19):
        g = PyXorshift(rank)
        ref = np.array([g.raw() for _ in range(5000)], dtype=np.float64)
        assert np.array_equal(po.rng_draw(rank, 4, 5000), ref)
        g = PyXorshift(rank)
        assert np.array_equal(po.rng_draw(rank, 0, 1000), np.array([g.rand_u() for _ in range(1000)]))
        g = PyXorshift(rank)
        assert np.array_equal(po.rng_draw(rank, 1, 1000), np.array([g.rand_u2() for _ in range(1000)]))


def test_ranges():
    u = po.rng_draw(3, 0, 200000)
    u2 = po.rng_draw(3, 1, 200000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert u2.min() > 0.0 and u2.max() < 1.0


def test_rand_g_and_rand_r_formulas():
    # rand_g = sqrt(-2 log v1) cos(2 pi v2), v1 first (:98-100); rand_r = sqrt(-2 log u) (:109-110)
    g = PyXorshift(2)
    ref = []
    for _ in range(500):
        v1, v2 = g.rand_u2(), g.rand_u2()
        ref.append(math.sqrt(-2.0 * math.log(v1)) * math.cos(2.0 * math.acos(-1.0) * v2))
    assert np.allclose(po.rng_draw(2, 2, 500), ref, rtol=0, atol=1e-15)
    g = PyXorshift(2)
    ref = [math.sqrt(-2.0 * math.log(g.rand_u2())) for _ in range(500)]
    assert np.allclose(po.rng_draw(2, 3, 500), ref, rtol=0, atol=1e-15)


def test_moments():
    g = po.rng_draw(1, 2, 400000)
    assert abs(g.mean()) < 0.01 and abs(g.std() - 1.0) < 0.01
    r = po.rng_draw(1, 3, 400000)
    assert abs(r.mean() - math.sqrt(math.pi / 2)) < 0.01  # Rayleigh(1)


def test_philox_random123_known_answers():
    def phx(c, k):
        return po.philox(k[0] | (k[1] << 32), *c)
    assert phx((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert phx((M32,) * 4, (M32, M32)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert phx((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
