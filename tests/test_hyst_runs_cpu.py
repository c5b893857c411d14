"""The bit tricks of hyst_local_kernel (image_b200/csrc/canny.cu), re-stated in Python and checked against
scipy's 8-connected labelling on random 32x32 tiles:
  * peeling the runs of a row mask lowest first:  low = rem & -rem;  x = rem + low;  run = rem & ~x;  rem &= x
  * the up-run that starts at bit us:  uhi = up >> us << us;  run = uhi & ~(uhi + (1 << us))
  * run_start(bits, x) = x - clz(~(bits << (31 - x))) + 1
  * linking every run to the touching runs of the row above gives exactly the 8-connected components, and a
    component is strong iff one of its runs overlaps the strong mask.
(The kernel does the unions lock-free and concurrently; the result of a union-find does not depend on the order.)"""
import numpy as np
import pytest

M = 0xFFFFFFFF


def ffs(x):
    return (x & -x).bit_length()


def clz(x):
    return 32 - x.bit_length()


def run_start(bits, x):
    return x - clz((~(bits << (31 - x))) & M) + 1


def find(lab, a):
    while lab[a] != a:
        a = lab[a]
    return a


def union(lab, a, b):
    a, b = find(lab, a), find(lab, b)
    if a != b:
        lab[max(a, b)] = min(a, b)


def label_tile(E, S):
    em = [int(sum(int(E[r, c]) << c for c in range(32))) for r in range(32)]
    sm = [int(sum(int(S[r, c]) << c for c in range(32))) for r in range(32)]
    lab, strong = list(range(1024)), [0] * 1024
    for lane in range(1, 32):
        up, rem = em[lane - 1], em[lane]
        while up and rem:
            low = rem & (-rem) & M; x = (rem + low) & M; rm = rem & ~x & M; rem &= x
            touch = up & (rm | (rm << 1) & M | rm >> 1)
            while touch:
                us = run_start(up, ffs(touch) - 1)
                uhi = (up >> us << us) & M
                touch &= ~(uhi & ~((uhi + (1 << us)) & M)) & M
                union(lab, lane * 32 + ffs(low) - 1, (lane - 1) * 32 + us)
    for lane in range(32):
        rem = em[lane]
        while rem:
            low = rem & (-rem) & M; x = (rem + low) & M; rm = rem & ~x & M; rem &= x
            root = find(lab, lane * 32 + ffs(low) - 1)
            if sm[lane] & rm:
                strong[root] = 1
    roots = np.full((32, 32), -1); st = np.zeros((32, 32), bool)
    for r in range(32):
        for c in range(32):
            if (em[r] >> c) & 1:
                roots[r, c] = find(lab, r * 32 + run_start(em[r], c))
                st[r, c] = bool(strong[roots[r, c]])
    return roots, st


def test_run_peeling_and_run_start():
    rng = np.random.default_rng(0)
    for bits in [0, 1, M, 0x80000000, 0xF0F0F0F0, 0x7FFFFFFE] + [int(v) for v in rng.integers(0, 2 ** 32, 200)]:
        runs, rem = [], bits
        while rem:
            low = rem & (-rem) & M; x = (rem + low) & M; runs.append(rem & ~x & M); rem &= x
        want, b = [], 0
        while b < 32:                                   # reference: scan the bits
            if (bits >> b) & 1:
                e = b
                while e + 1 < 32 and (bits >> (e + 1)) & 1:
                    e += 1
                want.append(((1 << (e - b + 1)) - 1) << b); b = e + 1
            else:
                b += 1
        assert runs == want
        for m in want:
            s = ffs(m) - 1
            for xx in range(s, s + bin(m).count("1")):
                assert run_start(bits, xx) == s


@pytest.mark.parametrize("seed", range(6))
def test_tile_components_equal_8_connected_labelling(seed):
    ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(seed)
    for _ in range(25):
        E = rng.random((32, 32)) < rng.uniform(0.03, 0.75)
        S = E & (rng.random((32, 32)) < 0.04)
        roots, st = label_tile(E, S)
        lab, n = ndimage.label(E, structure=np.ones((3, 3)))
        assert len(set(roots[E].tolist())) == n
        for k in range(1, n + 1):
            sel = lab == k
            assert len(set(roots[sel].tolist())) == 1
            assert st[sel].all() == st[sel].any() == bool(S[sel].any())
        assert len(set(roots[E].tolist())) <= 256       # HYST_MAX_ROOTS: components are at least two pixels apart
    iso = np.zeros((32, 32), bool); iso[::2, ::2] = True    # the extreme case: 256 isolated pixels
    roots, _ = label_tile(iso, iso)
    assert len(set(roots[iso].tolist())) == 256


def test_class_bytes_to_mask_bits():
    """One multiply turns the per-byte flags (bits 0, 8, 16, 24) of a word of class bytes into 4 adjacent bits."""
    import itertools
    for b in itertools.product((0, 1, 2), repeat=4):
        w = b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24
        s1 = (w >> 1) & 0x01010101
        e1 = (w & 0x01010101) | s1
        edge = ((e1 * 0x01020408) & M) >> 24
        strong = ((s1 * 0x01020408) & M) >> 24
        assert edge == sum((1 << k) for k in range(4) if b[k] != 0)
        assert strong == sum((1 << k) for k in range(4) if b[k] == 2)
