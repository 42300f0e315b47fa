import ctypes as C, random, sys
sys.path.insert(0,'/root/repo')
from oracle import plonk_oracle as po
import os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
subprocess.run(['/usr/bin/g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-o', '/tmp/f29.so', os.path.join(HERE, 'field29_shim.cpp')], check=True)
lib = C.CDLL('/tmp/f29.so')
P = {0: po.BN254.r, 1: po.BN254.p, 2: po.BLS12_381.r, 3: po.BLS12_381.p}
NW = {0: 8, 1: 8, 2: 8, 3: 12}
LW = {0: 261, 1: 261, 2: 261, 3: 392}
def words(x, n): return (C.c_uint32 * n)(*[(x >> (32*i)) & 0xffffffff for i in range(n)])
def val(w): return sum(int(v) << (32*i) for i, v in enumerate(w))
def op(f, o, a, b=0, c=0, d=0):
    n = NW[f]; out = (C.c_uint32 * n)()
    lib.h29_field_op(f, o, words(a, n), words(b, n), words(c, n), words(d, n), out)
    return val(out)
rng = random.Random(1)
for f in range(4):
    p = P[f]; Rp = 1 << LW[f]; Ri = pow(Rp, -1, p); R = 1 << (32 * NW[f])
    special = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (1 << 29) - 1, 1 << 29, (1 << 232) - 1, p >> 1]
    for it in range(3000):
        pick = lambda: rng.choice(special) if rng.random() < 0.25 else rng.randrange(p)
        a, b, c, d = pick(), pick(), pick(), pick()
        assert op(f, 0, a, b) == a * b * Ri % p, (f, 'mul')
        assert op(f, 1, a) == a * a * Ri % p, (f, 'sqr')
        assert op(f, 2, a, b) == (a + b) % p
        assert op(f, 3, a, b) == (a - b) % p
        assert op(f, 4, a, b) == (a - b) % p
        assert op(f, 5, a, b, c, d) == (a * b + c * d) * Ri % p, (f, 'mul_add')
        assert op(f, 6, a) == (-a) % p
        assert op(f, 7, a) == a, (f, 'roundtrip')
        assert op(f, 8, a) == a * Rp * pow(R, -1, p) % p
        assert op(f, 9, a, b, c, d) == ((a - b) + (c - d)) ** 2 * Ri % p, (f, 'lazy')
        assert op(f, 11, a, b, c, d) == (4 * a + 2 * b + c + d) % p
    # zero test on multiples of p (as plain integers) and near misses
    for k in range(0, 12):
        if k * p < (1 << (32 * NW[f])):
            assert op(f, 10, k * p) == 1, (f, k)
            if k: assert op(f, 10, k * p + 1) == 0 and op(f, 10, k * p - 1) == 0
    print('field', f, 'ok')
# curve: madd in the reduced-radix domain equals the 32-bit-limb formulas coordinate by coordinate
for curve, cv in ((0, po.BN254), (1, po.BLS12_381)):
    nw = 8 if curve == 0 else 12
    R = 1 << (32 * nw)
    def aff(Pt):
        if Pt is None: return (C.c_uint32 * (2*nw))()
        return (C.c_uint32 * (2*nw))(*(list(words(Pt[0]*R % cv.p, nw)) + list(words(Pt[1]*R % cv.p, nw))))
    def xyzz_from(Pt, z):
        # (x z^2, y z^3, z^2, z^3)
        if Pt is None: return (C.c_uint32 * (4*nw))()
        zz, zzz = z*z % cv.p, z*z*z % cv.p
        vals = [Pt[0]*zz % cv.p, Pt[1]*zzz % cv.p, zz, zzz]
        out = []
        for v in vals: out += list(words(v*R % cv.p, nw))
        return (C.c_uint32 * (4*nw))(*out)
    G = cv.g1
    pts = [po.g1_mul(cv, G, k) for k in (1, 2, 3, 5, 7, 1000003)]
    cases = []
    for A in pts + [None]:
        for B in pts + [None]:
            for neg in (0, 1):
                cases.append((A, B, neg))
    for A, B, neg in cases:
        for z in (1, 5, rng.randrange(1, cv.p)):
            acc = xyzz_from(A, z); pt = aff(B)
            o1 = (C.c_uint32 * (4*nw))(); o2 = (C.c_uint32 * (4*nw))()
            lib.h29_madd(curve, acc, pt, neg, o1); lib.h29_madd_ref(curve, acc, pt, neg, o2)
            assert list(o1) == list(o2), (curve, A, B, neg, z)
    print('curve', curve, 'madd ok', len(cases)*3)
