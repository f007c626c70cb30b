"""Seeded synthetic inputs for the whole-circuit shapes: points a_i*G, scalars from
SplitMix64, explicit blinding points r1/r2, expected results from independent plain math."""
import ecmath as em


def g2_flat(b):
    return [b[0][0], b[0][1], b[1][0], b[1][1]]


def msm_inputs(C, n, seed, identity_result=False):
    g = em.scalar_stream(seed, C.r)
    pts, sc, acc = [], [], None
    for _ in range(n):
        a, b = next(g) or 1, next(g)
        P = C.mul(C.g1, a, 1)
        pts.append(P)
        sc.append(b)
        acc = C.add(acc, C.mul(P, b, 1), 1)
    r1, r2 = C.mul(C.g1, next(g) or 1, 1), C.mul(C.g1, next(g) or 1, 1)
    inp = []
    for P in pts:
        inp += [P[0], P[1], 0]
    inp += sc
    inp += [r1[0], r1[1], r2[0], r2[1]]
    inp += [acc[0], acc[1], 0] if acc is not None else [0, 0, 1]
    return inp


def bn_check_pairing_inputs(alpha, beta):
    C = em.BN256
    a, b = C.mul(C.g1, alpha, 1), C.mul(C.g2, beta, 2)
    na = C.neg(a, 1)
    return g2_flat(b) + [na[0], na[1], 0, a[0], a[1], 0]


def bls_check_pairing_inputs(alpha, beta, c):
    C = em.BLS12_381
    a, b = C.mul(C.g1, alpha, 1), C.mul(C.g2, beta, 2)
    ac, bc, na = C.mul(a, c % C.r, 1), C.mul(b, c % C.r, 2), C.neg(a, 1)
    return g2_flat(b) + g2_flat(bc) + [na[0], na[1], 0, ac[0], ac[1], 0]
