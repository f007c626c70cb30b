"""Independent plain-math bn256 / bls12_381 arithmetic (python ints) for test-input generation and
expected values: G1/G2 group law, scalar multiplication, and optimal-ate pairings computed the
textbook way (affine line functions, final exponentiation as a plain power). Shares no code and no
formulas with the oracle's or the product's in-circuit tower."""


class Curve:
    def __init__(self, name, p, r, b, xi, g1, g2, x, x_is_neg, twist):
        self.name, self.p, self.r, self.b, self.xi, self.g1, self.g2 = name, p, r, b, xi, g1, g2
        self.x, self.x_is_neg, self.twist = x, x_is_neg, twist  # twist: 'D' (bn256) or 'M' (bls12_381)

    # ---- Fp2 = Fp[u]/(u^2+1), elements (a0, a1) ----
    def f2add(self, a, b): return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)
    def f2sub(self, a, b): return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)
    def f2neg(self, a): return ((-a[0]) % self.p, (-a[1]) % self.p)
    def f2mul(self, a, b): return ((a[0] * b[0] - a[1] * b[1]) % self.p, (a[0] * b[1] + a[1] * b[0]) % self.p)
    def f2inv(self, a):
        n = pow(a[0] * a[0] + a[1] * a[1], -1, self.p)
        return (a[0] * n % self.p, (-a[1]) * n % self.p)
    def f2scal(self, a, k): return (a[0] * k % self.p, a[1] * k % self.p)

    # ---- generic affine group law over Fp (g=1) or Fp2 (g=2); None = identity ----
    def add(self, P, Q, g):
        if P is None: return Q
        if Q is None: return P
        if g == 1:
            p = self.p
            if P[0] == Q[0]:
                if (P[1] + Q[1]) % p == 0: return None
                lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, p) % p
            else:
                lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
            x = (lam * lam - P[0] - Q[0]) % p
            return (x, (lam * (P[0] - x) - P[1]) % p)
        if P[0] == Q[0]:
            if self.f2add(P[1], Q[1]) == (0, 0): return None
            lam = self.f2mul(self.f2scal(self.f2mul(P[0], P[0]), 3), self.f2inv(self.f2scal(P[1], 2)))
        else:
            lam = self.f2mul(self.f2sub(Q[1], P[1]), self.f2inv(self.f2sub(Q[0], P[0])))
        x = self.f2sub(self.f2sub(self.f2mul(lam, lam), P[0]), Q[0])
        return (x, self.f2sub(self.f2mul(lam, self.f2sub(P[0], x)), P[1]))

    def neg(self, P, g):
        if P is None: return None
        return (P[0], (-P[1]) % self.p) if g == 1 else (P[0], self.f2neg(P[1]))

    def mul(self, P, k, g):
        R = None
        for bit in bin(k)[2:] if k else "":
            R = self.add(R, R, g)
            if bit == "1": R = self.add(R, P, g)
        return R

    def on_curve(self, P, g):
        if g == 1:
            return (P[1] * P[1] - P[0] ** 3 - self.b) % self.p == 0
        b2 = self.f2mul((self.b, 0), self.f2inv(self.xi)) if self.twist == "D" else self.f2mul((self.b, 0), self.xi)
        return self.f2sub(self.f2mul(P[1], P[1]), self.f2add(self.f2mul(self.f2mul(P[0], P[0]), P[0]), b2)) == (0, 0)

    # ---- Fp12 as Fp2[w]/(w^6 - xi): list of 6 Fp2 coefficients ----
    def f12one(self): return [(1, 0)] + [(0, 0)] * 5
    def f12mul(self, a, b):
        t = [(0, 0)] * 11
        for i in range(6):
            if a[i] == (0, 0): continue
            for j in range(6):
                if b[j] == (0, 0): continue
                t[i + j] = self.f2add(t[i + j], self.f2mul(a[i], b[j]))
        for k in range(10, 5, -1):
            t[k - 6] = self.f2add(t[k - 6], self.f2mul(t[k], self.xi))
        return t[:6]
    def f12pow(self, a, e):
        r = self.f12one()
        for bit in bin(e)[2:]:
            r = self.f12mul(r, r)
            if bit == "1": r = self.f12mul(r, a)
        return r
    def f12conj(self, a):  # w -> -w  (the p^6 Frobenius)
        return [a[i] if i % 2 == 0 else self.f2neg(a[i]) for i in range(6)]

    # Line through T and Q' (or tangent at T) on the twist, evaluated at P in G1, as a sparse Fp12.
    # Untwist: D-type psi(x', y') = (x' w^2, y' w^3); M-type psi(x', y') = (x' / w^2, y' / w^3).
    def _line(self, T, Q, P):
        if T[0] == Q[0] and T[1] == Q[1]:
            lam = self.f2mul(self.f2scal(self.f2mul(T[0], T[0]), 3), self.f2inv(self.f2scal(T[1], 2)))
        else:
            lam = self.f2mul(self.f2sub(Q[1], T[1]), self.f2inv(self.f2sub(Q[0], T[0])))
        xp, yp = P
        c = self.f2sub(self.f2mul(lam, T[0]), T[1])  # lam*xT - yT
        out = [(0, 0)] * 6
        if self.twist == "D":
            # l = yP - lam*xP*w + (lam*xT - yT)*w^3
            out[0] = (yp, 0)
            out[1] = self.f2neg(self.f2scal(lam, xp))
            out[3] = c
        else:
            # multiply the M-type line by w^3 (killed by the final exponentiation):
            # l = yP*w^3 - lam*xP*w^2 + (lam*xT - yT)    with w^6 = xi handled by f12mul
            out[3] = (yp, 0)
            out[2] = self.f2neg(self.f2scal(lam, xp))
            out[0] = c
        return out, lam

    def _frob_twist(self, Q, n=1):
        # p-power Frobenius on the D-type twist: (x^p * xi^((p-1)/3), y^p * xi^((p-1)/2))
        for _ in range(n):
            cx = self._f2pow(self.xi, (self.p - 1) // 3)
            cy = self._f2pow(self.xi, (self.p - 1) // 2)
            Q = (self.f2mul((Q[0][0], (-Q[0][1]) % self.p), cx), self.f2mul((Q[1][0], (-Q[1][1]) % self.p), cy))
        return Q

    def _f2pow(self, a, e):
        r = (1, 0)
        for bit in bin(e)[2:]:
            r = self.f2mul(r, r)
            if bit == "1": r = self.f2mul(r, a)
        return r

    def miller(self, P, Q):
        loop = 6 * self.x + 2 if self.name == "bn256" else self.x
        f = self.f12one()
        T = Q
        for bit in bin(loop)[3:]:
            l, _ = self._line(T, T, P)
            f = self.f12mul(self.f12mul(f, f), l)
            T = self.add(T, T, 2)
            if bit == "1":
                l, _ = self._line(T, Q, P)
                f = self.f12mul(f, l)
                T = self.add(T, Q, 2)
        if self.name == "bn256":
            Q1 = self._frob_twist(Q, 1)
            Q2 = self.neg(self._frob_twist(Q, 2), 2)
            l, _ = self._line(T, Q1, P)
            f = self.f12mul(f, l)
            T = self.add(T, Q1, 2)
            l, _ = self._line(T, Q2, P)
            f = self.f12mul(f, l)
        if self.x_is_neg:
            f = self.f12conj(f)
        return f

    def final_exp(self, f, extra=1):
        return self.f12pow(f, (self.p ** 12 - 1) // self.r * extra)

    def pairing(self, P, Q, extra=1):
        return self.final_exp(self.miller(P, Q), extra)

    def to_tower(self, f):
        """Fp2[w]/(w^6 - xi) coefficients -> the reference's ((c0.c0,c0.c1,c0.c2),(c1.c0,c1.c1,c1.c2))
        ordering with Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - xi): w^(2i+j) <-> c_j . v^i"""
        return [f[0], f[2], f[4], f[1], f[3], f[5]]


BN256 = Curve(
    "bn256",
    p=0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    r=0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    b=3, xi=(9, 1), g1=(1, 2),
    g2=((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    x=4965661367192848881, x_is_neg=False, twist="D")

BLS12_381 = Curve(
    "bls12_381",
    p=0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    r=0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    b=4, xi=(1, 1),
    g1=(0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
        0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1),
    g2=((0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
         0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
        (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
         0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)),
    x=0xD201000000010000, x_is_neg=True, twist="M")


def splitmix64(seed):
    """scalars from SplitMix64 -> 32 bytes -> mod r."""
    state = seed & 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    return nxt


def scalar_stream(seed, modulus):
    g = splitmix64(seed)
    while True:
        v = g() | (g() << 64) | (g() << 128) | (g() << 192)
        yield v % modulus
