"""Python big-int model of BLS12-381 Fr / Fq / G1 used to pin the C oracle (tests only)."""
FR_P = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FQ_P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
FR_R = (1 << 256) % FR_P
FQ_R = (1 << 384) % FQ_P
FR_RINV = pow(FR_R, -1, FR_P)
FQ_RINV = pow(FQ_R, -1, FQ_P)
GX = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
GY = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1


def aff_add(P, Q):
    """Affine add over Fq (plain ints), None = infinity."""
    if P is None: return Q
    if Q is None: return P
    x1, y1 = P; x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % FQ_P == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, FQ_P) % FQ_P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, FQ_P) % FQ_P
    x3 = (lam * lam - x1 - x2) % FQ_P
    return (x3, (lam * (x1 - x3) - y1) % FQ_P)


def aff_mul(P, k):
    R = None
    while k:
        if k & 1: R = aff_add(R, P)
        P = aff_add(P, P); k >>= 1
    return R


def aff_neg(P):
    return None if P is None else (P[0], (-P[1]) % FQ_P)


def jac_limbs_to_affine(row):
    """36 u32 limbs (Montgomery Jacobian) -> plain affine tuple or None."""
    def val(ws):
        v = 0
        for j, x in enumerate(ws): v |= int(x) << (32 * j)
        return v * FQ_RINV % FQ_P
    X, Y, Z = val(row[0:12]), val(row[12:24]), val(row[24:36])
    if Z == 0: return None
    zi = pow(Z, -1, FQ_P)
    return (X * zi * zi % FQ_P, Y * zi * zi * zi % FQ_P)


def affine_to_jac_limbs(P):
    import numpy as np
    out = np.zeros(36, dtype=np.uint32)
    if P is None:
        vals = (0, FQ_R, 0)
    else:
        vals = (P[0] * FQ_R % FQ_P, P[1] * FQ_R % FQ_P, FQ_R)
    for k, v in enumerate(vals):
        for j in range(12): out[12 * k + j] = (v >> (32 * j)) & 0xFFFFFFFF
    return out
