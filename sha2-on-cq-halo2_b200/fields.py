"""Host-side bn256 field constants and the few scalar computations the reference also does on the host
(EvaluationDomain::new, poly/domain.rs:39-142). Python integers; values cross to the device as Montgomery limbs."""
import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # bn256/fr.rs:16
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # bn256/fq.rs:18
MONT_R = 1 << 256
FR_S = 28                                                                    # bn256/fr.rs:72
FR_ROOT_OF_UNITY = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c  # fr.rs:77-82
FR_ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23           # fr.rs:112-117
FR_ONE_MONT = MONT_R % R_MOD
FQ_ONE_MONT = MONT_R % Q_MOD


def fr_to_limbs(x):
    """canonical int -> (4,) uint64 Montgomery limbs"""
    m = (x % R_MOD) * MONT_R % R_MOD
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def fr_from_limbs(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    m = sum(int(a[i]) << (64 * i) for i in range(4))
    return m * pow(MONT_R, -1, R_MOD) % R_MOD


def fq_mont_limbs(x):
    m = (x % Q_MOD) * MONT_R % Q_MOD
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
