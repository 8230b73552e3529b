"""Deterministic input generators shared by make_golden.py and the tests (no dependence on numpy's RNG streams)."""
import numpy as np


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n pseudo-random uint64 (SplitMix64), vectorised."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def multiset(seed: int, hb: int, n_base: int, n_single: int) -> np.ndarray:
    """Keys in [0, 2^hb): n_base keys of which half repeat 2x, a quarter 3x, five 300x, one 20000x; n_single singletons."""
    mask = np.uint64((1 << hb) - 1)
    base = splitmix64(seed, n_base) & mask
    single = splitmix64(seed + 1, n_single) & mask
    keys = np.concatenate([base, base[: n_base // 2], base[: n_base // 4], single,
                           np.repeat(base[:5], 300), np.repeat(base[5:6], 20000)])
    perm = np.argsort(splitmix64(seed + 2, len(keys)), kind="stable")
    return keys[perm]


def dna(seed: int, n: int, alphabet=b"ACGT") -> bytes:
    a = np.frombuffer(alphabet, np.uint8)
    return a[(splitmix64(seed, n) >> np.uint64(33)) % np.uint64(len(a))].tobytes()
