"""Formula inputs of tests/tools/make_ref_golden/main.go, reproduced bit for bit (integer arithmetic only; every
float is an exact small integer divided by a power of two)."""
import numpy as np

M64 = (1 << 64) - 1


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _tri(p, period):
    q = p % np.uint64(period)
    q = np.where(q > period // 2, np.uint64(period) - q, q)
    return q.astype(np.int64) - np.int64(period // 4)


def smooth_int(i, seed):
    i = np.asarray(i, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return 3 * _tri(i, 4096) + _tri(i * np.uint64(3), 667) + (splitmix64(np.uint64(seed) ^ i) & np.uint64(15)).astype(np.int64)


def ramp(n):
    return (np.arange(n) % 256).astype(np.uint8)


def smooth_f32(elems, seed):
    v = smooth_int(np.arange(elems, dtype=np.uint64), seed).astype(np.float32) / np.float32(1024)
    return v.view(np.uint8)


def smooth_f64(elems, seed):
    i = np.arange(elems, dtype=np.uint64)
    with np.errstate(over="ignore"):
        noise = (splitmix64(np.uint64(seed) + i) & np.uint64(0xFFFFF)).astype(np.float64) / 1099511627776.0
    return (smooth_int(i, seed).astype(np.float64) / 1024.0 + noise).view(np.uint8)


def lowent_i16(elems, seed):
    i = np.arange(elems, dtype=np.uint64)
    return (splitmix64(np.uint64(seed) ^ i) & np.uint64(7)).astype(np.uint16).view(np.uint8)


def random_bytes(n, seed):
    i = np.arange(0, n, 8, dtype=np.uint64)
    with np.errstate(over="ignore"):
        v = splitmix64(np.uint64(seed) + i)
    return v.view(np.uint8)[:n].copy()


def make(call):
    """'smooth_f32(65536, 3)' -> bytes; only the generators above are callable."""
    return eval(call, {"__builtins__": {}}, {"ramp": ramp, "smooth_f32": smooth_f32, "smooth_f64": smooth_f64,
                                              "lowent_i16": lowent_i16, "random_bytes": random_bytes})
