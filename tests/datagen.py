"""Deterministic test inputs: the reference's own test generators and the
synthetic typed-array fields of BASELINE.json's configs."""
import numpy as np

SIZES = [1, 2, 3, 10, 12, 13, 15, 16, 17, 28, 32, 35, 37, 64, 97, 100, 103, 127, 128, 256, 999, 1000,
         1003, 1024, 4096, 10000, 65536, 100000, 100003]
TYPESIZES = [1, 2, 3, 4, 5, 7, 8, 12, 16, 17, 32, 255]


def ramp(n):                         # makeTestData, blosc_test.go:352-359
    return (np.arange(n) % 256).astype(np.uint8)


def f32_ramp(n=1000, k=0.1):         # blosc_test.go:107-117
    return (np.arange(n, dtype=np.float32) * np.float32(k)).view(np.uint8)


def f64_ramp(n=1000, k=0.1):         # blosc_test.go:136-146
    return (np.arange(n, dtype=np.float64) * k).view(np.uint8)


def lcg_bytes(n, seed=12345):        # fuzz_test.go:462-471
    out = np.empty(n, dtype=np.uint8)
    x = seed
    for i in range(n):
        x = (x * 1103515245 + 12345) & 0xFFFFFFFF
        out[i] = (x >> 16) & 0xFF
    return out


def random_bytes(n, seed=0):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def smooth_f32(n_elems, seed=0, start=0):
    """BASELINE config C2/C3 field, float32."""
    i = np.arange(start, start + n_elems, dtype=np.float64)
    u = np.random.default_rng(seed).uniform(-1, 1, n_elems)
    return (np.sin(2 * np.pi * i / 4096) + 0.25 * np.sin(2 * np.pi * i / 333.3) + 1e-3 * u).astype(np.float32).view(np.uint8)


def smooth_f64(n_elems, seed=0, start=0):
    """BASELINE config C4 field, float64."""
    i = np.arange(start, start + n_elems, dtype=np.float64)
    u = np.random.default_rng(seed).uniform(-1, 1, n_elems)
    return (np.sin(2 * np.pi * i / 4096) + 0.25 * np.sin(2 * np.pi * i / 333.3) + 1e-3 * u).view(np.uint8)


def lowent_i16(n_elems, seed=0):
    """BASELINE config C5 low-entropy int16 (3 random bits per element)."""
    return np.random.default_rng(seed).integers(0, 8, n_elems).astype(np.int16).view(np.uint8)


def text_like(n, seed=0):
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 9)).astype(np.uint8)) for _ in range(200)]
    out = bytearray()
    while len(out) < n:
        out += words[int(rng.integers(0, len(words)))] + b" "
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()


def corpus(n):
    """Named inputs of n bytes exercising literals-only, long matches, short matches, overlaps."""
    return {
        "ramp": ramp(n),
        "zeros": np.zeros(n, dtype=np.uint8),
        "random": random_bytes(n, 3),
        "smooth_f32": smooth_f32((n + 3) // 4, 1)[:n].copy(),
        "lowent_i16": lowent_i16((n + 1) // 2, 2)[:n].copy(),
        "text": text_like(n, 4),
        "period3": np.tile(np.array([1, 2, 3], dtype=np.uint8), n // 3 + 1)[:n].copy(),
        "period7_noise": (np.tile(np.arange(7, dtype=np.uint8), n // 7 + 1)[:n]
                          ^ (random_bytes(n, 5) > 250).astype(np.uint8)),
    }
