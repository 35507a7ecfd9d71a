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


def strip_adversarial(n, seed=0):
    """Inputs aimed at the strip-parallel encoder (61-byte strips, 32 per step, 64 KiB segments)
    and the batched decoder (32-position token windows, <= 32 sequences per batch): periods equal
    to / around the strip and the step, matches that span strips, steps and segments, runs broken
    by single bytes, sequences with exactly one extension byte, dense 4-byte matches."""
    rng = np.random.default_rng(seed)
    out = {}
    for period in (1, 2, 3, 4, 5, 60, 61, 62, 64, 122, 183, 244, 1951, 1952, 1953, 4096):
        base = rng.integers(0, 256, period, dtype=np.uint8)
        out[f"period{period}"] = np.tile(base, n // period + 1)[:n].copy()
    # runs of random length (1..400) of random bytes: overlapping matches of every length class
    runs = []
    total = 0
    while total < n:
        k = int(rng.integers(1, 401)); runs.append(np.full(k, rng.integers(0, 256), dtype=np.uint8)); total += k
    out["runs"] = np.concatenate(runs)[:n].copy()
    # noise with a copy of an earlier chunk every ~200 bytes (matches 4..300 long at far offsets)
    a = rng.integers(0, 256, n, dtype=np.uint8)
    pos = 300
    while pos + 310 < n:
        ln = int(rng.integers(4, 300)); src = int(rng.integers(0, pos - ln)) if pos > ln else 0
        if pos - src < 65535:
            a[pos:pos + ln] = a[src:src + ln]
        pos += ln + int(rng.integers(1, 200))
    out["copies"] = a
    # literal runs of 15..40 bytes between 4..8-byte matches (one-extension-byte tokens)
    b = rng.integers(0, 256, n, dtype=np.uint8)
    pos = 64
    while pos + 64 < n:
        ln = int(rng.integers(4, 9)); b[pos:pos + ln] = b[pos - 50:pos - 50 + ln]; pos += ln + int(rng.integers(15, 41))
    out["lit_ext"] = b
    # two-symbol text: dense short matches, long dependency chains inside a decode batch
    out["binary_text"] = rng.integers(0, 2, n, dtype=np.uint8) * 7 + 65
    # a long match that starts in one strip / segment and ends far in a later one
    c = rng.integers(0, 256, n, dtype=np.uint8)
    if n > 3000:
        half = n // 2
        c[half:half + half // 2] = c[10:10 + half // 2] if half - 10 < 65535 else c[half - 65000:half - 65000 + half // 2]
    out["long_copy"] = c
    return out
