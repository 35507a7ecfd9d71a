"""Raw host memory copy bandwidth of the box with N threads (numpy releases the GIL in copyto): the ceiling of the
pageable staging path, which costs one extra pass over the data per direction."""
import sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
n = 1 << 30
src = np.ones(n, dtype=np.uint8); dst = np.zeros(n, dtype=np.uint8)
for threads in (1, 2, 4, 8, 12, 16):
    parts = [(i * n // threads, (i + 1) * n // threads) for i in range(threads)]
    with ThreadPoolExecutor(threads) as ex:
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            list(ex.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]]), parts))
            best = min(best, time.perf_counter() - t0)
    print(f"{threads:2d} threads: {n / best / 1e9:.1f} GB/s copied ({2 * n / best / 1e9:.1f} GB/s of traffic)")
