"""Two implementations behind one interface so the reference's own test cases (restated in
reference_suite.py) run unchanged against the CPU oracle and against the CUDA path."""
import numpy as np


class OracleImpl:
    """CPU oracle (checker)."""
    name = "oracle"

    def __init__(self, orc, pkg):
        self.orc, self.pkg = orc, pkg

    def _raise(self, rc):
        errs = {c.status: c for c in (self.pkg.ErrInvalidData, self.pkg.ErrInvalidHeader, self.pkg.ErrInvalidVersion,
                                      self.pkg.ErrInvalidCodec, self.pkg.ErrSizeMismatch, self.pkg.ErrDataTooLarge,
                                      self.pkg.ErrCompressionFailed, self.pkg.ErrDecompressionFailed,
                                      self.pkg.ErrUnsupported, self.pkg.ErrDstTooSmall)}
        raise errs[rc](f"oracle status {rc}")

    def compress(self, data, codec=1, level=5, shuffle=1, typesize=4, quirk=False):
        rc, fr = self.orc.compress(data, int(codec), int(level), int(shuffle), int(typesize),
                                   self.orc.MEMCPY_REF_QUIRK if quirk else self.orc.MEMCPY_SHUFFLED)
        if rc:
            self._raise(rc)
        return fr.tobytes()

    def decompress(self, frame, typesize=0):
        a = np.frombuffer(bytes(frame), dtype=np.uint8)
        cap = None
        if a.size >= 16:
            cap = min(int.from_bytes(a[4:8].tobytes(), "little"), 255 * a.size + 64)
        rc, out = self.orc.decompress(a, typesize, cap)
        if rc == self.orc.EDST_TOO_SMALL:
            rc = self.orc.ESIZE_MISMATCH
        if rc:
            self._raise(rc)
        return out.tobytes()

    def shuffle(self, data, typesize, mode=1, inverse=False):
        fn = {(1, False): self.orc.shuffle, (1, True): self.orc.unshuffle,
              (2, False): self.orc.bitshuffle, (2, True): self.orc.bitunshuffle}.get((int(mode), bool(inverse)))
        a = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        return fn(a, typesize) if fn else a.copy()


class GpuImpl:
    """CUDA path through the C ABI."""
    name = "gpu"

    def __init__(self, ctx, pkg):
        self.ctx, self.pkg = ctx, pkg

    def compress(self, data, codec=1, level=5, shuffle=1, typesize=4, quirk=False):
        if quirk:
            self.ctx.set_option(self.pkg.OPT_REF_MEMCPY_QUIRK, 1)
        try:
            if len(data) == 0:
                raise self.pkg.ErrInvalidData("empty")
            return self.ctx.compress(data, codec, min(max(level, 1), 9), shuffle, typesize)
        finally:
            if quirk:
                self.ctx.set_option(self.pkg.OPT_REF_MEMCPY_QUIRK, 0)

    def decompress(self, frame, typesize=0):
        return self.ctx.decompress(frame, typesize)

    def shuffle(self, data, typesize, mode=1, inverse=False):
        return self.ctx.shuffle(data, typesize, mode, inverse)
