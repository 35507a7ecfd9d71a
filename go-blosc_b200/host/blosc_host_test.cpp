// blosc_host_test.cpp -- the reference's quick-start and a few of its tests, written against
// the C++ host mirror (blosc.hpp).  Needs a GPU: tests/test_gpu_parity.py runs it on the box.
// `--abi` only checks host-side entry points (no device), so it also runs on the CPU.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "blosc.hpp"

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); fails++; } } while (0)

static blosc::Bytes ramp(size_t n) { blosc::Bytes b(n); for (size_t i = 0; i < n; i++) b[i] = (uint8_t)(i % 256); return b; }

int main(int argc, char **argv) {
    using namespace blosc;
    // host-only checks
    Header h{}; h.version = 2; h.versionlz = 1; h.flags = 5; h.typesize = 8; h.nbytes_orig = 1000; h.blocksize = 1000; h.nbytes_comp = 500;
    Bytes raw = h.Bytes();
    Header g = ParseHeader(raw);
    EXPECT(g.nbytes_orig == 1000 && g.ShuffleMode() == Shuffle::BitShuffle && !g.IsMemcpy());
    EXPECT(to_string(Codec::LZ4) == "lz4" && to_string(Shuffle::Shuffle1) == "shuffle");
    try { ParseHeader(Bytes(3)); EXPECT(false); } catch (const ErrInvalidHeader &) {}
    raw[0] = 9;
    try { ParseHeader(raw); EXPECT(false); } catch (const ErrInvalidVersion &) {}
    try { Compress(Bytes(), Codec::LZ4, 5, Shuffle::Shuffle1, 4); EXPECT(false); } catch (const ErrInvalidData &) {}
    if (argc > 1 && !std::strcmp(argv[1], "--abi")) { std::printf(fails ? "abi: FAIL\n" : "abi: ok\n"); return fails != 0; }

    // README quick start: 100 KB ramp, LZ4 level 5, Shuffle typesize 4 (BASELINE config C1)
    Bytes data = ramp(100000);
    Bytes frame = Compress(data, Codec::LZ4, 5, Shuffle::Shuffle1, 4);
    Header info = GetInfo(frame);
    EXPECT(info.version == 2 && info.versionlz == 1 && info.typesize == 4 && info.nbytes_orig == 100000);
    EXPECT(info.HasShuffle() && !info.HasBitShuffle() && info.nbytes_comp == frame.size() && frame.size() < 2000);
    EXPECT(Decompress(frame) == data);
    EXPECT(GetDecompressedSize(frame) == 100000);
    // float64 * 0.1 with BitShuffle (blosc_test.go:136-163)
    Bytes f64(8000);
    for (int i = 0; i < 1000; i++) { double v = i * 0.1; std::memcpy(&f64[8 * i], &v, 8); }
    EXPECT(Decompress(Compress(f64, Codec::LZ4, 5, Shuffle::BitShuffle, 8)) == f64);
    // in-place wrappers (shuffle_test.go:93-111)
    Bytes buf = data;
    ShuffleBuffer(buf, 4, Shuffle::Shuffle1);
    EXPECT(buf != data && buf[1] == 4);   // dst[1] = src[1*4+0]
    UnshuffleBuffer(buf, 4, Shuffle::Shuffle1);
    EXPECT(buf == data);
    // sentinels
    try { Compress(data, Codec::ZSTD, 5, Shuffle::Shuffle1, 4); EXPECT(false); } catch (const ErrUnsupportedOnGPU &) {}
    try { Compress(data, Codec::BloscLZ, 5, Shuffle::Shuffle1, 4); EXPECT(false); } catch (const ErrInvalidCodec &) {}
    Bytes bad = frame; for (size_t i = 16; i < bad.size(); i++) bad[i] ^= 0xFF;   // blosc_test.go:593-611
    try { Decompress(bad); EXPECT(false); } catch (const Error &) {}
    Bytes mism = Compress(ramp(1000), Codec::LZ4, 5, Shuffle::NoShuffle, 1); mism[4] = 0xD0; mism[5] = 0x07;   // 2000
    try { Decompress(mism); EXPECT(false); } catch (const ErrSizeMismatch &) {}
    // opt-in Blosc-1 multi-block frames: Options.blockSize is honoured here and only here
    Options ob; ob.blockSize = 16384;
    Bytes bf = CompressBlocks(f64, ob), bd = CompressBlocks(data, ob);
    Header bh = GetInfo(bd);
    EXPECT(bh.blocksize == 16384 && bh.nbytes_orig == 100000 && bh.nbytes_comp == bd.size() && (bh.flags & 0x10));
    EXPECT(DecompressBlocks(bd) == data && DecompressBlocks(bf) == f64);
    try { Decompress(bd); EXPECT(false); } catch (const Error &) {}   // the one-block decoder must not accept it
    std::printf(fails ? "blosc_host_test: %d FAILURES\n" : "blosc_host_test: all ok\n", fails);
    return fails != 0;
}
