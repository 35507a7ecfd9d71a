// blosc.hpp -- C++ host-side mirror of go-blosc's public API over the b2b C ABI.
//
// The reference is a Go package; with no Go toolchain in this image the compiled host layer
// is C++ (the Go source of the same layer is in ../go/blosc).  Names, argument meaning and
// error behaviour follow the reference: blosc.go:55-317 (types, Compress*, Decompress*,
// GetInfo, GetDecompressedSize, ParseHeader), shuffle.go:298-323 (ShuffleBuffer,
// UnshuffleBuffer).  Errors are exceptions carrying the reference's sentinel identity.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "b2b.h"

namespace blosc {

inline constexpr const char *Version = "1.0.0";   // blosc.go:49
inline constexpr int FormatVersion = 2;           // blosc.go:50
inline constexpr int HeaderSize = 16, MinHeaderSize = 16;

enum class Codec : uint8_t { BloscLZ = 0, LZ4, LZ4HC, Snappy, ZLIB, ZSTD };   // blosc.go:55-64
enum class Shuffle : uint8_t { NoShuffle = 0, Shuffle1 = 1, BitShuffle = 2 };  // blosc.go:86-92

inline std::string to_string(Codec c) {
    static const char *n[] = {"blosclz", "lz4", "lz4hc", "snappy", "zlib", "zstd"};
    return (unsigned)c < 6 ? n[(unsigned)c] : "unknown(" + std::to_string((unsigned)c) + ")";
}
inline std::string to_string(Shuffle s) {
    static const char *n[] = {"noshuffle", "shuffle", "bitshuffle"};
    return (unsigned)s < 3 ? n[(unsigned)s] : "unknown(" + std::to_string((unsigned)s) + ")";
}

// One exception type per sentinel of blosc.go:125-149; `status` is the B2B_* code.
struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};
#define BLOSC_SENTINEL(Name, Code) \
    struct Name : Error { explicit Name(const std::string &m = b2b_strerror(Code)) : Error(Code, m) {} }
BLOSC_SENTINEL(ErrInvalidData, B2B_EINVALID_DATA);
BLOSC_SENTINEL(ErrInvalidHeader, B2B_EINVALID_HEADER);
BLOSC_SENTINEL(ErrInvalidVersion, B2B_EINVALID_VERSION);
BLOSC_SENTINEL(ErrInvalidCodec, B2B_EINVALID_CODEC);
BLOSC_SENTINEL(ErrSizeMismatch, B2B_ESIZE_MISMATCH);
BLOSC_SENTINEL(ErrDataTooLarge, B2B_EDATA_TOO_LARGE);
BLOSC_SENTINEL(ErrCompressionFailed, B2B_ECOMPRESSION_FAILED);
BLOSC_SENTINEL(ErrDecompressionFailed, B2B_EDECOMPRESSION_FAILED);
BLOSC_SENTINEL(ErrUnsupportedOnGPU, B2B_EUNSUPPORTED);
#undef BLOSC_SENTINEL

[[noreturn]] inline void raise(int st, const std::string &detail = "") {
    const std::string m = std::string(b2b_strerror(st)) + (detail.empty() ? "" : ": " + detail);
    switch (st) {
        case B2B_EINVALID_DATA: throw ErrInvalidData(m);
        case B2B_EINVALID_HEADER: throw ErrInvalidHeader(m);
        case B2B_EINVALID_VERSION: throw ErrInvalidVersion(m);
        case B2B_EINVALID_CODEC: throw ErrInvalidCodec(m);
        case B2B_ESIZE_MISMATCH: throw ErrSizeMismatch(m);
        case B2B_EDATA_TOO_LARGE: throw ErrDataTooLarge(m);
        case B2B_ECOMPRESSION_FAILED: throw ErrCompressionFailed(m);
        case B2B_EDECOMPRESSION_FAILED: throw ErrDecompressionFailed(m);
        case B2B_EUNSUPPORTED: throw ErrUnsupportedOnGPU(m);
        default: throw Error(st, m);
    }
}

struct Header : b2b_header {   // blosc.go:154-224
    bool HasShuffle() const { return flags & B2B_FLAG_SHUFFLE; }
    bool HasBitShuffle() const { return flags & B2B_FLAG_BITSHUFFLE; }
    bool IsMemcpy() const { return flags & B2B_FLAG_MEMCPY; }
    Shuffle ShuffleMode() const {
        return HasBitShuffle() ? Shuffle::BitShuffle : HasShuffle() ? Shuffle::Shuffle1 : Shuffle::NoShuffle;
    }
    std::vector<uint8_t> Bytes() const { std::vector<uint8_t> o(16); b2b_header_bytes(this, o.data()); return o; }
};

struct Options {   // blosc.go:227-245
    Codec codec = Codec::LZ4;
    int level = 5;
    Shuffle shuffle = Shuffle::Shuffle1;
    int64_t typeSize = 4;
    int blockSize = 0, numThreads = 0;   // declared, never read (as in the reference)
};
inline Options DefaultOptions() { return Options{}; }

using Bytes = std::vector<uint8_t>;

inline Header ParseHeader(const Bytes &data) {   // blosc.go:165-185
    Header h{};
    if (int rc = b2b_parse_header(data.data(), data.size(), &h)) raise(rc);
    return h;
}
inline Header GetInfo(const Bytes &data) { return ParseHeader(data); }
inline size_t GetDecompressedSize(const Bytes &data) { return ParseHeader(data).nbytes_orig; }

// one process-wide context (a pool would serve concurrent callers, like the Go package's)
inline b2b_ctx *context() {
    static b2b_ctx *ctx = [] {
        b2b_ctx *c = nullptr;
        if (int rc = b2b_init(0, &c)) raise(rc, "b2b_init: no CUDA device (there is no CPU fallback)");
        return c;
    }();
    return ctx;
}

inline Bytes CompressWithOptions(const Bytes &data, Options o) {   // blosc.go:268-286
    if (data.empty()) throw ErrInvalidData();
    if (o.typeSize <= 0) o.typeSize = 1;
    o.level = o.level < 1 ? 1 : o.level > 9 ? 9 : o.level;
    Bytes out(b2b_max_frame_size(data.size()) + 64);
    size_t n = 0;
    if (int rc = b2b_compress(context(), data.data(), data.size(), (int)o.codec, o.level, (int)o.shuffle, o.typeSize,
                              out.data(), out.size(), &n))
        raise(rc, to_string(o.codec));
    out.resize(n);
    return out;
}
inline Bytes Compress(const Bytes &data, Codec c, int level, Shuffle s, int64_t typeSize) {   // blosc.go:257-265
    Options o; o.codec = c; o.level = level; o.shuffle = s; o.typeSize = typeSize;
    return CompressWithOptions(data, o);
}
inline Bytes DecompressWithSize(const Bytes &data, int64_t typeSize) {   // blosc.go:296-303
    if (data.size() < (size_t)HeaderSize) throw ErrInvalidHeader();
    Header h = ParseHeader(data);
    size_t cap = h.nbytes_orig, reach = 255 * data.size() + 64;
    if (cap > reach) cap = reach;
    Bytes out(cap + 1);
    size_t n = 0;
    int rc = b2b_decompress(context(), data.data(), data.size(), typeSize, out.data(), cap, &n);
    if (rc == B2B_EDST_TOO_SMALL && cap < h.nbytes_orig) rc = B2B_ESIZE_MISMATCH;
    if (rc) raise(rc);
    out.resize(n);
    return out;
}
inline Bytes Decompress(const Bytes &data) { return DecompressWithSize(data, 0); }   // blosc.go:291-293

// Blosc-1 multi-block frames: the one place where Options.blockSize (declared, never read by the
// reference: blosc.go:227-234) means something; 0 = 64 KiB.  The reference cannot decode these frames.
inline Bytes CompressBlocks(const Bytes &data, const Options &o) {
    if (data.empty()) throw ErrInvalidData();
    if (o.codec != Codec::LZ4) raise(B2B_EINVALID_CODEC, "multi-block frames are LZ4 only");
    Bytes out(data.size() + 16 + 64);
    size_t n = 0;
    if (int rc = b2b_compress_blocks(context(), data.data(), data.size(), (int)o.shuffle, o.typeSize,
                                     (uint32_t)o.blockSize, out.data(), out.size(), &n))
        raise(rc, "blocks");
    out.resize(n);
    return out;
}
inline Bytes DecompressBlocks(const Bytes &data) {
    if (data.size() < (size_t)HeaderSize) throw ErrInvalidHeader();
    Header h = ParseHeader(data);
    size_t cap = h.nbytes_orig, reach = 255 * data.size() + 64;
    if (cap > reach) cap = reach;
    Bytes out(cap + 1);
    size_t n = 0;
    int rc = b2b_decompress_blocks(context(), data.data(), data.size(), out.data(), cap, &n);
    if (rc == B2B_EDST_TOO_SMALL && cap < h.nbytes_orig) rc = B2B_EDECOMPRESSION_FAILED;
    if (rc) raise(rc);
    out.resize(n);
    return out;
}

inline void ShuffleBuffer(Bytes &data, int64_t typeSize, Shuffle mode) {   // shuffle.go:298-309
    if (data.empty() || (mode != Shuffle::Shuffle1 && mode != Shuffle::BitShuffle)) return;
    if (int rc = b2b_shuffle(context(), (int)mode, 0, typeSize, data.data(), data.data(), data.size())) raise(rc);
}
inline void UnshuffleBuffer(Bytes &data, int64_t typeSize, Shuffle mode) {   // shuffle.go:312-323
    if (data.empty() || (mode != Shuffle::Shuffle1 && mode != Shuffle::BitShuffle)) return;
    if (int rc = b2b_shuffle(context(), (int)mode, 1, typeSize, data.data(), data.data(), data.size())) raise(rc);
}

}  // namespace blosc
