/*
 * blosc_oracle.h -- CPU ORACLE for the go-blosc shuffle + LZ4 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (go-blosc_b200/, include/)
 * may include, link, import or execute anything under oracle/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and there only as the checker / the CPU arm that is timed beside the GPU.
 *
 * What it restates (reference = mrjoshuak/go-blosc v1.0.2, cited as file:line):
 *   shuffle.go:16-73     shuffleBytes      -> orc_shuffle
 *   shuffle.go:76-133    unshuffleBytes    -> orc_unshuffle
 *   shuffle.go:145-219   bitShuffle        -> orc_bitshuffle
 *   shuffle.go:222-295   bitUnshuffle      -> orc_bitunshuffle
 *   blosc.go:154-198     Header / ParseHeader / Bytes -> orc_header_*
 *   blosc.go:268-286     CompressWithOptions clamps   -> orc_compress
 *   blosc.go:320-374     compressBackend              -> orc_compress
 *   blosc.go:296-303,377-434 DecompressWithSize / decompressBackend -> orc_decompress
 *   codec.go:63-84       lz4Codec.{Compress,Decompress}            -> orc_lz4_*
 *
 * Third-party arithmetic that is NOT in /root/reference: github.com/pierrec/lz4/v4
 * v4.1.23 (go.mod:7).  orc_lz4_compress restates its published fast block compressor
 * (CompressBlock: 2^16-entry uint16 table, 6-byte multiplicative hash, probes at s,
 * s+1, s+2, adaptive skip >>7, backward extension, mfLimit 14) and orc_lz4_decompress
 * follows the public LZ4 block format.
 *
 * PARITY STATUS
 *   - shuffle / unshuffle / bitshuffle / bitunshuffle / header / frame logic / error
 *     selection: pinned against every formula, literal and header assertion the
 *     reference's own tests hold (tests/test_oracle_pins.py lists them file:line).
 *   - LZ4 *compressed bytes and sizes*: PARITY UNPINNED.  The reference's tests hold no
 *     compressed byte or size, there is no Go toolchain here to run it, and pierrec's
 *     source is not in the tree.  Streams are referee'd for format validity by the
 *     system liblz4 1.9.4 (LZ4_decompress_safe) in tests/.
 */
#ifndef BLOSC_ORACLE_H
#define BLOSC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: one per reference sentinel (blosc.go:125-149), same numbering as
 * include/b2b.h so tests can compare them directly. */
enum {
    ORC_OK = 0,
    ORC_EINVALID_DATA = 1,          /* ErrInvalidData          */
    ORC_EINVALID_HEADER = 2,        /* ErrInvalidHeader        */
    ORC_EINVALID_VERSION = 3,       /* ErrInvalidVersion       */
    ORC_EINVALID_CODEC = 4,         /* ErrInvalidCodec         */
    ORC_ESIZE_MISMATCH = 5,         /* ErrSizeMismatch         */
    ORC_EDATA_TOO_LARGE = 6,        /* ErrDataTooLarge (n >= 2^32-16: the u32 header cannot hold it) */
    ORC_ECOMPRESSION_FAILED = 7,    /* ErrCompressionFailed    */
    ORC_EDECOMPRESSION_FAILED = 8,  /* ErrDecompressionFailed  */
    ORC_EUNSUPPORTED = 10,          /* codec registered in the reference but outside this path */
    ORC_EDST_TOO_SMALL = 11
};

enum { ORC_NOSHUFFLE = 0, ORC_SHUFFLE = 1, ORC_BITSHUFFLE = 2 };
enum { ORC_BLOSCLZ = 0, ORC_LZ4 = 1, ORC_LZ4HC = 2, ORC_SNAPPY = 3, ORC_ZLIB = 4, ORC_ZSTD = 5 };
enum { ORC_FLAG_SHUFFLE = 0x1, ORC_FLAG_MEMCPY = 0x2, ORC_FLAG_BITSHUFFLE = 0x4 };

/* Encoder policy for memcpy frames (SURVEY F4 / DESIGN.md):
 * 0 = store the SHUFFLED bytes (round-trips through the reference decoder),
 * 1 = store the original bytes exactly as blosc.go:342-345 does (byte-identical to the
 *     reference, but such frames do not round-trip when a shuffle flag is set). */
enum { ORC_MEMCPY_SHUFFLED = 0, ORC_MEMCPY_REF_QUIRK = 1 };

typedef struct {
    uint8_t version, versionlz, flags, typesize;
    uint32_t nbytes_orig, blocksize, nbytes_comp;
} orc_header;

/* ---- filters (out of place; src and dst must not overlap) ---------------------- */
void orc_shuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t typesize);
void orc_unshuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t typesize);
void orc_bitshuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t typesize);
void orc_bitunshuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t typesize);
/* AVX2 T=4 byte shuffle pair used only by the CPU *baseline* so that it is as fast as the
 * reference's own amd64 path (shuffle_amd64.s:138-330); bit-identical to the scalar ones.
 * Fall back to the scalar loops when the CPU lacks AVX2. */
void orc_shuffle_fast(const uint8_t *src, uint8_t *dst, size_t n, int64_t typesize);
void orc_unshuffle_fast(const uint8_t *src, uint8_t *dst, size_t n, int64_t typesize);

/* ---- header ------------------------------------------------------------------- */
int orc_header_parse(const uint8_t *data, size_t len, orc_header *h);
void orc_header_bytes(const orc_header *h, uint8_t out[16]);

/* ---- LZ4 block codec ------------------------------------------------------------ */
size_t orc_lz4_bound(size_t n);
/* returns compressed size (>0) or 0 if dst is too small */
size_t orc_lz4_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap);
/* returns decoded size (>=0) or -1 on malformed input / overrun of dst */
int64_t orc_lz4_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap);

/* ---- frames --------------------------------------------------------------------- */
size_t orc_max_frame_size(size_t n);
int orc_compress(const uint8_t *data, size_t n, int codec, int level, int shuffle,
                 int64_t typesize, int memcpy_policy, uint8_t *dst, size_t cap, size_t *out_len);
int orc_decompress(const uint8_t *frame, size_t len, int64_t typesize_override, uint8_t *dst,
                   size_t cap, size_t *out_len);

/* ---- multi-threaded batch drivers (CPU baseline arm of bench.py) ------------------
 * One frame per task, `threads` pthreads pulling tasks from an atomic counter: the
 * goroutine-per-chunk arrangement BASELINE.md describes for the reference. Returns the
 * first non-zero status (or 0). `fast` selects the AVX2 T=4 shuffle like the reference. */
int orc_compress_batch_mt(const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len,
                          uint32_t nframes, int shuffle, int64_t typesize, uint8_t *dst,
                          const uint64_t *dst_off, uint32_t *dst_len, int threads, int fast);
int orc_decompress_batch_mt(const uint8_t *frames, const uint64_t *frame_off,
                            const uint32_t *frame_len, uint32_t nframes, uint8_t *dst,
                            const uint64_t *dst_off, uint32_t *out_len, int threads, int fast);
int orc_shuffle_mt(int mode, int inverse, int64_t typesize, const uint8_t *src, uint8_t *dst,
                   size_t n, int threads, int fast);

/* ---- opt-in Blosc-1 multi-block frames (blosc1_blocks.c) ------------
 * No reference counterpart (the reference ignores Options.BlockSize): PARITY
 * UNPINNED, the published Blosc-1 chunk layout restated.  blocksize_req 0 = 64 KiB;
 * split != 0 writes typesize streams per block where Blosc-1 would. */
enum { ORC_B1_FLAG_DONTSPLIT = 0x10, ORC_B1_LZ4_FORMAT = 1 };
uint32_t orc_blocks_blocksize(size_t n, int64_t typesize, uint32_t blocksize_req);
int orc_blocks_compress(const uint8_t *data, size_t n, int shuffle, int64_t typesize,
                        uint32_t blocksize_req, int split, uint8_t *dst, size_t cap, size_t *out_len);
int orc_blocks_decompress(const uint8_t *frame, size_t len, uint8_t *dst, size_t cap, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
