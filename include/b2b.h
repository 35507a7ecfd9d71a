/*
 * b2b.h -- C ABI of the B200-native shuffle + LZ4 backend for go-blosc ("b2b" = blosc-to-B200).
 *
 * This is the drop-in boundary: plain pointers and sizes, no CUDA or torch types.  It is
 * what the reference's own backend seam would bind through cgo:
 *
 *   reference (mrjoshuak/go-blosc v1.0.2)            entry point here
 *   ------------------------------------------------  --------------------------------------
 *   compressBackend      blosc.go:320-374            b2b_compress
 *   decompressBackend    blosc.go:377-434            b2b_decompress
 *   ShuffleBuffer        shuffle.go:298-309          b2b_shuffle(mode, inverse=0)
 *   UnshuffleBuffer      shuffle.go:312-323          b2b_shuffle(mode, inverse=1)
 *   shuffleBytes/unshuffleBytes/bitShuffle/bitUnshuffle and their amd64/arm64 assembly
 *                        shuffle.go:16-295, shuffle_amd64.s, shuffle_arm64.s
 *                                                     b2b_shuffle / b2b_shuffle_dev
 *   lz4Codec.Compress / .Decompress  codec.go:63-84  inside b2b_compress / b2b_decompress
 *                                                     (and b2b_lz4_block_* for the
 *                                                     CodecInterface plugin seam, codec.go:15-38)
 *   ParseHeader / Header.Bytes  blosc.go:165-198     b2b_parse_header / b2b_header_bytes
 *   (no counterpart: the reference is one frame per call)
 *                                                     *_batch and *_batch_dev: many independent
 *                                                     frames per call + the packed-offsets table
 *
 * The Go side (go-blosc_b200/go/blosc) keeps every exported identifier of package blosc
 * and maps the status codes below onto the reference's sentinels (blosc.go:125-149).
 *
 * Threading: a b2b_ctx serialises the calls made on it (internal mutex); use one ctx per
 * concurrent caller (the Go package keeps a pool).  All functions return a B2B_* status.
 * There is NO CPU fallback: without a usable CUDA device b2b_init fails with B2B_ECUDA.
 */
#ifndef B2B_H
#define B2B_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2B_API __attribute__((visibility("default")))
#else
#define B2B_API
#endif

/* ---- status codes: one per reference sentinel (blosc.go:125-149) -------------------- */
enum {
    B2B_OK = 0,
    B2B_EINVALID_DATA = 1,         /* ErrInvalidData        (returned bare by the reference)  */
    B2B_EINVALID_HEADER = 2,       /* ErrInvalidHeader      (bare)                             */
    B2B_EINVALID_VERSION = 3,      /* ErrInvalidVersion     (wrapped)                          */
    B2B_EINVALID_CODEC = 4,        /* ErrInvalidCodec       (wrapped)                          */
    B2B_ESIZE_MISMATCH = 5,        /* ErrSizeMismatch       (wrapped)                          */
    B2B_EDATA_TOO_LARGE = 6,       /* ErrDataTooLarge: n > 2^32-17 cannot be framed */
    B2B_ECOMPRESSION_FAILED = 7,   /* ErrCompressionFailed  (wrapped)                          */
    B2B_EDECOMPRESSION_FAILED = 8, /* ErrDecompressionFailed(wrapped)                          */
    B2B_ECUDA = 9,                 /* CUDA runtime failure; see b2b_last_error                 */
    B2B_EUNSUPPORTED = 10,         /* codec is registered in the reference (LZ4HC encode, Snappy,
                                      ZLIB, ZSTD) but is outside this path: the Go host keeps
                                      those on the reference codecs.  LZ4 never returns this.  */
    B2B_EDST_TOO_SMALL = 11,       /* caller's dst capacity is insufficient                    */
    B2B_EINVAL = 12                /* bad argument (null pointer, bad enum, ...)               */
};

/* ---- enums of the reference (blosc.go:55-64, 86-92, 110-115) ------------------------ */
enum { B2B_BLOSCLZ = 0, B2B_LZ4 = 1, B2B_LZ4HC = 2, B2B_SNAPPY = 3, B2B_ZLIB = 4, B2B_ZSTD = 5 };
enum { B2B_NOSHUFFLE = 0, B2B_SHUFFLE = 1, B2B_BITSHUFFLE = 2 };
enum { B2B_FLAG_SHUFFLE = 0x1, B2B_FLAG_MEMCPY = 0x2, B2B_FLAG_BITSHUFFLE = 0x4 };
#define B2B_HEADER_SIZE 16
#define B2B_FORMAT_VERSION 2
/* additional flag bits of the opt-in multi-block frames (b2b_*_blocks_*) */
enum { B2B_FLAG_DONTSPLIT = 0x10, B2B_BLOCKS_LZ4_FORMAT = 1 /* flags >> 5 */ };

/* 16-byte frame header, little-endian on the wire (blosc.go:154-162) */
typedef struct b2b_header {
    uint8_t version;      /* 2 */
    uint8_t versionlz;    /* Codec enum (go-blosc stores the codec id here) */
    uint8_t flags;        /* B2B_FLAG_* */
    uint8_t typesize;     /* uint8(opts.TypeSize) */
    uint32_t nbytes_orig;
    uint32_t blocksize;   /* always == nbytes_orig (one block per frame) */
    uint32_t nbytes_comp; /* 16 + payload bytes */
} b2b_header;

/* ---- context ------------------------------------------------------------------------- */
typedef struct b2b_ctx b2b_ctx;

B2B_API int b2b_init(int device, b2b_ctx **out);
B2B_API void b2b_destroy(b2b_ctx *ctx);
B2B_API const char *b2b_strerror(int status);
B2B_API const char *b2b_version(void);
/* text of the last CUDA failure seen on this ctx ("" if none) */
B2B_API const char *b2b_last_error(b2b_ctx *ctx);

/* options */
enum {
    /* 0 (default): a memcpy frame stores the SHUFFLED bytes, so it round-trips through the
     *    reference decoder (which un-shuffles memcpy payloads too, blosc.go:398-426);
     * 1: store the original bytes exactly like blosc.go:342-345 (byte-identical frames, but
     *    the reference itself cannot round-trip them when a shuffle flag is set). */
    B2B_OPT_REF_MEMCPY_QUIRK = 1,
    /* number of CTAs per SM for the persistent filter kernels (tuning; 0 = built-in default) */
    B2B_OPT_FILTER_CTAS_PER_SM = 2,
    /* host batch path: bytes of uncompressed data per pipeline chunk (0 = default 128 MiB; 4 chunks in flight) */
    B2B_OPT_HOST_STAGE_BYTES = 3,
    /* log2 of the LZ4 match-finder's per-warp shared-memory hash table: 10..13, or 0 = automatic
     * (10 behind a byte or bit shuffle; 12 for unshuffled input, where the history has to reach
     * further).  Larger tables find more matches (ratio) and cost resident warps (speed). */
    B2B_OPT_HASH_LOG = 4,
    /* 1: bracket every kernel launch with CUDA events on its stream (see b2b_kernel_stats) */
    B2B_OPT_KERNEL_TIMING = 5,
    /* bytes of input the match finder hashes per position: 4, 5 or 6, or 0 = automatic (4 behind a
     * shuffle, 5 for unshuffled input: the most recent occurrence of a 4-byte context is rarely the
     * longest one in text or low-entropy integers; the reference compressor hashes 6).  5 needs
     * B2B_OPT_HASH_LOG >= 11 and 6 needs >= 12 (smaller tables are raised to that). */
    B2B_OPT_HASH_BYTES = 6,
    /* host batch path, pageable caller buffers: host threads that move them into / out of the pinned staging
     * ring (1..64; 0 = automatic: half of the CPUs this process may run on, at most 8) */
    B2B_OPT_HOST_THREADS = 7,
    /* 1: hand pageable buffers to cudaMemcpyAsync directly (synchronous, staged by the driver) instead of
     * the library's pinned ring; for comparison only */
    B2B_OPT_NO_HOST_STAGING = 8,
    /* LZ4 decoder variant: -1 automatic (default: by the batch's shape), 0 chunk-parallel (a frame is spread over
     * many threads: frames over 512 KiB), 1 fused one-warp-per-frame kernel, 2 parse kernel + copy kernel (one warp
     * per frame: thousands of small frames), 3 one lane per frame (2^16 frames or more of at most 4 KiB), 4 the
     * chunk-parallel parse followed by pointer jumping over per-byte source indices (a FEW LARGE frames, e.g. the one
     * frame of b2b_decompress: the frame is spread over the whole device; needs 4 bytes of scratch per output byte,
     * at most 256 frames and under 4 GiB of output per call, otherwise 0 is used; automatic when the batch's output
     * is at most 128 times its largest frame).  All give identical results. */
    B2B_OPT_DECODER = 9,
    /* 1: in the one-warp-per-frame decoders the warp that decoded a byte-shuffled frame (typesize 2 or 4, 16-byte
     * aligned slots, element count a multiple of 16) also un-shuffles it, instead of a separate pass over the
     * batch; 0 (default): the separate pass.  Same results; measured equally fast on a B200 (the decoding warps
     * are latency-bound and the device is fully occupied by them, so the transpose does not hide behind them). */
    B2B_OPT_FUSE_UNSHUFFLE = 10,
    /* device-pointer decompress batches of 2048 frames or more run as N parts on N streams (the caller's and
     * internal ones, forked and joined with events, so the call still behaves as work on the caller's stream):
     * 0 = automatic (2), 1 = one stream only, 2..4 = that many */
    B2B_OPT_DECODE_STREAMS = 11
};
B2B_API int b2b_set_option(b2b_ctx *ctx, int option, int64_t value);
/* pre-size the device scratch arena so that later calls do not allocate */
B2B_API int b2b_reserve(b2b_ctx *ctx, uint64_t total_uncompressed_bytes, uint32_t nframes);
/* number of kernel launches issued through this ctx since creation (bench's gpu_launches) */
B2B_API uint64_t b2b_launch_count(b2b_ctx *ctx);
/* per-kernel launch counts and (with B2B_OPT_KERNEL_TIMING) summed device time; kernel = 0..19:
 * filter, lz4 encode, lz4 decode (copy half), offsets scan, pack, frame info, finalize, lz4 parse
 * (decode's parse half), block-frame tables, block-frame pack, block-frame decode, then the chunk-parallel
 * decoder's frame prep, chunk parse, stitch and copy engine, the lane decoder, and the pointer-jumping
 * decoder's map, rounds, gather and long-run (+ check) launches; B2B_EINVAL beyond.
 * Synchronises pending events.  Times are event spans on the launching stream: kernels of a decompress batch
 * that runs on two streams overlap (B2B_OPT_DECODE_STREAMS = 1 gives exclusive times). */
B2B_API int b2b_kernel_stats(b2b_ctx *ctx, int kernel, const char **name, uint64_t *launches,
                             double *total_ms);
B2B_API int b2b_kernel_stats_reset(b2b_ctx *ctx);

/* ---- sizes and headers (host only; GetInfo / GetDecompressedSize never touch the GPU) -- */
B2B_API size_t b2b_max_frame_size(size_t n); /* 16 + n: the memcpy frame bounds every frame */
B2B_API int b2b_parse_header(const void *frame, size_t len, b2b_header *out);
B2B_API void b2b_header_bytes(const b2b_header *h, uint8_t out[16]);

/* ---- host-pointer, one frame: what compressBackend / decompressBackend bind ------------
 * b2b_compress: options are taken as CompressWithOptions leaves them (blosc.go:268-286):
 *   n == 0 -> B2B_EINVALID_DATA; typesize <= 0 -> 1; level is clamped and then unused by
 *   LZ4 (codec.go:63); header typesize is uint8(typesize).  dst needs b2b_max_frame_size(n).
 * b2b_decompress: check order of blosc.go:296-303,377-434; typesize_override as
 *   DecompressWithSize; dst needs nbytes_orig bytes (b2b_parse_header gives it). */
B2B_API int b2b_compress(b2b_ctx *ctx, const void *src, size_t n, int codec, int level, int shuffle,
                         int64_t typesize, void *dst, size_t cap, size_t *out_len);
B2B_API int b2b_decompress(b2b_ctx *ctx, const void *frame, size_t len, int64_t typesize_override,
                           void *dst, size_t cap, size_t *out_len);
/* whole-buffer filter, out of place or in place (src == dst), any n (64-bit), any typesize.
 * mode: B2B_SHUFFLE | B2B_BITSHUFFLE (anything else: plain copy, like shuffle.go:300-307). */
B2B_API int b2b_shuffle(b2b_ctx *ctx, int mode, int inverse, int64_t typesize, const void *src,
                        void *dst, size_t n);

/* ---- CodecInterface plugin seam (codec.go:15-38): raw LZ4 block, host pointers ----------
 * For RegisterCodec(LZ4, gpuCodec): Compress returns one LZ4 block (cap >= b2b_lz4_bound(n));
 * Decompress decodes into exactly expected_size bytes and reports how many were produced. */
B2B_API size_t b2b_lz4_bound(size_t n); /* n + n/255 + 16, as CompressBlockBound */
B2B_API int b2b_lz4_block_compress(b2b_ctx *ctx, const void *src, size_t n, void *dst, size_t cap,
                                   size_t *out_len);
B2B_API int b2b_lz4_block_decompress(b2b_ctx *ctx, const void *src, size_t n, void *dst,
                                     size_t expected_size, size_t *out_len);

/* ---- host-pointer batches (pipelined H2D -> kernels -> D2H) ---------
 * nframes independent frames; frame f is src[src_off[f] .. +src_len[f]).  Output frames are
 * packed back to back into dst; frame_off/frame_len (host arrays, nframes entries) receive
 * the packed-offsets table, status[f] the per-frame B2B_* status.  Pinned (cudaHostAlloc /
 * cudaHostRegister) buffers are DMA'd directly, pageable ones go through the context's pinned staging
 * ring.  The source ranges of a compress batch must not overlap (B2B_EINVAL): a frame is filtered in
 * scratch at its own offset.  The output slots of a decompress batch (dst_off[f], NBytesOrig of frame f)
 * may come in any order and with gaps; slots that overlap are refused (B2B_EINVAL); only the bytes a
 * frame produced are written to dst (gaps and the slots of failed frames keep the caller's bytes). */
B2B_API int b2b_compress_batch(b2b_ctx *ctx, const void *src, const uint64_t *src_off,
                               const uint32_t *src_len, uint32_t nframes, int shuffle,
                               int64_t typesize, void *dst, uint64_t dst_cap, uint64_t *frame_off,
                               uint32_t *frame_len, uint32_t *status, uint64_t *total_out);
B2B_API int b2b_decompress_batch(b2b_ctx *ctx, const void *frames, const uint64_t *frame_off,
                                 const uint32_t *frame_len, uint32_t nframes,
                                 int64_t typesize_override, void *dst, uint64_t dst_cap,
                                 const uint64_t *dst_off, uint32_t *out_len, uint32_t *status);

/* ---- device-pointer entry points (the measured path; `stream` is a cudaStream_t) ---------
 * All pointers prefixed d_ are device pointers.  Work is enqueued on `stream`; nothing is
 * synchronised unless the scratch arena has to grow (avoid with b2b_reserve).  The calls of one
 * context share one scratch arena: a call on a different stream than the previous call first waits
 * (on the device) for that call to finish; use one context per stream for concurrent batches. */
B2B_API int b2b_shuffle_dev(b2b_ctx *ctx, int mode, int inverse, int64_t typesize,
                            const void *d_src, void *d_dst, size_t n, void *stream);

/* Compress nframes frames.  total_src_bytes / max_frame_len are host-known bounds used to
 * size scratch: the extent max(d_src_off[f] + d_src_len[f]) and max(d_src_len[f]).
 * Output is PACKED: frame f is written at d_dst + d_frame_off[f] (exclusive scan of d_frame_len, built on the device by a
 * single-pass decoupled-look-back scan of the frame lengths ROUNDED UP TO 16: frames start on 16-byte
 * boundaries), d_total_out[0] = total bytes incl. that padding.  dst_cap must be
 * >= total_src_bytes + 31*nframes (16 header bytes + at most 15 of padding per frame) and d_dst 16-byte aligned.
 * d_status[f] is a B2B_* code.
 * The source ranges must not overlap (a frame is filtered in scratch at its own offset); a frame that
 * does not fit the scratch sized from total_src_bytes reports B2B_EDST_TOO_SMALL, nothing is written
 * out of bounds. */
B2B_API int b2b_compress_batch_dev(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                                   const uint32_t *d_src_len, uint32_t nframes,
                                   uint64_t total_src_bytes, uint32_t max_frame_len, int shuffle,
                                   int64_t typesize, void *d_dst, uint64_t dst_cap,
                                   uint64_t *d_frame_off, uint32_t *d_frame_len,
                                   uint32_t *d_status, uint64_t *d_total_out, void *stream);

/* Parse nframes headers on the device: d_orig_len[f] = NBytesOrig (0 on a bad header),
 * d_dst_off = exclusive scan of d_orig_len ROUNDED UP TO 16 (output slots start on 16-byte boundaries, which keeps the
 * un-shuffle on its vector path), d_total[0] = their sum (the bytes a destination laid out that way needs),
 * d_status[f] = header status. */
B2B_API int b2b_frame_info_batch_dev(b2b_ctx *ctx, const void *d_frames,
                                     const uint64_t *d_frame_off, const uint32_t *d_frame_len,
                                     uint32_t nframes, uint32_t *d_orig_len, uint64_t *d_dst_off,
                                     uint64_t *d_total, uint32_t *d_status, void *stream);

/* Decompress nframes frames; frame f's output goes to d_dst + d_dst_off[f] and must fit in
 * d_dst_cap[f] bytes (pass the d_orig_len from b2b_frame_info_batch_dev).
 * total_dst_bytes / max_orig_len are host-known bounds for scratch sizing. */
B2B_API int b2b_decompress_batch_dev(b2b_ctx *ctx, const void *d_frames,
                                     const uint64_t *d_frame_off, const uint32_t *d_frame_len,
                                     uint32_t nframes, int64_t typesize_override, void *d_dst,
                                     const uint64_t *d_dst_off, const uint32_t *d_dst_cap,
                                     uint64_t total_dst_bytes, uint32_t max_orig_len,
                                     uint32_t *d_out_len, uint32_t *d_status, void *stream);

/* ---- multi-GPU: all-gather of the per-frame sizes + global packed-offsets table ----------
 * Frames are sharded over the ranks with no collective on the data path; this is the one exchange, needed only
 * when a table over EVERY rank's frames is wanted (a device-resident multi-rank container).  nccl_comm is an
 * ncclComm_t of `world` ranks; every rank contributes n_local sizes (equal counts: pad with zeros).  d_all_len
 * receives world * n_local sizes in rank order, d_all_off their exclusive scan (of the sizes rounded up to 16 when
 * align16 != 0, as b2b_compress_batch_dev packs frames), d_total[0] the sum.  ONE ncclAllGather and one scan
 * kernel on `stream`, no host synchronisation.  libb2b.so does not link NCCL: ncclAllGather is resolved at run
 * time from the NCCL library already loaded in the process (the one that created nccl_comm), else from
 * libnccl.so.2; B2B_EUNSUPPORTED when there is none. */
B2B_API int b2b_allgather_sizes(b2b_ctx *ctx, void *nccl_comm, const uint32_t *d_local_len, uint32_t n_local,
                                uint32_t world, uint32_t *d_all_len, uint64_t *d_all_off, uint64_t *d_total,
                                int align16, void *stream);

/* ---- side-car decode index (no reference counterpart) ------------------
 * The reference's wire format has ONE LZ4 block per frame (blosc.go:393), so a frame decodes on one
 * warp however large it is.  b2b_compress_batch_dev_indexed produces the same standard frames
 * (every decoder, the reference's included, reads them) but compresses each 64 KiB segment
 * without reference to the others and also returns, outside the frames, an index of sequence
 * boundaries: d_index[f * segs_per_frame + s] = payload offset of the token that opens segment s
 * | output position of its first literal << 32, or ~0 when segment s opens no sequence.
 * segs_per_frame >= b2b_index_segments(max_frame_len).  b2b_decompress_batch_dev_indexed decodes
 * such frames with one warp per index entry (d_dst_cap[f] must be >= NBytesOrig); results are
 * identical to b2b_decompress_batch_dev.  The index is advisory: frames travel without it. */
B2B_API uint32_t b2b_index_segments(uint32_t max_frame_len);
B2B_API int b2b_compress_batch_dev_indexed(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                                           const uint32_t *d_src_len, uint32_t nframes,
                                           uint64_t total_src_bytes, uint32_t max_frame_len,
                                           int shuffle, int64_t typesize, void *d_dst,
                                           uint64_t dst_cap, uint64_t *d_frame_off,
                                           uint32_t *d_frame_len, uint32_t *d_status,
                                           uint64_t *d_total_out, uint64_t *d_index,
                                           uint32_t segs_per_frame, void *stream);
B2B_API int b2b_decompress_batch_dev_indexed(b2b_ctx *ctx, const void *d_frames,
                                             const uint64_t *d_frame_off, const uint32_t *d_frame_len,
                                             uint32_t nframes, int64_t typesize_override, void *d_dst,
                                             const uint64_t *d_dst_off, const uint32_t *d_dst_cap,
                                             uint64_t total_dst_bytes, uint32_t max_orig_len,
                                             uint32_t *d_out_len, uint32_t *d_status,
                                             const uint64_t *d_index, uint32_t segs_per_frame,
                                             void *stream);

/* ---- Blosc-1 multi-block frames (opt-in, no reference counterpart) ------
 * The reference declares Options.BlockSize and never reads it (blosc.go:227-234; every frame it
 * writes or reads is ONE block, blosc.go:320-434), so these frames stay behind their own entry
 * points and the reference cannot decode them.  Layout: the published Blosc-1 chunk -- 16-byte
 * header (version 2, versionlz 1, flags 0x1 shuffle / 0x2 stored / 0x4 bitshuffle / 0x10 not
 * split / format 1 = LZ4 in bits 5..7, typesize, nbytes, blocksize, cbytes), int32 bstarts[nblocks],
 * then per block an int32 size and one LZ4 block (size == block length: stored raw).  The filter
 * runs per block; the bit shuffle is the reference's own arrangement (shuffle.go:145-295), so
 * frames with flag 0x4 are private to this library.  The encoder never splits blocks into
 * per-byte streams (flag 0x10); the decoder also reads split blocks.  A frame is at most
 * nbytes + 16 bytes (stored when it would not be).  blocksize: 0 = 64 KiB, otherwise >= 128;
 * the block size used is b2b_blocks_blocksize() (whole elements, at most the buffer).
 * typesize outside 1..255 counts as 1.  n == 0: B2B_EINVALID_DATA; n > 2^31 - 17: B2B_EDATA_TOO_LARGE.
 * Batch calls have the conventions of b2b_compress_batch_dev / b2b_decompress_batch_dev; on
 * decompression `blocksize` is what the frames were written with (it sizes the block table:
 * a frame whose blocks are smaller than that block size rounded down to whole elements
 * reports B2B_EUNSUPPORTED). */
B2B_API uint32_t b2b_blocks_blocksize(size_t n, int64_t typesize, uint32_t blocksize);
B2B_API int b2b_compress_blocks(b2b_ctx *ctx, const void *src, size_t n, int shuffle, int64_t typesize,
                                uint32_t blocksize, void *dst, size_t cap, size_t *out_len);
B2B_API int b2b_decompress_blocks(b2b_ctx *ctx, const void *frame, size_t len, void *dst, size_t cap,
                                  size_t *out_len);
/* host-pointer batches: the same pipeline as b2b_compress_batch / b2b_decompress_batch */
B2B_API int b2b_compress_blocks_batch(b2b_ctx *ctx, const void *src, const uint64_t *src_off,
                                      const uint32_t *src_len, uint32_t nframes, int shuffle,
                                      int64_t typesize, uint32_t blocksize, void *dst, uint64_t dst_cap,
                                      uint64_t *frame_off, uint32_t *frame_len, uint32_t *status,
                                      uint64_t *total_out);
B2B_API int b2b_decompress_blocks_batch(b2b_ctx *ctx, const void *frames, const uint64_t *frame_off,
                                        const uint32_t *frame_len, uint32_t nframes, uint32_t blocksize,
                                        void *dst, uint64_t dst_cap, const uint64_t *dst_off,
                                        uint32_t *out_len, uint32_t *status);
B2B_API int b2b_compress_blocks_batch_dev(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                                          const uint32_t *d_src_len, uint32_t nframes,
                                          uint64_t total_src_bytes, uint32_t max_frame_len, int shuffle,
                                          int64_t typesize, uint32_t blocksize, void *d_dst,
                                          uint64_t dst_cap, uint64_t *d_frame_off, uint32_t *d_frame_len,
                                          uint32_t *d_status, uint64_t *d_total_out, void *stream);
B2B_API int b2b_decompress_blocks_batch_dev(b2b_ctx *ctx, const void *d_frames,
                                            const uint64_t *d_frame_off, const uint32_t *d_frame_len,
                                            uint32_t nframes, void *d_dst, const uint64_t *d_dst_off,
                                            const uint32_t *d_dst_cap, uint64_t total_dst_bytes,
                                            uint32_t max_orig_len, uint32_t blocksize,
                                            uint32_t *d_out_len, uint32_t *d_status, void *stream);

/* K5 on its own: d_off = exclusive scan of d_len (u32 -> u64), d_total[0] = sum. */
B2B_API int b2b_scan_offsets_dev(b2b_ctx *ctx, const uint32_t *d_len, uint32_t n, uint64_t *d_off,
                                 uint64_t *d_total, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2B_H */
