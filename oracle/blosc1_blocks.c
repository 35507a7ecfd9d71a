/*
 * blosc1_blocks.c -- CPU ORACLE for the opt-in Blosc-1 multi-block frame.
 *
 * TEST INFRASTRUCTURE ONLY (see blosc_oracle.h).
 *
 * The reference (go-blosc v1.0.2) writes and reads ONE block per frame and ignores
 * Options.BlockSize (blosc.go:227-234, 320-374), so this container has no reference
 * counterpart: PARITY UNPINNED.  What is restated here is the published Blosc-1 chunk layout
 * (c-blosc 1.x README_CHUNK_FORMAT.rst / blosc.c blosc_c, blosc_d, recalled; c-blosc is not in
 * this image):
 *
 *   bytes 0..15   header: version 2 | versionlz 1 | flags | typesize | nbytes | blocksize | cbytes
 *                 flags: 0x1 byte shuffle, 0x2 memcpyed, 0x4 bit shuffle, 0x10 blocks are not split,
 *                 bits 5..7 compressor format (1 = LZ4)
 *   bytes 16..    bstarts: int32[nblocks], offset of each block from the start of the frame
 *   blocks        nsplits streams each: int32 csize, then csize bytes; csize == stream length means
 *                 the stream is stored raw, anything else is one LZ4 block
 *   memcpyed      header, then the nbytes original bytes (no bstarts, no filter)
 *
 * The filter runs per block (not over the whole buffer as in the reference's frame): the byte
 * shuffle is Blosc's; the bit shuffle is the reference's own shuffle.go:145-295 arrangement
 * (MSB first), which differs from c-blosc's bitshuffle, so frames with flag 0x4 are private to
 * this implementation.  LZ4 streams come from the same restated compressor as the rest of the
 * oracle (orc_lz4_compress).
 */
#include "blosc_oracle.h"

#include <stdlib.h>
#include <string.h>

#define B1_MIN_BUFFERSIZE 128u   /* buffers below this are stored, streams below this are not split */
#define B1_MAX_SPLITS 16u
#define B1_MAX_BUFFERSIZE (0x7FFFFFFFu - 16u)

static uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static void wr32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

uint32_t orc_blocks_blocksize(size_t n, int64_t typesize, uint32_t requested) {
    uint64_t T = (typesize <= 0 || typesize > 255) ? 1u : (uint64_t)typesize;
    uint64_t b = requested ? requested : 65536u;
    if (b > n) b = n;
    if (b > T) b = b / T * T;      /* whole elements per block */
    if (b < T) b = T;              /* (a buffer shorter than one element keeps b > n: one block) */
    return (uint32_t)b;
}

static int splits_for(uint32_t flags, uint32_t T, uint32_t blocksize, int leftover_block) {
    if ((flags & ORC_B1_FLAG_DONTSPLIT) || leftover_block) return 1;
    if (T <= B1_MAX_SPLITS && blocksize / T >= B1_MIN_BUFFERSIZE) return (int)T;
    return 1;
}

int orc_blocks_compress(const uint8_t *data, size_t n, int shuffle, int64_t typesize,
                        uint32_t blocksize_req, int split, uint8_t *dst, size_t cap, size_t *out_len) {
    if (n == 0) return ORC_EINVALID_DATA;
    if (n > B1_MAX_BUFFERSIZE) return ORC_EDATA_TOO_LARGE;
    const uint32_t T = (typesize <= 0 || typesize > 255) ? 1u : (uint32_t)typesize;
    const uint32_t bs = orc_blocks_blocksize(n, T, blocksize_req);
    uint32_t flags = (uint32_t)ORC_B1_LZ4_FORMAT << 5;
    if (shuffle == ORC_SHUFFLE) flags |= ORC_FLAG_SHUFFLE;
    else if (shuffle == ORC_BITSHUFFLE) flags |= ORC_FLAG_BITSHUFFLE;
    if (!split) flags |= ORC_B1_FLAG_DONTSPLIT;
    const uint32_t nblocks = (uint32_t)((n + bs - 1) / bs);
    const uint32_t leftover = (uint32_t)(n % bs);

    /* worst case: every stream stored raw behind its 4-byte size */
    const size_t worst = 16 + 4ull * nblocks + n + 4ull * nblocks * B1_MAX_SPLITS;
    uint8_t *out = (uint8_t *)malloc(worst);
    uint8_t *tmp = (uint8_t *)malloc(bs);
    uint8_t *cbuf = (uint8_t *)malloc(orc_lz4_bound(bs));
    if (!out || !tmp || !cbuf) { free(out); free(tmp); free(cbuf); return ORC_ECOMPRESSION_FAILED; }

    size_t pos = 16 + 4ull * nblocks;
    int stored = n < B1_MIN_BUFFERSIZE;
    for (uint32_t b = 0; b < nblocks && !stored; b++) {
        const int last_partial = (b == nblocks - 1) && leftover > 0;
        const uint32_t bsize = last_partial ? leftover : bs;
        const uint8_t *in = data + (size_t)b * bs;
        if ((flags & ORC_FLAG_SHUFFLE) && T > 1) { orc_shuffle(in, tmp, bsize, T); in = tmp; }
        else if ((flags & ORC_FLAG_BITSHUFFLE) && T > 1) { orc_bitshuffle(in, tmp, bsize, T); in = tmp; }
        wr32(out + 16 + 4ull * b, (uint32_t)pos);
        const int ns = splits_for(flags, T, bs, last_partial);
        const uint32_t ne = bsize / (uint32_t)ns;
        for (int k = 0; k < ns; k++) {
            const uint8_t *s = in + (size_t)k * ne;
            size_t c = orc_lz4_compress(s, ne, cbuf, orc_lz4_bound(ne));
            if (c == 0 || c >= ne) { wr32(out + pos, ne); memcpy(out + pos + 4, s, ne); pos += 4 + (size_t)ne; }
            else { wr32(out + pos, (uint32_t)c); memcpy(out + pos + 4, cbuf, c); pos += 4 + c; }
        }
    }
    if (pos > n + 16) stored = 1;             /* Blosc's bound: a frame never exceeds nbytes + 16 */
    int rc = ORC_OK;
    size_t total = stored ? n + 16 : pos;
    if (cap < total) rc = ORC_EDST_TOO_SMALL;
    else {
        orc_header h = {2, 1, (uint8_t)(flags | (stored ? ORC_FLAG_MEMCPY : 0)), (uint8_t)T, (uint32_t)n, bs,
                        (uint32_t)total};
        if (stored) memcpy(dst + 16, data, n);
        else memcpy(dst, out, pos);
        orc_header_bytes(&h, dst);
        *out_len = total;
    }
    free(out); free(tmp); free(cbuf);
    return rc;
}

int orc_blocks_decompress(const uint8_t *frame, size_t len, uint8_t *dst, size_t cap, size_t *out_len) {
    orc_header h;
    int rc = orc_header_parse(frame, len, &h);
    if (rc) return rc;
    if ((size_t)h.nbytes_comp > len || h.nbytes_comp < 16) return ORC_EINVALID_DATA;
    const uint32_t n = h.nbytes_orig, bs = h.blocksize, T = h.typesize, flags = h.flags;
    if (n > B1_MAX_BUFFERSIZE) return ORC_EINVALID_DATA;
    if (n == 0) { *out_len = 0; return ORC_OK; }
    if (T == 0 || bs == 0) return ORC_EINVALID_DATA;
    if (flags & ORC_FLAG_MEMCPY) {
        if (h.nbytes_comp != n + 16) return ORC_ESIZE_MISMATCH;
        if (cap < n) return ORC_EDST_TOO_SMALL;
        memcpy(dst, frame + 16, n);
        *out_len = n;
        return ORC_OK;
    }
    const uint32_t fmt = flags >> 5;
    if (fmt > 4) return ORC_EINVALID_CODEC;
    if (fmt != ORC_B1_LZ4_FORMAT) return ORC_EUNSUPPORTED;   /* blosclz, snappy, zlib, zstd */
    if (h.versionlz != 1) return ORC_EINVALID_CODEC;
    const uint32_t nblocks = (uint32_t)(((uint64_t)n + bs - 1) / bs);
    const uint32_t leftover = n % bs;
    const uint64_t table_end = 16 + 4ull * nblocks;
    if (table_end > h.nbytes_comp) return ORC_EINVALID_DATA;
    if (cap < n) return ORC_EDST_TOO_SMALL;
    const uint32_t maxb = bs < n ? bs : n;
    uint8_t *tmp = (uint8_t *)malloc(maxb);
    if (!tmp) return ORC_EDECOMPRESSION_FAILED;
    for (uint32_t b = 0; b < nblocks; b++) {
        const int last_partial = (b == nblocks - 1) && leftover > 0;
        const uint32_t bsize = last_partial ? leftover : bs;
        uint64_t pos = rd32(frame + 16 + 4ull * b);
        if (pos < table_end) { free(tmp); return ORC_EDECOMPRESSION_FAILED; }
        const int ns = splits_for(flags, T, bs, last_partial);
        const uint32_t ne = bsize / (uint32_t)ns;
        for (int k = 0; k < ns; k++) {
            if (pos + 4 > h.nbytes_comp) { free(tmp); return ORC_EDECOMPRESSION_FAILED; }
            const uint32_t c = rd32(frame + pos);
            pos += 4;
            if (c == 0 || c > 0x7FFFFFFFu || pos + c > h.nbytes_comp) { free(tmp); return ORC_EDECOMPRESSION_FAILED; }
            if (c == ne) memcpy(tmp + (size_t)k * ne, frame + pos, ne);
            else if (orc_lz4_decompress(frame + pos, c, tmp + (size_t)k * ne, ne) != (int64_t)ne) {
                free(tmp); return ORC_EDECOMPRESSION_FAILED;
            }
            pos += c;
        }
        uint8_t *o = dst + (size_t)b * bs;
        if ((flags & ORC_FLAG_BITSHUFFLE) && T > 1) orc_bitunshuffle(tmp, o, bsize, T);
        else if ((flags & ORC_FLAG_SHUFFLE) && T > 1) orc_unshuffle(tmp, o, bsize, T);
        else memcpy(o, tmp, bsize);
    }
    free(tmp);
    *out_len = n;
    return ORC_OK;
}
