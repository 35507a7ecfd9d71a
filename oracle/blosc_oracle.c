/*
 * blosc_oracle.c -- CPU ORACLE (test infrastructure, never the product path).
 * See blosc_oracle.h for scope, citations and parity status ("parity unpinned" for LZ4
 * compressed bytes/sizes; pinned for everything else).
 */
#include "blosc_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

/* ------------------------------------------------------------------------------- */
/* Filters                                                                          */
/* ------------------------------------------------------------------------------- */

/* shuffle.go:16-19: typeSize <= 1 or len(src) < typeSize returns src itself. */
static int filter_is_noop(size_t n, int64_t T) { return T <= 1 || (uint64_t)n < (uint64_t)T; }

/* shuffle.go:60-70: dst[j*E+i] = src[i*T+j]; bytes past E*T copied raw. */
void orc_shuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t T) {
    if (filter_is_noop(n, T)) { memcpy(dst, src, n); return; }
    size_t t = (size_t)T, E = n / t;
    for (size_t j = 0; j < t; j++) {
        uint8_t *plane = dst + j * E;
        const uint8_t *s = src + j;
        for (size_t i = 0; i < E; i++) plane[i] = s[i * t];
    }
    memcpy(dst + E * t, src + E * t, n - E * t);
}

/* shuffle.go:120-130: dst[i*T+j] = src[j*E+i]. */
void orc_unshuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t T) {
    if (filter_is_noop(n, T)) { memcpy(dst, src, n); return; }
    size_t t = (size_t)T, E = n / t;
    for (size_t j = 0; j < t; j++) {
        const uint8_t *plane = src + j * E;
        uint8_t *d = dst + j;
        for (size_t i = 0; i < E; i++) d[i * t] = plane[i];
    }
    memcpy(dst + E * t, src + E * t, n - E * t);
}

/* One 8x8 bit-matrix step of shuffle.go:184-200 (and of 261-276, which is the same map):
 * out[k] bit (7-m) = in[m] bit (7-k). */
static inline void bit_transpose_8(const uint8_t in[8], uint8_t out[8]) {
    for (int k = 0; k < 8; k++) {
        uint8_t o = 0;
        for (int m = 0; m < 8; m++)
            if (in[m] & (uint8_t)(1u << (7 - k))) o |= (uint8_t)(1u << (7 - m));
        out[k] = o;
    }
}

/* shuffle.go:176-216 */
void orc_bitshuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t T) {
    if (filter_is_noop(n, T)) { memcpy(dst, src, n); return; }
    size_t t = (size_t)T, E = n / t, G = E / 8;
    for (size_t g = 0; g < G; g++) {
        size_t base = g * 8 * t;
        for (size_t j = 0; j < t; j++) {
            uint8_t b[8], o[8];
            for (int m = 0; m < 8; m++) b[m] = src[base + (size_t)m * t + j];
            bit_transpose_8(b, o);
            memcpy(dst + base + j * 8, o, 8);
        }
    }
    /* leftover elements (E % 8) and the n % T tail are copied untouched */
    memcpy(dst + G * 8 * t, src + G * 8 * t, n - G * 8 * t);
}

/* shuffle.go:253-292 */
void orc_bitunshuffle(const uint8_t *src, uint8_t *dst, size_t n, int64_t T) {
    if (filter_is_noop(n, T)) { memcpy(dst, src, n); return; }
    size_t t = (size_t)T, E = n / t, G = E / 8;
    for (size_t g = 0; g < G; g++) {
        size_t base = g * 8 * t;
        for (size_t j = 0; j < t; j++) {
            uint8_t s[8], o[8];
            memcpy(s, src + base + j * 8, 8);
            bit_transpose_8(s, o);
            for (int e = 0; e < 8; e++) dst[base + (size_t)e * t + j] = o[e];
        }
    }
    memcpy(dst + G * 8 * t, src + G * 8 * t, n - G * 8 * t);
}

/* --- AVX2 T=4 pair for the baseline arm (shuffle_amd64.s:183-226 / 285-322 process 8
 *     elements per step with VPSHUFB + VPERMD; this does the same with intrinsics) ----- */
#if defined(__x86_64__)
__attribute__((target("avx2"))) static size_t shuffle4_avx2(const uint8_t *src, uint8_t *dst,
                                                            size_t E) {
    const __m256i bytesel = _mm256_setr_epi8(0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15,
                                             0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15);
    const __m256i lanesel = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    size_t i = 0;
    for (; i + 8 <= E; i += 8) {
        __m256i v = _mm256_loadu_si256((const __m256i *)(src + i * 4));
        v = _mm256_shuffle_epi8(v, bytesel);         /* per 128-bit lane: p0 p1 p2 p3 dwords */
        v = _mm256_permutevar8x32_epi32(v, lanesel); /* p0lo p0hi p1lo p1hi ...            */
        uint64_t q[4];
        _mm256_storeu_si256((__m256i *)q, v);
        memcpy(dst + 0 * E + i, &q[0], 8);
        memcpy(dst + 1 * E + i, &q[1], 8);
        memcpy(dst + 2 * E + i, &q[2], 8);
        memcpy(dst + 3 * E + i, &q[3], 8);
    }
    return i;
}
__attribute__((target("avx2"))) static size_t unshuffle4_avx2(const uint8_t *src, uint8_t *dst,
                                                              size_t E) {
    const __m256i bytesel = _mm256_setr_epi8(0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15,
                                             0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15);
    const __m256i lanesel = _mm256_setr_epi32(0, 2, 4, 6, 1, 3, 5, 7);
    size_t i = 0;
    for (; i + 8 <= E; i += 8) {
        uint64_t q[4];
        memcpy(&q[0], src + 0 * E + i, 8);
        memcpy(&q[1], src + 1 * E + i, 8);
        memcpy(&q[2], src + 2 * E + i, 8);
        memcpy(&q[3], src + 3 * E + i, 8);
        __m256i v = _mm256_loadu_si256((const __m256i *)q); /* p0lo p0hi p1lo p1hi p2lo ... */
        v = _mm256_permutevar8x32_epi32(v, lanesel);        /* p0lo p1lo p2lo p3lo | hi...  */
        v = _mm256_shuffle_epi8(v, bytesel);                /* 4x4 byte transpose per lane  */
        _mm256_storeu_si256((__m256i *)(dst + i * 4), v);
    }
    return i;
}
static int cpu_has_avx2(void) { return __builtin_cpu_supports("avx2"); }
#else
static int cpu_has_avx2(void) { return 0; }
#endif

void orc_shuffle_fast(const uint8_t *src, uint8_t *dst, size_t n, int64_t T) {
#if defined(__x86_64__)
    if (T == 4 && n >= 32 && cpu_has_avx2()) {
        size_t E = n / 4, done = shuffle4_avx2(src, dst, E);
        for (size_t i = done; i < E; i++)
            for (size_t j = 0; j < 4; j++) dst[j * E + i] = src[i * 4 + j];
        memcpy(dst + E * 4, src + E * 4, n - E * 4);
        return;
    }
#endif
    orc_shuffle(src, dst, n, T);
}

void orc_unshuffle_fast(const uint8_t *src, uint8_t *dst, size_t n, int64_t T) {
#if defined(__x86_64__)
    if (T == 4 && n >= 32 && cpu_has_avx2()) {
        size_t E = n / 4, done = unshuffle4_avx2(src, dst, E);
        for (size_t i = done; i < E; i++)
            for (size_t j = 0; j < 4; j++) dst[i * 4 + j] = src[j * E + i];
        memcpy(dst + E * 4, src + E * 4, n - E * 4);
        return;
    }
#endif
    orc_unshuffle(src, dst, n, T);
}

/* ------------------------------------------------------------------------------- */
/* Header (blosc.go:154-198)                                                        */
/* ------------------------------------------------------------------------------- */
static uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static void wr32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

int orc_header_parse(const uint8_t *data, size_t len, orc_header *h) {
    if (len < 16) return ORC_EINVALID_HEADER;          /* blosc.go:166-168 */
    h->version = data[0]; h->versionlz = data[1]; h->flags = data[2]; h->typesize = data[3];
    h->nbytes_orig = rd32(data + 4); h->blocksize = rd32(data + 8); h->nbytes_comp = rd32(data + 12);
    if (h->version != 2) return ORC_EINVALID_VERSION;  /* blosc.go:180-182 */
    return ORC_OK;
}

void orc_header_bytes(const orc_header *h, uint8_t out[16]) {
    out[0] = h->version; out[1] = h->versionlz; out[2] = h->flags; out[3] = h->typesize;
    wr32(out + 4, h->nbytes_orig); wr32(out + 8, h->blocksize); wr32(out + 12, h->nbytes_comp);
}

/* ------------------------------------------------------------------------------- */
/* LZ4 block compressor: restatement of pierrec/lz4 v4 CompressBlock                 */
/* (called at codec.go:65-66).  PARITY UNPINNED -- see header comment.               */
/* ------------------------------------------------------------------------------- */
#define PZ_HASHLOG 16
#define PZ_HTSIZE (1u << PZ_HASHLOG)
#define PZ_WINSIZE 65536
#define PZ_MFLIMIT 14
#define PZ_SKIPLOG 7

size_t orc_lz4_bound(size_t n) { return n + n / 255 + 16; } /* CompressBlockBound, codec.go:65 */

static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t pz_hash(uint64_t x) {
    return (uint32_t)(((x << 16) * 227718039650203ULL) >> (64 - PZ_HASHLOG));
}

typedef struct {
    uint16_t table[PZ_HTSIZE];
    uint32_t inuse[PZ_HTSIZE / 32];
} pz_state;

/* candidate position for hash h seen from position si: the low 16 bits are relative to
 * the 64 KiB boundary at or before si, else to the one before; unset entries read as 0 */
static inline int64_t pz_get(const pz_state *c, uint32_t h, int64_t si) {
    int64_t i = 0;
    if (c->inuse[h >> 5] & (1u << (h & 31))) i = c->table[h];
    i += si & ~(int64_t)(PZ_WINSIZE - 1);
    if (i >= si) i -= PZ_WINSIZE;
    return i;
}
static inline void pz_put(pz_state *c, uint32_t h, int64_t si) {
    c->table[h] = (uint16_t)si;
    c->inuse[h >> 5] |= 1u << (h & 31);
}

static size_t put_len_ext(uint8_t *dst, size_t di, size_t v) { /* v already minus 15 */
    while (v >= 255) { dst[di++] = 255; v -= 255; }
    dst[di++] = (uint8_t)v;
    return di;
}

size_t orc_lz4_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
    if (cap < orc_lz4_bound(n)) return 0; /* the adapter always passes a bound-sized dst */
    pz_state *c = (pz_state *)malloc(sizeof(pz_state));
    if (!c) return 0;
    memset(c->inuse, 0, sizeof(c->inuse));
    int64_t si = 0, anchor = 0, sn = (int64_t)n - PZ_MFLIMIT;
    size_t di = 0;
    while (si < sn) {
        uint64_t match = ld64(src + si);
        uint32_t h = pz_hash(match), h2 = pz_hash(match >> 8);
        int64_t ref = pz_get(c, h, si), ref2 = pz_get(c, h2, si + 1);
        pz_put(c, h, si);
        pz_put(c, h2, si + 1);
        int64_t offset = si - ref;
        if (offset <= 0 || offset >= PZ_WINSIZE || (uint32_t)match != ld32(src + ref)) {
            h = pz_hash(match >> 16);
            int64_t ref3 = pz_get(c, h, si + 2);
            si += 1;
            offset = si - ref2;
            if (offset <= 0 || offset >= PZ_WINSIZE || (uint32_t)(match >> 8) != ld32(src + ref2)) {
                si += 1;
                offset = si - ref3;
                pz_put(c, h, si);
                if (offset <= 0 || offset >= PZ_WINSIZE ||
                    (uint32_t)(match >> 16) != ld32(src + ref3)) {
                    si += 2 + ((si - anchor) >> PZ_SKIPLOG);
                    continue;
                }
            }
        }
        /* match of >= 4 bytes at si, `offset` back */
        int64_t llen = si - anchor, mstart_len = 4;
        int64_t toff = si - offset - 1;
        while (llen > 0 && toff >= 0 && src[si - 1] == src[toff]) { si--; toff--; llen--; mstart_len++; }
        int64_t mbase = si + 4; /* encoded length counts from here */
        si += mstart_len;
        while (si + 8 <= sn) {
            uint64_t x = ld64(src + si) ^ ld64(src + si - offset);
            if (x == 0) { si += 8; } else { si += __builtin_ctzll(x) >> 3; break; }
        }
        int64_t mlen = si - mbase;
        uint8_t tok = (uint8_t)(mlen < 15 ? mlen : 15);
        size_t tokpos = di++;
        if (llen < 15) { tok |= (uint8_t)(llen << 4); }
        else { tok |= 0xF0; di = put_len_ext(dst, di, (size_t)llen - 15); }
        dst[tokpos] = tok;
        memcpy(dst + di, src + anchor, (size_t)llen);
        di += (size_t)llen;
        dst[di++] = (uint8_t)offset; dst[di++] = (uint8_t)(offset >> 8);
        if (mlen >= 15) di = put_len_ext(dst, di, (size_t)mlen - 15);
        anchor = si;
        if (si >= sn) break;
        pz_put(c, pz_hash(ld64(src + si - 2)), si - 2);
    }
    /* last literals */
    size_t llen = n - (size_t)anchor;
    if (llen < 15) { dst[di++] = (uint8_t)(llen << 4); }
    else { dst[di++] = 0xF0; di = put_len_ext(dst, di, llen - 15); }
    memcpy(dst + di, src + anchor, llen);
    di += llen;
    free(c);
    return di;
}

/* ------------------------------------------------------------------------------- */
/* LZ4 block decoder (public block format; pierrec UncompressBlock at codec.go:79)   */
/* Strict: any read past src, write past dst, zero offset, offset beyond the output  */
/* produced so far, or a stream that does not end right after a literal run with a    */
/* zero match nibble is an error (-1).  Empty src decodes to 0 bytes.                */
/* ------------------------------------------------------------------------------- */
int64_t orc_lz4_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
    if (n == 0) return 0;
    size_t si = 0, di = 0;
    for (;;) {
        if (si >= n) return -1;
        unsigned tok = src[si++];
        size_t ll = tok >> 4;
        if (ll == 15) {
            unsigned x;
            do {
                if (si >= n) return -1;
                x = src[si++];
                ll += x;
            } while (x == 255);
        }
        if (ll > n - si || ll > cap - di) return -1;
        memcpy(dst + di, src + si, ll);
        si += ll; di += ll;
        size_t ml = tok & 15;
        if (si == n) { if (ml != 0) return -1; break; }
        if (n - si < 2) return -1;
        size_t off = (size_t)src[si] | ((size_t)src[si + 1] << 8);
        si += 2;
        if (off == 0 || off > di) return -1;
        if (ml == 15) {
            unsigned x;
            do {
                if (si >= n) return -1;
                x = src[si++];
                ml += x;
            } while (x == 255);
        }
        ml += 4;
        if (ml > cap - di) return -1;
        const uint8_t *m = dst + di - off;
        uint8_t *d = dst + di;
        if (off >= 8) {
            size_t k = 0;
            for (; k + 8 <= ml; k += 8) memcpy(d + k, m + k, 8);
            for (; k < ml; k++) d[k] = m[k];
        } else {
            for (size_t k = 0; k < ml; k++) d[k] = m[k];
        }
        di += ml;
    }
    return (int64_t)di;
}

/* ------------------------------------------------------------------------------- */
/* Frames                                                                           */
/* ------------------------------------------------------------------------------- */
size_t orc_max_frame_size(size_t n) { return 16 + n; } /* memcpy bounds every frame */

int orc_compress(const uint8_t *data, size_t n, int codec, int level, int shuffle, int64_t T,
                 int memcpy_policy, uint8_t *dst, size_t cap, size_t *out_len) {
    (void)level; /* clamped at blosc.go:277-282 and then never read by the LZ4 adapter (codec.go:63) */
    if (n == 0) return ORC_EINVALID_DATA;                  /* blosc.go:269-271 */
    if (T <= 0) T = 1;                                     /* blosc.go:274-276 */
    if (codec < 0 || codec > 255) return ORC_EINVALID_CODEC;
    if (codec == ORC_BLOSCLZ || codec > ORC_ZSTD) return ORC_EINVALID_CODEC; /* blosc.go:322-325 */
    if (codec != ORC_LZ4) return ORC_EUNSUPPORTED;
    if (n > 0xFFFFFFFFu - 16u) return ORC_EDATA_TOO_LARGE;

    uint8_t *shuf = NULL;
    const uint8_t *input = data;
    if (shuffle == ORC_SHUFFLE && T > 1) {                 /* blosc.go:329-333 */
        shuf = (uint8_t *)malloc(n); if (!shuf) return ORC_ECOMPRESSION_FAILED;
        orc_shuffle(data, shuf, n, T); input = shuf;
    } else if (shuffle == ORC_BITSHUFFLE && T > 1) {
        shuf = (uint8_t *)malloc(n); if (!shuf) return ORC_ECOMPRESSION_FAILED;
        orc_bitshuffle(data, shuf, n, T); input = shuf;
    }
    size_t bound = orc_lz4_bound(n);
    uint8_t *cbuf = (uint8_t *)malloc(bound);
    if (!cbuf) { free(shuf); return ORC_ECOMPRESSION_FAILED; }
    size_t c = orc_lz4_compress(input, n, cbuf, bound);
    const uint8_t *payload = cbuf;
    int use_memcpy = (c == 0) || c >= n;                   /* codec.go:70-73, blosc.go:342-345 */
    if (use_memcpy) { c = n; payload = (memcpy_policy == ORC_MEMCPY_REF_QUIRK) ? data : input; }

    uint8_t flags = 0;                                     /* blosc.go:348-356 */
    if (shuffle == ORC_SHUFFLE) flags |= ORC_FLAG_SHUFFLE;
    else if (shuffle == ORC_BITSHUFFLE) flags |= ORC_FLAG_BITSHUFFLE;
    if (use_memcpy) flags |= ORC_FLAG_MEMCPY;

    int rc = ORC_OK;
    if (cap < 16 + c) { rc = ORC_EDST_TOO_SMALL; }
    else {
        orc_header h = {2, (uint8_t)codec, flags, (uint8_t)T, (uint32_t)n, (uint32_t)n,
                        (uint32_t)(16 + c)};               /* blosc.go:358-366 */
        orc_header_bytes(&h, dst);
        memcpy(dst + 16, payload, c);
        *out_len = 16 + c;
    }
    free(cbuf); free(shuf);
    return rc;
}

static int codec_registered(unsigned id) { return id >= ORC_LZ4 && id <= ORC_ZSTD; } /* codec.go:27-33 */

int orc_decompress(const uint8_t *frame, size_t len, int64_t T_override, uint8_t *dst, size_t cap,
                   size_t *out_len) {
    orc_header h;
    int rc = orc_header_parse(frame, len, &h);             /* blosc.go:297-299, 379-382 */
    if (rc) return rc;
    if ((size_t)h.nbytes_comp > len) return ORC_EINVALID_DATA;  /* blosc.go:385-387 */
    if (h.nbytes_comp < 16) return ORC_EINVALID_DATA;           /* blosc.go:388-390 */
    const uint8_t *payload = frame + 16;
    size_t plen = (size_t)h.nbytes_comp - 16;
    size_t n = h.nbytes_orig, got;
    uint8_t *tmp = NULL;

    if (h.flags & ORC_FLAG_MEMCPY) {                       /* blosc.go:398-400 */
        got = plen;
        tmp = (uint8_t *)malloc(got ? got : 1); if (!tmp) return ORC_EDECOMPRESSION_FAILED;
        memcpy(tmp, payload, got);
    } else {
        if (!codec_registered(h.versionlz)) return ORC_EINVALID_CODEC;  /* blosc.go:403-407 */
        if (h.versionlz != ORC_LZ4 && h.versionlz != ORC_LZ4HC) return ORC_EUNSUPPORTED;
        tmp = (uint8_t *)malloc(n ? n : 1); if (!tmp) return ORC_EDECOMPRESSION_FAILED;
        int64_t d = orc_lz4_decompress(payload, plen, tmp, n);  /* codec.go:77-84 */
        if (d < 0) { free(tmp); return ORC_EDECOMPRESSION_FAILED; }   /* blosc.go:410-413 */
        got = (size_t)d;
    }
    int64_t T = T_override > 0 ? T_override : (int64_t)h.typesize;    /* blosc.go:417-419 */
    /* blosc.go:422-431: unshuffle, then the size check.  A short result fails the size
     * check whatever the unshuffle did, so check first and transform only good frames. */
    if (got != n) { free(tmp); return ORC_ESIZE_MISMATCH; }
    if (cap < n) { free(tmp); return ORC_EDST_TOO_SMALL; }
    if ((h.flags & ORC_FLAG_BITSHUFFLE) && T > 1) orc_bitunshuffle(tmp, dst, n, T);
    else if ((h.flags & ORC_FLAG_SHUFFLE) && T > 1) orc_unshuffle(tmp, dst, n, T);
    else memcpy(dst, tmp, n);
    free(tmp);
    *out_len = n;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------- */
/* Multi-threaded batch drivers for the CPU baseline arm                            */
/* ------------------------------------------------------------------------------- */
typedef struct {
    int kind; /* 0 compress, 1 decompress, 2 shuffle slice */
    const uint8_t *src; uint8_t *dst;
    const uint64_t *in_off; const uint32_t *in_len; const uint64_t *out_off; uint32_t *out_len;
    uint32_t nframes; int shuffle; int64_t T; int fast;
    atomic_uint next; atomic_int status;
    /* shuffle slice */
    int mode, inverse; size_t n; int nthreads;
} mt_job;

static void filter_apply(int shuffle, int inverse, int fast, const uint8_t *s, uint8_t *d, size_t n,
                         int64_t T) {
    if (shuffle == ORC_SHUFFLE) {
        if (inverse) (fast ? orc_unshuffle_fast : orc_unshuffle)(s, d, n, T);
        else (fast ? orc_shuffle_fast : orc_shuffle)(s, d, n, T);
    } else if (shuffle == ORC_BITSHUFFLE) {
        if (inverse) orc_bitunshuffle(s, d, n, T); else orc_bitshuffle(s, d, n, T);
    } else memcpy(d, s, n);
}

/* Same frame logic as orc_compress/orc_decompress, with per-thread scratch reuse and the
 * optional AVX2 shuffle, so the baseline is not handicapped by malloc per frame. */
static void *mt_worker(void *arg) {
    mt_job *J = (mt_job *)arg;
    uint8_t *shuf = NULL, *cbuf = NULL; size_t shuf_cap = 0, cbuf_cap = 0;
    for (;;) {
        uint32_t f = atomic_fetch_add(&J->next, 1);
        if (f >= J->nframes) break;
        size_t n = J->in_len[f];
        const uint8_t *in = J->src + J->in_off[f];
        uint8_t *out = J->dst + J->out_off[f];
        int rc = ORC_OK;
        if (J->kind == 0) {
            int64_t T = J->T <= 0 ? 1 : J->T;
            if (n == 0) { rc = ORC_EINVALID_DATA; goto done; }
            if (shuf_cap < n) { free(shuf); shuf = (uint8_t *)malloc(n); shuf_cap = n; }
            size_t bound = orc_lz4_bound(n);
            if (cbuf_cap < bound) { free(cbuf); cbuf = (uint8_t *)malloc(bound); cbuf_cap = bound; }
            const uint8_t *input = in;
            if ((J->shuffle == ORC_SHUFFLE || J->shuffle == ORC_BITSHUFFLE) && T > 1) {
                filter_apply(J->shuffle, 0, J->fast, in, shuf, n, T); input = shuf;
            }
            size_t c = orc_lz4_compress(input, n, cbuf, bound);
            const uint8_t *payload = cbuf;
            uint8_t flags = J->shuffle == ORC_SHUFFLE ? ORC_FLAG_SHUFFLE
                          : J->shuffle == ORC_BITSHUFFLE ? ORC_FLAG_BITSHUFFLE : 0;
            if (c == 0 || c >= n) { c = n; payload = input; flags |= ORC_FLAG_MEMCPY; }
            orc_header h = {2, ORC_LZ4, flags, (uint8_t)T, (uint32_t)n, (uint32_t)n, (uint32_t)(16 + c)};
            orc_header_bytes(&h, out);
            memcpy(out + 16, payload, c);
            J->out_len[f] = (uint32_t)(16 + c);
        } else {
            orc_header h;
            rc = orc_header_parse(in, n, &h);
            if (rc) goto done;
            if (h.nbytes_comp > n || h.nbytes_comp < 16) { rc = ORC_EINVALID_DATA; goto done; }
            size_t no = h.nbytes_orig, plen = h.nbytes_comp - 16, got;
            int64_t T = h.typesize;
            int filt = (h.flags & ORC_FLAG_BITSHUFFLE) ? ORC_BITSHUFFLE
                     : (h.flags & ORC_FLAG_SHUFFLE) ? ORC_SHUFFLE : ORC_NOSHUFFLE;
            int need = filt != ORC_NOSHUFFLE && T > 1;
            if (shuf_cap < no) { free(shuf); shuf = (uint8_t *)malloc(no ? no : 1); shuf_cap = no; }
            uint8_t *stage = need ? shuf : out;
            if (h.flags & ORC_FLAG_MEMCPY) {
                if (plen != no) { rc = ORC_ESIZE_MISMATCH; goto done; }
                memcpy(stage, in + 16, plen); got = plen;
            } else {
                int64_t d = orc_lz4_decompress(in + 16, plen, stage, no);
                if (d < 0) { rc = ORC_EDECOMPRESSION_FAILED; goto done; }
                got = (size_t)d;
            }
            if (got != no) { rc = ORC_ESIZE_MISMATCH; goto done; }
            if (need) filter_apply(filt, 1, J->fast, shuf, out, no, T);
            J->out_len[f] = (uint32_t)no;
        }
    done:
        if (rc) { int z = 0; atomic_compare_exchange_strong(&J->status, &z, rc); }
    }
    free(shuf); free(cbuf);
    return NULL;
}

static int run_mt(mt_job *J, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    atomic_store(&J->next, 0); atomic_store(&J->status, 0);
    int started = 0;
    for (int i = 0; i < threads; i++)
        if (pthread_create(&th[i], NULL, mt_worker, J) == 0) started++; else break;
    if (started == 0) mt_worker(J);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    return atomic_load(&J->status);
}

int orc_compress_batch_mt(const uint8_t *src, const uint64_t *src_off, const uint32_t *src_len,
                          uint32_t nframes, int shuffle, int64_t T, uint8_t *dst,
                          const uint64_t *dst_off, uint32_t *dst_len, int threads, int fast) {
    mt_job J; memset(&J, 0, sizeof J);
    J.kind = 0; J.src = src; J.dst = dst; J.in_off = src_off; J.in_len = src_len;
    J.out_off = dst_off; J.out_len = dst_len; J.nframes = nframes; J.shuffle = shuffle; J.T = T;
    J.fast = fast;
    return run_mt(&J, threads);
}

int orc_decompress_batch_mt(const uint8_t *frames, const uint64_t *frame_off,
                            const uint32_t *frame_len, uint32_t nframes, uint8_t *dst,
                            const uint64_t *dst_off, uint32_t *out_len, int threads, int fast) {
    mt_job J; memset(&J, 0, sizeof J);
    J.kind = 1; J.src = frames; J.dst = dst; J.in_off = frame_off; J.in_len = frame_len;
    J.out_off = dst_off; J.out_len = out_len; J.nframes = nframes; J.fast = fast;
    return run_mt(&J, threads);
}

/* Whole-buffer filter.  The reference runs this on one goroutine (shuffle.go:298-323);
 * `threads` > 1 splits the element range, which is the most favourable CPU arrangement. */
typedef struct { int mode, inverse, fast; int64_t T; const uint8_t *src; uint8_t *dst; size_t n, lo, hi; } slice_job;

static void *slice_worker(void *arg) {
    slice_job *S = (slice_job *)arg;
    size_t t = (size_t)S->T, E = S->n / t;
    if (S->mode == ORC_SHUFFLE) {
        for (size_t j = 0; j < t; j++)
            for (size_t i = S->lo; i < S->hi; i++) {
                if (!S->inverse) S->dst[j * E + i] = S->src[i * t + j];
                else S->dst[i * t + j] = S->src[j * E + i];
            }
    } else { /* bitshuffle: lo/hi are group indices */
        size_t bytes = (S->hi - S->lo) * 8 * t, off = S->lo * 8 * t;
        if (!S->inverse) orc_bitshuffle(S->src + off, S->dst + off, bytes, S->T);
        else orc_bitunshuffle(S->src + off, S->dst + off, bytes, S->T);
    }
    return NULL;
}

int orc_shuffle_mt(int mode, int inverse, int64_t T, const uint8_t *src, uint8_t *dst, size_t n,
                   int threads, int fast) {
    if ((mode != ORC_SHUFFLE && mode != ORC_BITSHUFFLE) || filter_is_noop(n, T)) {
        memcpy(dst, src, n); return ORC_OK;
    }
    if (threads <= 1) { filter_apply(mode, inverse, fast, src, dst, n, T); return ORC_OK; }
    if (threads > 1024) threads = 1024;
    size_t t = (size_t)T, E = n / t;
    size_t units = mode == ORC_SHUFFLE ? E : E / 8;
    size_t covered = mode == ORC_SHUFFLE ? E * t : (E / 8) * 8 * t;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    slice_job *jobs = (slice_job *)malloc(sizeof(slice_job) * (size_t)threads);
    int started = 0;
    for (int i = 0; i < threads; i++) {
        slice_job s = {mode, inverse, fast, T, src, dst, n, units * (size_t)i / (size_t)threads,
                       units * (size_t)(i + 1) / (size_t)threads};
        jobs[i] = s;
        if (pthread_create(&th[i], NULL, slice_worker, &jobs[i]) == 0) started++;
        else { slice_worker(&jobs[i]); th[i] = 0; }
    }
    for (int i = 0; i < threads; i++) if (th[i]) pthread_join(th[i], NULL);
    (void)started;
    memcpy(dst + covered, src + covered, n - covered);
    free(th); free(jobs);
    return ORC_OK;
}
