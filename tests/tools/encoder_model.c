/* encoder_model.c -- design-exploration model of the warp-cooperative LZ4 encoder (K3).
 *
 * NOT part of the product and not a fallback: it replays the kernel's parse on the CPU, lane
 * by lane, so that match-finder parameters (hash width, table size, skip schedule) can be
 * compared with the oracle's pierrec-style compressor without spending GPU time.
 * Build: gcc -O2 -o /tmp/encoder_model tests/tools/encoder_model.c oracle/blosc_oracle.c -lm -pthread
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/blosc_oracle.h"

static int INSERT_AFTER = 0, LAZY = 0, WAYS = 1, LAZYCAP = 1 << 30, SEG = 0, FIXD = 0, NOBACK = 0; static int HASHLOG = 12, HASHBYTES = 4, SKIPLOG = 7, SKIPDIV = 3, INSERT_END = 0, PREFER_TABLE = 0;

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t hashf(const uint8_t *p) {
    if (HASHBYTES == 4) return (ld32(p) * 2654435761u) >> (32 - HASHLOG);
    uint64_t x = ld64(p);
    if (HASHBYTES == 5) return (uint32_t)(((x << 24) * 889523592379ULL) >> (64 - HASHLOG));
    return (uint32_t)(((x << 16) * 227718039650203ULL) >> (64 - HASHLOG));
}

static size_t put_ext(uint8_t *o, size_t v) { size_t k = 0; while (v >= 255) { o[k++] = 255; v -= 255; } o[k++] = (uint8_t)v; return k; }

size_t model_encode(const uint8_t *src, uint32_t n, uint8_t *out, long *nseq) {
    uint32_t op = 0, anchor = 0;
    uint16_t *table = calloc(1u << HASHLOG, 2);
    *nseq = 0;
    if (n > 14) {
        const uint32_t mfl = n - 14, mlimit = n - 14;
        uint32_t si = 0;
        while (si < mfl) {
            uint32_t lits = si - anchor;
            uint32_t stride = (4u + (lits >> SKIPLOG)) / SKIPDIV; if (stride < 1) stride = 1;
            uint32_t p[32], seq[32], h[32], c16[32]; int valid[32]; int first_hit = 64;
            for (int l = 0; l < 32; l++) {
                uint64_t p64 = (uint64_t)si + (uint64_t)l * stride;
                valid[l] = p64 < mfl; p[l] = valid[l] ? (uint32_t)p64 : 0;
                seq[l] = valid[l] ? ld32(src + p[l]) : 0; h[l] = valid[l] ? hashf(src + p[l]) : 0; c16[l] = table[h[l]];
            }
            int first = -1; int64_t cand_f = 0; int best_score = -1000000;
            for (int l = 0; l < 32 && (first < 0 || (stride == 1 && l <= first_hit + LAZY)); l++) {
                if (!valid[l]) continue;
                int ok = 0; int64_t cand = 0;
                int srcl = -1;
                for (int m = l - 1; m >= 0; m--) if (valid[m] && seq[m] == seq[l]) { srcl = m; break; }
                int64_t ctab = (int64_t)((p[l] & ~0xFFFFu) + c16[l]); if (ctab >= (int64_t)p[l]) ctab -= 65536;
                int tab_ok = ctab >= 0 && ((int64_t)p[l] - ctab) < 65536 && ld32(src + ctab) == seq[l];
                if (srcl >= 0) { uint64_t dist = (uint64_t)(l - srcl) * stride; cand = (int64_t)p[l] - (int64_t)dist; ok = dist < 65536; }
                if (PREFER_TABLE && tab_ok) { cand = ctab; ok = 1; }
                if (!ok && tab_ok) { cand = ctab; ok = 1; }
                if (!ok && FIXD > 0 && p[l] >= (uint32_t)FIXD && ld32(src + p[l] - FIXD) == seq[l]) { cand = (int64_t)p[l] - FIXD; ok = 1; }
                if (ok) {
                    if (first < 0) first_hit = l;
                    uint32_t e = p[l] + 4, cc = (uint32_t)cand + 4;
                    while (e < mlimit && src[e] == src[cc]) { e++; cc++; }
                    int len_ = (int)(e - p[l]); if (len_ > LAZYCAP) len_ = LAZYCAP; int score = len_ - (int)(l - first_hit) * 1;
                    if (LAZY == 0) { first = l; cand_f = cand; }
                    else if (score > best_score + (first >= 0 ? 0 : 0)) { best_score = score; first = l; cand_f = cand; }
                }
            }
            /* positions up to the match start are recorded (last lane with the same 4 bytes wins) */
            for (int l = 0; l < 32; l++) {
                if (!valid[l] || (first >= 0 && l > first + INSERT_AFTER)) continue;
                table[h[l]] = (uint16_t)p[l];
            }
            if (first < 0) { uint64_t nx = (uint64_t)si + 32ull * stride; si = nx < mfl ? (uint32_t)nx : mfl; continue; }
            uint32_t mp = p[first], mc = (uint32_t)cand_f, offset = mp - mc;
            uint32_t mend = mp + 4, cpos = mc + 4;
            while (mend < mlimit && src[mend] == src[cpos]) { mend++; cpos++; }
            if (!(NOBACK && stride == 1)) while (mp > anchor && mc > 0 && src[mp - 1] == src[mc - 1]) { mp--; mc--; }
            uint32_t ll = mp - anchor, ml = mend - mp - 4;
            uint32_t tok = op++;
            if (ll >= 15) op += put_ext(out + op, ll - 15);
            memcpy(out + op, src + anchor, ll); op += ll;
            out[tok] = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (ml < 15 ? ml : 15));
            out[op++] = (uint8_t)offset; out[op++] = (uint8_t)(offset >> 8);
            if (ml >= 15) op += put_ext(out + op, ml - 15);
            if (INSERT_END && mend >= 2 && mend - 2 < mfl) table[hashf(src + mend - 2)] = (uint16_t)(mend - 2);
            si = mend; anchor = mend; (*nseq)++;
        }
    }
    uint32_t ll = n - anchor; uint32_t tok = op++;
    out[tok] = (uint8_t)((ll < 15 ? ll : 15) << 4);
    if (ll >= 15) op += put_ext(out + op, ll - 15);
    memcpy(out + op, src + anchor, ll); op += ll;
    free(table);
    return op;
}

/* independent segments: matches never leave the segment; trailing literals of a segment are
 * carried into the first sequence of the next one (what the stitch pass does on the GPU) */
size_t model_encode_seg(const uint8_t *src, uint32_t n, uint8_t *out, long *nseq) {
    if (!SEG || SEG >= (int)n) return model_encode(src, n, out, nseq);
    uint8_t *tmp = malloc(2 * SEG + 64);
    size_t op = 0; uint32_t carry_start = 0; *nseq = 0;
    for (uint32_t b = 0; b < n; b += SEG) {
        uint32_t len = n - b < (uint32_t)SEG ? n - b : (uint32_t)SEG; int last = b + len >= n;
        /* encode segment; for non-last segments matches may run to the very end of the segment */
        long ns = 0; size_t c;
        if (last) c = model_encode(src + b, len, tmp, &ns);
        else { /* emulate 'no end-of-block rules' by encoding len+14 bytes worth of limit: pad view */
            uint8_t *pad = malloc(len + 14); memcpy(pad, src + b, len); memset(pad + len, 0xA5, 14);
            c = model_encode(pad, len + 14, tmp, &ns); free(pad);
        }
        /* re-parse the segment stream and re-emit with the carry merged into the first sequence */
        size_t ip = 0; uint32_t pos = b; int firstseq = 1;
        while (ip < c) {
            uint32_t tok = tmp[ip++]; uint32_t ll = tok >> 4;
            if (ll == 15) { uint32_t x; do { x = tmp[ip++]; ll += x; } while (x == 255); }
            ip += ll;
            if (ip >= c) { /* trailing literals: carried (minus the 14 pad bytes) */ break; }
            uint32_t off = tmp[ip] | (tmp[ip + 1] << 8); ip += 2; uint32_t ml = tok & 15;
            if (ml == 15) { uint32_t x; do { x = tmp[ip++]; ml += x; } while (x == 255); }
            ml += 4;
            uint32_t mstart = pos + ll;
            if (mstart + ml > b + len) { /* clipped by the pad; should not happen (pad never matches) */ ml = b + len - mstart; }
            uint32_t L = mstart - carry_start; (void)firstseq;
            size_t t = op++; uint32_t mlc = ml - 4;
            if (L >= 15) op += put_ext(out + op, L - 15);
            memcpy(out + op, src + carry_start, L); op += L;
            out[t] = (uint8_t)(((L < 15 ? L : 15) << 4) | (mlc < 15 ? mlc : 15));
            out[op++] = (uint8_t)off; out[op++] = (uint8_t)(off >> 8);
            if (mlc >= 15) op += put_ext(out + op, mlc - 15);
            pos = mstart + ml; carry_start = pos; firstseq = 0; (*nseq)++;
        }
    }
    uint32_t L = n - carry_start; size_t t = op++;
    out[t] = (uint8_t)((L < 15 ? L : 15) << 4);
    if (L >= 15) op += put_ext(out + op, L - 15);
    memcpy(out + op, src + carry_start, L); op += L;
    free(tmp);
    return op;
}

static uint64_t sm64(uint64_t *s) { uint64_t z = (*s += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }

int main(int argc, char **argv) {
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "hashlog")) HASHLOG = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "hashbytes")) HASHBYTES = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "skiplog")) SKIPLOG = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "skipdiv")) SKIPDIV = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "insert_end")) INSERT_END = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "prefer_table")) PREFER_TABLE = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "insert_after")) INSERT_AFTER = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "lazy")) LAZY = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "ways")) WAYS = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "lazycap")) LAZYCAP = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "seg")) SEG = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "fixd")) FIXD = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "noback")) NOBACK = atoi(argv[i + 1]);
    }
    const uint32_t n = 262144;
    uint8_t *raw = malloc(n), *sh = malloc(n), *o1 = malloc(n * 2), *o2 = malloc(n * 2), *back = malloc(n);
    const char *names[] = {"smooth_f32+S4", "smooth_f64+B8", "lowent_i16+S2", "ramp+S4", "f32ramp.001+S4", "text", "smooth_f32+B4", "smooth_f64+S8"};
    printf("hashlog=%d hashbytes=%d skiplog=%d skipdiv=%d insert_end=%d prefer_table=%d\n", HASHLOG, HASHBYTES, SKIPLOG, SKIPDIV, INSERT_END, PREFER_TABLE);
    int NREP = getenv("MODEL_REPS") ? atoi(getenv("MODEL_REPS")) : 1;
    for (int k = 0; k < 8; k++) {
      size_t tot1 = 0, tot2 = 0; long totseq = 0; int allok = 1;
      for (int rep = 0; rep < NREP; rep++) {
        uint64_t s = 1234 + k + 7919ull * rep; uint32_t base = rep * 65536u * 3u + rep * 1237u;
        if (k == 0 || k == 6) { float *f = (float *)raw; for (uint32_t i = 0; i < n / 4; i++) { double u = (double)(sm64(&s) >> 11) / 9007199254740992.0 * 2 - 1; f[i] = (float)(sin(2 * M_PI * (i + base) / 4096) + 0.25 * sin(2 * M_PI * (i + base) / 333.3) + 1e-3 * u); } }
        if (k == 1 || k == 7) { double *f = (double *)raw; for (uint32_t i = 0; i < n / 8; i++) { double u = (double)(sm64(&s) >> 11) / 9007199254740992.0 * 2 - 1; f[i] = sin(2 * M_PI * (i + base) / 4096) + 0.25 * sin(2 * M_PI * (i + base) / 333.3) + 1e-3 * u; } }
        if (k == 2) { int16_t *f = (int16_t *)raw; for (uint32_t i = 0; i < n / 2; i++) f[i] = (int16_t)(sm64(&s) & 7); }
        if (k == 3) for (uint32_t i = 0; i < n; i++) raw[i] = (uint8_t)i;
        if (k == 4) { float *f = (float *)raw; for (uint32_t i = 0; i < n / 4; i++) f[i] = (float)(i + base) * 0.001f; }
        if (k == 5) { uint32_t i = 0; while (i < n) { uint64_t w = sm64(&s) % 200; uint64_t ws = w * 7919; int len = 2 + (int)(ws % 7); for (int j = 0; j < len && i < n; j++) raw[i++] = (uint8_t)('a' + (ws >> (j * 3)) % 26); if (i < n) raw[i++] = ' '; } }
        if (k == 0 || k == 3 || k == 4) orc_shuffle(raw, sh, n, 4);
        else if (k == 1) orc_bitshuffle(raw, sh, n, 8);
        else if (k == 2) orc_shuffle(raw, sh, n, 2);
        else if (k == 6) orc_bitshuffle(raw, sh, n, 4);
        else if (k == 7) orc_shuffle(raw, sh, n, 8);
        else memcpy(sh, raw, n);
        long nseq = 0;
        size_t c1 = model_encode_seg(sh, n, o1, &nseq);
        size_t c2 = orc_lz4_compress(sh, n, o2, orc_lz4_bound(n));
        int64_t d = orc_lz4_decompress(o1, c1, back, n);
        int okk = d == (int64_t)n && !memcmp(back, sh, n);
        if (getenv("MODEL_DUMP")) { char fn[64]; sprintf(fn, "/tmp/model_%d.bin", k); FILE *f = fopen(fn, "wb"); fwrite(o1, 1, c1, f); fclose(f);
            sprintf(fn, "/tmp/oracle_%d.bin", k); f = fopen(fn, "wb"); fwrite(o2, 1, c2, f); fclose(f); }
        tot1 += c1; tot2 += c2; totseq += nseq; allok &= okk;
      }
      printf("%-16s model %8zu  oracle %8zu  ratio %.4f  seqs/frame %6ld  valid=%d\n", names[k], tot1, tot2, (double)tot1 / tot2, totseq / NREP, allok);
    }
    return 0;
}
