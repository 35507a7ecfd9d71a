/* encoder_model.c -- design-exploration model of the warp-cooperative LZ4 encoder (K3).
 *
 * NOT part of the product and not a fallback: it replays the kernel's parse on the CPU, lane
 * by lane, so that match-finder parameters (hash width, table size, skip schedule) can be
 * compared with the oracle's pierrec-style compressor without spending GPU time.
 * Build: gcc -O2 -o /tmp/strip_model tests/tools/strip_model.c oracle/blosc_oracle.c -lm -pthread
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


static int MINLL = 0; static int STRIP = 64, CAPX = 2, SM = 0, NOTRIMBACK = 0, BACKX = 0, REP = 0, GROUP = 4, WAYS2 = 0, UPTO = 0, INGRP = 0, MINM = 4, MINM_TABLE_ONLY = 0, PHASED = 0, RECMAX = 16, HALVES = 1, SKIPCONT = 0, NOSHIFT = 0, REPFIRST = 0, LCMP = 16, SHORTRULE = 0, SHORTLEN = 6, SHORTLL = 15, REPWIN = 0;
static double COST_A = 0, COST_B = 0; static long NSTEPS = 0, NOPEN = 0, NDROP = 0, NTRIM = 0;
typedef struct { uint32_t ms, me, off; int open; } mt_t;

size_t model_encode(const uint8_t *src, uint32_t n, uint8_t *out, long *nseq) {
    uint32_t op = 0, anchor = 0;
    uint32_t *table = calloc(1u << HASHLOG, 4);   /* chk<<17 | pos, pos < 2^17 */
    *nseq = 0;
    if (n > 14) {
        const uint32_t mfl = n - 14, mlimit = n - 14;
        uint32_t si = 0;
        static mt_t lists[32][128]; int cnt[32];
        while (si < mfl) {
            uint32_t pos[32], send[32], lanch[32];
            for (int l = 0; l < 32; l++) { uint64_t a = (uint64_t)si + (uint64_t)l * STRIP; pos[l] = a < mfl ? (uint32_t)a : mfl; uint64_t e = a + STRIP; send[l] = e < mfl ? (uint32_t)e : mfl; { uint32_t lo = si > anchor ? si : anchor; lanch[l] = pos[l] > lo + BACKX ? pos[l] - BACKX : lo; if (lanch[l] > pos[l]) lanch[l] = pos[l]; } cnt[l] = 0; }
            uint32_t region_end = send[31];
            int lane_iters[32] = {0}; double costA = 0; static uint32_t rep[32];
            if (PHASED) {
                static uint32_t rpos[32][64], rcand[32][64]; int rn[32];
                int lastk[32]; uint32_t lastoff[32]; for (int l = 0; l < 32; l++) { lastk[l] = -2; lastoff[l] = 0; }
                uint32_t pe_l[32]; for (int l = 0; l < 32; l++) pe_l[l] = lanch[l];
                int nturns = (STRIP + 3) / 4; int HT = HALVES > 1 ? (nturns + HALVES - 1) / HALVES : nturns;
                for (int t0 = 0; t0 < nturns; t0 += HT) {
                for (int l = 0; l < 32; l++) rn[l] = 0;
                for (uint32_t t = t0; t < (uint32_t)(t0 + HT) && t < (uint32_t)nturns; t++) {
                    static uint32_t e4[32][4], h4[32][4], s4[32][4], c4[32][4]; int nv[32];
                    for (int l = 0; l < 32; l++) { uint32_t p0 = pos[l] + 4 * t; nv[l] = 0; if (p0 >= send[l]) continue; nv[l] = send[l] - p0 < 4 ? (int)(send[l] - p0) : 4;
                        for (int k = 0; k < nv[l]; k++) { s4[l][k] = ld32(src + p0 + k); uint32_t hv = s4[l][k] * 2654435761u; h4[l][k] = hv >> (32 - HASHLOG); c4[l][k] = (hv >> (17 - HASHLOG)) & 0x7FFF; e4[l][k] = table[h4[l][k]]; } }
                    if (!SKIPCONT) for (int l = 0; l < 32; l++) for (int k = 0; k < nv[l]; k++) table[h4[l][k]] = (c4[l][k] << 17) | (pos[l] + 4 * t + k);
                    static int doins[32][4];
                    for (int l = 0; l < 32; l++) for (int k = 0; k < nv[l]; k++) {
                        doins[l][k] = 1;
                        uint32_t p = pos[l] + 4 * t + k; uint32_t e_ = e4[l][k]; uint32_t c = e_ & 0x1FFFF;
                        int ok = (e_ >> 17) == c4[l][k] && c < p && p - c < 65536;
                        uint32_t cand = c;
                        if (!ok && REP && rep[l] && p >= rep[l] && ld32(src + p - rep[l]) == s4[l][k]) { ok = 1; cand = p - rep[l]; }
                        if (!ok && INGRP) for (int j = k - 1; j >= 0; j--) if (s4[l][j] == s4[l][k]) { ok = 1; cand = p - (k - j); break; }
                        if (!ok) continue;
                        int idx = (int)(4 * t + k);
                        if (idx == lastk[l] + 1 && p - cand == lastoff[l]) { lastk[l] = idx; doins[l][k] = 0; continue; }
                        lastk[l] = idx; lastoff[l] = p - cand;
                        if (rn[l] < RECMAX) { rpos[l][rn[l]] = p; rcand[l][rn[l]] = cand; rn[l]++; }
                    }
                    if (SKIPCONT) for (int l = 0; l < 32; l++) for (int k = 0; k < nv[l]; k++) if (doins[l][k] || SKIPCONT == 2) table[h4[l][k]] = (c4[l][k] << 17) | (pos[l] + 4 * t + k);
                }
                for (int l = 0; l < 32; l++) { uint32_t pe = pe_l[l];
                    for (int r = 0; r < rn[l]; r++) {
                        uint32_t p = rpos[l][r], cand = rcand[l][r];
                        if (p < pe) { if (NOSHIFT) continue; uint32_t d = pe - p; p += d; cand += d; if (p >= send[l]) continue; }
                        if (p + 4 > n || ld32(src + p) != ld32(src + cand)) continue;
                        uint32_t cap = send[l] + CAPX * STRIP; if (cap > mlimit) cap = mlimit;
                        uint32_t e = p + 4, cc = cand + 4;
                        while (e < cap && src[e] == src[cc]) { e++; cc++; }
                        int open = (e >= cap && cap < mlimit);
                        uint32_t ms = p, mc = cand;
                        while (ms > pe && mc > 0 && src[ms - 1] == src[mc - 1]) { ms--; mc--; }
                        rep[l] = ms - mc;
                        mt_t m = { ms, e, ms - mc, open }; if (cnt[l] < 127) lists[l][cnt[l]++] = m; pe = e;
                    }
                    pe_l[l] = pe;
                }
                }
            } else
            for (;;) {
                int any = 0; static uint32_t ent[32][4][2], hh[32][4], sq[32][4], ck[32][4]; int act[32], nv[32];
                const uint32_t HM = (1u << HASHLOG) - 1; (void)HM;
                for (int l = 0; l < 32; l++) { act[l] = pos[l] < send[l]; if (act[l]) { any = 1; nv[l] = send[l] - pos[l] < (uint32_t)GROUP ? (int)(send[l] - pos[l]) : GROUP;
                    for (int k = 0; k < nv[l]; k++) { sq[l][k] = ld32(src + pos[l] + k); uint32_t hv = sq[l][k] * 2654435761u; if (HASHBYTES == 5) hv = (uint32_t)(((ld64(src + pos[l] + k) << 24) * 889523592379ULL) >> 32); if (HASHBYTES == 6) hv = (uint32_t)(((ld64(src + pos[l] + k) << 16) * 227718039650203ULL) >> 32); if (HASHBYTES == 45) hv = (sq[l][k] ^ (src[pos[l] + k + 4] * 0x9E3779B1u)) * 2654435761u; hh[l][k] = hv >> (32 - HASHLOG); ck[l][k] = (hv >> (17 - HASHLOG)) & 0x7FFF;
                        if (WAYS2) { uint32_t b = hh[l][k] & ~1u; ent[l][k][0] = table[b]; ent[l][k][1] = table[b + 1]; } else { ent[l][k][0] = table[hh[l][k]]; ent[l][k][1] = 0; } } } }
                if (!any) break;
                for (int l = 0; l < 32; l++) if (act[l] && !UPTO) for (int k = 0; k < nv[l]; k++) { uint32_t v = (ck[l][k] << 17) | (pos[l] + k);
                    if (WAYS2) { uint32_t b = hh[l][k] & ~1u; uint32_t lowbit = hh[l][k] & 1u; v = (((ck[l][k] << 1) | lowbit) & 0x7FFF) << 17 | (pos[l] + k); table[b + 1] = table[b]; table[b] = v; } else table[hh[l][k]] = v; }
                int maxext = 0, anyok = 0;
                for (int l = 0; l < 32; l++) if (act[l]) {
                    lane_iters[l]++;
                    int pick = -1; uint32_t cand = 0; int from_rep = 0;
                    if (REPFIRST == 3 && REP && rep[l]) for (int k = 0; k < nv[l] && pick < 0; k++) { uint32_t p = pos[l] + k; if (p >= rep[l] && ld32(src + p - rep[l]) == sq[l][k]) { pick = k; cand = p - rep[l]; } }
                    for (int k = 0; k < nv[l] && pick < 0; k++) {
                        uint32_t p = pos[l] + k;
                        if ((REPFIRST == 1 || REPFIRST == 3) && REP && rep[l] && p >= rep[l] && (REPWIN == 0 || pos[l] - lanch[l] < (uint32_t)REPWIN) && ld32(src + p - rep[l]) == sq[l][k]) { pick = k; cand = p - rep[l]; from_rep = 1; break; }
                        if (REPFIRST == 2) { /* longest of table / rep / fixed offsets, compared over LCMP bytes */
                            uint32_t best = 0, bc = 0; uint32_t cs[4]; int nc = 0;
                            uint32_t e_ = ent[l][k][0]; uint32_t c = e_ & 0x1FFFF;
                            if ((e_ >> 17) == ck[l][k] && c < p && p - c < 65536 && ld32(src + c) == sq[l][k]) cs[nc++] = c;
                            if (REP && rep[l] && p >= rep[l] && ld32(src + p - rep[l]) == sq[l][k]) cs[nc++] = p - rep[l];
                            for (int q = 0; q < nc; q++) { uint32_t len = 4; while (len < (uint32_t)LCMP && p + len < mlimit && src[p + len] == src[cs[q] + len]) len++; if (len > best) { best = len; bc = cs[q]; } }
                            if (nc) { pick = k; cand = bc; break; }
                            continue;
                        }
                        for (int w = 0; w < (WAYS2 ? 2 : 1) && pick < 0; w++) {
                            uint32_t e_ = ent[l][k][w]; uint32_t chk = WAYS2 ? (((ck[l][k] << 1) | (hh[l][k] & 1u)) & 0x7FFF) : ck[l][k];
                            uint32_t c = e_ & 0x1FFFF; int ok = (e_ >> 17) == chk && c < p && p - c < 65536 && ld32(src + c) == sq[l][k];
                            if (ok) { pick = k; cand = c; }
                        }
                        if (pick < 0 && REP && rep[l] && p >= rep[l] && (REPWIN == 0 || pos[l] - lanch[l] < (uint32_t)REPWIN) && ld32(src + p - rep[l]) == sq[l][k]) { pick = k; cand = p - rep[l]; }
                    }
                    if (pick < 0 && INGRP) { /* repeats inside the group: nearest earlier position of the group with the same 4 bytes */
                        for (int k = 1; k < nv[l] && pick < 0; k++) for (int j = k - 1; j >= 0; j--) if (sq[l][j] == sq[l][k]) { pick = k; cand = pos[l] + j; break; } }
                    if (UPTO) { int lim = pick < 0 ? nv[l] - 1 : pick; for (int k = 0; k <= lim; k++) table[hh[l][k]] = (ck[l][k] << 17) | (pos[l] + k); }
                    if (pick < 0) { pos[l] += nv[l]; continue; }
                    anyok = 1; pos[l] += pick;
                    uint32_t cap = send[l] + CAPX * STRIP; if (cap > mlimit) cap = mlimit;
                    uint32_t e = pos[l] + 4, cc = cand + 4;
                    while (e < cap && src[e] == src[cc]) { e++; cc++; }
                    int open = (e >= cap && cap < mlimit);
                    uint32_t ms = pos[l], mc = cand;
                    while (ms > lanch[l] && mc > 0 && src[ms - 1] == src[mc - 1]) { ms--; mc--; }
                    if ((int)(e - ms) < MINM && !(MINM_TABLE_ONLY && from_rep) && (int)(ms - lanch[l]) >= MINLL) { pos[l] += 1; continue; }
                    int ext = (int)((e - pos[l]) / 4) + (int)(pos[l] - ms); if (ext > maxext) maxext = ext;
                    lane_iters[l] += (int)((e - pos[l] - 4 + 7) / 8) + 1;
                    rep[l] = ms - mc;
                    mt_t m = { ms, e, ms - mc, open }; lists[l][cnt[l]++] = m;
                    pos[l] = e; lanch[l] = e;
                }
                costA += 30 + (anyok ? 15 + 12.0 * maxext : 0);
            }
            int mx = 0; for (int l = 0; l < 32; l++) if (lane_iters[l] > mx) mx = lane_iters[l];
            COST_A += costA + 400; COST_B += 28.0 * mx + 400; NSTEPS++;
            /* resolution: serial exact version (GPU uses prefix max) */
            uint32_t E = si > anchor ? si : anchor; if (E < anchor) E = anchor;
            uint32_t next_si = region_end;
            int stop = 0;
            for (int l = 0; l < 32 && !stop; l++) for (int k = 0; k < cnt[l]; k++) {
                mt_t m = lists[l][k];
                if (m.open) { /* true end by cooperative extension */
                    uint32_t e = m.me, cc = m.me - m.off; while (e < mlimit && src[e] == src[cc]) { e++; cc++; } m.me = e; NOPEN++; }
                if (m.me <= E) { NDROP++; continue; }
                if (m.ms < E) { NTRIM++; m.ms = E; if (m.me - m.ms < 4) { NDROP++; continue; } }
                if (SHORTRULE && (int)(m.me - m.ms) < SHORTLEN && m.ms - anchor >= (uint32_t)SHORTLL) { NDROP++; continue; }
                /* emit */
                uint32_t ll = m.ms - anchor, ml = m.me - m.ms - 4;
                uint32_t tok = op++;
                if (ll >= 15) op += put_ext(out + op, ll - 15);
                memcpy(out + op, src + anchor, ll); op += ll;
                out[tok] = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (ml < 15 ? ml : 15));
                out[op++] = (uint8_t)m.off; out[op++] = (uint8_t)(m.off >> 8);
                if (ml >= 15) op += put_ext(out + op, ml - 15);
                anchor = m.me; E = m.me; (*nseq)++;
            }
            si = E > next_si ? E : next_si;
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
        if (!strcmp(argv[i], "strip")) STRIP = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "capx")) CAPX = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "backx")) BACKX = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "rep")) REP = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "group")) GROUP = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "ways2")) WAYS2 = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "upto")) UPTO = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "minm")) MINM = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "minll")) MINLL = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "minm_table_only")) MINM_TABLE_ONLY = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "phased")) PHASED = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "recmax")) RECMAX = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "halves")) HALVES = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "skipcont")) SKIPCONT = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "noshift")) NOSHIFT = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "repfirst")) REPFIRST = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "lcmp")) LCMP = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "shortrule")) SHORTRULE = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "repwin")) REPWIN = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "shortlen")) SHORTLEN = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "shortll")) SHORTLL = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "ingrp")) INGRP = atoi(argv[i + 1]);
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
      printf("   costA/B instr per byte %.2f %.2f  steps %ld open %ld drop %ld trim %ld\n", COST_A / ((double)n * NREP), COST_B / ((double)n * NREP), NSTEPS / NREP, NOPEN / NREP, NDROP / NREP, NTRIM / NREP); COST_A = COST_B = 0; NSTEPS = NOPEN = NDROP = NTRIM = 0;
      printf("%-16s model %8zu  oracle %8zu  ratio %.4f  seqs/frame %6ld  valid=%d\n", names[k], tot1, tot2, (double)tot1 / tot2, totseq / NREP, allok);
    }
    return 0;
}
