// tests/emu/emu_codec.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the CUDA kernels of go-blosc_b200/csrc on the
// CPU through the coroutine shim in this directory (cuda_runtime.h), one frame at a time, mirroring the
// launch sequences of compress_batch_dev_locked / decompress_batch_dev_locked in csrc/b2b.cu.  Used by
// tests/test_kernel_emulation.py (kernel logic against the oracle without a GPU) and by the design
// probes under tests/tools.  Not part of the product: the product has no CPU path.
//
// Build: g++ -O2 -std=c++17 -I tests/emu -shared -fPIC -o tests/_build/libemu_codec.so \
//            tests/emu/emu_codec.cpp tests/emu/emu_runtime.cpp
#include "cuda_runtime.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace b2b_emu_stats {
uint64_t counters[32];
}
#define B2B_STAT(i, v) (b2b_emu_stats::counters[(i)] += (uint64_t)(v))
namespace b2b_emu_stats {
std::vector<int64_t> trace;
bool trace_on = false;
inline void tr(int tag, int64_t a, int64_t b, int64_t c, int64_t d, int64_t e) {
    if (trace_on) { trace.push_back(tag); trace.push_back(a); trace.push_back(b); trace.push_back(c); trace.push_back(d); trace.push_back(e); }
}
}
#define B2B_TRACE(tag, a, b, c, d, e) b2b_emu_stats::tr(tag, (int64_t)(a), (int64_t)(b), (int64_t)(c), (int64_t)(d), (int64_t)(e))

#include "../../go-blosc_b200/csrc/common.cuh"
#include "../../go-blosc_b200/csrc/filter_kernels.cuh"
#include "../../go-blosc_b200/csrc/lz4_encode.cuh"
#include "../../go-blosc_b200/csrc/lz4_kernels.cuh"
#include "../../go-blosc_b200/csrc/scan.cuh"
#include "../../go-blosc_b200/csrc/lz4_decode2.cuh"
#include "../../go-blosc_b200/csrc/lz4_decode3.cuh"
#include "../../go-blosc_b200/csrc/lz4_decode4.cuh"

using namespace b2b;

namespace {

uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

void run_scan(const uint32_t *in, uint32_t n, uint64_t *out, uint64_t *total, int op) {
    const uint32_t tiles = (n + kScanTile - 1) / kScanTile;
    std::vector<uint64_t> state(tiles + 1, 0);
    uint32_t ticket = 0;
    ScanWork w;
    w.tile_state = state.data();
    w.ticket = &ticket;
    emu::launch(tiles, kScanThreads, [&] { scan_offsets_kernel(in, n, out, total, w, op); });
}

template <int HL, int HB, bool PH = false> void run_encode(const EncodeArgs &e, uint32_t grid) {
    emu::launch(grid, kEncThreads, [&] { lz4_encode_kernel<HL, HB, PH>(e); });
}

}  // namespace

extern "C" {

void emu_stat_reset() { for (auto &c : b2b_emu_stats::counters) c = 0; }
void emu_trace(int on) { b2b_emu_stats::trace_on = on != 0; b2b_emu_stats::trace.clear(); }
uint64_t emu_trace_len(void) { return b2b_emu_stats::trace.size(); }
const int64_t *emu_trace_data(void) { return b2b_emu_stats::trace.data(); }
uint64_t emu_stat(int i) { return b2b_emu_stats::counters[i & 31]; }
void emu_stats_reset(void) { memset(b2b_emu_stats::counters, 0, sizeof b2b_emu_stats::counters); }

// whole-buffer filter (b2b_shuffle_dev): mode 1 byte shuffle, 2 bit shuffle
int emu_filter(const uint8_t *src, uint8_t *dst, uint64_t n, int mode, uint32_t typesize, int inverse) {
    if (n == 0) return 0;
    FilterArgs a;
    a.src = src; a.dst = dst;
    a.ft.off = nullptr; a.ft.len = nullptr; a.ft.uniform_len = n; a.ft.nframes = 1;
    a.ft.tiles_per_frame = (uint32_t)std::max<uint64_t>(1, (n + kTileBytes - 1) / kTileBytes);
    a.meta = nullptr; a.uniform.mode = (uint32_t)mode; a.uniform.typesize = typesize; a.status = nullptr;
    a.inverse = inverse; a.copy_inactive = 1;
    emu::launch(a.ft.tiles_per_frame, kFilterThreads, [&] { filter_batch_kernel(a); });
    return 0;
}

// one frame through K1/K2 -> K3 -> finalize -> K5 -> pack, like compress_batch_dev_locked with nframes = 1.
// tune: the four encoder experiment knobs (0 = default).  Returns the frame's status word.
int emu_compress_frame(const uint8_t *src, uint32_t n, int shuffle, int64_t typesize, int hash_log, int hash_bytes, const uint32_t *tune,
                       int independent, int quirk, uint8_t *dst, uint64_t dst_cap, uint32_t *out_len) {
    if (typesize <= 0) typesize = 1;
    FrameMeta fm{0, 0};
    if ((shuffle == 1 || shuffle == 2) && typesize > 1) { fm.mode = (uint32_t)shuffle; fm.typesize = (uint32_t)typesize; }
    const bool filtered = fm.mode != 0;
    const uint32_t shuffle_flag = shuffle == 1 ? 0x1u : shuffle == 2 ? 0x4u : 0u;
    std::vector<uint8_t> raw(src, src + n);                       // 16-byte aligned copies like device buffers
    void *p_in = nullptr, *p_shuf = nullptr;
    if (posix_memalign(&p_in, 256, n + 256) || posix_memalign(&p_shuf, 256, n + 256)) return -1;
    memcpy(p_in, src, n);
    memset((uint8_t *)p_in + n, 0, 256);
    memset(p_shuf, 0, n + 256);
    const uint64_t src_off = 0;
    const uint32_t src_len = n;
    const uint8_t *in = (const uint8_t *)p_in;
    if (filtered) {
        FilterArgs a;
        a.src = in; a.dst = (uint8_t *)p_shuf;
        a.ft.off = &src_off; a.ft.len = &src_len; a.ft.uniform_len = 0; a.ft.nframes = 1;
        a.ft.tiles_per_frame = (uint32_t)std::max<uint64_t>(1, ((uint64_t)n + kTileBytes - 1) / kTileBytes);
        a.meta = nullptr; a.uniform = fm; a.status = nullptr; a.inverse = 0; a.copy_inactive = 1;
        emu::launch(a.ft.tiles_per_frame, kFilterThreads, [&] { filter_batch_kernel(a); });
        in = (const uint8_t *)p_shuf;
    }
    const uint64_t comp_bytes = (uint64_t)seg_count(n) * kSegSlot + kSegSlot + (uint64_t)n / 255 + 4096;
    const uint64_t max_segs = seg_count(n) + 2;
    std::vector<uint8_t> comp(comp_bytes + 64);
    std::vector<SegMeta> meta(max_segs);
    std::vector<SegPlace> place(max_segs);
    uint64_t comp_off = 0, seg_base = 0, frame_off = 0, total = 0;
    uint32_t comp_len = 0, frame_len = 0, flags = 0, final_ll = 0, final_off = 0, status = 0;
    unsigned long long ticket = 0;
    run_scan(&src_len, 1, &comp_off, nullptr, kScanSegSlot);
    run_scan(&src_len, 1, &seg_base, nullptr, kScanSegCount);
    const uint32_t segs_grid = std::max<uint32_t>(1, seg_count(n));
    EncodeArgs e;
    e.in = in; e.src_off = &src_off; e.src_len = &src_len; e.nframes = 1; e.segs_grid = segs_grid;
    e.comp = comp.data(); e.comp_off = &comp_off; e.seg_base = &seg_base; e.meta = meta.data(); e.ticket = &ticket;
    for (int i = 0; i < 4; i++) e.tune[i] = tune ? tune[i] : 0;
    e.independent = independent ? 1u : 0u;
    e.planes = fm.mode == 1 ? fm.typesize : 0u;
    e.phase_mask = (fm.mode == 2 && (fm.typesize & (fm.typesize - 1)) == 0 && fm.typesize <= 512) ? 8u * fm.typesize - 1u : 0u;
    e.comp_cap = comp_bytes; e.seg_cap = max_segs;
    const uint32_t grid = (segs_grid + kEncWarps - 1) / kEncWarps;
    // the automatic policy of launch_encode (csrc/b2b.cu)
    int hb = hash_bytes ? hash_bytes : (filtered ? 4 : 5);
    int hl = hash_log ? hash_log : (filtered ? kHashLogDefault : 12);
    if (hb == 5 && hl < 11) hl = 11;
    if (hb == 6 && hl < 12) hl = 12;
    if (e.phase_mask && hb == 4) {
        switch (hl) {
            case 10: run_encode<10, 4, true>(e, grid); break;
            case 11: run_encode<11, 4, true>(e, grid); break;
            case 12: run_encode<12, 4, true>(e, grid); break;
            default: run_encode<13, 4, true>(e, grid); break;
        }
    } else
    switch (hb * 100 + hl) {
        case 410: run_encode<10, 4>(e, grid); break;
        case 411: run_encode<11, 4>(e, grid); break;
        case 412: run_encode<12, 4>(e, grid); break;
        case 413: run_encode<13, 4>(e, grid); break;
        case 511: run_encode<11, 5>(e, grid); break;
        case 512: run_encode<12, 5>(e, grid); break;
        case 513: run_encode<13, 5>(e, grid); break;
        case 612: run_encode<12, 6>(e, grid); break;
        case 613: run_encode<13, 6>(e, grid); break;
        default: free(p_in); free(p_shuf); return -2;
    }
    FinalizeArgs fa;
    fa.src_len = &src_len; fa.seg_base = &seg_base; fa.meta = meta.data(); fa.place = place.data(); fa.nframes = 1;
    fa.shuffle_flag = shuffle_flag; fa.keep_raw = 0; fa.comp_len = &comp_len; fa.frame_len = &frame_len; fa.flags = &flags;
    fa.final_ll = &final_ll; fa.final_off = &final_off; fa.status = &status; fa.index = nullptr; fa.segs_per_frame = 0;
    fa.comp_off = &comp_off; fa.comp_cap = comp_bytes; fa.seg_cap = max_segs;
    emu::launch(1, 128, [&] { finalize_frames_kernel(fa); });
    run_scan(&frame_len, 1, &frame_off, &total, kScanAlign16);
    if (status == 0 && frame_len > dst_cap) { free(p_in); free(p_shuf); return 11; }
    void *p_out = nullptr;
    if (posix_memalign(&p_out, 256, (uint64_t)n + 512)) return -1;
    PackArgs p;
    p.in = in; p.raw = quirk ? (const uint8_t *)p_in : in; p.src_off = &src_off; p.src_len = &src_len; p.comp = comp.data();
    p.comp_off = &comp_off; p.seg_base = &seg_base; p.meta = meta.data(); p.place = place.data(); p.comp_len = &comp_len;
    p.flags = &flags; p.final_ll = &final_ll; p.final_off = &final_off; p.status = &status; p.frame_off = &frame_off;
    p.dst = (uint8_t *)p_out; p.nframes = 1; p.segs_grid = segs_grid; p.codec = 1; p.typesize_u8 = (uint32_t)(uint8_t)typesize;
    p.header = 1;
    emu::launch(segs_grid, kFilterThreads, [&] { pack_frames_kernel(p); });
    if (status == 0) { memcpy(dst, p_out, frame_len); *out_len = frame_len; } else *out_len = 0;
    free(p_in); free(p_shuf); free(p_out);
    return (int)status;
}

static uint32_t g_dst_misalign = 0, g_parse_grid = 1;
static uint32_t g_jump_taken = 0;
uint32_t emu_jump_taken() { return g_jump_taken; }     // frames the pointer-jumping engine decoded itself (state 1) so far
void emu_set_parse_grid(uint32_t g) { g_parse_grid = g ? g : 1; }
void emu_set_dst_misalign(uint32_t m) { g_dst_misalign = m & 15u; }

// one frame through K4 (split = parse kernel + copy kernel, else the fused kernel) and the inverse filter,
// like decompress_batch_dev_locked with nframes = 1.  Returns the frame's status word.
int emu_decompress_frame(const uint8_t *frame, uint32_t len, int64_t typesize_override, int split, uint8_t *dst,
                         uint32_t cap, uint32_t *out_len) {
    void *p_fr = nullptr, *p_stage = nullptr, *p_dst = nullptr;
    if (posix_memalign(&p_fr, 256, (uint64_t)len + 512) || posix_memalign(&p_stage, 256, (uint64_t)cap + 512) ||
        posix_memalign(&p_dst, 256, (uint64_t)cap + 512)) return -1;
    memset(p_fr, 0, (uint64_t)len + 512);
    memcpy(p_fr, frame, len);
    const uint64_t frame_off = 0, dst_off = g_dst_misalign;
    uint32_t out = 0, status = 0, cap_eff = 0, nrec = 0;
    uint64_t table_off = 0;
    FrameMeta meta{0, 0};
    emu::launch(1, 256, [&] { clip_caps_kernel(&dst_off, &cap, (uint64_t)cap + 16, 1, &cap_eff); });
    const uint64_t nrec_max = (uint64_t)cap / 4 + (kSeqSlack + 1) + 64;
    std::vector<uint64_t> table(nrec_max);
    DecodeArgs a;
    a.frames = (const uint8_t *)p_fr; a.frame_off = &frame_off; a.frame_len = &len; a.nframes = 1;
    a.typesize_override = typesize_override; a.dst = (uint8_t *)p_dst; a.scratch = (uint8_t *)p_stage; a.dst_off = &dst_off;
    a.dst_cap = &cap_eff; a.out_len = &out; a.status = &status; a.meta = &meta;
    a.table = nullptr; a.table_off = nullptr; a.nrec = nullptr; a.only = nullptr; a.fuse_unshuffle = 1; a.ticket = nullptr;
    if (split == 3) {
        emu::launch(1, kLaneThreads, [&] { lz4_lane_decode_kernel(a); });
    } else if (split == 2 || split == 4 || split == 5) {
        // split 5: split 4 with 1 KiB parse chunks
        const bool jumping = split >= 4;
        const uint32_t cshift = split == 5 ? kChunkShiftSmall : kChunkShift;
        // split 4: the pointer-jumping engine (lz4_decode4.cuh) in front of the tile copy engine
        // chunk-parallel decoder (lz4_decode2.cuh): prep -> K5 -> chunk parse -> stitch -> tile copy (+ fallback)
        FrameDec fd;
        uint32_t plen_eff = 0, last_chunk = 0, fallback = 0;
        uint64_t chunk_base = 0, total_chunks = 0;
        Prep2Args pa;
        pa.frames = a.frames; pa.frame_off = &frame_off; pa.frame_len = &len; pa.dst_cap = &cap_eff; pa.nframes = 1;
        pa.typesize_override = typesize_override; pa.fd = &fd; pa.plen_eff = &plen_eff; pa.out_len = &out; pa.status = &status;
        pa.meta = &meta; pa.keep_sparse = jumping ? 1u : 0u;
        emu::launch(1, 128, [&] { frame_prep_kernel(pa); });
        run_scan(&plen_eff, 1, &chunk_base, &total_chunks, cshift == kChunkShift ? kScanChunks : kScanChunksSmall);
        const uint64_t table_chunks = ((uint64_t)cap >> cshift) + ((uint64_t)cap / 255ull >> cshift) + 2 + 16;
        const uint32_t kSlotRecs = chunk_slot_records(cshift);
        std::vector<uint2> tab((table_chunks + 1) * kSlotRecs);   // + the frame's spare slot (warp stitch)
        std::vector<ChunkMeta> cmeta(table_chunks);
        std::vector<ChunkDesc> cdesc(table_chunks);
        Parse2Args pp;
        pp.frames = a.frames; pp.frame_off = &frame_off; pp.fd = &fd; pp.nframes = 1; pp.chunk_base = &chunk_base;
        pp.total_chunks = &total_chunks; pp.table = tab.data(); pp.meta = cmeta.data(); pp.table_chunks = table_chunks; pp.chunk_shift = cshift;
        std::vector<uint32_t> dead(table_chunks + 4, 0);
        unsigned long long ticket = 0;
        pp.dead = dead.data(); pp.ticket = &ticket;
        emu::launch(g_parse_grid, kParse2Threads, [&] { lz4_chunk_parse_kernel(pp); });
        {
            Repair2Args ra;
            ra.frames = a.frames; ra.frame_off = &frame_off; ra.fd = &fd; ra.nframes = 1; ra.chunk_base = &chunk_base;
            ra.total_chunks = &total_chunks; ra.table = tab.data(); ra.meta = cmeta.data(); ra.table_chunks = table_chunks; ra.chunk_shift = cshift;
            if (!getenv("EMU_NO_REPAIR")) emu::launch((uint32_t)((table_chunks + 127) / 128), 128, [&] { lz4_chunk_repair_kernel(ra); });
        }
        Stitch2Args sa;
        sa.frames = a.frames; sa.frame_off = &frame_off; sa.fd = &fd; sa.nframes = 1; sa.chunk_base = &chunk_base;
        sa.table = tab.data(); sa.meta = cmeta.data(); sa.desc = cdesc.data(); sa.last_chunk = &last_chunk;
        sa.fallback = &fallback; sa.table_chunks = table_chunks; sa.chunk_shift = cshift; sa.scratch = tab.data() + table_chunks * kSlotRecs;
        if (jumping) emu::launch(1, 64, [&] { lz4_stitch_warp_kernel(sa); });      // a warp per frame, 32 chunks per step
        else emu::launch(1, 128, [&] { lz4_stitch_kernel(sa); });
        if (getenv("EMU_DUMP_DESC"))
            for (uint64_t q = 0; q < total_chunks; q++)
                fprintf(stderr, "desc %llu: a=%lld b=%lld start=%u count=%u split=%u end=%u | meta entry=%u exit=%u count=%u out=%u pad=%u %x %u\n",
                        (unsigned long long)q, cdesc[q].base_a, cdesc[q].base_b, cdesc[q].start, cdesc[q].count, cdesc[q].split, cdesc[q].end,
                        cmeta[q].entry, cmeta[q].exit, cmeta[q].count, cmeta[q].out, cmeta[q].pad[0], cmeta[q].pad[1], cmeta[q].pad[2]);
        if (getenv("EMU_DUMP_REC"))
            for (uint64_t q = 0; q < total_chunks; q++)
                for (uint32_t i = 0; i <= cdesc[q].count && cdesc[q].count; i++)
                    fprintf(stderr, "rec %llu %u: %u %u\n", (unsigned long long)q, i, tab[q * kSlotRecs + cdesc[q].start + i].x, tab[q * kSlotRecs + cdesc[q].start + i].y);
        uint32_t jstate = 0, jtotal = 0;
        std::vector<uint32_t> jS;
        std::vector<JumpLong> jq;
        std::vector<uint8_t> jdone;
        uint32_t jctl[1 + kJumpRounds];
        if (jumping) {
            JumpArgs ja; ja.chunk_shift = cshift;
            ja.frames = a.frames; ja.frame_off = &frame_off; ja.fd = &fd; ja.nframes = 1; ja.dst = a.dst; ja.scratch = a.scratch;
            ja.dst_off = &dst_off; ja.chunk_base = &chunk_base; ja.total_chunks = &total_chunks; ja.table_chunks = table_chunks;
            ja.desc = cdesc.data(); ja.last_chunk = &last_chunk; ja.table = tab.data(); ja.fallback = &fallback;
            jS.assign((uint64_t)cap + 64, 0xDEADBEEFu); ja.S = jS.data();
            ja.state = &jstate; ja.total = &jtotal;
            ja.long_cap = cap / kJumpLong + 1 + 16; jq.resize(ja.long_cap); ja.longq = jq.data();
            ja.nlong = jctl; ja.changed = jctl + 1;
            ja.blocks_per_frame = cap / kJumpBlock + 1; jdone.assign(ja.blocks_per_frame, 0); ja.blockdone = jdone.data();
            ja.blocks_grid = g_parse_grid > 1 ? 3 : 1;
            ja.out_len = &out; ja.status = &status; ja.meta = &meta;
            emu::launch(1, 64, [&] { lz4_jump_select_kernel(ja); });
            emu::launch(g_parse_grid > 1 ? 3 : 1, kJumpThreads, [&] { lz4_jump_map_kernel(ja); });
            emu::launch(g_parse_grid > 1 ? 3 : 1, kJumpThreads, [&] { lz4_jump_long_kernel(ja); });
            emu::launch(1, 64, [&] { lz4_jump_check_kernel(ja); });
            for (uint32_t r = 0; r < kJumpRounds; r++) emu::launch(ja.blocks_grid, kJumpThreads, [&] { lz4_jump_round_kernel(ja, r); });
            emu::launch(ja.blocks_grid, kJumpThreads, [&] { lz4_jump_gather_kernel(ja); });
            if (jstate == 1) g_jump_taken++;
        }
        Copy2Args ca;
        ca.jump_state = jumping ? &jstate : nullptr; ca.chunk_shift = cshift;
        ca.frames = a.frames; ca.frame_off = &frame_off; ca.fd = &fd; ca.nframes = 1; ca.dst = a.dst; ca.scratch = a.scratch;
        ca.dst_off = &dst_off; ca.chunk_base = &chunk_base; ca.desc = cdesc.data(); ca.last_chunk = &last_chunk;
        ca.table = tab.data(); ca.fallback = &fallback; ca.out_len = &out; ca.status = &status; ca.meta = &meta;
        emu::launch(1, kCopy2Threads, [&] { lz4_copy2_kernel(ca); });
        a.only = &fallback;
        emu::launch(1, kCodecThreads, [&] { lz4_decode_kernel<false>(a); });
    } else if (split) {
        run_scan(&cap_eff, 1, &table_off, nullptr, kScanSeqSlots);
        ParseArgs pa;
        pa.frames = a.frames; pa.frame_off = &frame_off; pa.frame_len = &len; pa.dst_cap = &cap_eff; pa.nframes = 1;
        pa.table = table.data(); pa.table_off = &table_off; pa.nrec = &nrec; pa.table_cap = nrec_max;
        a.table = table.data(); a.table_off = &table_off; a.nrec = &nrec;
        emu::launch(1, kCodecThreads, [&] { lz4_parse_kernel(pa); });
        emu::launch(1, kCodecThreads, [&] { lz4_decode_kernel<true>(a); });
    } else {
        emu::launch(1, kCodecThreads, [&] { lz4_decode_kernel<false>(a); });
    }
    FilterArgs fa;
    fa.src = (const uint8_t *)p_stage; fa.dst = (uint8_t *)p_dst;
    fa.ft.off = &dst_off; fa.ft.len = &out; fa.ft.uniform_len = 0; fa.ft.nframes = 1;
    fa.ft.tiles_per_frame = (uint32_t)std::max<uint64_t>(1, ((uint64_t)cap + kTileBytes - 1) / kTileBytes);
    fa.meta = &meta; fa.uniform = FrameMeta{0, 0}; fa.status = &status; fa.inverse = 1; fa.copy_inactive = 0;
    emu::launch(fa.ft.tiles_per_frame, kFilterThreads, [&] { filter_batch_kernel(fa); });
    if (out) memcpy(dst, (uint8_t *)p_dst + dst_off, out);
    *out_len = out;
    free(p_fr); free(p_stage); free(p_dst);
    return (int)status;
}

}  // extern "C"
