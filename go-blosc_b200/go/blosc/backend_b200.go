//go:build b200

// cgo bridge: the reference's backend seam (compressBackend / decompressBackend, reference
// blosc.go:320-434; ShuffleBuffer / UnshuffleBuffer, shuffle.go:298-323) bound to include/b2b.h.
package blosc

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../lib -lb2b -Wl,-rpath,${SRCDIR}/../../lib
#include <stdlib.h>
#include "b2b.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// A cgo call pins an OS thread for its duration and a b2b_ctx serialises its callers, so concurrent
// goroutines draw contexts from a bounded free list (one stream + scratch arena + pinned staging ring
// each).  Contexts are created on demand up to maxContexts and then reused for the life of the process:
// a sync.Pool would let the garbage collector drop a context -- and its device arena -- between two calls.
const maxContexts = 4

type gpuCtx struct{ h *C.b2b_ctx }

var (
	ctxFree    = make(chan *gpuCtx, maxContexts)
	ctxMu      sync.Mutex
	ctxCreated int
)

func deviceIndex() int { return 0 } // one process per GPU: set CUDA_VISIBLE_DEVICES per rank

func acquireCtx() (*gpuCtx, error) {
	select {
	case c := <-ctxFree:
		return c, nil
	default:
	}
	ctxMu.Lock()
	if ctxCreated < maxContexts {
		var h *C.b2b_ctx
		if rc := C.b2b_init(C.int(deviceIndex()), &h); rc != C.B2B_OK {
			ctxMu.Unlock()
			return nil, fmt.Errorf("b2b_init: %s (no CPU fallback)", C.GoString(C.b2b_strerror(rc)))
		}
		ctxCreated++
		ctxMu.Unlock()
		return &gpuCtx{h: h}, nil
	}
	ctxMu.Unlock()
	return <-ctxFree, nil // all contexts are busy: wait for one
}

func withCtx(f func(*gpuCtx) error) error {
	c, err := acquireCtx()
	if err != nil {
		return err
	}
	defer func() { ctxFree <- c }()
	runtime.LockOSThread() // the CUDA calls of one batch stay on one thread
	defer runtime.UnlockOSThread()
	return f(c)
}

// Shutdown destroys the idle contexts (device arenas, pinned staging); call it when no call is in flight.
func Shutdown() {
	for {
		select {
		case c := <-ctxFree:
			C.b2b_destroy(c.h)
			ctxMu.Lock()
			ctxCreated--
			ctxMu.Unlock()
		default:
			return
		}
	}
}

// statusErr maps a B2B_* status to the reference's sentinel, bare or wrapped as the
// reference does (SURVEY 8(b) "Errors").
func statusErr(rc C.int, detail string) error {
	switch rc {
	case C.B2B_OK:
		return nil
	case C.B2B_EINVALID_DATA:
		return ErrInvalidData // bare (blosc.go:269-271, 385-390)
	case C.B2B_EINVALID_HEADER:
		return ErrInvalidHeader // bare (blosc.go:297-299)
	case C.B2B_EINVALID_VERSION:
		return fmt.Errorf("%w: %s", ErrInvalidVersion, detail)
	case C.B2B_EINVALID_CODEC:
		return fmt.Errorf("%w: %s", ErrInvalidCodec, detail)
	case C.B2B_ESIZE_MISMATCH:
		return fmt.Errorf("%w: %s", ErrSizeMismatch, detail)
	case C.B2B_EDATA_TOO_LARGE:
		return fmt.Errorf("%w: %s", ErrDataTooLarge, detail)
	case C.B2B_ECOMPRESSION_FAILED:
		return fmt.Errorf("%w: %s", ErrCompressionFailed, detail)
	case C.B2B_EDECOMPRESSION_FAILED:
		return fmt.Errorf("%w: %s", ErrDecompressionFailed, detail)
	}
	return fmt.Errorf("b2b: %s (%s)", C.GoString(C.b2b_strerror(rc)), detail)
}

func compressBackend(data []byte, opts Options) ([]byte, error) {
	if opts.Codec != LZ4 { // codecs that stay on the reference's CPU implementations
		return compressOnCPU(data, opts)
	}
	dst := make([]byte, int(C.b2b_max_frame_size(C.size_t(len(data))))+64)
	var n C.size_t
	err := withCtx(func(c *gpuCtx) error {
		rc := C.b2b_compress(c.h, unsafe.Pointer(&data[0]), C.size_t(len(data)), C.int(opts.Codec), C.int(opts.Level),
			C.int(opts.Shuffle), C.int64_t(opts.TypeSize), unsafe.Pointer(&dst[0]), C.size_t(len(dst)), &n)
		return statusErr(rc, opts.Codec.String())
	})
	if err != nil {
		return nil, err
	}
	return dst[:int(n):int(n)], nil
}

func decompressBackend(data []byte, typeSize int) ([]byte, error) {
	h, err := ParseHeader(data)
	if err != nil {
		return nil, err
	}
	if !h.IsMemcpy() && h.VersionLZ != uint8(LZ4) && h.VersionLZ != uint8(LZ4HC) {
		return decompressOnCPU(data, typeSize, h) // Snappy / ZLIB / ZSTD (or ErrInvalidCodec)
	}
	capacity := int(h.NBytesOrig)
	if reach := 255*len(data) + 64; capacity > reach { // an LZ4 block cannot expand more than 255x
		capacity = reach
	}
	dst := make([]byte, capacity+1)
	var n C.size_t
	err = withCtx(func(c *gpuCtx) error {
		rc := C.b2b_decompress(c.h, unsafe.Pointer(&data[0]), C.size_t(len(data)), C.int64_t(typeSize),
			unsafe.Pointer(&dst[0]), C.size_t(capacity), &n)
		if rc == C.B2B_EDST_TOO_SMALL && capacity < int(h.NBytesOrig) {
			rc = C.B2B_ESIZE_MISMATCH
		}
		return statusErr(rc, fmt.Sprintf("expected %d", h.NBytesOrig))
	})
	if err != nil {
		return nil, err
	}
	return dst[:int(n):int(n)], nil
}

// CompressBlocks writes a Blosc-1 MULTI-BLOCK frame: the one place where Options.BlockSize (declared and
// never read by the reference, blosc.go:227-234) means something.  0 = 64 KiB blocks.  Such frames have
// int32 bstarts and one LZ4 block per block; the reference's Decompress cannot read them (it decodes one
// block per frame), DecompressBlocks does.  opts.Codec must be LZ4.
func CompressBlocks(data []byte, opts Options) ([]byte, error) {
	if len(data) == 0 {
		return nil, ErrInvalidData
	}
	if opts.Codec != LZ4 {
		return nil, fmt.Errorf("%w: multi-block frames are LZ4 only", ErrInvalidCodec)
	}
	dst := make([]byte, len(data)+16+64)
	var n C.size_t
	err := withCtx(func(c *gpuCtx) error {
		rc := C.b2b_compress_blocks(c.h, unsafe.Pointer(&data[0]), C.size_t(len(data)), C.int(opts.Shuffle),
			C.int64_t(opts.TypeSize), C.uint32_t(opts.BlockSize), unsafe.Pointer(&dst[0]), C.size_t(len(dst)), &n)
		return statusErr(rc, "blocks")
	})
	if err != nil {
		return nil, err
	}
	return dst[:int(n):int(n)], nil
}

// DecompressBlocks reads a Blosc-1 multi-block LZ4 frame (split or unsplit blocks, stored frames).
func DecompressBlocks(data []byte) ([]byte, error) {
	h, err := ParseHeader(data)
	if err != nil {
		return nil, err
	}
	capacity := int(h.NBytesOrig)
	if reach := 255*len(data) + 64; capacity > reach {
		capacity = reach
	}
	dst := make([]byte, capacity+1)
	var n C.size_t
	err = withCtx(func(c *gpuCtx) error {
		rc := C.b2b_decompress_blocks(c.h, unsafe.Pointer(&data[0]), C.size_t(len(data)), unsafe.Pointer(&dst[0]),
			C.size_t(capacity), &n)
		if rc == C.B2B_EDST_TOO_SMALL && capacity < int(h.NBytesOrig) {
			rc = C.B2B_EDECOMPRESSION_FAILED
		}
		return statusErr(rc, fmt.Sprintf("expected %d", h.NBytesOrig))
	})
	if err != nil {
		return nil, err
	}
	return dst[:int(n):int(n)], nil
}

func filterInPlace(data []byte, typeSize int, mode Shuffle, inverse int) {
	if len(data) == 0 || (mode != Shuffle1 && mode != BitShuffle) {
		return // default arm of the reference's switch: untouched
	}
	_ = withCtx(func(c *gpuCtx) error {
		p := unsafe.Pointer(&data[0])
		C.b2b_shuffle(c.h, C.int(mode), C.int(inverse), C.int64_t(typeSize), p, p, C.size_t(len(data)))
		return nil
	})
}

// ShuffleBuffer / UnshuffleBuffer mirror reference shuffle.go:298-323 (in place).
func ShuffleBuffer(data []byte, typeSize int, mode Shuffle)   { filterInPlace(data, typeSize, mode, 0) }
func UnshuffleBuffer(data []byte, typeSize int, mode Shuffle) { filterInPlace(data, typeSize, mode, 1) }

// GPULZ4Codec implements CodecInterface over the raw-block entry points, for callers that
// want only the codec stage on the device: RegisterCodec(LZ4, GPULZ4Codec{}) (codec.go:36-38).
type GPULZ4Codec struct{}

func (GPULZ4Codec) Name() string { return "lz4" }

func (GPULZ4Codec) Compress(data []byte, level int) ([]byte, error) {
	if len(data) == 0 {
		return []byte{0}, nil
	}
	dst := make([]byte, int(C.b2b_lz4_bound(C.size_t(len(data)))))
	var n C.size_t
	err := withCtx(func(c *gpuCtx) error {
		return statusErr(C.b2b_lz4_block_compress(c.h, unsafe.Pointer(&data[0]), C.size_t(len(data)),
			unsafe.Pointer(&dst[0]), C.size_t(len(dst)), &n), "lz4 compress")
	})
	if err != nil {
		return nil, err
	}
	return dst[:int(n)], nil
}

func (GPULZ4Codec) Decompress(data []byte, expectedSize int) ([]byte, error) {
	buf := make([]byte, expectedSize+1)
	if len(data) == 0 {
		return buf[:0], nil
	}
	var n C.size_t
	err := withCtx(func(c *gpuCtx) error {
		return statusErr(C.b2b_lz4_block_decompress(c.h, unsafe.Pointer(&data[0]), C.size_t(len(data)),
			unsafe.Pointer(&buf[0]), C.size_t(expectedSize), &n), "lz4 decompress")
	})
	if err != nil {
		return nil, err
	}
	return buf[:int(n)], nil
}

// CompressChunks / DecompressChunks: many independent frames per call (no reference
// counterpart; SURVEY 8(f) rank 1).  Frames come back as sub-slices of one packed buffer.
// The chunks are gathered into one slice first: cgo does not allow a Go slice of Go pointers to cross
// the boundary, and one contiguous source is what b2b_compress_batch pipelines (its pinned staging ring
// takes pageable memory at full PCIe speed).  Callers that already hold one buffer should use
// CompressPacked, which skips the gather.
func CompressChunks(chunks [][]byte, shuffle Shuffle, typeSize int) ([][]byte, error) {
	n := len(chunks)
	if n == 0 {
		return nil, nil
	}
	var total uint64
	off := make([]C.uint64_t, n)
	ln := make([]C.uint32_t, n)
	for i, c := range chunks {
		off[i], ln[i] = C.uint64_t(total), C.uint32_t(len(c))
		total += uint64(len(c))
	}
	src := make([]byte, total)
	for i, c := range chunks {
		copy(src[off[i]:], c)
	}
	dst := make([]byte, total+uint64(32*n)+64)
	foff := make([]C.uint64_t, n)
	flen := make([]C.uint32_t, n)
	st := make([]C.uint32_t, n)
	var out C.uint64_t
	err := withCtx(func(c *gpuCtx) error {
		return statusErr(C.b2b_compress_batch(c.h, unsafe.Pointer(&src[0]), &off[0], &ln[0], C.uint32_t(n), C.int(shuffle),
			C.int64_t(typeSize), unsafe.Pointer(&dst[0]), C.uint64_t(len(dst)), &foff[0], &flen[0], &st[0], &out), "batch")
	})
	if err != nil {
		return nil, err
	}
	frames := make([][]byte, n)
	for i := range frames {
		if st[i] != 0 {
			return nil, statusErr(C.int(st[i]), fmt.Sprintf("chunk %d", i))
		}
		frames[i] = dst[foff[i] : uint64(foff[i])+uint64(flen[i])]
	}
	return frames, nil
}

// CompressPacked compresses the frames src[off[i] : off[i]+ln[i]] of ONE buffer (no gather copy) and returns the
// packed output with its offsets table.
func CompressPacked(src []byte, off []uint64, ln []uint32, shuffle Shuffle, typeSize int) (dst []byte, frameOff []uint64, frameLen []uint32, err error) {
	n := len(off)
	if n == 0 || len(ln) != n {
		return nil, nil, nil, nil
	}
	dst = make([]byte, uint64(len(src))+uint64(32*n)+64)
	frameOff = make([]uint64, n)
	frameLen = make([]uint32, n)
	st := make([]C.uint32_t, n)
	var out C.uint64_t
	err = withCtx(func(c *gpuCtx) error {
		return statusErr(C.b2b_compress_batch(c.h, unsafe.Pointer(&src[0]), (*C.uint64_t)(unsafe.Pointer(&off[0])),
			(*C.uint32_t)(unsafe.Pointer(&ln[0])), C.uint32_t(n), C.int(shuffle), C.int64_t(typeSize), unsafe.Pointer(&dst[0]),
			C.uint64_t(len(dst)), (*C.uint64_t)(unsafe.Pointer(&frameOff[0])), (*C.uint32_t)(unsafe.Pointer(&frameLen[0])), &st[0], &out), "batch")
	})
	if err != nil {
		return nil, nil, nil, err
	}
	for i := range st {
		if st[i] != 0 {
			return nil, nil, nil, statusErr(C.int(st[i]), fmt.Sprintf("chunk %d", i))
		}
	}
	return dst[:out], frameOff, frameLen, nil
}

// DecompressChunks decompresses many frames in one call; the results are sub-slices of one buffer.  Frames of
// other codecs than LZ4 / LZ4HC (or malformed headers) are reported per frame like Decompress would.
func DecompressChunks(frames [][]byte) ([][]byte, error) {
	n := len(frames)
	if n == 0 {
		return nil, nil
	}
	var total, outTotal uint64
	foff := make([]C.uint64_t, n)
	flen := make([]C.uint32_t, n)
	doff := make([]C.uint64_t, n)
	want := make([]uint64, n)
	for i, f := range frames {
		h, err := ParseHeader(f)
		if err != nil {
			return nil, fmt.Errorf("chunk %d: %w", i, err)
		}
		if !h.IsMemcpy() && h.VersionLZ != uint8(LZ4) && h.VersionLZ != uint8(LZ4HC) {
			return nil, fmt.Errorf("chunk %d: %w: DecompressChunks is the LZ4 path", i, ErrInvalidCodec)
		}
		foff[i], flen[i], doff[i] = C.uint64_t(total), C.uint32_t(len(f)), C.uint64_t(outTotal)
		want[i] = uint64(h.NBytesOrig)
		if reach := 255*uint64(len(f)) + 64; want[i] > reach { // an LZ4 block cannot expand more than 255x
			want[i] = reach
		}
		total += uint64(len(f))
		outTotal += (want[i] + 15) &^ 15
	}
	src := make([]byte, total)
	for i, f := range frames {
		copy(src[foff[i]:], f)
	}
	dst := make([]byte, outTotal+1)
	olen := make([]C.uint32_t, n)
	st := make([]C.uint32_t, n)
	err := withCtx(func(c *gpuCtx) error {
		return statusErr(C.b2b_decompress_batch(c.h, unsafe.Pointer(&src[0]), &foff[0], &flen[0], C.uint32_t(n), 0,
			unsafe.Pointer(&dst[0]), C.uint64_t(outTotal), &doff[0], &olen[0], &st[0]), "batch")
	})
	if err != nil {
		return nil, err
	}
	out := make([][]byte, n)
	for i := range out {
		if st[i] != 0 {
			return nil, statusErr(C.int(st[i]), fmt.Sprintf("chunk %d", i))
		}
		out[i] = dst[doff[i] : uint64(doff[i])+uint64(olen[i]) : uint64(doff[i])+uint64(olen[i])]
	}
	return out, nil
}
