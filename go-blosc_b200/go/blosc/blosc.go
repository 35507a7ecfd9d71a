// Package blosc is the Go host side of the B200 backend: it keeps every exported identifier
// of github.com/mrjoshuak/go-blosc (blosc.go, shuffle.go, codec.go of v1.0.2) and routes the
// shuffle + LZ4 hot path through the b2b C ABI (include/b2b.h) with cgo.
//
// This file holds what never touches the device: types, constants, sentinels, the 16-byte
// header and option clamping.  backend_b200.go (build tag `b200`) supplies compressBackend,
// decompressBackend and the shuffle entry points over cgo; LZ4-path frames never reach a Go
// codec or the SIMD shuffle code.  Codecs other than LZ4 (and LZ4HC decode) return
// errUnsupportedOnGPU from the bridge and are served by the reference's own Go codecs
// registered through RegisterCodec, exactly as before.
//
// NOTE: there is no Go toolchain in the image this was authored in; the file is written
// against the C ABI that the Python/ctypes and C++ harnesses exercise symbol by symbol.
package blosc

import (
	"encoding/binary"
	"errors"
	"fmt"
)

// Version constants (reference blosc.go:49-52).
const (
	Version       = "1.0.0"
	FormatVersion = 2
)

// Codec identifies the compression algorithm (reference blosc.go:55-64).
type Codec uint8

const (
	BloscLZ Codec = iota
	LZ4
	LZ4HC
	Snappy
	ZLIB
	ZSTD
)

func (c Codec) String() string {
	names := [...]string{"blosclz", "lz4", "lz4hc", "snappy", "zlib", "zstd"}
	if int(c) < len(names) {
		return names[c]
	}
	return fmt.Sprintf("unknown(%d)", c)
}

// Shuffle mode (reference blosc.go:86-92).
type Shuffle uint8

const (
	NoShuffle  Shuffle = 0x0
	Shuffle1   Shuffle = 0x1
	BitShuffle Shuffle = 0x2
)

func (s Shuffle) String() string {
	switch s {
	case NoShuffle:
		return "noshuffle"
	case Shuffle1:
		return "shuffle"
	case BitShuffle:
		return "bitshuffle"
	}
	return fmt.Sprintf("unknown(%d)", s)
}

const (
	flagShuffle    = 0x1
	flagMemcpy     = 0x2
	flagBitShuffle = 0x4

	HeaderSize    = 16
	MinHeaderSize = 16
)

// Sentinels (reference blosc.go:125-149).  The bridge maps B2B_* status codes onto them; the
// ones the reference returns bare stay bare, the others are wrapped with %w.
var (
	ErrInvalidData         = errors.New("blosc: invalid compressed data")
	ErrInvalidHeader       = errors.New("blosc: invalid header")
	ErrInvalidVersion      = errors.New("blosc: unsupported format version")
	ErrInvalidCodec        = errors.New("blosc: unsupported codec")
	ErrSizeMismatch        = errors.New("blosc: decompressed size mismatch")
	ErrDataTooLarge        = errors.New("blosc: data too large")
	ErrCompressionFailed   = errors.New("blosc: compression failed")
	ErrDecompressionFailed = errors.New("blosc: decompression failed")
)

// Header is the 16-byte frame header (reference blosc.go:154-162).
type Header struct {
	Version    uint8
	VersionLZ  uint8
	Flags      uint8
	TypeSize   uint8
	NBytesOrig uint32
	BlockSize  uint32
	NBytesComp uint32
}

// ParseHeader mirrors reference blosc.go:165-185 (host only).
func ParseHeader(data []byte) (*Header, error) {
	if len(data) < HeaderSize {
		return nil, ErrInvalidHeader
	}
	h := &Header{
		Version: data[0], VersionLZ: data[1], Flags: data[2], TypeSize: data[3],
		NBytesOrig: binary.LittleEndian.Uint32(data[4:8]),
		BlockSize:  binary.LittleEndian.Uint32(data[8:12]),
		NBytesComp: binary.LittleEndian.Uint32(data[12:16]),
	}
	if h.Version != FormatVersion {
		return nil, fmt.Errorf("%w: got %d, expected %d", ErrInvalidVersion, h.Version, FormatVersion)
	}
	return h, nil
}

// Bytes mirrors reference blosc.go:188-198.
func (h *Header) Bytes() []byte {
	buf := make([]byte, HeaderSize)
	buf[0], buf[1], buf[2], buf[3] = h.Version, h.VersionLZ, h.Flags, h.TypeSize
	binary.LittleEndian.PutUint32(buf[4:8], h.NBytesOrig)
	binary.LittleEndian.PutUint32(buf[8:12], h.BlockSize)
	binary.LittleEndian.PutUint32(buf[12:16], h.NBytesComp)
	return buf
}

func (h *Header) HasShuffle() bool    { return h.Flags&flagShuffle != 0 }
func (h *Header) HasBitShuffle() bool { return h.Flags&flagBitShuffle != 0 }
func (h *Header) IsMemcpy() bool      { return h.Flags&flagMemcpy != 0 }

// ShuffleMode mirrors reference blosc.go:216-224 (bit shuffle wins).
func (h *Header) ShuffleMode() Shuffle {
	if h.HasBitShuffle() {
		return BitShuffle
	}
	if h.HasShuffle() {
		return Shuffle1
	}
	return NoShuffle
}

// Options mirrors reference blosc.go:227-234; BlockSize and NumThreads stay unread (SURVEY F1).
type Options struct {
	Codec      Codec
	Level      int
	Shuffle    Shuffle
	TypeSize   int
	BlockSize  int
	NumThreads int
}

func DefaultOptions() Options {
	return Options{Codec: LZ4, Level: 5, Shuffle: Shuffle1, TypeSize: 4, BlockSize: 0}
}

// Compress mirrors reference blosc.go:257-265.
func Compress(data []byte, codec Codec, level int, shuffle Shuffle, typeSize int) ([]byte, error) {
	return CompressWithOptions(data, Options{Codec: codec, Level: level, Shuffle: shuffle, TypeSize: typeSize})
}

// CompressWithOptions mirrors reference blosc.go:268-286.
func CompressWithOptions(data []byte, opts Options) ([]byte, error) {
	if len(data) == 0 {
		return nil, ErrInvalidData
	}
	if opts.TypeSize <= 0 {
		opts.TypeSize = 1
	}
	if opts.Level < 1 {
		opts.Level = 1
	}
	if opts.Level > 9 {
		opts.Level = 9
	}
	return compressBackend(data, opts)
}

func Decompress(data []byte) ([]byte, error) { return DecompressWithSize(data, 0) }

// DecompressWithSize mirrors reference blosc.go:296-303.
func DecompressWithSize(data []byte, typeSize int) ([]byte, error) {
	if len(data) < HeaderSize {
		return nil, ErrInvalidHeader
	}
	return decompressBackend(data, typeSize)
}

func GetInfo(data []byte) (*Header, error) { return ParseHeader(data) }

func GetDecompressedSize(data []byte) (int, error) {
	h, err := ParseHeader(data)
	if err != nil {
		return 0, err
	}
	return int(h.NBytesOrig), nil
}

// CodecInterface and the registry mirror reference codec.go:14-53.  LZ4 frames never consult
// it (the device does the whole frame); it serves the codecs that stay on the CPU and lets a
// caller plug GPULZ4Codec in explicitly (RegisterCodec(LZ4, GPULZ4Codec{})).
type CodecInterface interface {
	Compress(data []byte, level int) ([]byte, error)
	Decompress(data []byte, expectedSize int) ([]byte, error)
	Name() string
}

var codecs = map[Codec]CodecInterface{}

func RegisterCodec(id Codec, codec CodecInterface) { codecs[id] = codec }

func GetCodec(id Codec) (CodecInterface, bool) {
	c, ok := codecs[id]
	return c, ok
}

func ListCodecs() []Codec {
	out := make([]Codec, 0, len(codecs)+1)
	out = append(out, LZ4)
	for id := range codecs {
		if id != LZ4 {
			out = append(out, id)
		}
	}
	return out
}
