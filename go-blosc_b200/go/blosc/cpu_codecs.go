//go:build b200

package blosc

import "fmt"

// Codecs outside the GPU path (LZ4HC encode, Snappy, ZLIB, ZSTD) keep running on whatever
// CodecInterface the application registered -- in a drop-in deployment the reference's own
// codec.go adapters, unchanged.  The frame logic below is the reference's (blosc.go:320-434)
// with the shuffle stage on the device.
func compressOnCPU(data []byte, opts Options) ([]byte, error) {
	codec, ok := codecs[opts.Codec]
	if !ok {
		return nil, fmt.Errorf("%w: %s", ErrInvalidCodec, opts.Codec)
	}
	shuffled := data
	if (opts.Shuffle == Shuffle1 || opts.Shuffle == BitShuffle) && opts.TypeSize > 1 {
		shuffled = append([]byte(nil), data...)
		ShuffleBuffer(shuffled, opts.TypeSize, opts.Shuffle)
	}
	comp, err := codec.Compress(shuffled, opts.Level)
	if err != nil {
		return nil, fmt.Errorf("%w: %v", ErrCompressionFailed, err)
	}
	flags := uint8(0)
	if opts.Shuffle == Shuffle1 {
		flags |= flagShuffle
	} else if opts.Shuffle == BitShuffle {
		flags |= flagBitShuffle
	}
	if len(comp) >= len(data) {
		comp, flags = shuffled, flags|flagMemcpy
	}
	h := Header{Version: FormatVersion, VersionLZ: uint8(opts.Codec), Flags: flags, TypeSize: uint8(opts.TypeSize),
		NBytesOrig: uint32(len(data)), BlockSize: uint32(len(data)), NBytesComp: uint32(HeaderSize + len(comp))}
	return append(h.Bytes(), comp...), nil
}

func decompressOnCPU(data []byte, typeSize int, h *Header) ([]byte, error) {
	if int(h.NBytesComp) > len(data) || h.NBytesComp < HeaderSize {
		return nil, ErrInvalidData
	}
	codec, ok := codecs[Codec(h.VersionLZ)]
	if !ok {
		return nil, fmt.Errorf("%w: %s", ErrInvalidCodec, Codec(h.VersionLZ))
	}
	out, err := codec.Decompress(data[HeaderSize:h.NBytesComp], int(h.NBytesOrig))
	if err != nil {
		return nil, fmt.Errorf("%w: %v", ErrDecompressionFailed, err)
	}
	if typeSize <= 0 {
		typeSize = int(h.TypeSize)
	}
	if typeSize > 1 && (h.HasBitShuffle() || h.HasShuffle()) {
		UnshuffleBuffer(out, typeSize, h.ShuffleMode())
	}
	if len(out) != int(h.NBytesOrig) {
		return nil, fmt.Errorf("%w: got %d, expected %d", ErrSizeMismatch, len(out), h.NBytesOrig)
	}
	return out, nil
}
