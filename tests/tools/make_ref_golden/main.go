// make_ref_golden -- runs the UNMODIFIED reference (github.com/mrjoshuak/go-blosc, pierrec/lz4 v4.1.23 per its
// go.mod) over formula inputs that tests/ref_golden_inputs.py reproduces bit for bit, and writes
// tests/golden/ref_v1.json: for every case the reference's frame (hex), NBytesComp and SHA-256 of input and frame.
//
// This image has no Go toolchain (SURVEY F10), so the file it writes does not exist yet and the compressed sizes
// of the oracle are "parity unpinned".  On any machine with Go >= 1.23 and network access to the module proxy:
//
//     cd tests/tools/make_ref_golden && go mod init refgolden && go get github.com/mrjoshuak/go-blosc@v1.0.2 \
//         && go run . > ../../golden/ref_v1.json
//
// tests/test_ref_golden.py picks the file up when it is there: the oracle must decode every frame to the input and
// reproduce NBytesComp exactly (that pins its compressor), the CUDA path must decode every frame bit-exactly and
// compress within 1 % of NBytesComp with identical header fields.
package main

import (
	"crypto/sha256"
	"encoding/binary"
	"encoding/hex"
	"encoding/json"
	"fmt"
	"math"
	"os"

	blosc "github.com/mrjoshuak/go-blosc"
)

func splitmix64(x uint64) uint64 {
	x += 0x9E3779B97F4A7C15
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9
	x = (x ^ (x >> 27)) * 0x94D049BB133111EB
	return x ^ (x >> 31)
}

// integer-only signals: every value is an exact small integer, so float conversion is exact in any language
func smoothInt(i uint64, seed uint64) int64 {
	tri := func(p, period uint64) int64 { // triangle wave in [-period/2, period/2]
		q := p % period
		if q > period/2 {
			q = period - q
		}
		return int64(q) - int64(period/4)
	}
	return 3*tri(i, 4096) + tri(i*3, 667) + int64(splitmix64(seed^i)&15)
}

func ramp(n int) []byte {
	b := make([]byte, n)
	for i := range b {
		b[i] = byte(i % 256)
	}
	return b
}

func smoothF32(elems int, seed uint64) []byte {
	b := make([]byte, 4*elems)
	for i := 0; i < elems; i++ {
		v := float32(smoothInt(uint64(i), seed)) / 1024
		binary.LittleEndian.PutUint32(b[4*i:], math.Float32bits(v))
	}
	return b
}

func smoothF64(elems int, seed uint64) []byte {
	b := make([]byte, 8*elems)
	for i := 0; i < elems; i++ {
		v := float64(smoothInt(uint64(i), seed))/1024 + float64(splitmix64(seed+uint64(i))&0xFFFFF)/1099511627776
		binary.LittleEndian.PutUint64(b[8*i:], math.Float64bits(v))
	}
	return b
}

func lowentI16(elems int, seed uint64) []byte {
	b := make([]byte, 2*elems)
	for i := 0; i < elems; i++ {
		binary.LittleEndian.PutUint16(b[2*i:], uint16(splitmix64(seed^uint64(i))&7))
	}
	return b
}

func randomBytes(n int, seed uint64) []byte {
	b := make([]byte, n)
	for i := 0; i < n; i += 8 {
		v := splitmix64(seed + uint64(i))
		for k := 0; k < 8 && i+k < n; k++ {
			b[i+k] = byte(v >> (8 * k))
		}
	}
	return b
}

type entry struct {
	Name       string `json:"name"`
	Input      string `json:"input"` // generator call, reproduced by tests/ref_golden_inputs.py
	Shuffle    int    `json:"shuffle"`
	TypeSize   int    `json:"typesize"`
	N          int    `json:"n"`
	InputSHA   string `json:"input_sha256"`
	NBytesComp uint32 `json:"nbytes_comp"`
	Flags      uint8  `json:"flags"`
	FrameSHA   string `json:"frame_sha256"`
	FrameHex   string `json:"frame_hex"`
	RoundTrip  bool   `json:"reference_round_trip"` // false for the memcpy + shuffle quirk (SURVEY F4)
}

func main() {
	type c struct {
		name, gen string
		data      []byte
		sh        blosc.Shuffle
		T         int
	}
	cases := []c{
		{"C1 README ramp", "ramp(100000)", ramp(100000), blosc.Shuffle1, 4},
		{"C3 frame: smooth f32, Shuffle1 T=4", "smooth_f32(65536, 3)", smoothF32(65536, 3), blosc.Shuffle1, 4},
		{"C4 frame: smooth f64, BitShuffle T=8", "smooth_f64(32768, 4)", smoothF64(32768, 4), blosc.BitShuffle, 8},
		{"C5 frame: low-entropy int16, Shuffle1 T=2", "lowent_i16(131072, 5)", lowentI16(131072, 5), blosc.Shuffle1, 2},
		{"C5 frame: random bytes, Shuffle1 T=2 (memcpy flag + shuffle quirk)", "random_bytes(65536, 6)", randomBytes(65536, 6), blosc.Shuffle1, 2},
		{"random bytes, NoShuffle (memcpy flag)", "random_bytes(65536, 7)", randomBytes(65536, 7), blosc.NoShuffle, 1},
		{"low-entropy int16, NoShuffle", "lowent_i16(131072, 8)", lowentI16(131072, 8), blosc.NoShuffle, 1},
		{"smooth f32, BitShuffle T=4", "smooth_f32(65536, 9)", smoothF32(65536, 9), blosc.BitShuffle, 4},
		{"short: 13 bytes", "ramp(13)", ramp(13), blosc.Shuffle1, 4},
	}
	var out []entry
	for _, k := range cases {
		fr, err := blosc.Compress(k.data, blosc.LZ4, 5, k.sh, k.T)
		if err != nil {
			fmt.Fprintln(os.Stderr, k.name, err)
			os.Exit(1)
		}
		h, _ := blosc.GetInfo(fr)
		back, derr := blosc.Decompress(fr)
		in := sha256.Sum256(k.data)
		fs := sha256.Sum256(fr)
		out = append(out, entry{k.name, k.gen, int(k.sh), k.T, len(k.data), hex.EncodeToString(in[:]), h.NBytesComp, h.Flags,
			hex.EncodeToString(fs[:]), hex.EncodeToString(fr), derr == nil && string(back) == string(k.data)})
	}
	enc := json.NewEncoder(os.Stdout)
	enc.SetIndent("", " ")
	_ = enc.Encode(map[string]any{"reference": "github.com/mrjoshuak/go-blosc v1.0.2 (" + blosc.Version + ")", "frames": out})
}
