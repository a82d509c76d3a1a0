// Command cubeec_ref_harness runs the UNMODIFIED CubeFS erasure-coding path (blobstore/common/ec on top of the
// vendored klauspost/reedsolomon) on seeded inputs and prints golden vectors as JSON.
//
// It exists to pin the oracle (oracle/cubeec_oracle.c) and the CUDA engine to bytes produced by the
// reference itself (SURVEY.md section 8c, last bullet).  There is no Go toolchain in the build image, so
// this file is committed ready to run: `tools/ref_harness/run.sh` builds it inside the reference module
// (-mod=vendor, via -overlay, nothing is written into the reference tree) and writes
// tests/golden/rs_golden.json, which tests/test_reference_golden.py consumes.
//
// Inputs are reproducible without Go: byte i of a case's data is byte (i % 8) of the (i / 8)-th output of
// splitmix64 seeded with `seed` (little endian) -- tests/refgolden.py holds the same generator.
//
// Reference entry points exercised:
//   ec.NewEncoder / Split / Encode / Verify / Reconstruct / ReconstructData  blobstore/common/ec/encoder.go:78-151
//   lrcEncoder (LRC code modes)                                              blobstore/common/ec/lrcencoder.go:35-200
//   reedsolomon.New(k, m).Encode / Reconstruct (codes that are no code mode) vendor/github.com/klauspost/reedsolomon/reedsolomon.go:413,609,1368
//   crc32.ChecksumIEEE of every shard                                         blobstore/access/stream/stream_put.go:265-269
package main

import (
	"crypto/sha256"
	"encoding/hex"
	"encoding/json"
	"fmt"
	"hash/crc32"
	"os"
	"sort"

	"github.com/klauspost/reedsolomon"

	"github.com/cubefs/cubefs/blobstore/common/codemode"
	"github.com/cubefs/cubefs/blobstore/common/ec"
)

type splitmix struct{ s uint64 }

func (r *splitmix) next() uint64 {
	r.s += 0x9E3779B97F4A7C15
	z := r.s
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

func fill(b []byte, seed uint64) {
	r := splitmix{s: seed}
	for i := 0; i < len(b); i += 8 {
		v := r.next()
		for j := 0; j < 8 && i+j < len(b); j++ {
			b[i+j] = byte(v >> (8 * uint(j)))
		}
	}
}

type recon struct {
	Bad      []int    `json:"bad"`
	DataOnly bool     `json:"data_only"`
	Sha256   []string `json:"sha256"` // of shards[bad[i]] after the call ("" when left missing)
}

type vector struct {
	Name      string   `json:"name"`
	Kind      string   `json:"kind"` // "rs" = reedsolomon.New(k, m); "codemode" = ec.NewEncoder(code mode)
	CodeMode  string   `json:"codemode,omitempty"`
	N         int      `json:"n"`
	M         int      `json:"m"`
	L         int      `json:"l"`
	AZCount   int      `json:"az_count"`
	Seed      uint64   `json:"seed"`
	DataLen   int      `json:"data_len"`   // bytes filled from the generator (the blob)
	ShardSize int      `json:"shard_size"` // len of every shard after Split / as allocated
	Sha256    []string `json:"sha256"`     // every shard after Encode (data, parity, local parity)
	Head      []string `json:"head"`       // first 16 bytes of every shard, hex
	Crc32     []uint32 `json:"crc32_ieee"` // crc32.ChecksumIEEE of every shard
	VerifyOK  bool     `json:"verify_ok"`
	Recon     []recon  `json:"reconstruct"`
}

func sum(b []byte) string {
	h := sha256.Sum256(b)
	return hex.EncodeToString(h[:])
}

func describe(v *vector, shards [][]byte) {
	for _, s := range shards {
		v.Sha256 = append(v.Sha256, sum(s))
		n := 16
		if len(s) < n {
			n = len(s)
		}
		v.Head = append(v.Head, hex.EncodeToString(s[:n]))
		v.Crc32 = append(v.Crc32, crc32.ChecksumIEEE(s))
	}
}

func clone(shards [][]byte) [][]byte {
	out := make([][]byte, len(shards))
	for i, s := range shards {
		out[i] = append(make([]byte, 0, len(s)), s...)
	}
	return out
}

func check(err error) {
	if err != nil {
		fmt.Fprintln(os.Stderr, "ref harness:", err)
		os.Exit(1)
	}
}

// plain reedsolomon.New(k, m): BASELINE configs that are not predefined code modes (C1 RS(4,2), C4 RS(20,4))
func rsCase(name string, k, m, shardSize int, seed uint64, bads [][]int) vector {
	enc, err := reedsolomon.New(k, m)
	check(err)
	data := make([]byte, k*shardSize)
	fill(data, seed)
	shards := make([][]byte, k+m)
	for i := 0; i < k; i++ {
		shards[i] = data[i*shardSize : (i+1)*shardSize]
	}
	for i := k; i < k+m; i++ {
		shards[i] = make([]byte, shardSize)
	}
	check(enc.Encode(shards))
	v := vector{Name: name, Kind: "rs", N: k, M: m, AZCount: 1, Seed: seed, DataLen: len(data), ShardSize: shardSize}
	describe(&v, shards)
	ok, err := enc.Verify(shards)
	check(err)
	v.VerifyOK = ok
	for _, bad := range bads {
		work := clone(shards)
		for _, i := range bad {
			for j := range work[i] {
				work[i][j] = 0
			}
			work[i] = work[i][:0]
		}
		check(enc.Reconstruct(work))
		r := recon{Bad: bad}
		for _, i := range bad {
			r.Sha256 = append(r.Sha256, sum(work[i]))
		}
		v.Recon = append(v.Recon, r)
	}
	return v
}

// a predefined code mode through ec.NewEncoder, the way access does it (stream_put.go:128-146):
// Split a blob into shards, Encode, then the cumulative-erasure loop of encoder_test.go:249-307
func modeCase(cm codemode.CodeMode, dataLen int, seed uint64) vector {
	t := cm.Tactic()
	enc, err := ec.NewEncoder(ec.Config{CodeMode: t, EnableVerify: true})
	check(err)
	size, err := ec.GetBufferSizes(dataLen, t)
	check(err)
	buf := make([]byte, size.ECSize) // what ec.NewBuffer hands out: room for every shard behind the data
	fill(buf[:dataLen], seed)
	shards, err := enc.Split(buf[:size.ECDataSize])
	check(err)
	check(enc.Encode(shards))
	v := vector{Name: "mode_" + cm.String(), Kind: "codemode", CodeMode: cm.String(), N: t.N, M: t.M, L: t.L,
		AZCount: t.AZCount, Seed: seed, DataLen: dataLen, ShardSize: len(shards[0])}
	describe(&v, shards)
	ok, err := enc.Verify(shards)
	check(err)
	v.VerifyOK = ok
	// cumulative erasures over the global stripe: first the data side, then parity, as long as M allows
	order := []int{}
	for i := 0; i < t.M; i++ {
		if i%2 == 0 {
			order = append(order, (i/2*5+1)%t.N)
		} else {
			order = append(order, t.N+(i/2*3)%t.M)
		}
	}
	seen := map[int]bool{}
	bads := []int{}
	for _, idx := range order {
		if seen[idx] {
			continue
		}
		seen[idx] = true
		bads = append(bads, idx)
		for _, dataOnly := range []bool{false, true} {
			work := clone(shards)
			for _, i := range bads {
				for j := range work[i] {
					work[i][j] = 0
				}
				work[i] = work[i][:0]
			}
			if dataOnly {
				check(enc.ReconstructData(work, bads))
			} else {
				check(enc.Reconstruct(work, bads))
			}
			r := recon{Bad: append([]int{}, bads...), DataOnly: dataOnly}
			for _, i := range bads {
				if len(work[i]) == 0 {
					r.Sha256 = append(r.Sha256, "")
				} else {
					r.Sha256 = append(r.Sha256, sum(work[i]))
				}
			}
			v.Recon = append(v.Recon, r)
		}
	}
	return v
}

func main() {
	out := []vector{}
	// C1: RS(4,2), 64 KiB shards, one stripe
	out = append(out, rsCase("C1_rs_4_2_64KiB", 4, 2, 65536, 0xC0BEF5, [][]int{{0, 1}, {1, 4}, {4, 5}}))
	// C2 / C3: EC12P4, 4 MiB blob -> shards of 349,526 bytes; three seeded stripes, 3 erasures each
	for s := uint64(0); s < 3; s++ {
		v := modeCase(codemode.EC12P4, 4<<20, 0xC0BEF5+s)
		v.Name = fmt.Sprintf("C2_EC12P4_4MiB_stripe%d", s)
		out = append(out, v)
	}
	out = append(out, rsCase("C3_rs_12_4_3erasures", 12, 4, 349526, 0xC0BEF5+100,
		[][]int{{1, 7, 13}, {0, 5, 11}, {2, 12, 15}, {13, 14, 15}, {3, 4, 9}}))
	// C4: RS(20,4), 1 MiB shards
	out = append(out, rsCase("C4_rs_20_4_1MiB", 20, 4, 1<<20, 0xC0BEF5+200, [][]int{{0, 19, 23}, {7}, {20, 21, 22, 23}}))
	// C5 corners: RS(6,3) and RS(12,4) at the sweep's smallest and an unaligned size
	out = append(out, rsCase("C5_rs_6_3_4KiB", 6, 3, 4096, 0xC0BEF5+300, [][]int{{0, 3, 8}}))
	out = append(out, rsCase("C5_rs_12_4_70001", 12, 4, 70001, 0xC0BEF5+301, [][]int{{0, 11, 12, 15}}))
	// every predefined EC code mode (encoder_test.go:249-307 walks the same list), blob of 64 KiB + 777 bytes
	modes := codemode.GetECCodeModes() // map order: sort for a stable file
	sort.Slice(modes, func(i, j int) bool { return modes[i] < modes[j] })
	for _, cm := range modes {
		out = append(out, modeCase(cm, (1<<16)+777, 0xC0BEF5+1000+uint64(cm)))
	}
	e := json.NewEncoder(os.Stdout)
	e.SetIndent("", " ")
	check(e.Encode(map[string]interface{}{
		"generator": "tools/ref_harness/main.go against cubefs blobstore/common/ec + vendored klauspost/reedsolomon",
		"prng":      "splitmix64(seed), little-endian bytes",
		"vectors":   out,
	}))
}
