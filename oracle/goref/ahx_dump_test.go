// ahx_dump_test.go -- one-command pin of allhic-b200's oracle against the Go reference.
//
// TEST INFRASTRUCTURE of allhic-b200 (not part of the product, not part of the reference).
// Drop this file into a checkout of github.com/tanghaibao/allhic (package allhic, next to
// clm.go) on any machine that has Go and the module cache, and run
//
//	AHX_GOLDEN=/path/to/allhic-b200/tests/golden go test -run TestAhxDump -count=1 .
//
// (oracle/goref/run.sh does exactly that).  For every case of tests/golden/known_answers.json
// -- 36 (tour, signs) pairs on the reference's own fixture tests/simulation/test.{ids,clm} --
// it calls the reference's own Tour.Evaluate (evaluate.go:116), CLM.EvaluateQ
// (orientation.go:159), CLM.M (clm.go:437), GoldenArray / SumLog (base.go:138-167) and
// math.Log, and writes the results bit-exactly (IEEE-754 bits as hex) to
// $AHX_GOLDEN/go_reference_dump.json.  tests/test_oracle.py::test_go_reference_dump_if_present
// compares the C oracle with that file whenever it exists; committing the file turns
// "parity unpinned" into "pinned against the Go binary".
package allhic

import (
	"encoding/json"
	"fmt"
	"io/ioutil"
	"math"
	"os"
	"path/filepath"
	"strconv"
	"strings"
	"testing"
)

type ahxCase struct {
	Tour  []int  `json:"tour"`
	Signs string `json:"signs"`
}

type ahxKnown struct {
	Cases []ahxCase `json:"cases"`
}

type ahxOutCase struct {
	Evaluate      float64 `json:"evaluate"`
	EvaluateBits  string  `json:"evaluate_bits"`
	EvaluateQ     float64 `json:"evaluate_q"`
	EvaluateQBits string  `json:"evaluate_q_bits"`
}

type ahxLine struct {
	Line        int     `json:"line"`
	GoldenArray []int   `json:"golden_array"`
	SumLogBits  string  `json:"sumlog_bits"`
	SumLog      float64 `json:"sumlog"`
}

type ahxDump struct {
	GoVersion    string            `json:"go_version"`
	N            int               `json:"N"`
	NContacts    int               `json:"ncontacts"`
	NOriented    int               `json:"noriented"`
	M01          int               `json:"M01"`
	M12          int               `json:"M12"`
	NnzM         int               `json:"nnzM"`
	SumUpperM    int               `json:"sum_upper_M"`
	SumSizes     int               `json:"sum_sizes"`
	Lines        []ahxLine         `json:"lines"`
	LogBits      map[string]string `json:"log_bits"`
	Cases        []ahxOutCase      `json:"cases"`
	OTrace       float64           `json:"O_upper_sum"`
	ContactsNl   map[string]int    `json:"contacts_nlinks_sample"`
	ContactsSign map[string]int    `json:"contacts_strandedness_sample"`
}

func ahxBits(x float64) string { return fmt.Sprintf("%016x", math.Float64bits(x)) }

func TestAhxDump(t *testing.T) {
	dir := os.Getenv("AHX_GOLDEN")
	if dir == "" {
		t.Skip("AHX_GOLDEN not set (path of allhic-b200/tests/golden)")
	}
	raw, err := ioutil.ReadFile(filepath.Join(dir, "known_answers.json"))
	if err != nil {
		t.Fatal(err)
	}
	var known ahxKnown
	if err := json.Unmarshal(raw, &known); err != nil {
		t.Fatal(err)
	}
	clm := NewCLM(filepath.Join(dir, "test.clm"), filepath.Join(dir, "test.ids"))
	M := clm.M()
	N := len(clm.Tigs)
	out := ahxDump{GoVersion: "see `go version`", N: N, NContacts: len(clm.contacts), NOriented: len(clm.orientedContacts),
		M01: M[0][1], M12: M[1][2], LogBits: map[string]string{}, ContactsNl: map[string]int{}, ContactsSign: map[string]int{}}
	for i := 0; i < N; i++ {
		out.SumSizes += clm.Tigs[i].Size
		for j := 0; j < N; j++ {
			if M[i][j] != 0 {
				out.NnzM++
			}
			if j > i {
				out.SumUpperM += M[i][j]
			}
		}
	}
	O := clm.O()
	for i := 0; i < N; i++ {
		for j := i + 1; j < N; j++ {
			out.OTrace += O.At(i, j)
		}
	}
	for pair, c := range clm.contacts {
		if pair.ai < 3 {
			key := fmt.Sprintf("%d,%d", pair.ai, pair.bi)
			out.ContactsNl[key] = c.nlinks
			out.ContactsSign[key] = c.strandedness
		}
	}
	// GoldenArray / SumLog of the first 64 CLM lines
	text, err := ioutil.ReadFile(filepath.Join(dir, "test.clm"))
	if err != nil {
		t.Fatal(err)
	}
	for ln, row := range strings.Split(string(text), "\n") {
		if ln >= 64 || row == "" {
			break
		}
		f := strings.Split(row, "\t")
		var links []int
		for _, w := range strings.Split(f[2], " ") {
			v, _ := strconv.Atoi(w)
			links = append(links, v)
		}
		ga := GoldenArray(links)
		sl := SumLog(links)
		out.Lines = append(out.Lines, ahxLine{Line: ln, GoldenArray: ga[:], SumLogBits: ahxBits(sl), SumLog: sl})
	}
	for _, x := range []float64{1, 2, 3, 10, 843, 5778, 9349, 123456, 1149851, 1e7, 10005778, 2147483647, 0.5, 1.0000000001} {
		out.LogBits[strconv.FormatFloat(x, 'g', -1, 64)] = ahxBits(math.Log(x))
	}
	for _, c := range known.Cases {
		tour := Tour{Tigs: make([]Tig, len(c.Tour)), M: M}
		for i, idx := range c.Tour {
			tour.Tigs[i] = Tig{Idx: idx, Size: clm.Tigs[idx].Size}
		}
		f, _ := tour.Evaluate()
		clm.Tour = tour
		clm.Signs = []byte(c.Signs)
		q := clm.EvaluateQ()
		out.Cases = append(out.Cases, ahxOutCase{Evaluate: f, EvaluateBits: ahxBits(f), EvaluateQ: q, EvaluateQBits: ahxBits(q)})
	}
	js, _ := json.MarshalIndent(out, "", " ")
	if err := ioutil.WriteFile(filepath.Join(dir, "go_reference_dump.json"), js, 0644); err != nil {
		t.Fatal(err)
	}
	t.Logf("wrote %d cases to %s", len(out.Cases), filepath.Join(dir, "go_reference_dump.json"))
}
