//go:build ahx

package allhic

/*
#cgo CFLAGS: -I${SRCDIR}/../allhic-b200/include
#cgo LDFLAGS: -L${SRCDIR}/../allhic-b200/allhic_b200/lib -lahx -Wl,-rpath,${SRCDIR}/../allhic-b200/allhic_b200/lib
#include <stdlib.h>
#include "ahx.h"

// cgo cannot call a Go closure through a C function pointer directly:
// gaProgress is exported from Go below and handed to ahx_ga_run via this shim.
extern void gaProgress(void *user, int32_t phase, int64_t gen, double best, int32_t *tour, int32_t L);
static int ahx_ga_run_go(ahx_ctx *ctx, const ahx_ga_params *p, int32_t L, const int32_t *init,
                         int32_t *best, ahx_ga_stats *st, void *user) {
  return ahx_ga_run(ctx, p, L, init, best, st, (ahx_ga_callback)gaProgress, user);
}
// the island model (one population per GPU, section 3): the same shim around ahx_ga_run_islands
static int ahx_ga_run_islands_go(ahx_ctx **ctxs, int32_t n, const ahx_ga_params *p, int32_t L, const int32_t *init,
                                 int32_t *best, ahx_ga_stats *st, void *user) {
  return ahx_ga_run_islands(ctxs, n, p, L, init, best, st, (ahx_ga_callback)gaProgress, user);
}
*/
import "C"

import (
	"fmt"
	"os"
	"runtime"
	"runtime/cgo"
	"unsafe"
)

// ahxState hangs off CLM (add the field `ahx *ahxState` to the struct at clm.go:32-41).
type ahxState struct {
	ctx *C.ahx_ctx
	clm *C.ahx_clm
}

func ahxCheck(ctx *C.ahx_ctx, rc C.int, what string) {
	if rc != C.AHX_OK { // reference convention: fatal log (base.go:349-362)
		log.Fatalf("%s: libahx error %d: %s", what, int(rc), C.GoString(C.ahx_last_error(ctx)))
	}
}

// ahxUpload is called at the end of NewCLM (clm.go:121-138): the Go parser has
// filled r.Tigs; replay the CLM lines into the flat pair table and upload it.
func (r *CLM) ahxUpload(device int, lines []CLMLine) {
	runtime.LockOSThread() // a context is bound to one OS thread / GPU
	st := &ahxState{}
	sizes := make([]C.int64_t, len(r.Tigs))
	for i, t := range r.Tigs {
		sizes[i] = C.int64_t(t.Size)
	}
	ahxCheck(nil, C.ahx_clm_new(C.int32_t(len(sizes)), &sizes[0], &st.clm), "ahx_clm_new")
	for _, line := range lines { // same filter and order as readClm (clm.go:208-239)
		ai, aok := r.tigToIdx[line.at]
		bi, bok := r.tigToIdx[line.bt]
		if !aok || !bok {
			continue
		}
		d := make([]C.int64_t, len(line.links))
		for k, x := range line.links {
			d[k] = C.int64_t(x)
		}
		var dp *C.int64_t
		if len(d) > 0 {
			dp = &d[0]
		}
		ahxCheck(nil, C.ahx_clm_add_line(st.clm, C.int32_t(ai), C.int32_t(bi), C.uint8_t(line.ao), C.uint8_t(line.bo),
			dp, C.int64_t(len(d))), "ahx_clm_add_line")
	}
	ahxCheck(nil, C.ahx_clm_finalize(st.clm), "ahx_clm_finalize")
	ahxCheck(nil, C.ahx_create(C.int(device), &st.ctx), "ahx_create") // AHX_ERR_NODEVICE: there is no CPU fallback
	ahxCheck(st.ctx, C.ahx_load_clm(st.ctx, st.clm), "ahx_load_clm")
	r.ahx = st
}

func (r *CLM) ahxClose() {
	if r.ahx != nil {
		C.ahx_destroy(r.ahx.ctx)
		C.ahx_clm_free(r.ahx.clm)
		r.ahx = nil
	}
}

func (r *CLM) tourIdx() []C.int32_t {
	t := make([]C.int32_t, r.Tour.Len())
	for i, tig := range r.Tour.Tigs {
		t[i] = C.int32_t(tig.Idx)
	}
	return t
}

type gaUser struct {
	r      *CLM
	fwtour *os.File
}

//export gaProgress
func gaProgress(user unsafe.Pointer, phase C.int32_t, gen C.int64_t, best C.double, tour *C.int32_t, L C.int32_t) {
	u := cgo.Handle(user).Value().(*gaUser) // invoked on the calling thread, between generation batches
	fmt.Printf("Current iteration GA%d-%d: max_score=%.5f\n", int(phase), int64(gen), float64(best)) // evaluate.go:295
	idx := unsafe.Slice(tour, int(L))
	cur := Tour{Tigs: make([]Tig, int(L)), M: u.r.Tour.M}
	for i, c := range idx {
		cur.Tigs[i] = Tig{int(c), u.r.Tigs[int(c)].Size}
	}
	u.r.printTour(u.fwtour, cur, fmt.Sprintf("GA%d-%d-%.5f", int(phase), int64(gen), float64(best))) // evaluate.go:298-299
}

// GARun replaces evaluate.go:258-315.
func (r *CLM) GARun(fwtour *os.File, opt *Optimizer, phase int) Tour {
	var p C.ahx_ga_params
	C.ahx_ga_default_params(&p) // tournament 3, 1e6 generations, log every 500 (evaluate.go:270-273,294)
	p.npop, p.ngen, p.mut_prob = C.int32_t(opt.NPop), C.int32_t(opt.NGen), C.double(opt.MutProb)
	p.seed, p.phase = C.uint64_t(opt.Seed)*1000003+C.uint64_t(phase), C.int32_t(phase) // one Philox stream per phase
	log.Noticef("GA initialized (npop: %v, ngen: %v, mu: %.2f, rng: %d, break: %d)",
		opt.NPop, opt.NGen, opt.MutProb, opt.Seed, LIMIT)
	init := r.tourIdx()
	best := make([]C.int32_t, len(init))
	var st C.ahx_ga_stats
	h := cgo.NewHandle(&gaUser{r, fwtour})
	defer h.Delete()
	ahxCheck(r.ahx.ctx, C.ahx_ga_run_go(r.ahx.ctx, &p, C.int32_t(len(init)), &init[0], &best[0], &st,
		unsafe.Pointer(h)), "ahx_ga_run")
	for i, c := range best { // r.Tour = ga.HallOfFame[0] (evaluate.go:313)
		r.Tour.Tigs[i] = Tig{int(c), r.Tigs[int(c)].Size}
	}
	return r.Tour
}

// EvaluateQ replaces orientation.go:159-193.
func (r *CLM) EvaluateQ() float64 {
	t := r.tourIdx()
	var out C.double
	ahxCheck(r.ahx.ctx, C.ahx_eval_q(r.ahx.ctx, 1, C.int32_t(len(t)), &t[0], 1,
		(*C.uint8_t)(unsafe.Pointer(&r.Signs[0])), &out), "ahx_eval_q")
	return float64(out)
}

func tagOf(accepted C.int32_t) string {
	if accepted != 0 {
		return ACCEPT
	}
	return REJECT
}

// flipWhole replaces orientation.go:67-84 (r.Signs is updated in place by the library).
func (r *CLM) flipWhole() string {
	t := r.tourIdx()
	var a, b C.double
	var acc C.int32_t
	ahxCheck(r.ahx.ctx, C.ahx_flip_whole(r.ahx.ctx, C.int32_t(len(t)), &t[0],
		(*C.uint8_t)(unsafe.Pointer(&r.Signs[0])), &a, &b, &acc), "ahx_flip_whole")
	flipLog("FLIPWHOLE", float64(a), float64(b), tagOf(acc))
	return tagOf(acc)
}

// flipOne replaces orientation.go:87-120; trace holds (score, newScore) per step for the every-50 log lines.
func (r *CLM) flipOne() string {
	t := r.tourIdx()
	L := len(t)
	trace := make([]C.double, 2*L)
	steps := make([]C.uint8_t, L)
	var fs C.double
	var na, nr C.int64_t
	var acc C.int32_t
	ahxCheck(r.ahx.ctx, C.ahx_flip_one(r.ahx.ctx, C.int32_t(L), &t[0], (*C.uint8_t)(unsafe.Pointer(&r.Signs[0])),
		&fs, &na, &nr, &trace[0], &steps[0], &acc), "ahx_flip_one")
	for i := 49; i < L; i += 50 {
		flipLog(fmt.Sprintf("FLIPONE (%d/%d)", i+1, L), float64(trace[2*i]), float64(trace[2*i+1]), tagOf(C.int32_t(steps[i])))
	}
	log.Noticef("FLIPONE: N_accepts=%d N_rejects=%d", int64(na), int64(nr))
	return tagOf(acc)
}

// flipAll replaces orientation.go:31-64 (Lanczos on the sparse O instead of mat64.EigenSym).
func (r *CLM) flipAll() string {
	t := r.tourIdx()
	var a, b C.double
	var acc C.int32_t
	ahxCheck(r.ahx.ctx, C.ahx_flip_all(r.ahx.ctx, C.int32_t(len(t)), &t[0],
		(*C.uint8_t)(unsafe.Pointer(&r.Signs[0])), &a, &b, &acc), "ahx_flip_all")
	flipLog("FLIPALL", float64(a), float64(b), tagOf(acc))
	return tagOf(acc)
}
