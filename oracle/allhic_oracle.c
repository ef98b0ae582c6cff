/*
 * allhic_oracle.c -- CPU restatement ("oracle") of the `allhic optimize` hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under allhic_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU baseline -- never as the product.
 *
 * PARITY UNPINNED: the reference (tanghaibao/allhic, Go) cannot be built in
 * this image (no Go toolchain, un-vendored eaopt/gonum modules, no network)
 * and the reference's own tests hold no golden vector for this path (the only
 * test that runs `optimize` asserts the string "Success",
 * functional-tests.sh:10-14).  This file is therefore a line-by-line
 * restatement of the Go sources, cross-checked against an independent
 * pure-Python restatement (tests/pyref.py) and the known-answer values in
 * BASELINE.md section 2 -- not against outputs of the Go binary.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repo root).  Loop order, break/continue semantics, integer
 * widths (Go int = int64) and floating-point evaluation order are kept
 * identical; compile with -O2 -ffp-contract=off and without -ffast-math.
 *
 * Third-party arithmetic that is not in the reference checkout:
 *   - Go stdlib math.Log (pure-Go port of FreeBSD/FDLIBM e_log.c): restated
 *     here as orc_go_log from the published algorithm.
 *   - github.com/MaxHalford/eaopt v0.4.2 (go.mod:6): GA engine; holds no score
 *     arithmetic.  orc_ga_run restates its published behaviour
 *     (ModGenerational + SelTournament{3} + HallOfFame(1) + early stop) as
 *     summarised in SURVEY.md 3.2.  The RNG stream is NOT Go's math/rand ALFG.
 *   - github.com/gonum/matrix mat64.EigenSym (go.mod:17): restated with a
 *     cyclic Jacobi eigensolver (orc_flip_all); eigenvector sign is
 *     implementation defined in both.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants: base.go:26-103, evaluate.go:21-25 ------------------------ */
#define ORC_LB 18                     /* base.go:30 */
#define ORC_UB 29                     /* base.go:32 */
#define ORC_BB (ORC_UB - ORC_LB + 1)  /* base.go:34 */
static const double ORC_PHI = 0.4812118250596684; /* base.go:36 */
#define ORC_LIMIT 10000000            /* evaluate.go:22 */
static const int64_t ORC_GR[ORC_BB] = {5778,   9349,   15127,  24476,
                                       39603,  64079,  103682, 167761,
                                       271443, 439204, 710647, 1149851}; /* base.go:101-103 */

typedef struct { int64_t v[ORC_BB]; } orc_garray; /* base.go:98 GArray [BB]int */

/* ---- Go math.Log (FDLIBM e_log.c algorithm, as ported in Go's math/log.go) */
double orc_go_log(double x) {
  static const double Ln2Hi = 6.93147180369123816490e-01, /* 3fe62e42 fee00000 */
      Ln2Lo = 1.90821492927058770002e-10,                 /* 3dea39ef 35793c76 */
      L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
      L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
      L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
      L7 = 1.479819860511658591e-01;
  if (isnan(x) || (isinf(x) && x > 0)) return x;
  if (x < 0) return NAN;
  if (x == 0) return -INFINITY;
  int ki;
  double f1 = frexp(x, &ki);
  if (f1 < 0.70710678118654752440 /* Sqrt2/2 */) {
    f1 *= 2;
    ki--;
  }
  double f = f1 - 1;
  double k = (double)ki;
  double s = f / (2 + f);
  double s2 = s * s;
  double s4 = s2 * s2;
  double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
  double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
  double R = t1 + t2;
  double hfsq = 0.5 * f * f;
  return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

/* base.go:130-135 Round: half away from zero */
static double orc_round(double input) {
  if (input < 0) return ceil(input - 0.5);
  return floor(input + 0.5);
}

/* base.go:138-144 SumLog */
double orc_sumlog(const int64_t *a, int64_t n) {
  double sum = 0.0;
  for (int64_t i = 0; i < n; i++) sum += orc_go_log((double)a[i]);
  return sum;
}

/* base.go:156-167 GoldenArray */
void orc_golden_array(const int64_t *a, int64_t n, int64_t *counts /*[12]*/) {
  for (int k = 0; k < ORC_BB; k++) counts[k] = 0;
  for (int64_t i = 0; i < n; i++) {
    double r = orc_round(orc_go_log((double)a[i]) / ORC_PHI);
    int64_t c;
    /* Go int(NaN/-Inf) on amd64 yields math.MinInt64 => clamps to LB */
    if (!(r >= -9.0e18)) c = ORC_LB;
    else if (r > 9.0e18) c = ORC_UB;
    else c = (int64_t)r;
    if (c < ORC_LB) c = ORC_LB;
    else if (c > ORC_UB) c = ORC_UB;
    counts[c - ORC_LB]++;
  }
}

/* ---- small containers ---------------------------------------------------- */
typedef struct {
  uint64_t *keys;
  int64_t *vals;
  uint64_t cap; /* power of two */
  uint64_t n;
} orc_map;

static uint64_t orc_mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
#define ORC_EMPTY 0xffffffffffffffffULL
static void orc_map_init(orc_map *m, uint64_t cap) {
  m->cap = 16;
  while (m->cap < cap) m->cap <<= 1;
  m->keys = (uint64_t *)malloc(m->cap * 8);
  m->vals = (int64_t *)malloc(m->cap * 8);
  memset(m->keys, 0xff, m->cap * 8);
  m->n = 0;
}
static void orc_map_free(orc_map *m) { free(m->keys); free(m->vals); m->keys = NULL; m->vals = NULL; }
static int64_t orc_map_get(const orc_map *m, uint64_t key) {
  uint64_t h = orc_mix(key) & (m->cap - 1);
  while (m->keys[h] != ORC_EMPTY) {
    if (m->keys[h] == key) return m->vals[h];
    h = (h + 1) & (m->cap - 1);
  }
  return -1;
}
static void orc_map_put(orc_map *m, uint64_t key, int64_t val);
static void orc_map_grow(orc_map *m) {
  orc_map o = *m;
  orc_map_init(m, o.cap * 2);
  for (uint64_t i = 0; i < o.cap; i++)
    if (o.keys[i] != ORC_EMPTY) orc_map_put(m, o.keys[i], o.vals[i]);
  orc_map_free(&o);
}
static void orc_map_put(orc_map *m, uint64_t key, int64_t val) {
  if ((m->n + 1) * 10 > m->cap * 7) orc_map_grow(m);
  uint64_t h = orc_mix(key) & (m->cap - 1);
  while (m->keys[h] != ORC_EMPTY) {
    if (m->keys[h] == key) { m->vals[h] = val; return; }
    h = (h + 1) & (m->cap - 1);
  }
  m->keys[h] = key; m->vals[h] = val; m->n++;
}

/* ---- domain model: clm.go:32-91 ------------------------------------------ */
typedef struct { /* clm.go:67-71 Contact */
  int32_t ai, bi;
  int64_t strandedness, nlinks;
  double meanDist;
} orc_contact;

typedef struct { /* clm.go:59-64 OrientedPair + its GArray value */
  int32_t ai, bi;
  uint8_t ao, bo;
  orc_garray g;
} orc_oriented;

typedef struct orc_clm {
  int64_t N;          /* len(r.Tigs) */
  char **names;       /* TigF.Name */
  int64_t *sizes;     /* TigF.Size */
  uint8_t *active;    /* TigF.IsActive */
  orc_map name2idx;   /* tigToIdx (hash of the name string -> idx, verified by strcmp) */
  /* contacts map[Pair]Contact */
  orc_contact *contacts; int64_t ncontacts, ccap; orc_map cmap;
  /* orientedContacts map[OrientedPair]GArray */
  orc_oriented *oriented; int64_t noriented, ocap; orc_map omap;
  /* Tour (clm.go:88-91): Tigs {Idx,Size}, M shared */
  int64_t L; int64_t *tour_idx; int64_t *tour_size;
  int64_t **M;        /* [][]int, row-pointer layout like Go (clm.go:437) */
  uint8_t *signs;     /* r.Signs, '+'/'-' indexed by contig */
  char err[256];
} orc_clm;

static uint64_t orc_strhash(const char *s, size_t n) {
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 1099511628211ULL; }
  return h & 0x7fffffffffffffffULL;
}
/* name -> idx with linear probing on hash collisions of distinct strings */
static int64_t orc_lookup_name(const orc_clm *c, const char *s, size_t n) {
  uint64_t h = orc_strhash(s, n);
  for (;;) {
    int64_t idx = orc_map_get(&c->name2idx, h);
    if (idx < 0) return -1;
    if (strlen(c->names[idx]) == n && memcmp(c->names[idx], s, n) == 0) return idx;
    h = (h + 1) & 0x7fffffffffffffffULL;
  }
}
static void orc_insert_name(orc_clm *c, const char *s, size_t n, int64_t idx) {
  /* Go map assignment: a repeated name re-points the key at the later idx */
  uint64_t h = orc_strhash(s, n);
  for (;;) {
    int64_t cur = orc_map_get(&c->name2idx, h);
    if (cur < 0 || (strlen(c->names[cur]) == n && memcmp(c->names[cur], s, n) == 0)) {
      orc_map_put(&c->name2idx, h, idx);
      return;
    }
    h = (h + 1) & 0x7fffffffffffffffULL;
  }
}

/* strconv.Atoi with the error discarded: invalid input -> 0 (clm.go:152,186,190) */
static int64_t orc_atoi(const char *s, size_t n) {
  if (n == 0) return 0;
  size_t i = 0; int neg = 0;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; if (n == 1) return 0; }
  int64_t v = 0;
  for (; i < n; i++) {
    if (s[i] < '0' || s[i] > '9') return 0;
    v = v * 10 + (s[i] - '0');
  }
  return neg ? -v : v;
}

static int orc_isspace(char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\v' || ch == '\f'; }

static char *orc_slurp(const char *path, size_t *len) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)sz + 1);
  if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return NULL; }
  fclose(f);
  buf[sz] = 0;
  *len = (size_t)sz;
  return buf;
}

/* clm.go:141-157 readRE: name = first field, size = LAST field, '#' skipped */
static int orc_read_re(orc_clm *c, const char *refile) {
  size_t len; char *buf = orc_slurp(refile, &len);
  if (!buf) { snprintf(c->err, sizeof c->err, "open %s failed", refile); return -1; }
  int64_t cap = 1024;
  c->names = (char **)malloc(cap * sizeof(char *));
  c->sizes = (int64_t *)malloc(cap * 8);
  orc_map_init(&c->name2idx, 4096);
  size_t p = 0; int64_t idx = 0;
  while (p < len) {
    size_t e = p;
    while (e < len && buf[e] != '\n') e++;
    /* strings.Fields */
    size_t q = p; const char *first = NULL; size_t firstn = 0; const char *last = NULL; size_t lastn = 0;
    while (q < e) {
      while (q < e && orc_isspace(buf[q])) q++;
      if (q >= e) break;
      size_t w = q;
      while (q < e && !orc_isspace(buf[q])) q++;
      if (!first) { first = buf + w; firstn = q - w; }
      last = buf + w; lastn = q - w;
    }
    if (first && first[0] != '#') { /* Go would panic on an empty line; we skip it */
      if (idx == cap) { cap *= 2; c->names = (char **)realloc(c->names, cap * sizeof(char *)); c->sizes = (int64_t *)realloc(c->sizes, cap * 8); }
      c->names[idx] = strndup(first, firstn);
      c->sizes[idx] = orc_atoi(last, lastn);
      orc_insert_name(c, first, firstn, idx);
      idx++;
    }
    p = e + 1;
  }
  free(buf);
  c->N = idx;
  c->active = (uint8_t *)malloc((size_t)(idx ? idx : 1));
  memset(c->active, 1, (size_t)idx); /* TigF{idx,tig,size,true} */
  return 0;
}

static uint8_t orc_rr(uint8_t b) { return b == '-' ? '+' : '-'; } /* clm.go:160-165 */

static uint64_t orc_pair_key(int64_t ai, int64_t bi) { return ((uint64_t)ai << 32) | (uint64_t)bi; }
static uint64_t orc_okey(int64_t ai, int64_t bi, uint8_t ao, uint8_t bo) {
  return ((uint64_t)ai << 40) | ((uint64_t)bi << 16) | ((uint64_t)ao << 8) | bo;
}

static void orc_set_oriented(orc_clm *c, int64_t ai, int64_t bi, uint8_t ao, uint8_t bo, const orc_garray *g) {
  uint64_t k = orc_okey(ai, bi, ao, bo);
  int64_t at = orc_map_get(&c->omap, k);
  if (at < 0) {
    if (c->noriented == c->ocap) { c->ocap = c->ocap ? c->ocap * 2 : 1024; c->oriented = (orc_oriented *)realloc(c->oriented, c->ocap * sizeof(orc_oriented)); }
    at = c->noriented++;
    orc_map_put(&c->omap, k, at);
    c->oriented[at].ai = (int32_t)ai; c->oriented[at].bi = (int32_t)bi;
    c->oriented[at].ao = ao; c->oriented[at].bo = bo;
  }
  c->oriented[at].g = *g; /* map assignment: last writer wins */
}

/* clm.go:168-240 readClmLines + readClm, fused (lines are consumed in file order) */
static int orc_read_clm(orc_clm *c, const char *clmfile) {
  size_t len; char *buf = orc_slurp(clmfile, &len);
  if (!buf) { snprintf(c->err, sizeof c->err, "open %s failed", clmfile); return -1; }
  orc_map_init(&c->cmap, 4096);
  orc_map_init(&c->omap, 16384);
  int64_t *dists = NULL; int64_t dcap = 0;
  size_t p = 0;
  while (p < len) {
    size_t e = p;
    while (e < len && buf[e] != '\n') e++;
    size_t next = e + 1;
    /* strings.TrimSpace(row) */
    size_t a = p, b = e;
    while (a < b && orc_isspace(buf[a])) a++;
    while (b > a && orc_isspace(buf[b - 1])) b--;
    if (a == b) { /* Go: empty row at EOF ends the loop; elsewhere it would panic */
      if (next >= len) break;
      snprintf(c->err, sizeof c->err, "empty CLM line (the reference panics here)");
      free(buf); free(dists); return -2;
    }
    /* words = Split(row, "\t"); abtig = Split(words[0], " ") */
    size_t t1 = a; while (t1 < b && buf[t1] != '\t') t1++;
    size_t sp = a; while (sp < t1 && buf[sp] != ' ') sp++;
    if (sp >= t1 || sp == a || sp + 1 >= t1 || t1 >= b) { snprintf(c->err, sizeof c->err, "malformed CLM line (the reference panics here)"); free(buf); free(dists); return -2; }
    /* abtig[1] = text up to the next space inside words[0] */
    size_t sp2 = sp + 1; while (sp2 < t1 && buf[sp2] != ' ') sp2++;
    const char *at = buf + a; size_t atn = sp - a - 1; uint8_t ao = (uint8_t)buf[sp - 1];
    const char *bt = buf + sp + 1; size_t btn = sp2 - (sp + 1) - 1; uint8_t bo = (uint8_t)buf[sp2 - 1];
    size_t t2 = t1 + 1; while (t2 < b && buf[t2] != '\t') t2++;
    /* words[1] = nlinks (only used for the Malformed warning), words[2] = distances */
    int64_t nl = 0;
    size_t q = (t2 < b) ? t2 + 1 : b;
    size_t wend = q; while (wend < b && buf[wend] != '\t') wend++;
    if (t2 >= b) { snprintf(c->err, sizeof c->err, "malformed CLM line: no distance column (the reference panics here)"); free(buf); free(dists); return -2; }
    /* Split(words[2], " "): every separator yields a token (empty tokens -> Atoi error -> 0) */
    for (;;) {
      size_t w = q;
      while (q < wend && buf[q] != ' ') q++;
      if (nl == dcap) { dcap = dcap ? dcap * 2 : 256; dists = (int64_t *)realloc(dists, dcap * 8); }
      dists[nl++] = orc_atoi(buf + w, q - w);
      if (q >= wend) break;
      q++;
    }
    /* readClm body, clm.go:209-239 */
    int64_t ai = orc_lookup_name(c, at, atn);
    int64_t bi = (ai >= 0) ? orc_lookup_name(c, bt, btn) : -1;
    if (ai >= 0 && bi >= 0) {
      orc_garray g;
      orc_golden_array(dists, nl, g.v);
      double meanDist = orc_sumlog(dists, nl);
      int64_t strandedness = (ao != bo) ? -1 : 1;
      uint64_t pk = orc_pair_key(ai, bi);
      int64_t at_c = orc_map_get(&c->cmap, pk);
      if (at_c >= 0) {
        if (meanDist < c->contacts[at_c].meanDist) {
          c->contacts[at_c].strandedness = strandedness;
          c->contacts[at_c].nlinks = nl;
          c->contacts[at_c].meanDist = meanDist;
        }
      } else {
        if (c->ncontacts == c->ccap) { c->ccap = c->ccap ? c->ccap * 2 : 1024; c->contacts = (orc_contact *)realloc(c->contacts, c->ccap * sizeof(orc_contact)); }
        at_c = c->ncontacts++;
        orc_map_put(&c->cmap, pk, at_c);
        c->contacts[at_c].ai = (int32_t)ai; c->contacts[at_c].bi = (int32_t)bi;
        c->contacts[at_c].strandedness = strandedness;
        c->contacts[at_c].nlinks = nl;
        c->contacts[at_c].meanDist = meanDist;
      }
      orc_set_oriented(c, ai, bi, ao, bo, &g);
      orc_set_oriented(c, bi, ai, orc_rr(bo), orc_rr(ao), &g);
    }
    p = next;
  }
  free(buf); free(dists);
  return 0;
}

static char orc_global_err[256];
const char *orc_last_error(void) { return orc_global_err; }

void orc_free_clm(orc_clm *c);

/* clm.go:121-134 NewCLM */
orc_clm *orc_new_clm(const char *clmfile, const char *refile) {
  orc_clm *c = (orc_clm *)calloc(1, sizeof(orc_clm));
  if (orc_read_re(c, refile) != 0 || orc_read_clm(c, clmfile) != 0) {
    snprintf(orc_global_err, sizeof orc_global_err, "%s", c->err);
    orc_free_clm(c);
    return NULL;
  }
  c->signs = (uint8_t *)malloc((size_t)(c->N ? c->N : 1));
  memset(c->signs, '+', (size_t)c->N);
  return c;
}

void orc_free_clm(orc_clm *c) {
  if (!c) return;
  for (int64_t i = 0; i < c->N; i++) free(c->names[i]);
  free(c->names); free(c->sizes); free(c->active);
  if (c->name2idx.keys) orc_map_free(&c->name2idx);
  if (c->cmap.keys) orc_map_free(&c->cmap);
  if (c->omap.keys) orc_map_free(&c->omap);
  free(c->contacts); free(c->oriented);
  free(c->tour_idx); free(c->tour_size);
  if (c->M) { for (int64_t i = 0; i < c->N; i++) free(c->M[i]); free(c->M); }
  free(c->signs);
  free(c);
}

/* ---- accessors used by the tests ----------------------------------------- */
int64_t orc_ntigs(const orc_clm *c) { return c->N; }
const char *orc_tig_name(const orc_clm *c, int64_t i) { return c->names[i]; }
int64_t orc_tig_size(const orc_clm *c, int64_t i) { return c->sizes[i]; }
void orc_get_sizes(const orc_clm *c, int64_t *out) { memcpy(out, c->sizes, (size_t)c->N * 8); }
int64_t orc_ncontacts(const orc_clm *c) { return c->ncontacts; }
void orc_get_contacts(const orc_clm *c, int32_t *ai, int32_t *bi, int64_t *strand, int64_t *nlinks, double *meandist) {
  for (int64_t i = 0; i < c->ncontacts; i++) {
    ai[i] = c->contacts[i].ai; bi[i] = c->contacts[i].bi;
    strand[i] = c->contacts[i].strandedness; nlinks[i] = c->contacts[i].nlinks; meandist[i] = c->contacts[i].meanDist;
  }
}
int64_t orc_noriented(const orc_clm *c) { return c->noriented; }
void orc_get_oriented(const orc_clm *c, int32_t *ai, int32_t *bi, uint8_t *ao, uint8_t *bo, int64_t *g /* n x 12 */) {
  for (int64_t i = 0; i < c->noriented; i++) {
    ai[i] = c->oriented[i].ai; bi[i] = c->oriented[i].bi; ao[i] = c->oriented[i].ao; bo[i] = c->oriented[i].bo;
    memcpy(g + i * ORC_BB, c->oriented[i].g.v, ORC_BB * 8);
  }
}
void orc_set_active(orc_clm *c, const uint8_t *active) { memcpy(c->active, active, (size_t)c->N); }
void orc_set_signs(orc_clm *c, const uint8_t *signs) { memcpy(c->signs, signs, (size_t)c->N); }
void orc_get_signs(const orc_clm *c, uint8_t *signs) { memcpy(signs, c->signs, (size_t)c->N); }
int64_t orc_tour_len(const orc_clm *c) { return c->L; }
void orc_get_tour(const orc_clm *c, int64_t *idx) { memcpy(idx, c->tour_idx, (size_t)c->L * 8); }

/* clm.go:437-447 M(): dense N x N [][]int, symmetric fill from contacts */
static void orc_build_M(orc_clm *c) {
  if (c->M) return;
  c->M = (int64_t **)malloc((size_t)(c->N ? c->N : 1) * sizeof(int64_t *)); /* base.go:279-285 Make2DSlice */
  for (int64_t i = 0; i < c->N; i++) c->M[i] = (int64_t *)calloc((size_t)c->N, 8);
  /* Go iterates the map in random order; for the (never produced by `extract`)
   * case of both (a,b) and (b,a) keys we apply them in first-insertion order. */
  for (int64_t i = 0; i < c->ncontacts; i++) {
    c->M[c->contacts[i].ai][c->contacts[i].bi] = c->contacts[i].nlinks;
    c->M[c->contacts[i].bi][c->contacts[i].ai] = c->contacts[i].nlinks;
  }
}
void orc_get_M(orc_clm *c, int64_t *out /* N*N */) {
  orc_build_M(c);
  for (int64_t i = 0; i < c->N; i++) memcpy(out + i * c->N, c->M[i], (size_t)c->N * 8);
}

/* orientation.go:124-132 O(): strandedness * nlinks, symmetric */
void orc_get_O(const orc_clm *c, double *out /* N*N */) {
  memset(out, 0, (size_t)(c->N * c->N) * 8);
  for (int64_t i = 0; i < c->ncontacts; i++) {
    double s = (double)(c->contacts[i].strandedness * c->contacts[i].nlinks);
    int64_t a = c->contacts[i].ai, b = c->contacts[i].bi;
    /* mat64.SymDense.SetSym stores the upper triangle; (a,b) and (b,a) alias */
    out[a * c->N + b] = s; out[b * c->N + a] = s;
  }
}

/* Set r.Tour from a list of contig indices (sizes looked up), as
 * clm.go:399-407 builds it from the active tigs. */
void orc_set_tour(orc_clm *c, int64_t L, const int64_t *idx) {
  orc_build_M(c);
  free(c->tour_idx); free(c->tour_size);
  c->L = L;
  c->tour_idx = (int64_t *)malloc((size_t)(L ? L : 1) * 8);
  c->tour_size = (int64_t *)malloc((size_t)(L ? L : 1) * 8);
  for (int64_t i = 0; i < L; i++) { c->tour_idx[i] = idx[i]; c->tour_size[i] = c->sizes[idx[i]]; }
}

/* evaluate.go:116-143 Tour.Evaluate ("EvaluateM"); literal loop order. */
static double orc_eval_core(int64_t *const *M, const int64_t *sizes, int64_t size, const int64_t *tigs, double *mid) {
  double cumSum = 0.0;
  for (int64_t i = 0; i < size; i++) {
    double tsize = (double)sizes[tigs[i]];
    mid[i] = cumSum + tsize / 2;
    cumSum += tsize;
  }
  double score = 0.0;
  for (int64_t i = 0; i < size; i++) {
    int64_t a = tigs[i];
    for (int64_t j = i + 1; j < size; j++) {
      int64_t b = tigs[j];
      int64_t nlinks = M[a][b];
      double dist = mid[j] - mid[i];
      if (dist > ORC_LIMIT) break;
      score -= (double)nlinks / dist;
    }
  }
  return score;
}
double orc_tour_evaluate(orc_clm *c, int64_t L, const int64_t *idx) {
  orc_build_M(c);
  double *mid = (double *)malloc((size_t)(L ? L : 1) * 8);
  double s = orc_eval_core(c->M, c->sizes, L, idx, mid);
  free(mid);
  return s;
}
double orc_evaluate(orc_clm *c) { return orc_tour_evaluate(c, c->L, c->tour_idx); }

/* W(t): number of (i<j) pair terms Tour.Evaluate visits before the LIMIT break */
int64_t orc_count_W(const orc_clm *c, int64_t L, const int64_t *idx) {
  double *mid = (double *)malloc((size_t)(L ? L : 1) * 8);
  double cum = 0;
  for (int64_t i = 0; i < L; i++) { double t = (double)c->sizes[idx[i]]; mid[i] = cum + t / 2; cum += t; }
  int64_t W = 0;
  for (int64_t i = 0; i < L; i++)
    for (int64_t j = i + 1; j < L; j++) { if (mid[j] - mid[i] > ORC_LIMIT) break; W++; }
  free(mid);
  return W;
}

/* Population evaluate on `nthreads` threads, contiguous chunks of
 * ceil(P/nthreads) individuals: what eaopt's ParallelEval=true does
 * (evaluate.go:279).  tours: P x L int64. */
typedef struct { orc_clm *c; int64_t L; const int64_t *tours; double *out; int64_t lo, hi; } orc_pe_arg;
static void *orc_pe_worker(void *p) {
  orc_pe_arg *a = (orc_pe_arg *)p;
  double *mid = (double *)malloc((size_t)(a->L ? a->L : 1) * 8);
  for (int64_t i = a->lo; i < a->hi; i++) a->out[i] = orc_eval_core(a->c->M, a->c->sizes, a->L, a->tours + i * a->L, mid);
  free(mid);
  return NULL;
}
void orc_population_evaluate(orc_clm *c, int64_t P, int64_t L, const int64_t *tours, double *out, int nthreads) {
  orc_build_M(c);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > P) nthreads = (int)(P > 0 ? P : 1);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  orc_pe_arg *args = (orc_pe_arg *)malloc(sizeof(orc_pe_arg) * (size_t)nthreads);
  int64_t chunk = (P + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    int64_t lo = t * chunk, hi = lo + chunk; if (hi > P) hi = P; if (lo > P) lo = P;
    args[t] = (orc_pe_arg){c, L, tours, out, lo, hi};
    pthread_create(&th[t], NULL, orc_pe_worker, &args[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(args);
}

/* ---- mutation operators: evaluate.go:146-232 ------------------------------
 * The random draws are passed in explicitly so tests can replay the exact
 * same operator on the device. */
void orc_mut_permute(int64_t *t, int64_t L, int64_t p, int64_t q) { /* evaluate.go:200-208 */
  if (L <= 1) return;
  int64_t x = t[p]; t[p] = t[q]; t[q] = x;
}
void orc_mut_splice(int64_t *t, int64_t L, int64_t k) { /* evaluate.go:212-218: b.Append(a) */
  int64_t *tmp = (int64_t *)malloc((size_t)L * 8);
  memcpy(tmp, t + k, (size_t)(L - k) * 8);
  memcpy(tmp + (L - k), t, (size_t)k * 8);
  memcpy(t, tmp, (size_t)L * 8);
  free(tmp);
}
void orc_mut_insertion(int64_t *t, int64_t L, int64_t p, int64_t q, int front) { /* evaluate.go:172-197 */
  (void)L;
  if (p == q) return;
  if (front) { /* rng.Float64() < .5: pop q, insert at p */
    int64_t cq = t[q];
    for (int64_t i = q; i > p; i--) t[i] = t[i - 1];
    t[p] = cq;
  } else {
    int64_t cp = t[p];
    for (int64_t i = p; i < q; i++) t[i] = t[i + 1];
    t[q] = cp;
  }
}
void orc_mut_inversion(int64_t *t, int64_t L, int64_t p, int64_t q) { /* evaluate.go:157-169 */
  (void)L;
  if (p == q) return;
  for (int64_t i = p, j = q; i < j; i++, j--) { int64_t x = t[i]; t[i] = t[j]; t[j] = x; }
}

/* ---- RNG for the GA emulation (NOT Go's math/rand stream) ----------------- */
typedef struct { uint64_t s[4]; } orc_rng;
static uint64_t orc_splitmix(uint64_t *x) {
  uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static void orc_rng_seed(orc_rng *r, uint64_t seed) { for (int i = 0; i < 4; i++) r->s[i] = orc_splitmix(&seed); }
static uint64_t orc_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t orc_rng_next(orc_rng *r) { /* xoshiro256** */
  uint64_t res = orc_rotl(r->s[1] * 5, 7) * 9, t = r->s[1] << 17;
  r->s[2] ^= r->s[0]; r->s[3] ^= r->s[1]; r->s[1] ^= r->s[2]; r->s[0] ^= r->s[3];
  r->s[2] ^= t; r->s[3] = orc_rotl(r->s[3], 45);
  return res;
}
static int64_t orc_intn(orc_rng *r, int64_t n) { /* uniform in [0,n) by rejection */
  uint64_t un = (uint64_t)n, lim = UINT64_MAX - (UINT64_MAX % un);
  uint64_t v;
  do { v = orc_rng_next(r); } while (v >= lim);
  return (int64_t)(v % un);
}
static double orc_float64(orc_rng *r) { return (double)(orc_rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* evaluate.go:248-255 Shuffle (Fisher-Yates, j = i + Intn(N-i)) */
static void orc_shuffle_rng(int64_t *t, int64_t N, orc_rng *r) {
  for (int64_t i = 0; i < N; i++) {
    int64_t j = i + orc_intn(r, N - i);
    int64_t x = t[j]; t[j] = t[i]; t[i] = x;
  }
}
void orc_shuffle(int64_t *t, int64_t N, uint64_t seed) { orc_rng r; orc_rng_seed(&r, seed); orc_shuffle_rng(t, N, &r); }

/* evaluate.go:146-154 randomTwoInts */
static void orc_two_ints(orc_rng *r, int64_t n, int64_t *p, int64_t *q) {
  *p = orc_intn(r, n); *q = orc_intn(r, n);
  if (*p > *q) { int64_t x = *p; *p = *q; *q = x; }
}
/* evaluate.go:221-232 Tour.Mutate: the draws (which operator, where), then the operator.
 * op: 0 MutPermute, 1 MutSplice (p = k), 2 MutInsertion, 3 MutInversion, -1 nothing to do. */
static void orc_mutate_draw(orc_rng *r, int64_t L, int *op, int64_t *p, int64_t *q, int *front) {
  double rd = orc_float64(r);
  *op = -1; *p = 0; *q = 0; *front = 0;
  if (rd < 0.2) { if (L <= 1) return; orc_two_ints(r, L, p, q); *op = 0; }
  else if (rd < .4) { *p = orc_intn(r, L - 1) + 1; *op = 1; }
  else if (rd < .7) { orc_two_ints(r, L, p, q); *op = 2; if (*p == *q) return; *front = orc_float64(r) < .5; }
  else { orc_two_ints(r, L, p, q); *op = 3; }
}
static void orc_mutate(int64_t *t, int64_t L, orc_rng *r) {
  int op, front; int64_t p, q;
  orc_mutate_draw(r, L, &op, &p, &q, &front);
  if (op == 0) orc_mut_permute(t, L, p, q);
  else if (op == 1) orc_mut_splice(t, L, p);
  else if (op == 2) { if (p != q) orc_mut_insertion(t, L, p, q, front); }
  else if (op == 3) orc_mut_inversion(t, L, p, q);
}
/* test support: n draws of Tour.Mutate's decisions for a tour of length L; out: n x 4 = op, p, q, front */
void orc_mutate_sample(int64_t L, uint64_t seed, int64_t n, int64_t *out) {
  orc_rng r; orc_rng_seed(&r, seed);
  for (int64_t i = 0; i < n; i++) {
    int op, front; int64_t p, q;
    orc_mutate_draw(&r, L, &op, &p, &q, &front);
    out[4 * i] = op; out[4 * i + 1] = p; out[4 * i + 2] = q; out[4 * i + 3] = front;
  }
}

/* ---- GA: evaluate.go:258-315 wired onto an eaopt v0.4.2 emulation ---------- */
typedef struct {
  double best_score;      /* -HallOfFame[0].Fitness */
  int64_t generations;    /* ga.Generations at stop */
  int64_t evaluations;    /* Tour.Evaluate calls made */
  double seconds;         /* wall time of the run */
} orc_ga_result;

#include <time.h>
static double orc_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

/* eaopt SelTournament{NContestants:k}.Apply(2,...): two distinct winners; each
 * the fittest of k contestants sampled without replacement from the not yet
 * selected individuals. */
static void orc_tournament2(const double *fit, int64_t P, int k, orc_rng *r, int64_t *w1, int64_t *w2) {
  int64_t banned = -1;
  for (int round = 0; round < 2; round++) {
    int64_t pool = P - (banned >= 0 ? 1 : 0);
    int64_t c[16]; int kk = k > 16 ? 16 : k; if (kk > pool) kk = (int)pool;
    int64_t best = -1;
    for (int i = 0; i < kk; i++) {
      int64_t v;
      for (;;) { /* sample without replacement */
        v = orc_intn(r, pool);
        int dup = 0;
        for (int j = 0; j < i; j++) if (c[j] == v) dup = 1;
        if (!dup) break;
      }
      c[i] = v;
      int64_t id = (banned >= 0 && v >= banned) ? v + 1 : v;
      if (best < 0 || fit[id] < fit[best]) best = id;
    }
    if (round == 0) { *w1 = best; banned = best; } else *w2 = best;
  }
}

/* test support: n draws of SelTournament{k}.Apply(2, ...) over fit[P]; w: n x 2 winners */
void orc_tournament_sample(const double *fit, int64_t P, int k, uint64_t seed, int64_t n, int64_t *w) {
  orc_rng r; orc_rng_seed(&r, seed);
  for (int64_t i = 0; i < n; i++) orc_tournament2(fit, P, k, &r, &w[2 * i], &w[2 * i + 1]);
}

/* Runs GARun for one phase starting from c's current tour; on return c's tour
 * is HallOfFame[0] (evaluate.go:313).  `max_gens` bounds ga.NGenerations
 * (1e6 in the reference, evaluate.go:270); `max_seconds` (<=0: none) lets the
 * bench take a bounded sample.  log_every=500 in the reference. */
int orc_ga_run(orc_clm *c, int64_t npop, int64_t ngen, double mutprob, uint64_t seed, int64_t max_gens,
               double max_seconds, int nthreads, int verbose_phase, orc_ga_result *res) {
  orc_build_M(c);
  int64_t L = c->L, P = npop;
  if (P < 4 || L < 2) { snprintf(orc_global_err, sizeof orc_global_err, "GA needs npop>=4 and L>=2"); return -1; }
  orc_rng rng; orc_rng_seed(&rng, seed);
  int64_t *pop = (int64_t *)malloc((size_t)(P * L) * 8), *off = (int64_t *)malloc((size_t)(P * L) * 8);
  double *fit = (double *)malloc((size_t)P * 8), *ofit = (double *)malloc((size_t)P * 8);
  uint8_t *need = (uint8_t *)malloc((size_t)P);
  int64_t *todo = (int64_t *)malloc((size_t)P * 8);
  int64_t *ctours = (int64_t *)malloc((size_t)(P * L) * 8);
  double *cout = (double *)malloc((size_t)P * 8);
  int64_t *hof = (int64_t *)malloc((size_t)L * 8);
  double t0 = orc_now();
  int64_t evals = 0;
  /* ga.init: PopSize clones of the same tour (evaluate.go:259-262), evaluated */
  for (int64_t i = 0; i < P; i++) memcpy(pop + i * L, c->tour_idx, (size_t)L * 8);
  orc_population_evaluate(c, P, L, pop, fit, nthreads); evals += P;
  int64_t bi = 0; for (int64_t i = 1; i < P; i++) if (fit[i] < fit[bi]) bi = i;
  double hof_fit = fit[bi]; memcpy(hof, pop + bi * L, (size_t)L * 8);
  double best = -1.7976931348623157e308; int64_t updated = 0, gen = 0; /* evaluate.go:281-285 */
  for (;;) {
    /* Callback, evaluate.go:287-301 */
    double currentBest = -hof_fit;
    if (currentBest > best) { best = currentBest; updated = gen; }
    if (verbose_phase > 0 && gen % 500 == 0) printf("Current iteration GA%d-%lld: max_score=%.5f\n", verbose_phase, (long long)gen, currentBest);
    if (gen >= max_gens) break;
    if (gen - updated > ngen) break; /* EarlyStop, evaluate.go:304-306 */
    if (max_seconds > 0 && orc_now() - t0 > max_seconds) break;
    /* ModGenerational.Apply */
    for (int64_t i = 0; i < P; i += 2) {
      int64_t w1, w2; orc_tournament2(fit, P, 3, &rng, &w1, &w2);
      memcpy(off + i * L, pop + w1 * L, (size_t)L * 8); ofit[i] = fit[w1];
      if (i + 1 < P) { memcpy(off + (i + 1) * L, pop + w2 * L, (size_t)L * 8); ofit[i + 1] = fit[w2]; }
    }
    int64_t ntodo = 0;
    for (int64_t i = 0; i < P; i++) {
      need[i] = 0;
      if (orc_float64(&rng) < mutprob) { orc_mutate(off + i * L, L, &rng); need[i] = 1; todo[ntodo++] = i; }
    }
    /* Individuals.Evaluate: only individuals with Evaluated=false are re-scored */
    for (int64_t k = 0; k < ntodo; k++) memcpy(ctours + k * L, off + todo[k] * L, (size_t)L * 8);
    orc_population_evaluate(c, ntodo, L, ctours, cout, nthreads); evals += ntodo;
    for (int64_t k = 0; k < ntodo; k++) ofit[todo[k]] = cout[k];
    { int64_t *tp = pop; pop = off; off = tp; double *tf = fit; fit = ofit; ofit = tf; }
    /* SortByFitness + updateHallOfFame(size 1): keep the best ever */
    bi = 0; for (int64_t i = 1; i < P; i++) if (fit[i] < fit[bi]) bi = i;
    if (fit[bi] < hof_fit) { hof_fit = fit[bi]; memcpy(hof, pop + bi * L, (size_t)L * 8); }
    gen++;
  }
  memcpy(c->tour_idx, hof, (size_t)L * 8); /* r.Tour = ga.HallOfFame[0] */
  for (int64_t i = 0; i < L; i++) c->tour_size[i] = c->sizes[c->tour_idx[i]];
  if (res) { res->best_score = -hof_fit; res->generations = gen; res->evaluations = evals; res->seconds = orc_now() - t0; }
  free(pop); free(off); free(fit); free(ofit); free(need); free(todo); free(ctours); free(cout); free(hof);
  return 0;
}

/* ---- orientation: orientation.go ------------------------------------------ */

/* orientation.go:137-153 Q(): dense N x N x 12 rebuilt from the oriented map
 * for the CURRENT signs; sentinel -1 in slot 0 when there is no entry. */
static orc_garray *orc_build_Q(const orc_clm *c) {
  int64_t N = c->N;
  orc_garray *P = (orc_garray *)calloc((size_t)(N > 0 ? N * N : 1), sizeof(orc_garray)); /* Make2DGArraySlice */
  for (int64_t i = 0; i < N * N; i++) P[i].v[0] = -1;
  for (int64_t k = 0; k < c->noriented; k++) {
    const orc_oriented *o = &c->oriented[k];
    if (c->signs[o->ai] == o->ao && c->signs[o->bi] == o->bo) P[(int64_t)o->ai * N + o->bi] = o->g;
  }
  return P;
}

/* orientation.go:159-193 EvaluateQ, literal (dense Q rebuilt on every call). */
double orc_evaluate_q(const orc_clm *c) {
  orc_garray *Q = orc_build_Q(c);
  double score = 0.0;
  int64_t size = c->L, N = c->N;
  int64_t *cumsize = (int64_t *)malloc((size_t)(size ? size : 1) * 8);
  int64_t cumSum = 0;
  for (int64_t i = 0; i < size; i++) { cumsize[i] = cumSum; cumSum += c->tour_size[i]; }
  for (int64_t i = 0; i < size; i++) {
    int64_t a = c->tour_idx[i];
    for (int64_t j = i + 1; j < size; j++) {
      int64_t b = c->tour_idx[j];
      if (Q[a * N + b].v[0] == -1) continue; /* Entire GArray is empty */
      int64_t dist = cumsize[j - 1] - cumsize[i];
      if (dist > ORC_LIMIT) break;
      for (int k = 0; k < ORC_BB; k++)
        score -= (double)Q[a * N + b].v[k] * orc_go_log((double)(ORC_GR[k] + dist));
    }
  }
  free(cumsize); free(Q);
  return score;
}

/* Same loop, same order, same arithmetic as orc_evaluate_q, but the
 * (a,b,Signs[a],Signs[b]) entry is fetched straight from the oriented map
 * instead of materialising the 96*N^2-byte Q() first.  Bit-identical result
 * (tests/test_oracle.py checks it); usable where dense Q is infeasible.
 * A stored GArray whose slot 0 is literally -1 cannot occur (counts >= 0). */
double orc_evaluate_q_lookup(const orc_clm *c, int64_t *n_terms) {
  double score = 0.0;
  int64_t size = c->L, terms = 0;
  int64_t *cumsize = (int64_t *)malloc((size_t)(size ? size : 1) * 8);
  int64_t cumSum = 0;
  for (int64_t i = 0; i < size; i++) { cumsize[i] = cumSum; cumSum += c->tour_size[i]; }
  for (int64_t i = 0; i < size; i++) {
    int64_t a = c->tour_idx[i];
    for (int64_t j = i + 1; j < size; j++) {
      int64_t b = c->tour_idx[j];
      int64_t at = orc_map_get(&c->omap, orc_okey(a, b, c->signs[a], c->signs[b]));
      if (at < 0) continue;
      int64_t dist = cumsize[j - 1] - cumsize[i];
      if (dist > ORC_LIMIT) break;
      terms++;
      const int64_t *g = c->oriented[at].g.v;
      for (int k = 0; k < ORC_BB; k++) score -= (double)g[k] * orc_go_log((double)(ORC_GR[k] + dist));
    }
  }
  free(cumsize);
  if (n_terms) *n_terms = terms;
  return score;
}

/* orientation.go:67-84 flipWhole; returns 1 for ACCEPT, 0 for REJECT */
int orc_flip_whole(orc_clm *c, double *old_score, double *new_score) {
  uint8_t *oldSigns = (uint8_t *)malloc((size_t)(c->N ? c->N : 1));
  memcpy(oldSigns, c->signs, (size_t)c->N);
  double score = orc_evaluate_q_lookup(c, NULL);
  for (int64_t i = 0; i < c->N; i++) c->signs[i] = orc_rr(c->signs[i]);
  double newScore = orc_evaluate_q_lookup(c, NULL);
  int accept = 1;
  if (newScore <= score) { memcpy(c->signs, oldSigns, (size_t)c->N); accept = 0; }
  free(oldSigns);
  if (old_score) *old_score = score;
  if (new_score) *new_score = newScore;
  return accept;
}

/* orientation.go:87-120 flipOne; returns 1 if any flip was accepted.
 * trace (optional, L x 2 doubles): score before / newScore of every step;
 * accepted (optional, L bytes): per-step tag. */
int orc_flip_one(orc_clm *c, int64_t *n_accepts, int64_t *n_rejects, double *final_score, double *trace, uint8_t *accepted) {
  int64_t nAccepts = 0, nRejects = 0; int any = 0;
  double score = orc_evaluate_q_lookup(c, NULL);
  for (int64_t i = 0; i < c->L; i++) {
    int64_t idx = c->tour_idx[i];
    c->signs[idx] = orc_rr(c->signs[idx]);
    double newScore = orc_evaluate_q_lookup(c, NULL);
    int tag;
    if (newScore > score) { nAccepts++; tag = 1; }
    else { c->signs[idx] = orc_rr(c->signs[idx]); nRejects++; tag = 0; }
    if (trace) { trace[2 * i] = score; trace[2 * i + 1] = newScore; }
    if (accepted) accepted[i] = (uint8_t)tag;
    if (tag) { any = 1; score = newScore; }
  }
  if (n_accepts) *n_accepts = nAccepts;
  if (n_rejects) *n_rejects = nRejects;
  if (final_score) *final_score = score;
  return any;
}

/* Cyclic Jacobi eigen-decomposition of a dense symmetric matrix (restates the
 * role of mat64.EigenSym, orientation.go:33-44).  A: n x n (destroyed),
 * V: n x n eigenvectors in columns, w: eigenvalues (unsorted). */
static void orc_jacobi(double *A, double *V, double *w, int64_t n) {
  for (int64_t i = 0; i < n; i++) for (int64_t j = 0; j < n; j++) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int64_t i = 0; i < n; i++) for (int64_t j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-22) break;
    for (int64_t p = 0; p < n; p++)
      for (int64_t q = p + 1; q < n; q++) {
        double apq = A[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        double cs = 1 / sqrt(t * t + 1), sn = t * cs;
        for (int64_t k = 0; k < n; k++) {
          double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = cs * akp - sn * akq; A[k * n + q] = sn * akp + cs * akq;
        }
        for (int64_t k = 0; k < n; k++) {
          double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = cs * apk - sn * aqk; A[q * n + k] = sn * apk + cs * aqk;
        }
        for (int64_t k = 0; k < n; k++) {
          double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = cs * vkp - sn * vkq; V[k * n + q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int64_t i = 0; i < n; i++) w[i] = A[i * n + i];
}

/* Top eigenvector of O (largest eigenvalue), orientation.go:42-44 ColView(N-1).
 * O(N^3): small N only. */
void orc_top_eigvec(const orc_clm *c, double *v /* N */, double *eigval) {
  int64_t N = c->N;
  double *A = (double *)malloc((size_t)(N * N) * 8), *V = (double *)malloc((size_t)(N * N) * 8), *w = (double *)malloc((size_t)N * 8);
  orc_get_O(c, A);
  orc_jacobi(A, V, w, N);
  int64_t top = 0; for (int64_t i = 1; i < N; i++) if (w[i] > w[top]) top = i;
  for (int64_t i = 0; i < N; i++) v[i] = V[i * N + top];
  if (eigval) *eigval = w[top];
  free(A); free(V); free(w);
}

/* orientation.go:31-64 flipAll; returns 1 ACCEPT / 0 REJECT */
int orc_flip_all(orc_clm *c, double *old_score, double *new_score) {
  int64_t N = c->N;
  uint8_t *oldSigns = (uint8_t *)malloc((size_t)(N ? N : 1));
  memcpy(oldSigns, c->signs, (size_t)N);
  double score = orc_evaluate_q_lookup(c, NULL);
  double *v = (double *)malloc((size_t)(N ? N : 1) * 8);
  orc_top_eigvec(c, v, NULL);
  for (int64_t i = 0; i < N; i++) c->signs[i] = (v[i] < 0) ? '-' : '+';
  double newScore = orc_evaluate_q_lookup(c, NULL);
  int accept = 1;
  if (newScore < score) { memcpy(c->signs, oldSigns, (size_t)N); accept = 0; }
  free(oldSigns); free(v);
  if (old_score) *old_score = score;
  if (new_score) *new_score = newScore;
  return accept;
}

/* clm.go:392-418 Activate(shuffle, rng) with our RNG; flipAll included when
 * do_flip_all != 0. */
void orc_activate(orc_clm *c, int shuffle, uint64_t seed, int do_flip_all) {
  int64_t cnt = 0;
  for (int64_t i = 0; i < c->N; i++) if (c->active[i]) cnt++;
  int64_t *idx = (int64_t *)malloc((size_t)(cnt ? cnt : 1) * 8);
  int64_t k = 0;
  for (int64_t i = 0; i < c->N; i++) if (c->active[i]) idx[k++] = i;
  if (shuffle) orc_shuffle(idx, cnt, seed);
  orc_set_tour(c, cnt, idx);
  free(idx);
  memset(c->signs, '+', (size_t)c->N);
  if (do_flip_all) orc_flip_all(c, NULL, NULL);
}

/* optimize.go:176-184 printTour: ">LABEL\nname± name± ...\n" into buf */
int64_t orc_format_tour(const orc_clm *c, const char *label, char *buf, int64_t cap) {
  int64_t n = 0;
  n += snprintf(buf + n, (size_t)(cap - n), ">%s\n", label);
  for (int64_t i = 0; i < c->L && n < cap; i++) {
    int64_t idx = c->tour_idx[i];
    n += snprintf(buf + n, (size_t)(cap - n > 0 ? cap - n : 0), "%s%s%c", i ? " " : "", c->names[idx], c->signs[idx]);
  }
  if (n < cap) n += snprintf(buf + n, (size_t)(cap - n), "\n");
  return n;
}
