// ahx_clm.cc -- host-side CLM ingestion for libahx (no CUDA in this file).
//
// Product code: implements NewCLM/readRE/readClmLines/readClm of the reference
// (clm.go:121-240) plus GoldenArray/SumLog (base.go:130-167) with the same map
// semantics, then freezes the two Go maps (`contacts`, `orientedContacts`)
// into the flat pair table documented in include/ahx.h, which is what gets
// uploaded to the GPU.  Parsing is multi-threaded: lines are tokenised and
// histogrammed in parallel, then applied to the maps in file order so that
// "last line wins" / "smallest SumLog wins" behave exactly like the
// single-threaded Go loop.
#include "ahx_internal.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <chrono>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace ahx {

static thread_local std::string g_err;
void set_global_error(const std::string &s) { g_err = s; }
const char *global_error() { return g_err.c_str(); }

// Go's math.Log (FDLIBM e_log algorithm).  Used instead of libm so that SumLog
// ties (clm.go:230 `meanDist < p.meanDist`) resolve as they do in the Go
// binary on amd64, where the compiler never fuses multiply-adds.
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
double go_log(double x) {
  constexpr double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;
  constexpr double kL[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                            2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                            1.479819860511658591e-01};
  if (std::isnan(x) || (std::isinf(x) && x > 0)) return x;
  if (x < 0) return std::nan("");
  if (x == 0) return -INFINITY;
  int e;
  double m = std::frexp(x, &e);
  if (m < M_SQRT1_2) { m *= 2; --e; }
  const double f = m - 1, k = static_cast<double>(e);
  const double s = f / (2 + f), z = s * s, w = z * z;
  const double odd = z * (kL[0] + w * (kL[2] + w * (kL[4] + w * kL[6])));
  const double even = w * (kL[1] + w * (kL[3] + w * kL[5]));
  const double r = odd + even, hfsq = 0.5 * f * f;
  return k * kLn2Hi - ((hfsq - (s * (hfsq + r) + k * kLn2Lo)) - f);
}

// base.go:156-167 GoldenArray + base.go:138-144 SumLog in one pass.
void golden_and_sumlog(const int64_t *d, int64_t n, int32_t *bins, double *sumlog) {
  for (int k = 0; k < AHX_BB; ++k) bins[k] = 0;
  double sum = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double lg = go_log(static_cast<double>(d[i]));
    sum += lg;
    const double q = lg / 0.4812118250596684;                              // PHI, base.go:36
    const double r = q < 0 ? std::ceil(q - 0.5) : std::floor(q + 0.5);      // Round, base.go:130
    int c = 18;                                                            // LB; also Go's int(NaN/-Inf)
    if (r >= 29.0) c = 29; else if (r > 18.0) c = static_cast<int>(r);
    bins[c - 18]++;
  }
  *sumlog = sum;
}
#pragma GCC pop_options

static inline uint8_t rr(uint8_t b) { return b == '-' ? '+' : '-'; }  // clm.go:160
static inline uint64_t pair_key(int32_t a, int32_t b) { return (static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b); }
static inline uint64_t okey(int32_t a, int32_t b, uint8_t ao, uint8_t bo) {
  return (static_cast<uint64_t>(static_cast<uint32_t>(a)) << 40) | (static_cast<uint64_t>(static_cast<uint32_t>(b)) << 16) | (static_cast<uint64_t>(ao) << 8) | bo;
}

}  // namespace ahx

using namespace ahx;

// uint64 -> int64 map for the two Go maps' key -> insertion index: open addressing, linear probing,
// no erase.  (std::unordered_map's node allocations made the serial apply phase 2 us per CLM line.)
struct FlatMap {
  std::vector<uint64_t> k;   // key + 1; 0 = empty
  std::vector<int64_t> v;
  size_t n = 0, mask = 0;
  static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
  void rehash(size_t cap) {
    std::vector<uint64_t> ok; std::vector<int64_t> ov;
    ok.swap(k); ov.swap(v);
    k.assign(cap, 0); v.assign(cap, 0); mask = cap - 1; n = 0;
    for (size_t i = 0; i < ok.size(); ++i) if (ok[i]) { bool ins; *insert(ok[i] - 1, ov[i], &ins) = ov[i]; }
  }
  void reserve(size_t want) { size_t cap = 16; while (cap < want + want / 2) cap <<= 1; if (cap > k.size()) rehash(cap); }
  const int64_t *find(uint64_t key) const {
    if (k.empty()) return nullptr;
    for (size_t i = mix(key) & mask;; i = (i + 1) & mask) {
      if (k[i] == key + 1) return &v[i];
      if (!k[i]) return nullptr;
    }
  }
  // slot of `key`; a new key is entered with `val` (*inserted = true)
  int64_t *insert(uint64_t key, int64_t val, bool *inserted) {
    if ((n + 1) * 3 > k.size() * 2) rehash(k.empty() ? 16 : k.size() * 2);   // load factor <= 2/3
    for (size_t i = mix(key) & mask;; i = (i + 1) & mask) {
      if (k[i] == key + 1) { *inserted = false; return &v[i]; }
      if (!k[i]) { k[i] = key + 1; v[i] = val; ++n; *inserted = true; return &v[i]; }
    }
  }
};

// std::allocator whose default construction leaves the element uninitialised (vector::resize without the zero fill)
template <typename T>
struct NoInit : std::allocator<T> {
  template <typename U> struct rebind { using other = NoInit<U>; };
  using std::allocator<T>::allocator;
  template <typename U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
  template <typename U, typename... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};

struct ahx_clm {
  int32_t n = 0;
  std::vector<std::string> names;
  std::vector<int64_t> sizes;
  std::unordered_map<std::string, int32_t> tig2idx;
  struct Contact { int32_t ai, bi; int32_t strand; int64_t nlinks; double mean_dist; int64_t order; };   // order: stamp of the line that created the key
  struct Oriented { int32_t ai, bi; uint8_t ao, bo; std::array<int32_t, AHX_BB> g; };
  // contacts and orientedContacts, both sharded by key hash so that ahx_clm_parse can fill the shards from several
  // threads; within a shard entries keep the file's order ("smallest SumLog wins, the first of equals" and
  // "later line wins" are per-key rules); a contact remembers the stamp of the line that created its key, which
  // restores the one cross-key rule there is (finalize: of the keys (a,b) and (b,a) the later-created wins)
  struct CShard { FlatMap map; std::vector<Contact> v; };
  struct OShard { FlatMap map; std::vector<Oriented> v; };
  std::vector<CShard> csh = std::vector<CShard>(1);
  std::vector<OShard> osh = std::vector<OShard>(1);
  int64_t next_order = 0;                                // builder API: lines arrive one by one
  // every key of a CLM line -- contacts[(a,b)], orientedContacts[(a,b,ao,bo)] and its mirror -- lives in the shard of the
  // unordered contig pair, so the parser can bucket whole lines by shard while it tokenises
  static size_t pair_shard(int32_t a, int32_t b, size_t S) {
    return S <= 1 ? 0 : static_cast<size_t>((FlatMap::mix(pair_key(std::min(a, b), std::max(a, b))) >> 40) % S);
  }
  bool finalized = false;
  std::vector<int32_t> pa, pb, nl, slots;
  std::vector<int32_t, NoInit<int32_t>> hist;   // 48 bytes per oriented contact: sized without a zero fill, first touched by the threads that pack it
  std::vector<int8_t> strand;
  std::vector<double> diag;   // O[a][a] from self-pair lines (empty when there are none: `extract` never writes them)

  // sh: pair_shard(ai, bi, number of shards) (one thread per shard; csh and osh have the same size)
  void apply_contact(int32_t ai, int32_t bi, uint8_t ao, uint8_t bo, double mean_dist, int64_t nlinks, int64_t order, size_t sh) {
    const uint64_t key = pair_key(ai, bi);
    CShard &cs = csh[sh];
    const int32_t st = (ao != bo) ? -1 : 1;                                 // clm.go:224-227
    bool fresh;
    const int64_t at = *cs.map.insert(key, static_cast<int64_t>(cs.v.size()), &fresh);
    if (fresh) {
      cs.v.push_back({ai, bi, st, nlinks, mean_dist, order});
    } else {
      Contact &c = cs.v[at];
      if (mean_dist < c.mean_dist) { c.strand = st; c.nlinks = nlinks; c.mean_dist = mean_dist; }   // clm.go:229-236
    }
  }
  void apply_line(int32_t ai, int32_t bi, uint8_t ao, uint8_t bo, const std::array<int32_t, AHX_BB> &g,
                  double mean_dist, int64_t nlinks) {
    finalized = false;
    const size_t sh = pair_shard(ai, bi, csh.size());
    apply_contact(ai, bi, ao, bo, mean_dist, nlinks, next_order++, sh);
    put_oriented(ai, bi, ao, bo, g, sh);
    put_oriented(bi, ai, rr(bo), rr(ao), g, sh);                            // clm.go:238
  }
  void put_oriented(int32_t ai, int32_t bi, uint8_t ao, uint8_t bo, const std::array<int32_t, AHX_BB> &g, size_t sh) {
    const uint64_t key = okey(ai, bi, ao, bo);
    OShard &o = osh[sh];
    bool fresh;
    const int64_t at = *o.map.insert(key, static_cast<int64_t>(o.v.size()), &fresh);
    if (fresh) o.v.push_back({ai, bi, ao, bo, g});
    else o.v[at].g = g;
  }
};

namespace {

struct Mapped {
  const char *p = nullptr; size_t n = 0; int fd = -1;
  bool open(const char *path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) return false;
    n = static_cast<size_t>(st.st_size);
    if (n == 0) { p = ""; return true; }
    void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) return false;
    p = static_cast<const char *>(m);
    return true;
  }
  ~Mapped() { if (p && n) munmap(const_cast<char *>(p), n); if (fd >= 0) close(fd); }
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// strconv.Atoi with the error dropped (clm.go:152,186,190): invalid -> 0.
inline int64_t atoi_or_zero(const char *s, size_t n) {
  if (n == 0) return 0;
  size_t i = 0; bool neg = false;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; if (n == 1) return 0; }
  int64_t v = 0;
  for (; i < n; ++i) { const unsigned d = static_cast<unsigned>(s[i] - '0'); if (d > 9) return 0; v = v * 10 + d; }
  return neg ? -v : v;
}

struct ParsedLine { int32_t ai, bi; uint8_t ao, bo; std::array<int32_t, AHX_BB> g; double sumlog; int64_t nlinks; };

// Parses lines of [beg,end) (beg at a line start).  Returns false on a line the
// reference would panic on.  resolve(name_a, name_b, hits) lists the (group, ai, bi) that keep
// the line -- both contigs in that group's REfile, clm.go:211-218; the line's distances are
// tokenised and histogrammed once, whatever the number of groups.
struct Hit { int32_t group, ai, bi; };
template <typename Resolve>
bool parse_chunk(const Resolve &resolve, const char *buf, size_t beg, size_t end, size_t total, size_t S,
                 std::vector<std::vector<ParsedLine>> *out /* [group * S + shard] */, std::string *err) {
  std::vector<Hit> hits;
  std::vector<int64_t> dists;
  size_t p = beg;
  while (p < end) {
    size_t e = p;
    while (e < total && buf[e] != '\n') ++e;
    const size_t next = e + 1;
    size_t a = p, b = e;
    while (a < b && is_space(buf[a])) ++a;                                  // strings.TrimSpace
    while (b > a && is_space(buf[b - 1])) --b;
    if (a == b) {
      if (next >= total) break;                                            // trailing blank at EOF
      *err = "empty line in CLM file (the reference panics on it)"; return false;
    }
    size_t t1 = a; while (t1 < b && buf[t1] != '\t') ++t1;
    size_t s1 = a; while (s1 < t1 && buf[s1] != ' ') ++s1;
    if (s1 == a || s1 + 1 >= t1 || t1 >= b) { *err = "malformed CLM line: expected `a± b±\\tN\\td1 d2 ...`"; return false; }
    size_t s2 = s1 + 1; while (s2 < t1 && buf[s2] != ' ') ++s2;
    size_t t2 = t1 + 1; while (t2 < b && buf[t2] != '\t') ++t2;
    if (t2 >= b) { *err = "malformed CLM line: no distance column"; return false; }
    ParsedLine pl;
    pl.ao = static_cast<uint8_t>(buf[s1 - 1]);
    pl.bo = static_cast<uint8_t>(buf[s2 - 1]);
    hits.clear();
    resolve(std::string(buf + a, s1 - 1 - a), std::string(buf + s1 + 1, s2 - 1 - (s1 + 1)), &hits);
    if (!hits.empty()) {                                                   // clm.go:211-218
      size_t q = t2 + 1, wend = q;
      while (wend < b && buf[wend] != '\t') ++wend;
      dists.clear();
      for (;;) {                                                           // Split(words[2], " ")
        const size_t w = q;
        while (q < wend && buf[q] != ' ') ++q;
        dists.push_back(atoi_or_zero(buf + w, q - w));
        if (q >= wend) break;
        ++q;
      }
      pl.nlinks = static_cast<int64_t>(dists.size());                      // len(line.links), clm.go:228
      golden_and_sumlog(dists.data(), pl.nlinks, pl.g.data(), &pl.sumlog);
      for (const Hit &h : hits) { pl.ai = h.ai; pl.bi = h.bi; (*out)[h.group * S + ahx_clm::pair_shard(h.ai, h.bi, S)].push_back(pl); }
    }
    p = next;
  }
  return true;
}

int read_re(ahx_clm *clm, const char *refile) {                           // clm.go:141-157
  Mapped m;
  if (!m.open(refile)) { set_global_error(std::string("cannot open REfile `") + refile + "`"); return AHX_ERR_IO; }
  size_t p = 0;
  while (p < m.n) {
    size_t e = p;
    while (e < m.n && m.p[e] != '\n') ++e;
    size_t q = p; std::string_view first, last;
    while (q < e) {                                                        // strings.Fields
      while (q < e && is_space(m.p[q])) ++q;
      if (q >= e) break;
      const size_t w = q;
      while (q < e && !is_space(m.p[q])) ++q;
      if (first.empty()) first = std::string_view(m.p + w, q - w);
      last = std::string_view(m.p + w, q - w);
    }
    if (!first.empty() && first[0] != '#') {
      clm->names.emplace_back(first);
      clm->sizes.push_back(atoi_or_zero(last.data(), last.size()));
      clm->tig2idx[clm->names.back()] = static_cast<int32_t>(clm->names.size() - 1);
    }
    p = e + 1;
  }
  clm->n = static_cast<int32_t>(clm->names.size());
  return AHX_OK;
}

}  // namespace

extern "C" {

// One pass over the CLM text for any number of groups: threads tokenise chunks of the file (names
// resolved against every group's REfile, GoldenArray / SumLog once per kept line), then every group's
// two maps are filled in file order -- "smallest SumLog wins" / "last line wins" exactly like the
// single-threaded Go loop (clm.go:221-238) run on that group alone.
static int finalize_impl(ahx_clm *clm, int nthreads);
static int parse_groups(const char *clmfile, const char *const *refiles, int32_t n_groups, int nthreads, ahx_clm **out) {
  std::vector<ahx_clm *> clms(n_groups, nullptr);
  auto drop = [&] { for (auto *c : clms) delete c; };
  for (int32_t g = 0; g < n_groups; ++g) {
    clms[g] = new ahx_clm();
    const int rc = read_re(clms[g], refiles[g]);
    if (rc != AHX_OK) { drop(); return rc; }
  }
  Mapped m;
  if (!m.open(clmfile)) { set_global_error(std::string("cannot open clmfile `") + clmfile + "`"); drop(); return AHX_ERR_IO; }
  const auto t_start = std::chrono::steady_clock::now();
  if (nthreads <= 0) nthreads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
  const int napply = nthreads;
  nthreads = static_cast<int>(std::min<size_t>(static_cast<size_t>(nthreads), m.n / (1 << 16) + 1));
  std::vector<size_t> cut(nthreads + 1);
  cut[0] = 0; cut[nthreads] = m.n;
  for (int t = 1; t < nthreads; ++t) {
    size_t c = m.n / nthreads * t;
    while (c < m.n && m.p[c] != '\n') ++c;
    cut[t] = std::min(m.n, c + 1);
  }
  // contig name -> every (group, index) that lists it (a partition lists a contig once; overlapping REfiles still work)
  std::unordered_map<std::string, std::vector<std::pair<int32_t, int32_t>>> where;
  if (n_groups > 1)
    for (int32_t g = 0; g < n_groups; ++g)
      for (auto &kv : clms[g]->tig2idx) where[kv.first].emplace_back(g, kv.second);
  auto resolve1 = [&](const std::string &a, const std::string &b, std::vector<Hit> *hits) {
    const auto &idx = clms[0]->tig2idx;
    auto ia = idx.find(a);
    if (ia == idx.end()) return;
    auto ib = idx.find(b);
    if (ib != idx.end()) hits->push_back({0, ia->second, ib->second});
  };
  auto resolveN = [&](const std::string &a, const std::string &b, std::vector<Hit> *hits) {
    auto ia = where.find(a);
    if (ia == where.end()) return;
    auto ib = where.find(b);
    if (ib == where.end()) return;
    for (auto &x : ia->second)
      for (auto &y : ib->second) if (x.first == y.first) hits->push_back({x.first, x.second, y.second});
  };
  // shards per group: one group is applied by up to 8 threads, several groups by one thread each
  const size_t S = n_groups == 1 ? static_cast<size_t>(std::max(1, std::min(napply, 8))) : 1;
  std::vector<std::vector<std::vector<ParsedLine>>> parts(nthreads, std::vector<std::vector<ParsedLine>>(n_groups * S));   // [chunk][group * S + shard]
  std::vector<std::string> errs(nthreads);
  std::vector<char> ok(nthreads, 1);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] {
        ok[t] = n_groups == 1 ? parse_chunk(resolve1, m.p, cut[t], cut[t + 1], m.n, S, &parts[t], &errs[t])
                              : parse_chunk(resolveN, m.p, cut[t], cut[t + 1], m.n, S, &parts[t], &errs[t]);
      });
    for (auto &x : th) x.join();
  }
  const auto t_parse = std::chrono::steady_clock::now();
  for (int t = 0; t < nthreads; ++t)
    if (!ok[t]) { set_global_error(errs[t]); drop(); return AHX_ERR_IO; }
  // both maps of a group are sharded by key hash: S threads each walk the group's lines in file order and keep
  // the keys of their shard (`contacts`: smallest SumLog wins, `orientedContacts`: last line wins -- per-key rules);
  // several groups: the groups are spread over the threads instead, one shard each
  std::vector<int> rcs(n_groups, AHX_OK);
  std::vector<std::string> ferr(n_groups);
  auto apply_group = [&](int32_t g) {
    ahx_clm *clm = clms[g];
    size_t nlines = 0;
    for (auto &part : parts) for (size_t sh = 0; sh < S; ++sh) nlines += part[g * S + sh].size();
    clm->osh.assign(S, ahx_clm::OShard{});
    clm->csh.assign(S, ahx_clm::CShard{});
    std::vector<std::thread> ap;
    // thread sh applies the lines the tokenisers bucketed into shard sh, in file order: contacts and oriented contacts
    auto shard = [&, clm, g](size_t sh) {
      size_t mine = 0;
      for (auto &part : parts) mine += part[g * S + sh].size();
      ahx_clm::OShard &o = clm->osh[sh];
      ahx_clm::CShard &c = clm->csh[sh];
      o.map.reserve(2 * mine + mine / 4 + 16); o.v.reserve(2 * mine + mine / 4 + 16);
      c.map.reserve(mine / 2 + mine / 8 + 16); c.v.reserve(mine / 2 + mine / 8 + 16);
      int64_t order = 0;   // stamps are only compared between the keys (a,b) and (b,a), which share a shard
      for (auto &part : parts)
        for (auto &pl : part[g * S + sh]) {
          clm->apply_contact(pl.ai, pl.bi, pl.ao, pl.bo, pl.sumlog, pl.nlinks, order++, sh);
          clm->put_oriented(pl.ai, pl.bi, pl.ao, pl.bo, pl.g, sh);
          clm->put_oriented(pl.bi, pl.ai, rr(pl.bo), rr(pl.ao), pl.g, sh);
        }
    };
    if (S == 1) shard(0);
    else {
      for (size_t sh = 0; sh < S; ++sh) ap.emplace_back(shard, sh);
      for (auto &x : ap) x.join();
    }
    clm->next_order = static_cast<int64_t>(nlines);
  };
  std::chrono::steady_clock::time_point t_apply;
  if (n_groups == 1) {
    apply_group(0);
    t_apply = std::chrono::steady_clock::now();
    rcs[0] = finalize_impl(clms[0], napply);
    if (rcs[0] != AHX_OK) ferr[0] = global_error();
  } else {
    std::vector<std::thread> gt;
    std::atomic<int32_t> next{0};
    for (int t = 0; t < std::min<int>(napply, n_groups); ++t)
      gt.emplace_back([&] {
        for (int32_t g = next++; g < n_groups; g = next++) {
          apply_group(g);
          for (auto &part : parts) std::vector<ParsedLine>().swap(part[g]);
          rcs[g] = ahx_clm_finalize(clms[g]);
          if (rcs[g] != AHX_OK) ferr[g] = global_error();
        }
      });
    for (auto &x : gt) x.join();
    t_apply = std::chrono::steady_clock::now();
  }
  if (getenv("AHX_CLM_TIMING")) {
    const auto t_fin = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "ahx_clm_parse: %.1f MB, %d group(s), %d threads: tokenise+log %.1f ms, apply %.1f ms, finalize %.1f ms\n", m.n / 1e6, n_groups, nthreads,
            ms(t_start, t_parse), ms(t_parse, t_apply), ms(t_apply, t_fin));
  }
  for (int32_t g = 0; g < n_groups; ++g)
    if (rcs[g] != AHX_OK) { set_global_error(ferr[g]); drop(); return rcs[g]; }
  for (int32_t g = 0; g < n_groups; ++g) out[g] = clms[g];
  return AHX_OK;
}

int ahx_clm_parse(const char *clmfile, const char *refile, int nthreads, ahx_clm **out) {
  if (!clmfile || !refile || !out) { set_global_error("ahx_clm_parse: null argument"); return AHX_ERR_ARG; }
  return parse_groups(clmfile, &refile, 1, nthreads, out);
}

int ahx_clm_parse_groups(const char *clmfile, const char *const *refiles, int32_t n_groups, int nthreads, ahx_clm **out) {
  if (!clmfile || !refiles || !out || n_groups < 1) { set_global_error("ahx_clm_parse_groups: bad argument"); return AHX_ERR_ARG; }
  for (int32_t g = 0; g < n_groups; ++g) if (!refiles[g]) { set_global_error("ahx_clm_parse_groups: null REfile"); return AHX_ERR_ARG; }
  return parse_groups(clmfile, refiles, n_groups, nthreads, out);
}

int ahx_clm_new(int32_t n_tigs, const int64_t *sizes, ahx_clm **out) {
  if (n_tigs < 0 || (n_tigs > 0 && !sizes) || !out) { set_global_error("ahx_clm_new: bad argument"); return AHX_ERR_ARG; }
  auto *clm = new ahx_clm();
  clm->n = n_tigs;
  clm->sizes.assign(sizes, sizes + n_tigs);
  clm->names.assign(n_tigs, std::string());
  *out = clm;
  return AHX_OK;
}

int ahx_clm_add_line(ahx_clm *clm, int32_t ai, int32_t bi, uint8_t ao, uint8_t bo, const int64_t *dists, int64_t n) {
  if (!clm || ai < 0 || bi < 0 || ai >= clm->n || bi >= clm->n || n < 0 || (n > 0 && !dists)) {
    set_global_error("ahx_clm_add_line: bad argument"); return AHX_ERR_ARG;
  }
  std::array<int32_t, AHX_BB> g; double sl;
  golden_and_sumlog(dists, n, g.data(), &sl);
  clm->apply_line(ai, bi, ao, bo, g, sl, n);
  return AHX_OK;
}

// nthreads > 1: the oriented entries are looked up and the histograms packed by several threads (the result is the
// same table: rows and slots are positions, not insertion orders)
static int finalize_impl(ahx_clm *clm, int nthreads) {
  if (!clm) { set_global_error("ahx_clm_finalize: null"); return AHX_ERR_ARG; }
  if (clm->finalized) return AHX_OK;
  const auto tf0 = std::chrono::steady_clock::now();
  auto tf = [&](const char *what) { if (getenv("AHX_CLM_TIMING2")) fprintf(stderr, "  finalize %s at %.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tf0).count()); };
  // rows = unordered pairs {a<b}, sorted by (a,b) so that consecutive rows share `a`
  std::vector<uint64_t> keys;
  { size_t nc = 0; for (auto &cs : clm->csh) nc += cs.v.size(); keys.reserve(nc); }
  for (auto &cs : clm->csh)
    for (auto &c : cs.v) if (c.ai != c.bi) keys.push_back(pair_key(std::min(c.ai, c.bi), std::max(c.ai, c.bi)));
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  const size_t E = keys.size();
  FlatMap rows;
  rows.reserve(E);
  for (size_t e = 0; e < E; ++e) { bool fresh; rows.insert(keys[e], static_cast<int64_t>(e), &fresh); }
  auto row_of = [&](int32_t a, int32_t b) -> int64_t {
    const int64_t *r = rows.find(pair_key(std::min(a, b), std::max(a, b)));
    return r ? *r : -1;
  };
  clm->pa.resize(E); clm->pb.resize(E); clm->nl.assign(E, 0); clm->strand.assign(E, 0);
  clm->slots.assign(E * 8, -1); clm->hist.clear(); { size_t no = 0; for (auto &sh : clm->osh) no += sh.v.size(); clm->hist.reserve(no * AHX_BB); }
  for (size_t e = 0; e < E; ++e) { clm->pa[e] = static_cast<int32_t>(keys[e] >> 32); clm->pb[e] = static_cast<int32_t>(keys[e] & 0xffffffffu); }
  // M(): P[ai][bi] = P[bi][ai] = nlinks per map entry (clm.go:440-445).  Go's map
  // order is random; when both (a,b) and (b,a) keys exist (never written by
  // `extract`, extract.go:477-481) the later-inserted key wins here.
  clm->diag.clear();
  std::vector<int64_t> won(E, -1);                       // stamp of the key that currently owns the row
  for (auto &cs : clm->csh)
    for (auto &c : cs.v) {
      if (c.ai == c.bi) {   // a self-pair line: never read by M() / Q() (i < j are distinct contigs), but O() puts it on the diagonal (SetSym(a, a, ...))
        if (clm->diag.empty()) clm->diag.assign(clm->n, 0.0);
        clm->diag[c.ai] = static_cast<double>(c.strand) * static_cast<double>(c.nlinks);
        continue;
      }
      const int64_t e = row_of(c.ai, c.bi);
      if (c.nlinks > INT32_MAX) { set_global_error("nlinks exceeds int32"); return AHX_ERR_ARG; }
      if (c.order < won[e]) continue;                    // the key created later wins the row
      won[e] = c.order;
      clm->nl[e] = static_cast<int32_t>(c.nlinks);
      clm->strand[e] = static_cast<int8_t>(c.strand);
    }
  tf("rows+contacts");
  // histograms in (row, slot) order, whatever the insertion order (and the sharding) was
  std::vector<const ahx_clm::Oriented *> ref(E * 8, nullptr);
  auto fill_refs = [&](size_t sh0, size_t step) {       // a key lives in exactly one shard: no two threads write one slot
    for (size_t si = sh0; si < clm->osh.size(); si += step)
      for (auto &o : clm->osh[si].v) {
        if (o.ai == o.bi) continue;                       // Q[a][a] is never read (i<j are distinct contigs)
        const bool sa = o.ao == '-', sb = o.bo == '-';
        if ((!sa && o.ao != '+') || (!sb && o.bo != '+')) continue;  // can never equal a Signs byte
        const int64_t e = row_of(o.ai, o.bi);
        if (e < 0) continue;
        const int dir = o.ai < o.bi ? 0 : 1;
        ref[e * 8 + dir * 4 + 2 * (sa ? 1 : 0) + (sb ? 1 : 0)] = &o;
      }
  };
  const size_t nt = static_cast<size_t>(std::max(1, std::min<int>(nthreads, 16)));
  if (nt == 1 || ref.size() < (1u << 16)) {
    fill_refs(0, 1);
    for (size_t i = 0; i < ref.size(); ++i) {
      if (!ref[i]) continue;
      clm->slots[i] = static_cast<int32_t>(clm->hist.size() / AHX_BB);
      clm->hist.insert(clm->hist.end(), ref[i]->g.begin(), ref[i]->g.end());
    }
  } else {
    {
      std::vector<std::thread> th;
      const size_t step = std::min(nt, std::max<size_t>(1, clm->osh.size()));
      for (size_t t = 0; t < step; ++t) th.emplace_back(fill_refs, t, step);
      for (auto &x : th) x.join();
    }
    tf("refs");
    // pack the histograms in slot order: count per chunk, prefix, copy
    std::vector<size_t> cnt(nt + 1, 0);
    const size_t chunk = (ref.size() + nt - 1) / nt;
    auto count = [&](size_t t) { size_t c = 0; for (size_t i = t * chunk; i < std::min(ref.size(), (t + 1) * chunk); ++i) c += ref[i] != nullptr; cnt[t + 1] = c; };
    { std::vector<std::thread> th; for (size_t t = 0; t < nt; ++t) th.emplace_back(count, t); for (auto &x : th) x.join(); }
    for (size_t t = 0; t < nt; ++t) cnt[t + 1] += cnt[t];
    clm->hist.resize(cnt[nt] * AHX_BB);
    auto pack = [&](size_t t) {
      size_t h = cnt[t];
      for (size_t i = t * chunk; i < std::min(ref.size(), (t + 1) * chunk); ++i) {
        if (!ref[i]) continue;
        clm->slots[i] = static_cast<int32_t>(h);
        std::copy(ref[i]->g.begin(), ref[i]->g.end(), clm->hist.begin() + static_cast<std::ptrdiff_t>(h * AHX_BB));
        ++h;
      }
    };
    { std::vector<std::thread> th; for (size_t t = 0; t < nt; ++t) th.emplace_back(pack, t); for (auto &x : th) x.join(); }
  }
  tf("packed");
  clm->finalized = true;
  return AHX_OK;
}

int ahx_clm_finalize(ahx_clm *clm) { return finalize_impl(clm, 1); }

void ahx_clm_free(ahx_clm *clm) { delete clm; }
int32_t ahx_clm_ntigs(const ahx_clm *clm) { return clm ? clm->n : 0; }
const char *ahx_clm_tig_name(const ahx_clm *clm, int32_t i) { return (clm && i >= 0 && i < clm->n) ? clm->names[i].c_str() : ""; }
int64_t ahx_clm_tig_size(const ahx_clm *clm, int32_t i) { return (clm && i >= 0 && i < clm->n) ? clm->sizes[i] : 0; }
int32_t ahx_clm_tig_index(const ahx_clm *clm, const char *name) {
  if (!clm || !name) return -1;
  auto it = clm->tig2idx.find(name);
  return it == clm->tig2idx.end() ? -1 : it->second;
}
int64_t ahx_clm_npairs(const ahx_clm *clm) { return (clm && clm->finalized) ? static_cast<int64_t>(clm->pa.size()) : -1; }
int64_t ahx_clm_nhists(const ahx_clm *clm) { return (clm && clm->finalized) ? static_cast<int64_t>(clm->hist.size() / AHX_BB) : -1; }
int ahx_clm_get_pairs(const ahx_clm *clm, int32_t *pa, int32_t *pb, int32_t *nl, int8_t *strand, int32_t *slots) {
  if (!clm || !clm->finalized) { set_global_error("ahx_clm_get_pairs: not finalized"); return AHX_ERR_STATE; }
  const size_t E = clm->pa.size();
  if (pa) std::memcpy(pa, clm->pa.data(), E * 4);
  if (pb) std::memcpy(pb, clm->pb.data(), E * 4);
  if (nl) std::memcpy(nl, clm->nl.data(), E * 4);
  if (strand) std::memcpy(strand, clm->strand.data(), E);
  if (slots) std::memcpy(slots, clm->slots.data(), E * 8 * 4);
  return AHX_OK;
}
int ahx_clm_get_hists(const ahx_clm *clm, int32_t *hist) {
  if (!clm || !clm->finalized) { set_global_error("ahx_clm_get_hists: not finalized"); return AHX_ERR_STATE; }
  if (hist) std::memcpy(hist, clm->hist.data(), clm->hist.size() * 4);
  return AHX_OK;
}

}  // extern "C"

// Internal accessors for ahx_load_clm (ahx_core.cu).
namespace ahx {
ClmView view_of(const ahx_clm *clm) {
  ClmView v{};
  v.n = clm->n; v.sizes = clm->sizes.data(); v.E = static_cast<int64_t>(clm->pa.size());
  v.pa = clm->pa.data(); v.pb = clm->pb.data(); v.nl = clm->nl.data(); v.strand = clm->strand.data();
  v.slots = clm->slots.data(); v.H = static_cast<int64_t>(clm->hist.size() / AHX_BB); v.hist = clm->hist.data();
  v.finalized = clm->finalized;
  v.diag = clm->diag.empty() ? nullptr : clm->diag.data();
  return v;
}
}  // namespace ahx
