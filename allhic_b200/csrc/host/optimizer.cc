// optimizer.cc -- see optimizer.h.  Orchestration only: every score, mutation,
// selection and flip runs inside libahx on the GPU.
#include "optimizer.h"

#include <sys/stat.h>

#include <cstdarg>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>

namespace allhic {

static thread_local std::string g_tag;
void logSetTag(const std::string &tag) { g_tag = tag; }

static std::string vlog(const char *level, const char *fmt, va_list ap) {
  // same shape as the reference's go-logging format (base.go:106-108): time, level, message on stderr;
  // one write per line so that lines of concurrent groups do not interleave
  char ts[16], msg[4096];
  const time_t now = time(nullptr);
  struct tm tmv;
  localtime_r(&now, &tmv);
  strftime(ts, sizeof ts, "%H:%M:%S", &tmv);
  vsnprintf(msg, sizeof msg, fmt, ap);
  std::string line = std::string(ts) + " | " + level + " " + g_tag + msg + "\n";
  fwrite(line.data(), 1, line.size(), stderr);
  return msg;
}
void logNotice(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vlog("NOTICE", fmt, ap); va_end(ap); }
void logFatal(const char *fmt, ...) {   // log.Fatal: message, then the run ends (base.go:349-362)
  va_list ap; va_start(ap, fmt); std::string m = vlog("CRITIC", fmt, ap); va_end(ap);
  throw FatalError{m};
}

std::string RemoveExt(const std::string &filename) {   // strings.TrimSuffix(filename, path.Ext(filename))
  const size_t slash = filename.find_last_of('/');
  const size_t dot = filename.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return filename;
  return filename.substr(0, dot);
}

void CLM::check(int rc, const char *what) {
  if (rc != AHX_OK) logFatal("%s: %s", what, ahx_last_error(ctx_));
}

CLM::CLM(const std::string &clmfile, const std::string &refile, int device, const std::vector<ahx_ctx *> &shared, ahx_clm *parsed, FILE *out)
    : REfile(refile), Clmfile(clmfile), clm_(parsed), out_(out) {
  logNotice("Parse REfile `%s`", refile.c_str());
  logNotice("Parse clmfile `%s`", clmfile.c_str());
  if (!clm_ && ahx_clm_parse(clmfile.c_str(), refile.c_str(), 0, &clm_) != AHX_OK) logFatal("%s", ahx_last_error(nullptr));   // mustOpen -> log.Fatal
  const int n = ahx_clm_ntigs(clm_);
  for (int i = 0; i < n; ++i) Tigs.push_back({i, ahx_clm_tig_name(clm_, i), ahx_clm_tig_size(clm_, i), true});
  if (!shared.empty()) { islands_ = shared; own_ctx_ = false; }
  else {
    ahx_ctx *c = nullptr;
    if (ahx_create(device, &c) != AHX_OK) logFatal("%s", ahx_last_error(nullptr));
    islands_.push_back(c);
  }
  ctx_ = islands_[0];
  for (ahx_ctx *c : islands_)
    if (ahx_load_clm(c, clm_) != AHX_OK) logFatal("ahx_load_clm: %s", ahx_last_error(c));
}

CLM::~CLM() {
  if (own_ctx_) for (ahx_ctx *c : islands_) ahx_destroy(c);
  if (clm_) ahx_clm_free(clm_);
}

void CLM::reportActive(bool verbose, int *activeCounts, int64_t *sumLength) {
  int cnt = 0; int64_t sum = 0;
  for (auto &t : Tigs) if (t.IsActive) { ++cnt; sum += t.Size; }
  if (verbose) logNotice("Active tigs: %d (length=%lld)", cnt, static_cast<long long>(sum));
  if (activeCounts) *activeCounts = cnt;
  if (sumLength) *sumLength = sum;
}

void CLM::Activate(bool shuffle, std::mt19937_64 &rng) {
  reportActive(true, nullptr, nullptr);
  Tour.Tigs.clear();
  for (auto &t : Tigs) if (t.IsActive) Tour.Tigs.push_back(t.Idx);
  if (shuffle) {   // Tour.Shuffle, evaluate.go:248-255 (Fisher-Yates, j = i + Intn(N-i))
    const int N = Tour.Len();
    for (int i = 0; i < N; ++i) {
      const int j = i + static_cast<int>(rng() % static_cast<uint64_t>(N - i));
      std::swap(Tour.Tigs[i], Tour.Tigs[j]);
    }
  }
  Signs.assign(Tigs.size(), '+');
  flipAll();   // Initialize with the signs of the tigs
}

double CLM::EvaluateQ() {
  double out = 0.0;
  check(ahx_eval_q(ctx_, 1, Tour.Len(), Tour.Tigs.data(), 1, Signs.data(), &out), "ahx_eval_q");
  return out;
}

double CLM::Evaluate() {
  double out = 0.0;
  check(ahx_eval_m(ctx_, 1, Tour.Len(), Tour.Tigs.data(), &out), "ahx_eval_m");
  return out;
}

static void flipLog(const char *method, double score, double scoreFlipped, const char *tag) {   // orientation.go:25-28
  logNotice("%s: %.5f => %.5f %s", method, score, scoreFlipped, tag);
}

const char *CLM::flipAll() {
  double a = 0, b = 0; int32_t acc = 0;
  check(ahx_flip_all(ctx_, Tour.Len(), Tour.Tigs.data(), Signs.data(), &a, &b, &acc), "ahx_flip_all");
  const char *tag = acc ? ACCEPT : REJECT;
  flipLog("FLIPALL", a, b, tag);
  double resid = 0; int32_t conv = 1;
  ahx_flip_all_info(ctx_, &resid, &conv, nullptr, nullptr);
  if (!conv) logNotice("FLIPALL: the eigensolver stopped at residual %.3g (near-degenerate top eigenvalues); flipWhole / flipOne refine the signs", resid);
  return tag;
}

const char *CLM::flipWhole() {
  double a = 0, b = 0; int32_t acc = 0;
  check(ahx_flip_whole(ctx_, Tour.Len(), Tour.Tigs.data(), Signs.data(), &a, &b, &acc), "ahx_flip_whole");
  const char *tag = acc ? ACCEPT : REJECT;
  flipLog("FLIPWHOLE", a, b, tag);
  return tag;
}

const char *CLM::flipOne() {
  const int L = Tour.Len();
  std::vector<double> trace(static_cast<size_t>(L) * 2 + 2);
  std::vector<uint8_t> steps(L + 1);
  double fs = 0; int64_t na = 0, nr = 0; int32_t acc = 0;
  check(ahx_flip_one(ctx_, L, Tour.Tigs.data(), Signs.data(), &fs, &na, &nr, trace.data(), steps.data(), &acc), "ahx_flip_one");
  for (int i = 0; i < L; ++i)
    if ((i + 1) % 50 == 0) {   // orientation.go:103-106
      char m[64];
      snprintf(m, sizeof m, "FLIPONE (%d/%d)", i + 1, L);
      flipLog(m, trace[2 * i], trace[2 * i + 1], steps[i] ? ACCEPT : REJECT);
    }
  logNotice("FLIPONE: N_accepts=%lld N_rejects=%lld", static_cast<long long>(na), static_cast<long long>(nr));
  return acc ? ACCEPT : REJECT;
}

void CLM::printTour(FILE *fw, const allhic::Tour &tour, const std::string &label) {
  if (!fw) return;
  fprintf(fw, ">%s\n", label.c_str());
  for (int i = 0; i < tour.Len(); ++i) {
    const int idx = tour.Tigs[i];
    fprintf(fw, "%s%s%c", i ? " " : "", Tigs[idx].Name.c_str(), Signs[idx]);
  }
  fputc('\n', fw);
  fflush(fw);
}

void CLM::parseTourFile(const std::string &filename) {
  logNotice("Parse tour file `%s`", filename.c_str());
  std::ifstream f(filename);
  if (!f) logFatal("open %s: no such file or directory", filename.c_str());
  std::string row, last;
  while (std::getline(f, row)) {
    while (!row.empty() && isspace(static_cast<unsigned char>(row.back()))) row.pop_back();
    size_t a = 0;
    while (a < row.size() && isspace(static_cast<unsigned char>(row[a]))) ++a;
    row = row.substr(a);
    if (row.empty() || row[0] == '>') continue;
    last = row;   // only the last record is retained (optimize.go:98-117)
  }
  Signs.assign(Tigs.size(), 0);                       // prepareTour, optimize.go:120-125
  for (auto &t : Tigs) t.IsActive = false;
  Tour.Tigs.clear();
  std::istringstream ss(last);
  std::string word;
  while (std::getline(ss, word, ' ')) {
    if (word.empty()) continue;
    const std::string name = word.substr(0, word.size() - 1);
    const int idx = ahx_clm_tig_index(clm_, name.c_str());
    if (idx < 0) { fprintf(stderr, "ERROR Contig %s not found!\n", name.c_str()); continue; }
    Tour.Tigs.push_back(idx);
    Signs[idx] = static_cast<uint8_t>(word.back());
    Tigs[idx].IsActive = true;
  }
  printTour(out_, Tour, "INIT");
}

struct GaCbCtx { CLM *clm; FILE *fw; FILE *out; };

static void ga_progress(void *user, int32_t phase, int64_t gen, double best, const int32_t *tour, int32_t L) {
  auto *c = static_cast<GaCbCtx *>(user);
  fprintf(c->out, "Current iteration GA%d-%lld: max_score=%.5f\n", phase, static_cast<long long>(gen), best);   // evaluate.go:295
  allhic::Tour t;
  t.Tigs.assign(tour, tour + L);
  char label[96];
  snprintf(label, sizeof label, "GA%d-%lld-%.5f", phase, static_cast<long long>(gen), best);
  c->clm->printTour(c->fw, t, label);
}

allhic::Tour CLM::GARun(FILE *fwtour, Optimizer *opt, int phase) {
  ahx_ga_params prm;
  ahx_ga_default_params(&prm);
  prm.npop = opt->NPop; prm.ngen = opt->NGen; prm.mut_prob = opt->MutProb;
  prm.seed = static_cast<uint64_t>(opt->Seed) * 1000003ULL + static_cast<uint64_t>(phase);   // one stream per phase
  prm.phase = phase;
  logNotice("GA initialized (npop: %d, ngen: %d, mu: %.2f, rng: %lld, break: %d)", opt->NPop, opt->NGen, opt->MutProb,
            static_cast<long long>(opt->Seed), AHX_LIMIT);
  std::vector<int32_t> best(Tour.Len());
  ahx_ga_stats st;
  GaCbCtx cb{this, fwtour, out_};
  if (islands_.size() > 1) {   // one island per GPU; the exchange runs inside the GA kernels (include/ahx.h "Islands")
    prm.migrate_every = opt->MigrateEvery; prm.n_elite = opt->NElite;
    logNotice("Island model: %zu populations of %d, %d elites to every other island every %d generations", islands_.size(), opt->NPop, opt->NElite, opt->MigrateEvery);
    check(ahx_ga_run_islands(islands_.data(), static_cast<int32_t>(islands_.size()), &prm, Tour.Len(), Tour.Tigs.data(), best.data(), &st, ga_progress, &cb), "ahx_ga_run_islands");
  } else {
    check(ahx_ga_run(ctx_, &prm, Tour.Len(), Tour.Tigs.data(), best.data(), &st, ga_progress, &cb), "ahx_ga_run");
  }
  Tour.Tigs = best;   // r.Tour = ga.HallOfFame[0]
  return Tour;
}

void CLM::OptimizeOrdering(FILE *fwtour, Optimizer *opt, int phase) { GARun(fwtour, opt, phase); }

void CLM::OptimizeOrientations(FILE *fwtour, int phase, const char **tag1, const char **tag2) {
  *tag1 = flipWhole();
  printTour(fwtour, Tour, "FLIPWHOLE" + std::to_string(phase));
  *tag2 = flipOne();
  printTour(fwtour, Tour, "FLIPONE" + std::to_string(phase));
}

void Optimizer::Run() {
  std::mt19937_64 rng(static_cast<uint64_t>(Seed));
  std::vector<ahx_ctx *> own;   // --gpus N without borrowed contexts: devices Device .. Device + N - 1
  struct Cleanup { std::vector<ahx_ctx *> &v; ~Cleanup() { for (ahx_ctx *c : v) ahx_destroy(c); } } cleanup{own};
  if (Ctxs.empty() && Gpus > 1) {
    for (int i = 0; i < Gpus; ++i) {
      ahx_ctx *c = nullptr;
      if (ahx_create(Device + i, &c) != AHX_OK) logFatal("%s", ahx_last_error(nullptr));
      own.push_back(c);
    }
  }
  ahx_clm *parsed = Parsed;
  Parsed = nullptr;
  CLM clm(Clmfile, REfile, Device, Ctxs.empty() ? own : Ctxs, parsed, Out);
  const std::string tourfile = RemoveExt(REfile) + ".tour";
  struct stat sb;
  if (Resume && stat(tourfile.c_str(), &sb) == 0) {   // optimize.go:44-51
    logNotice("Found existing tour file `%s`", tourfile.c_str());
    clm.parseTourFile(tourfile);
    const std::string backup = tourfile + ".sav";
    rename(tourfile.c_str(), backup.c_str());
    logNotice("Backup `%s` to `%s`", tourfile.c_str(), backup.c_str());
  }
  clm.Activate(true, rng);
  logNotice("Optimization history logged to `%s`", tourfile.c_str());
  FILE *fwtour = fopen(tourfile.c_str(), "w");   // os.Create; the reference ignores the error
  OutTourFile = tourfile;
  clm.printTour(Out, clm.Tour, "INIT");
  clm.printTour(fwtour, clm.Tour, "INIT");
  if (RunGA)
    for (int phase = 1; phase < 3; ++phase) clm.OptimizeOrdering(fwtour, this, phase);
  for (int phase = 1;; ++phase) {
    const char *tag1, *tag2;
    clm.OptimizeOrientations(fwtour, phase, &tag1, &tag2);
    if (tag1 == REJECT && tag2 == REJECT) {
      logNotice("Terminating ... no more %s", ACCEPT);
      break;
    }
  }
  clm.printTour(Out, clm.Tour, "FINAL");
  logNotice("Success");
  if (fwtour) fclose(fwtour);
}

}  // namespace allhic
