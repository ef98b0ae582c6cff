// optimizer.h -- C++ host side of `allhic optimize`, written against the C ABI
// only (include/ahx.h).  It mirrors the reference's Go types for this path --
// Optimizer (optimize.go:21-35), CLM (clm.go:32-41), Tour (clm.go:88-91) -- with
// the same names, argument meaning and error behaviour, so that it is what the
// Go host becomes once its three call sites are bound to libahx through cgo
// (INTEGRATION.md).  The Go toolchain is not available in the build image; this
// translation unit is the stand-in that is compiled and tested.
#ifndef AHX_HOST_OPTIMIZER_H_
#define AHX_HOST_OPTIMIZER_H_

#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "ahx.h"

namespace allhic {

constexpr const char *ACCEPT = "ACCEPT";   // orientation.go:19
constexpr const char *REJECT = "REJECT";   // orientation.go:22

struct Optimizer;

// Tour: Tigs[i].Idx (sizes live in CLM.Tigs; M lives on the device)
struct Tour {
  std::vector<int32_t> Tigs;
  int Len() const { return static_cast<int>(Tigs.size()); }
};

struct TigF { int Idx; std::string Name; int64_t Size; bool IsActive; };   // clm.go:74-79

class CLM {
 public:
  std::string REfile, Clmfile;
  std::vector<TigF> Tigs;
  allhic::Tour Tour;
  std::vector<uint8_t> Signs;

  // NewCLM (clm.go:121): parse both files and upload the group to `device`.  `shared`: contexts the
  // caller owns and reuses from group to group (`optimize-many`; more than one = one island per context);
  // else one is created and destroyed here.  `parsed`: this group's share of a genome-wide CLM that the
  // caller parsed once for all groups (ahx_clm_parse_groups); ownership passes to the CLM.
  CLM(const std::string &clmfile, const std::string &refile, int device, const std::vector<ahx_ctx *> &shared = {},
      ahx_clm *parsed = nullptr, FILE *out = stdout);
  ~CLM();
  CLM(const CLM &) = delete;
  CLM &operator=(const CLM &) = delete;

  void Activate(bool shuffle, std::mt19937_64 &rng);                 // clm.go:392
  allhic::Tour GARun(FILE *fwtour, Optimizer *opt, int phase);       // evaluate.go:258
  double EvaluateQ();                                                // orientation.go:159
  double Evaluate();                                                 // Tour.Evaluate, evaluate.go:116 (returns -score like Go)
  const char *flipAll();                                             // orientation.go:31
  const char *flipWhole();                                           // orientation.go:67
  const char *flipOne();                                             // orientation.go:87
  void OptimizeOrdering(FILE *fwtour, Optimizer *opt, int phase);    // optimize.go:82
  void OptimizeOrientations(FILE *fwtour, int phase, const char **tag1, const char **tag2);  // optimize.go:88
  void printTour(FILE *fw, const allhic::Tour &tour, const std::string &label);             // optimize.go:176
  void parseTourFile(const std::string &filename);                   // optimize.go:129
  void reportActive(bool verbose, int *activeCounts, int64_t *sumLength);                   // clm.go:421

 private:
  ahx_clm *clm_ = nullptr;
  ahx_ctx *ctx_ = nullptr;                 // islands_[0]: scores, flips, the single-population GA
  std::vector<ahx_ctx *> islands_;         // one context per GPU of the island model (size 1: none)
  bool own_ctx_ = true;
  FILE *out_ = stdout;                     // where the reference writes to os.Stdout
  void check(int rc, const char *what);
};

struct Optimizer {       // optimize.go:21-35
  std::string REfile, Clmfile;
  bool RunGA = true, Resume = false;
  int64_t Seed = 42;     // base.go:65
  int NPop = 100;        // base.go:67
  int NGen = 5000;       // base.go:69
  double MutProb = 0.2;  // base.go:71
  int Device = 0;
  // Island model over several GPUs (--gpus N; eaopt's NPops, which the reference pins to 1, evaluate.go:269):
  // NPop individuals per island, an exchange of NElite elites every MigrateEvery generations.
  int Gpus = 1, MigrateEvery = 50, NElite = 2;
  std::vector<ahx_ctx *> Ctxs;   // borrowed device contexts (optimize-many / islands); empty: Run() makes its own
  ahx_clm *Parsed = nullptr;     // pre-parsed group (shared genome-wide parse); Run() hands it to the CLM
  FILE *Out = stdout;            // os.Stdout of this run (a buffer when several groups run at once)
  std::string OutTourFile;
  void Run();            // optimize.go:38
};

std::string RemoveExt(const std::string &filename);   // base.go:125
void logNotice(const char *fmt, ...);
// log.Fatal (base.go:349-362): message, then the run ends.  Thrown as an exception so that
// `optimize-many` loses the failing group only; `optimize` turns it into exit(1) like the reference.
struct FatalError { std::string msg; };
[[noreturn]] void logFatal(const char *fmt, ...);
// tag prepended to this thread's log lines (the group, when several run at once)
void logSetTag(const std::string &tag);

}  // namespace allhic

#endif  // AHX_HOST_OPTIMIZER_H_
