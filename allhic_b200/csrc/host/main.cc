// allhic-b200: `optimize` sub-command with the reference's flags (allhic.go:160-190), and
// `optimize-many`: the same run for a list of groups in ONE process -- what `pipeline` does with
// its sequential loop over the groups' REfiles (allhic.go:285-293), except that
//   * the process drives every GPU of the box (--gpus N) and keeps --jobs J groups in flight per
//     GPU, each on its own context and stream: at the CLI's default population (100) one group
//     fills 25 of a B200's 148 SMs, so several groups run side by side; one CUDA start-up for all;
//   * groups that name the same clmfile (pipeline hands the genome-wide CLM to every group) share
//     ONE parse of it (ahx_clm_parse_groups).
// Every group's .tour file and stdout text are the same as a separate `optimize` run's.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ahx.h"
#include "optimizer.h"

static void usage() {
  fprintf(stderr,
          "Usage:\n  allhic-b200 optimize counts_RE.txt clmfile [flags]\n"
          "  allhic-b200 optimize-many listfile|- [flags]     one `counts_RE.txt clmfile` pair per line (- = stdin);\n"
          "                                                   prints `AHX_GROUP_DONE <line> <counts_RE.txt>` after each\n\nFlags:\n"
          "      --mutapb float   Mutation prob in GA (default 0.2)\n"
          "      --ngen int       Number of generations for convergence (default 5000)\n"
          "      --npop int       Population size (default 100)\n"
          "      --resume         Resume from existing tour file\n"
          "      --seed int       Random seed (default 42)\n"
          "      --skipGA         Skip GA step\n"
          "      --device int     first CUDA device (default 0)\n"
          "      --gpus int       optimize: island model, one population of --npop per GPU (default 1);\n"
          "                       optimize-many: farm the groups over this many GPUs (default 1)\n"
          "      --migrate int    islands: generations between exchanges (default 50)\n"
          "      --elite int      islands: elites sent to every other island per exchange (default 2)\n"
          "      --jobs int       optimize-many: groups in flight per GPU (default 1)\n");
}

namespace {

struct Group { int idx; std::string re, clm; long ntigs = 0; ahx_clm *parsed = nullptr; };

long count_contigs(const std::string &refile) {
  FILE *f = fopen(refile.c_str(), "r");
  if (!f) return 0;
  long n = 0; char line[65536];
  while (fgets(line, sizeof line, f)) { const char *p = line; while (*p == ' ' || *p == '\t') ++p; if (*p && *p != '\n' && *p != '#') ++n; }
  fclose(f);
  return n;
}

// Runs one group; its stdout text goes to `text` (emitted in one piece by the caller).
bool run_group(const allhic::Optimizer &flags, Group &g, const std::vector<ahx_ctx *> &ctxs, std::string *text) {
  char *buf = nullptr; size_t len = 0;
  FILE *mem = open_memstream(&buf, &len);
  bool ok = true;
  try {
    allhic::Optimizer q = flags;                                   // same flags (and seed) for every group, like `pipeline`
    q.REfile = g.re; q.Clmfile = g.clm; q.Ctxs = ctxs; q.Out = mem; q.Parsed = g.parsed;
    g.parsed = nullptr;
    q.Run();
  } catch (const allhic::FatalError &) {
    ok = false;                                                    // log.Fatal of this group only
  }
  fclose(mem);
  text->assign(buf ? buf : "", len);
  free(buf);
  return ok;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2 || (strcmp(argv[1], "optimize") != 0 && strcmp(argv[1], "optimize-many") != 0)) { usage(); return 2; }
  const bool many = strcmp(argv[1], "optimize-many") == 0;
  allhic::Optimizer p;
  int jobs = 1;
  std::vector<std::string> pos;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i], v;
    const size_t eq = a.find('=');
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); }
    auto val = [&]() -> std::string { if (!v.empty()) return v; if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
    if (a == "--skipGA") p.RunGA = false;
    else if (a == "--resume") p.Resume = true;
    else if (a == "--seed") p.Seed = atoll(val().c_str());
    else if (a == "--npop") p.NPop = atoi(val().c_str());
    else if (a == "--ngen") p.NGen = atoi(val().c_str());
    else if (a == "--mutapb") p.MutProb = atof(val().c_str());
    else if (a == "--device") p.Device = atoi(val().c_str());
    else if (a == "--gpus") p.Gpus = atoi(val().c_str());
    else if (a == "--migrate") p.MigrateEvery = atoi(val().c_str());
    else if (a == "--elite") p.NElite = atoi(val().c_str());
    else if (a == "--jobs") jobs = atoi(val().c_str());
    else if (a.rfind("--", 0) == 0) { fprintf(stderr, "unknown flag: %s\n", a.c_str()); usage(); return 2; }
    else pos.push_back(a);
  }
  if (p.Gpus < 1 || jobs < 1) { usage(); return 2; }
  if (!many) {
    if (pos.size() != 2) { fprintf(stderr, "accepts 2 arg(s), received %zu\n", pos.size()); usage(); return 2; }   // cobra.ExactArgs(2)
    p.REfile = pos[0];
    p.Clmfile = pos[1];
    try { p.Run(); } catch (const allhic::FatalError &) { return 1; }   // log.Fatal -> os.Exit(1)
    return 0;
  }

  if (pos.size() != 1) { fprintf(stderr, "optimize-many accepts 1 arg, received %zu\n", pos.size()); usage(); return 2; }
  FILE *list = pos[0] == "-" ? stdin : fopen(pos[0].c_str(), "r");
  if (!list) { fprintf(stderr, "cannot open %s\n", pos[0].c_str()); return 1; }
  const int ngpus = p.Gpus;
  p.Gpus = 1;                                                      // here --gpus farms groups; every group is one population
  const int nworkers = ngpus * jobs;
  char line[8192];

  if (nworkers == 1) {
    // one group at a time, as they arrive (a driver may feed stdin and wait for each AHX_GROUP_DONE)
    ahx_ctx *ctx = nullptr;
    if (ahx_create(p.Device, &ctx) != AHX_OK) { fprintf(stderr, "%s\n", ahx_last_error(nullptr)); return 1; }
    int failed = 0;
    for (int idx = 0; fgets(line, sizeof line, list);) {
      char re[4096], clm[4096];
      if (sscanf(line, "%4095s %4095s", re, clm) != 2) continue;   // blank line
      Group g{idx, re, clm};
      std::string text;
      const bool ok = run_group(p, g, {ctx}, &text);
      fwrite(text.data(), 1, text.size(), stdout);
      printf("%s %d %s\n", ok ? "AHX_GROUP_DONE" : "AHX_GROUP_FAILED", idx++, re);
      fflush(stdout);
      failed += !ok;
    }
    ahx_destroy(ctx);
    if (list != stdin) fclose(list);
    return failed ? 1 : 0;
  }

  // ---- several groups in flight: read the whole list, share CLM parses, largest group first
  std::vector<Group> groups;
  while (fgets(line, sizeof line, list)) {
    char re[4096], clm[4096];
    if (sscanf(line, "%4095s %4095s", re, clm) != 2) continue;
    Group g{static_cast<int>(groups.size()), re, clm};
    g.ntigs = count_contigs(g.re);
    groups.push_back(g);
  }
  if (list != stdin) fclose(list);
  {
    std::map<std::string, std::vector<int>> by_clm;
    for (auto &g : groups) by_clm[g.clm].push_back(g.idx);
    for (auto &kv : by_clm) {
      if (kv.second.size() < 2) continue;
      const auto t0 = std::chrono::steady_clock::now();
      std::vector<const char *> res;
      for (int i : kv.second) res.push_back(groups[i].re.c_str());
      std::vector<ahx_clm *> out(res.size(), nullptr);
      if (ahx_clm_parse_groups(kv.first.c_str(), res.data(), static_cast<int32_t>(res.size()), 0, out.data()) != AHX_OK) {
        fprintf(stderr, "%s\n", ahx_last_error(nullptr));                    // the groups then fail one by one on their own parse
        continue;
      }
      for (size_t k = 0; k < out.size(); ++k) groups[kv.second[k]].parsed = out[k];
      allhic::logNotice("Parsed clmfile `%s` once for %zu groups in %.2f s", kv.first.c_str(), out.size(),
                        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
  }
  std::vector<int> order(groups.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int>(i);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return groups[a].ntigs > groups[b].ntigs; });   // ~N^2 work: longest first
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  // AHX_FARM_TIMING=1: wall-clock offsets of every worker's milestones on stderr
  const bool farm_timing = getenv("AHX_FARM_TIMING") && atoi(getenv("AHX_FARM_TIMING"));
  const auto farm_t0 = std::chrono::steady_clock::now();
  auto farm_now = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - farm_t0).count(); };
  std::mutex out_mu;
  std::vector<std::thread> workers;
  for (int w = 0; w < nworkers; ++w)
    workers.emplace_back([&, w] {
      const int dev = p.Device + w % ngpus;
      ahx_ctx *ctx = nullptr;
      if (ahx_create(dev, &ctx) != AHX_OK) { fprintf(stderr, "%s\n", ahx_last_error(nullptr)); return; }
      if (farm_timing) fprintf(stderr, "farm: worker %d (device %d) has its context at %.3f s\n", w, dev, farm_now());
      for (size_t k = next++; k < order.size(); k = next++) {
        Group &g = groups[order[k]];
        allhic::logSetTag("[group " + std::to_string(g.idx) + "] ");
        allhic::Optimizer q = p;
        q.Device = dev;
        std::string text;
        const auto t0 = std::chrono::steady_clock::now();
        const bool ok = run_group(q, g, {ctx}, &text);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (farm_timing) fprintf(stderr, "farm: worker %d group %d from %.3f to %.3f s\n", w, g.idx, farm_now() - dt, farm_now());
        std::lock_guard<std::mutex> lk(out_mu);
        fwrite(text.data(), 1, text.size(), stdout);
        printf("%s %d %s device=%d seconds=%.3f\n", ok ? "AHX_GROUP_DONE" : "AHX_GROUP_FAILED", g.idx, g.re.c_str(), dev, dt);
        fflush(stdout);
        failed += !ok;
      }
      ahx_destroy(ctx);
    });
  for (auto &t : workers) t.join();
  if (farm_timing) fprintf(stderr, "farm: all workers joined at %.3f s\n", farm_now());   // what follows (1.3 s on an 8-GPU box) is the driver tearing the process's CUDA contexts down; ending the process with _exit instead did not shorten it
  for (auto &g : groups) if (g.parsed) ahx_clm_free(g.parsed);   // groups no worker reached (every worker failed to start)
  return failed ? 1 : (next.load() >= order.size() ? 0 : 1);
}
