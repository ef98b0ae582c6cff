// allhic-b200: `optimize` sub-command with the reference's flags (allhic.go:160-190), and
// `optimize-many`: the same run for a list of groups in ONE process and one device context --
// what `pipeline` does with its sequential loop over the groups' REfiles (allhic.go:285-293).
// Creating a CUDA context costs about a second (several when 8 processes start at once), as
// much as the whole optimisation of a small group.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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
          "      --device int     CUDA device (default 0)\n");
}

int main(int argc, char **argv) {
  if (argc < 2 || (strcmp(argv[1], "optimize") != 0 && strcmp(argv[1], "optimize-many") != 0)) { usage(); return 2; }
  const bool many = strcmp(argv[1], "optimize-many") == 0;
  allhic::Optimizer p;
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
    else if (a.rfind("--", 0) == 0) { fprintf(stderr, "unknown flag: %s\n", a.c_str()); usage(); return 2; }
    else pos.push_back(a);
  }
  if (many) {
    if (pos.size() != 1) { fprintf(stderr, "optimize-many accepts 1 arg, received %zu\n", pos.size()); usage(); return 2; }
    FILE *list = pos[0] == "-" ? stdin : fopen(pos[0].c_str(), "r");
    if (!list) { fprintf(stderr, "cannot open %s\n", pos[0].c_str()); return 1; }
    ahx_ctx *ctx = nullptr;
    if (ahx_create(p.Device, &ctx) != AHX_OK) { fprintf(stderr, "%s\n", ahx_last_error(nullptr)); return 1; }
    char line[8192];
    for (int idx = 0; fgets(line, sizeof line, list);) {
      char re[4096], clm[4096];
      if (sscanf(line, "%4095s %4095s", re, clm) != 2) continue;   // blank line
      allhic::Optimizer q = p;                                       // same flags (and seed) for every group, like `pipeline`
      q.REfile = re; q.Clmfile = clm; q.Ctx = ctx;
      q.Run();
      printf("AHX_GROUP_DONE %d %s\n", idx++, re);
      fflush(stdout);
    }
    ahx_destroy(ctx);
    if (list != stdin) fclose(list);
    return 0;
  }
  if (pos.size() != 2) { fprintf(stderr, "accepts 2 arg(s), received %zu\n", pos.size()); usage(); return 2; }   // cobra.ExactArgs(2)
  p.REfile = pos[0];
  p.Clmfile = pos[1];
  p.Run();
  return 0;
}
