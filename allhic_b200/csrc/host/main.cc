// allhic-b200: `optimize` sub-command with the reference's flags (allhic.go:160-190).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "optimizer.h"

static void usage() {
  fprintf(stderr,
          "Usage:\n  allhic-b200 optimize counts_RE.txt clmfile [flags]\n\nFlags:\n"
          "      --mutapb float   Mutation prob in GA (default 0.2)\n"
          "      --ngen int       Number of generations for convergence (default 5000)\n"
          "      --npop int       Population size (default 100)\n"
          "      --resume         Resume from existing tour file\n"
          "      --seed int       Random seed (default 42)\n"
          "      --skipGA         Skip GA step\n"
          "      --device int     CUDA device (default 0)\n");
}

int main(int argc, char **argv) {
  if (argc < 2 || strcmp(argv[1], "optimize") != 0) { usage(); return 2; }
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
  if (pos.size() != 2) { fprintf(stderr, "accepts 2 arg(s), received %zu\n", pos.size()); usage(); return 2; }   // cobra.ExactArgs(2)
  p.REfile = pos[0];
  p.Clmfile = pos[1];
  p.Run();
  return 0;
}
