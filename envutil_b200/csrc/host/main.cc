// main.cc - envutil_b200_cli: envutil's command line on the B200 back-end.
//   envutil_b200_cli [options] --output OUT      one job        (envutil_main.cc:1922-1947)
//   envutil_b200_cli [common options] -          pipe mode: one argument line per job on stdin,
//                                                sources stay staged in HBM between jobs
//                                                (envutil_main.cc:1948-1982)
#include <cstdio>
#include <iostream>

#include "envutil_host.h"

int main(int argc, const char** argv) {
  using namespace eu_host;
  if (argc < 2) {
    fprintf(stderr, "usage: envutil_b200_cli [options...] --output OUTPUT   (or a trailing '-' for pipe mode)\n");
    return 1;
  }
  if (std::string(argv[argc - 1]) == "+") {
    fprintf(stderr, "envutil_b200: tethered (visor) mode is outside the built path\n");
    return 1;
  }
  int rc = 0;
  if (std::string(argv[argc - 1]) != "-") {
    rc = core(argc, argv);
  } else {
    argc--;
    std::string line;
    while (std::getline(std::cin, line)) {
      std::vector<std::string> sv = tokenize(line);
      if (sv.empty()) continue;
      std::vector<const char*> av(argv, argv + argc);
      for (const auto& t : sv) av.push_back(t.c_str());
      int r = core((int)av.size(), av.data());
      if (r) rc = r;
    }
    printf("pipe has reached EOF\n");
  }
  eu_shutdown();
  return rc ? 1 : 0;
}
