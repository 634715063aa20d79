// Stand-in for the reference's visor.h (interactive-viewer tether over boost::interprocess,
// out of scope - SURVEY.md section 2 row 13). boost is not installed; the oracle build resolves
// `#include "visor.h"` to this file so that envutil_main.cc compiles unmodified. Only the
// names envutil_main.cc:1755-1944 mentions are declared.
//
// Tethered mode ('+' as the last argument) runs ONE job instead of visor's queue, so that the
// reference's own tethered pixel pipeline (handle_job -> core(tethered = true) -> work():
// act + to_screen_t, envutil_payload.cc:298-413,524-531) can be pinned by a golden frame:
//   EU_TETHER_SPEC  "width height yaw pitch roll hfov brighten refine"  (what visor puts into a spec_t)
//   EU_TETHER_ARGS  file with one argument per line (what visor keeps in ipc.flat_args; line 0 = argv[0])
//   EU_TETHER_OUT   file that receives the width*height uint32 sRGBA frame buffer
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

struct spec_t {
  std::size_t serial_no = 0;
  int buffer_index = 0;
  bool snapshot = false, refine = false;
  int width_cam = 0, height_cam = 0;
  double yaw_cam = 0, pitch_cam = 0, roll_cam = 0, hfov_cam = 0, brighten = 1;
  std::string filename;
};

struct ipc_data_t {
  struct store_t { bool get(int) { return true; } } store;
  struct ptr_t {
    std::byte* p = nullptr;
    std::byte* get() { return p; }
  };
  struct flat_args_t {
    std::vector<std::string> lines;
    void extract(std::size_t& argc, std::vector<const char*>& argv) {
      argv.clear();
      for (auto& s : lines) argv.push_back(s.c_str());
      argc = argv.size();
    }
  } flat_args;
  std::vector<spec_t> spec_array;
  std::vector<std::uint32_t> frame;
  int desktop_width = 0, desktop_height = 0;
  ptr_t get_buffer_address(int) { return ptr_t{reinterpret_cast<std::byte*>(frame.data())}; }
};

struct visor_protocol {
  static bool render_loop(std::function<bool(ipc_data_t&, int)> job_handler) {
    const char* spec_s = std::getenv("EU_TETHER_SPEC");
    const char* args_p = std::getenv("EU_TETHER_ARGS");
    const char* out_p = std::getenv("EU_TETHER_OUT");
    if (!spec_s || !args_p || !out_p) {
      std::cerr << "visor tether: the oracle build runs one job described by EU_TETHER_SPEC / _ARGS / _OUT" << std::endl;
      return false;
    }
    ipc_data_t ipc;
    spec_t spec;
    int refine = 0;
    std::istringstream is(spec_s);
    is >> spec.width_cam >> spec.height_cam >> spec.yaw_cam >> spec.pitch_cam >> spec.roll_cam >> spec.hfov_cam >>
        spec.brighten >> refine;
    spec.refine = refine != 0;
    spec.serial_no = 1;
    std::ifstream af(args_p);
    for (std::string line; std::getline(af, line);) ipc.flat_args.lines.push_back(line);
    ipc.desktop_width = spec.width_cam;
    ipc.desktop_height = spec.height_cam;
    ipc.frame.assign((std::size_t)spec.width_cam * spec.height_cam, 0u);
    ipc.spec_array.push_back(spec);
    bool ok = job_handler(ipc, 0);
    std::FILE* f = std::fopen(out_p, "wb");
    if (!f) return false;
    std::fwrite(ipc.frame.data(), sizeof(std::uint32_t), ipc.frame.size(), f);
    std::fclose(f);
    return ok;
  }
};
