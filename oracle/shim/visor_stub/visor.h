// Stand-in for the reference's visor.h (interactive-viewer tether over boost::interprocess,
// out of scope - SURVEY.md section 2 row 13). boost is not installed; the oracle build resolves
// `#include "visor.h"` to this file so that envutil_main.cc compiles unmodified. Only the
// names envutil_main.cc:1755-1944 mentions are declared; tethered mode ('+') is refused.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <iostream>
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
  struct store_t { bool get(int) { return false; } } store;
  struct ptr_t { std::byte* get() { return nullptr; } };
  struct flat_args_t {
    void extract(std::size_t& argc, std::vector<const char*>& argv) { argc = 0; argv.clear(); }
  } flat_args;
  std::vector<spec_t> spec_array;
  int desktop_width = 0, desktop_height = 0;
  ptr_t get_buffer_address(int) { return ptr_t(); }
};

struct visor_protocol {
  static void render_loop(std::function<bool(ipc_data_t&, int)>) {
    std::cerr << "visor tether is not available in the oracle build" << std::endl;
  }
};
