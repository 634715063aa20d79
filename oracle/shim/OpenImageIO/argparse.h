// Minimal stand-in for <OpenImageIO/argparse.h> (test infrastructure only).
// Implements just the subset of OIIO::ArgParse behaviour that envutil_main.cc:104-378
// relies on: "--name METAVAR" single-value options retrievable through ap["name"],
// flag options bound to bool*, and printf-like multi-value specs with %s %d %f %F %L.
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace OIIO {

class ArgParse {
 public:
  struct Arg {
    std::string flag;                 // e.g. "--facet"
    std::vector<char> codes;          // per value: 's','d','f','F','L', or 'v' (stored by name)
    std::vector<void*> ptrs;          // destination per value (nullptr for 'v')
    bool* flagptr = nullptr;          // for valueless flags
    std::string helptext, meta;
    Arg& help(const std::string& h) { helptext = h; return *this; }
    Arg& metavar(const std::string& m) { meta = m; return *this; }
    Arg& action(int) { return *this; }
  };

  struct Value {
    bool present = false;
    std::string s;
    std::string as_string(const std::string& dflt = std::string()) const { return present ? s : dflt; }
    template <class T> T get(const T& dflt = T()) const {
      if (!present) return dflt;
      std::istringstream is(s);
      T v = dflt;
      is >> v;
      return v;
    }
  };

  ArgParse() {}
  ~ArgParse() { for (auto* a : m_args) delete a; }
  ArgParse& intro(const std::string&) { return *this; }
  ArgParse& usage(const std::string&) { return *this; }
  ArgParse& description(const std::string&) { return *this; }
  ArgParse& separator(const std::string&) { return *this; }

  Arg& arg(const char* spec) { return add(spec, {}); }
  template <class... P> Arg& arg(const char* spec, P*... p) { return add(spec, {(void*)p...}, sizeof...(P) == 1 ? first_is_bool(p...) : false); }
  Arg& add_argument(const char* spec) { return add(spec, {}); }
  template <class... P> Arg& add_argument(const char* spec, P*... p) { return add(spec, {(void*)p...}, sizeof...(P) == 1 ? first_is_bool(p...) : false); }

  int parse(int argc, const char** argv) {
    for (int i = 1; i < argc; i++) {
      std::string tok = argv[i];
      Arg* a = find(tok);
      if (!a) {
        m_error = "Invalid option \"" + tok + "\"";
        return -1;
      }
      if (a->flagptr && a->codes.empty()) {
        *a->flagptr = true;
        continue;
      }
      for (std::size_t k = 0; k < a->codes.size(); k++) {
        if (++i >= argc) {
          m_error = "Missing parameter for \"" + tok + "\"";
          return -1;
        }
        std::string val = argv[i];
        void* p = a->ptrs[k];
        switch (a->codes[k]) {
          case 'v': { Value& v = m_values[a->flag.substr(a->flag.find_first_not_of('-'))]; v.present = true; v.s = val; break; }
          case 's': *(std::string*)p = val; break;
          case 'd': *(int*)p = std::atoi(val.c_str()); break;
          case 'f': *(float*)p = float(std::atof(val.c_str())); break;
          case 'F': *(double*)p = std::atof(val.c_str()); break;
          case 'L': ((std::vector<std::string>*)p)->push_back(val); break;
        }
      }
    }
    return 0;
  }
  std::string geterror() const { return m_error; }
  void print_help() const {
    for (auto* a : m_args) std::cout << "  " << a->flag << " " << a->meta << "   " << a->helptext << std::endl;
  }
  Value operator[](const std::string& name) const {
    auto it = m_values.find(name);
    return it == m_values.end() ? Value() : it->second;
  }

 private:
  static bool first_is_bool(bool*) { return true; }
  template <class T, class... R> static bool first_is_bool(T*, R*...) { return false; }

  Arg& add(const char* spec, std::vector<void*> ptrs, bool boolflag = false) {
    Arg* a = new Arg;
    std::istringstream is(spec);
    std::string tok;
    is >> a->flag;
    std::size_t np = 0;
    while (is >> tok) {
      if (tok[0] == '%') {
        a->codes.push_back(tok[1]);
        a->ptrs.push_back(np < ptrs.size() ? ptrs[np] : nullptr);
        np++;
      } else {
        a->codes.push_back('v');
        a->ptrs.push_back(nullptr);
        a->meta = tok;
      }
    }
    if (a->codes.empty() && boolflag) a->flagptr = (bool*)ptrs[0];
    m_args.push_back(a);
    return *a;
  }
  Arg* find(const std::string& flag) {
    // later registrations do not shadow earlier ones; first match wins
    for (auto* a : m_args)
      if (a->flag == flag) return a;
    return nullptr;
  }
  std::vector<Arg*> m_args;
  std::map<std::string, Value> m_values;
  std::string m_error;
};

}  // namespace OIIO
