// pto.cc - the PTO subset envutil honours (reference pto.h:82-240): a line is a one-letter head
// followed by items `NAMEvalue`, NAME = letters, value = a quoted string or a run of
// non-blanks; `=N` in an i-line refers back to the same field of i-line N (:137-147).
#include <cctype>
#include <fstream>

#include "envutil_host.h"

namespace eu_host {

const std::string& pto_line::get(const std::string& key) const {
  static const std::string empty;
  // the reference keeps a map: a later duplicate overwrites an earlier one
  const std::string* hit = &empty;
  for (const auto& kv : fields)
    if (kv.first == key) hit = &kv.second;
  return *hit;
}

bool parse_pto_line(const std::string& raw, std::vector<pto_line>& lines) {
  std::string s = raw;
  while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
  // pto_line_regex "([a-zA-Z])\s(.+)": anything else (comments, blank lines) is skipped
  if (s.size() < 3 || !isalpha((unsigned char)s[0]) || !isspace((unsigned char)s[1])) return true;
  pto_line ln;
  ln.head = s[0];
  size_t i = 2, n = s.size();
  while (i < n) {
    // pto_item_regex "([A-Za-z]+)((\"[^\"]+\")|(\S*))" scanned left to right
    while (i < n && !isalpha((unsigned char)s[i])) i++;
    if (i >= n) break;
    size_t a = i;
    while (i < n && isalpha((unsigned char)s[i])) i++;
    std::string name = s.substr(a, i - a), value;
    if (i < n && s[i] == '"') {
      size_t q = s.find('"', i + 1);
      if (q != std::string::npos && q > i + 1) {
        value = s.substr(i, q - i + 1);
        i = q + 1;
      } else {  // no closing quote: the \S* alternative takes the run of non-blanks
        size_t b = i;
        while (i < n && !isspace((unsigned char)s[i])) i++;
        value = s.substr(b, i - b);
      }
    } else {
      size_t b = i;
      while (i < n && !isspace((unsigned char)s[i])) i++;
      value = s.substr(b, i - b);
    }
    if (!value.empty() && value[0] == '=' && name != "j") {
      int ref = atoi(value.c_str() + 1);
      int seen = -1;
      for (const auto& l : lines)
        if (l.head == 'i' && ++seen == ref) {
          value = l.get(name);
          break;
        }
    }
    ln.fields.emplace_back(name, value);
  }
  lines.push_back(ln);
  return true;
}

bool read_pto_file(const std::string& fn, const std::vector<std::string>& addenda, std::vector<pto_line>& lines,
                   std::string& err) {
  if (!fn.empty()) {
    std::ifstream str(fn);
    if (!str) {
      err = "could not open pto file " + fn;
      return false;
    }
    std::string line;
    while (std::getline(str, line)) parse_pto_line(line, lines);
  }
  for (const auto& l : addenda) parse_pto_line(l, lines);
  return true;
}

}  // namespace eu_host
