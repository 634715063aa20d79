// Minimal stand-in for <OpenImageIO/color.h> (test infrastructure only): no OCIO config.
#pragma once
#include <string>
namespace OIIO {
class ColorConfig {
 public:
  static const ColorConfig& default_colorconfig() { static ColorConfig c; return c; }
  std::string getColorSpaceNameByRole(const std::string&) const { return "linear"; }
  int getNumColorSpaces() const { return 1; }
};
}  // namespace OIIO
