// Minimal stand-in for <OpenImageIO/filesystem.h> (test infrastructure only).
#pragma once
namespace OIIO {
namespace Filesystem {
inline void convert_native_arguments(int, const char**) {}
}  // namespace Filesystem
}  // namespace OIIO
