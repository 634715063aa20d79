// Minimal stand-in for <OpenImageIO/imagebufalgo.h> (test infrastructure only).
// colorconvert is the identity: the synthetic data are already in the working space
// (envutil_basic.h:960 only converts when the spaces differ).
#pragma once
#include "imagebuf.h"
namespace OIIO {
namespace ImageBufAlgo {
inline bool colorconvert(ImageBuf&, const ImageBuf&, const std::string&, const std::string&) { return true; }
}  // namespace ImageBufAlgo
}  // namespace OIIO
