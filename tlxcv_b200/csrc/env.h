// Environment switches of libtlxcv_b200.so (host side).
#pragma once

#include <stdlib.h>

namespace tlxcv {

// Switches read from the environment.  The shipped library honours only RESULT-PRESERVING ones (kernel choice, tiling,
// launch features: the parity tests use them to run both variants of a layer) through tuning_env().  Anything that changes
// what a kernel computes (ablation masks) or dumps timelines goes through debug_env(), which is compiled out unless the
// library is built with -DTLXCV_DEBUG_TOOLS (`make debug` -> libtlxcv_b200_debug.so, loaded via TLXCV_B200_LIB by tools/):
// a stray environment variable can never corrupt a production forward.
inline const char* tuning_env(const char* name) { return getenv(name); }
inline const char* debug_env(const char* name) {
#ifdef TLXCV_DEBUG_TOOLS
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

}  // namespace tlxcv
