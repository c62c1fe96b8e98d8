// core/ApiVersion.h -- version macros reference scripts include (src/core/ApiVersion.h.in of the reference)
#pragma once
#define DEME_VERSION_MAJOR 2
#define DEME_VERSION_MINOR 1
#define DEME_VERSION_PATCH 0
#define DEME_B200_NATIVE 1
