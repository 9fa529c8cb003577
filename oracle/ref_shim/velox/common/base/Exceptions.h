// TEST INFRASTRUCTURE. Stand-in for velox/common/base/Exceptions.h when compiling the reference's
// vendored dbgen (velox/tpch/gen/dbgen/build.cpp:39,96 uses one macro from it) without folly/fmt.
#pragma once
#include <cstdio>
#include <cstdlib>
#define VELOX_CHECK_GT(a, b) \
  do {                       \
    if (!((a) > (b))) std::abort(); \
  } while (0)
