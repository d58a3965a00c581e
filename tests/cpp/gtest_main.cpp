// main() for the reference's test translation units compiled against the engine (oracle/Makefile: ref_tests_on_b200):
// select the GPU, run every registered TEST.
#include <gtest/gtest.h>

#include "../../include/ecb200.h"

int main() {
  if (ecb200_init(0) != ECB200_OK) {
    std::printf("ecb200_init failed: %s\n", ecb200_last_error());
    return 2;
  }
  return gtest_shim::run_all();
}
