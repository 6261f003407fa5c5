#include "common.cuh"
extern "C" int tiger_abi_version(void) { return 1; }
