// Single translation unit of libcald_b200.so (kernels are defined in headers).
#include "ops_abi.cu"
#include "engine.cu"
