// error plumbing shared by the C-ABI entry points
#pragma once
#include <cuda_runtime.h>
int ppp_fail(int code, const char* msg);
int ppp_check(const char* where);      // cudaGetLastError() -> code + message
