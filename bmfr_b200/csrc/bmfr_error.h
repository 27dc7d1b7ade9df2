// Thread-local last-error plumbing behind bmfr_last_error() (include/bmfr_b200.h).
// Replaces the reference's cl::Error exceptions (bmfr.cpp:564-576) with status codes.
#pragma once
#include <cuda_runtime.h>

int bmfr_set_error(int status, const char* fmt, ...);
int bmfr_check_cuda(cudaError_t e, const char* what);

#define BMFR_CUDA_TRY(expr)                                    \
    do {                                                       \
        int _st = bmfr_check_cuda((expr), #expr);              \
        if (_st != 0) return _st;                              \
    } while (0)
