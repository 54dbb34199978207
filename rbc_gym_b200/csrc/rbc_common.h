// rbc_common.h — error plumbing shared by the translation units of librbc_b200.so
#pragma once
#include <cuda_runtime.h>

#include <string>

int rbc_fail(const std::string& m);   // records the message for rbc_last_error() and returns -1

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return rbc_fail(std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    } while (0)
