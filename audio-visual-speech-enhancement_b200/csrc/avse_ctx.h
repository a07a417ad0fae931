// Internal: the opaque context behind include/avse_b200.h and shared error helpers.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "avse_tables.h"
#include "avse_fwd_stages.cuh"

struct avse_ctx {
    int device = 0;
    int num_sms = 148;
    bool std_tables = false;      // mel round widths equal AVSE_STD_ROUNDW -> unrolled kernels
    avse::HostTables host;
    void* dbase = nullptr;        // one device allocation holding every table
    avse::FwdTables fwd{};
    const float* d_tri_w = nullptr;
    const float* d_tri_ipiv = nullptr;
    const float* d_tri_sup = nullptr;
    const int* d_col_band = nullptr;
    const float* d_col_w = nullptr;
};

int avse_fail(int code, const std::string& msg);
int avse_cuda_fail(cudaError_t e, const char* where);
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return avse_cuda_fail(e_, #x); } while (0)
