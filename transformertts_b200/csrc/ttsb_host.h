// Host-side helpers shared by the C-ABI translation units (error text, launch counter, TMA descriptor encode).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttsb {
void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch();
int num_sms();
int current_device();
int set_salt_gemm(const uint32_t* salt_dev, cudaStream_t stream);
int set_salt_rowops(const uint32_t* salt_dev, cudaStream_t stream);
int set_salt_train_ops(const uint32_t* salt_dev, cudaStream_t stream);
int set_salt_bgemm(const uint32_t* salt_dev, cudaStream_t stream);
int set_salt_attn_probs(const uint32_t* salt_dev, cudaStream_t stream);

// State that CUDA keeps PER DEVICE (cudaFuncSetAttribute, __constant__ uploads, the SM count) must be initialised once per
// device, not once per process: a process that touches a second GPU would otherwise run there with default limits / zero tables.
template <typename T>
struct PerDevice {
  T v[64] = {};
  T& get() { return v[current_device() & 63]; }
};

// bf16 tensor (dim0 contiguous, dim1 rows, dim2 batches); strides in ELEMENTS; 128-byte swizzle; OOB reads give zeros.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1,
                      uint64_t stride2, uint32_t box0, uint32_t box1);
// the same with the 64-byte swizzle (boxes of 32 bf16 columns: row stride 64 B, 16-byte chunk index ^= (row >> 1) & 3)
int make_tmap_bf16_3d_sw64(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1,
                           uint64_t stride2, uint32_t box0, uint32_t box1);
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1, uint32_t box0,
                      uint32_t box1);
}  // namespace ttsb
