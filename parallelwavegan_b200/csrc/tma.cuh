// Tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG): host-side descriptor encoding through the driver entry
// point (no link against libcuda) and the device-side load wrappers.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace pwgb {

typedef CUresult (*tma_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tma_encode_tiled_fn tma_encoder() {
  static tma_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (tma_encode_tiled_fn)p;
  }
  return fn;
}

// packed (dense) tensor of `rank` <= 5 dimensions of `elem_bytes`-sized elements, dims[0] fastest.  Out-of-bounds box
// elements read as zero.  Keep the innermost box extent long (>= 128 B): TMA requests whole 32-byte sectors per inner
// row, so a 16-byte inner extent doubles the L2 traffic (measured: profiles/ncu_r2_wavenet_fused_v2_*).
inline int tma_make(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, int rank, const void* base, const unsigned long long* dims,
                    const unsigned* box) {
  tma_encode_tiled_fn enc = tma_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return PWGB_CUDA_ERROR;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  unsigned long long stride = elem_bytes;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return PWGB_CUDA_ERROR;
  }
  return PWGB_OK;
}

__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// L2 prefetch of a box (no shared memory, no barrier): later loads of the same box hit L2 instead of HBM
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared -> global box store / element-wise reduction (fp32 add performed by the L2), tracked by the issuing thread's
// bulk async-group: tma_store_commit() after issuing, tma_store_wait_read<N>() before the staging memory is rewritten
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, unsigned src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, unsigned src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, unsigned src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

}  // namespace pwgb
