// Host helpers for the tensor-core kernels: TMA tensor-map encoding via the driver entry point.
#include "common.cuh"
#include "tc.cuh"

namespace tgfr {

PFN_tgfr_encodeTiled get_encode_tiled() {
  static PFN_tgfr_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return nullptr;
  }
  fn = reinterpret_cast<PFN_tgfr_encodeTiled>(p);
  return fn;
}

int make_tmap_3d(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t d0, uint64_t d1,
                 uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes, uint64_t pitch_elems, bool overlap) {
  PFN_tgfr_encodeTiled enc = get_encode_tiled();
  if (!enc) return TGFR_E_CUDA;
  // The encoder is a DRIVER call: it fails with CUDA_ERROR_INVALID_CONTEXT (201) on a thread that has not touched the
  // runtime yet -- e.g. autograd's backward thread when a tensor-map-building backward is the first one of the process.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    TGFR_CUDA_OK(cudaFree(nullptr));        // binds the device's primary context to this thread
    ctx_bound = true;
  }
  TGFR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base address must be 16-byte aligned");
  const uint64_t pitch = pitch_elems ? pitch_elems : d0;
  // overlap: rows may share memory (pitch < row length): the sliding n-gram windows of TextHeading
  TGFR_REQUIRE((overlap || pitch >= d0) && (pitch * elem_bytes) % 16 == 0, "tensor map: row pitch must be a multiple of 16 bytes");
  TGFR_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 64, "tensor map: swizzle must be 64 or 128 bytes");
  TGFR_REQUIRE(b0 * elem_bytes <= (uint32_t)swizzle_bytes && b1 <= 256 && b2 <= 256, "tensor map: box too large for the swizzle");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch * (uint64_t)elem_bytes, pitch * d1 * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu x %llu, box %u x %u x %u)", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1, b2);
    return TGFR_E_CUDA;
  }
  return TGFR_OK;
}

}  // namespace tgfr
