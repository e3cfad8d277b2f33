// Host-side CUtensorMap construction.  cuTensorMapEncodeTiled is a driver-API symbol; it is looked up
// through the runtime (cudaGetDriverEntryPoint) so the library has no link-time dependency on libcuda
// and still loads on a box without a driver (the CPU test tier).
#include <mutex>

#include "umma.cuh"

namespace dkd {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elt_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const char* what) {
  EncodeTiledFn enc = get_encode();
  DKD_REQUIRE(enc != nullptr, DKD_E_ARCH, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
  DKD_REQUIRE((((uintptr_t)base) & 15) == 0, DKD_E_ALIGN, "%s: tensor base must be 16-byte aligned", what);
  DKD_REQUIRE(elt_bytes == 2 || elt_bytes == 4, DKD_E_DTYPE, "%s: element size %d", what, elt_bytes);
  DKD_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 64, DKD_E_UNSUPPORTED, "%s: swizzle %d", what, swizzle_bytes);
  DKD_REQUIRE((int)box[0] * elt_bytes <= swizzle_bytes, DKD_E_SHAPE, "%s: inner box exceeds the swizzle span", what);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    DKD_REQUIRE(gstr[i] % 16 == 0, DKD_E_ALIGN, "%s: tensor stride %llu is not a multiple of 16 bytes", what, (unsigned long long)gstr[i]);
  }
  CUresult r = enc(out, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                   const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DKD_REQUIRE(r == CUDA_SUCCESS, DKD_E_LAUNCH, "%s: cuTensorMapEncodeTiled failed (CUresult %d)", what, (int)r);
  return DKD_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char* what) {
  return make_tmap(out, base, 2, 128, rank, dims, strides_bytes, box, what);
}

}  // namespace dkd
