// Host side of the tensor-map (TMA) staging of the chunked WKV-7 kernels: one CUtensorMap per activation tensor of a
// launch, describing the reference's [B, T, H, 64] bf16 layout (model/llm/cuda/wkv7_cuda.cu:10-16 indexes it the same
// way) as a 3-D tensor (channel, head, token) whose box is one 16-token chunk of one head: 16 x 64 bf16 = 2 KB.
// The maps travel to the kernel as a __grid_constant__ parameter; the driver entry point is looked up at run time
// (no link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rwkvtts {

struct TmaMaps {
    CUtensorMap m[7];      // forward: w q k v a b; backward: + dy
};

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<TensorMapEncodeTiledFn>(p);
    }();
    return fn;
}

// [n_tok, H, 64] bf16 contiguous at `base` (16-byte aligned); box = {64 channels, 1 head, 16 tokens}
inline cudaError_t make_chunk_map(CUtensorMap *map, const void *base, long long n_tok, int H) {
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (enc == nullptr) return cudaErrorNotSupported;
    const cuuint64_t dims[3] = {64, (cuuint64_t)H, (cuuint64_t)n_tok};
    const cuuint64_t strides[2] = {128, (cuuint64_t)H * 128};          // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {64, 1, 16};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace rwkvtts
