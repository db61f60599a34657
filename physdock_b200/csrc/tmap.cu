// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point, so
// the library has no link-time dependency on libcuda) with a small cache keyed on the map's arguments.
#include <cstring>
#include <unordered_map>

#include "umma.cuh"

namespace pdk {

namespace {

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeFn)p;
    }
    return fn;
}

struct Key {
    const void* ptr;
    uint64_t rows, cols, ld;
    uint32_t box_rows, box_cols;
    int swizzle, dtype;
    bool operator==(const Key& o) const { return std::memcmp(this, &o, sizeof(Key)) == 0; }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        uint64_t h = 1469598103934665603ull;
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
        for (size_t i = 0; i < sizeof(Key); ++i) { h ^= p[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};

cudaError_t get_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                    int swizzle_bytes, int dtype_bytes, CUtensorMap* out) {
    static thread_local std::unordered_map<Key, CUtensorMap, KeyHash> cache;
    Key k;
    std::memset(&k, 0, sizeof(k));
    k.ptr = ptr; k.rows = rows; k.cols = cols; k.ld = ld; k.box_rows = box_rows; k.box_cols = box_cols;
    k.swizzle = swizzle_bytes; k.dtype = dtype_bytes;
    auto it = cache.find(k);
    if (it != cache.end()) { *out = it->second; return cudaSuccess; }
    EncodeFn enc = encode_fn();
    if (!enc) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * dtype_bytes) & 15) ||
        (swizzle_bytes && (uint64_t)box_cols * dtype_bytes > (uint64_t)swizzle_bytes) || box_rows > 256 || box_cols > 256)
        return cudaErrorInvalidValue;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * (uint64_t)dtype_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUtensorMap m;
    CUresult r = enc(&m, dtype_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    if (cache.size() > 8192) cache.clear();
    cache.emplace(k, m);
    *out = m;
    return cudaSuccess;
}

}  // namespace

cudaError_t get_tensor_map_f16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                               uint32_t box_cols, int swizzle_bytes, CUtensorMap* out) {
    return get_map(ptr, rows, cols, ld, box_rows, box_cols, swizzle_bytes, 2, out);
}
cudaError_t get_tensor_map_f32(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                               uint32_t box_cols, int swizzle_bytes, CUtensorMap* out) {
    return get_map(ptr, rows, cols, ld, box_rows, box_cols, swizzle_bytes, 4, out);
}

}  // namespace pdk
