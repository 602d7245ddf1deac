// Embedding gather / concat of the speech layouts in ONE kernel  [HBM roofline: reads + writes every output row once].
//
// Reference (caller side of the hot path): the batch builders loop over the samples
// and, per sample, run one nn.Embedding lookup per piece ([tag2, text.., tag0, global.., tag1, semantic..]), two
// torch.cat and a pad_sequence -- data/utils/spark_dataset.py:163-239, utils/multiple_jsonl.py:4-75,
// inference/rwkv7speech_inference.py:35-67, model/llm/cosy_llm.py:64-73 -- about 20-25 kernels per sample.  Here the
// host computes, once per batch, which (table, row) every output position comes from (rwkvtts_b200/batch.py); this
// kernel then writes every row of the padded [rows, D] batch exactly once: the source row of its table, or zeros for a
// padding position (row_src < 0), so there is no memset and no scatter.  One warp per output row, 16-byte accesses.
#include "wkv7_common.cuh"

namespace rwkvtts {

constexpr int kMaxTables = 8;
struct GatherTables { const bf16 *tab[kMaxTables]; };
constexpr int kTableShift = 40;      // row_src = (table << 40) | row

__global__ void __launch_bounds__(256) embed_rows_kernel(GatherTables tabs, const long long *__restrict__ row_src,
                                                         long long rows, int D, bf16 *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int nvec = D >> 3;                                 // uint4 pieces per row
    for (long long r = warp; r < rows; r += nwarps) {
        const long long s = row_src[r];
        uint4 *dst = reinterpret_cast<uint4 *>(out + r * D);
        if (s < 0) {
            for (int i = lane; i < nvec; i += 32) dst[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            const bf16 *src_row = tabs.tab[s >> kTableShift] + (s & ((1ll << kTableShift) - 1)) * D;
            const uint4 *src = reinterpret_cast<const uint4 *>(src_row);
            for (int i = lane; i < nvec; i += 32) dst[i] = ldg_nc_v4(src + i);
        }
    }
}

cudaError_t launch_embed_rows(const void *const *tables, int ntab, const long long *row_src, long long rows, int D,
                              void *out, cudaStream_t st) {
    GatherTables t{};
    for (int i = 0; i < ntab; i++) t.tab[i] = static_cast<const bf16 *>(tables[i]);
    long long blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    count_launch();
    embed_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(t, row_src, rows, D, static_cast<bf16 *>(out));
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Gradients into the flat ZeRO buffer, one launch per bucket  [HBM roofline: every gradient read once, written once].
// The reference's DeepSpeed engine copies every parameter's gradient into its contiguous gradient buffer as it arrives
// (`contiguous_gradients: True`, train_scripts/train_spark_rwkv7speech.py:483-516): one small kernel per parameter, ~1000
// per step for the 0.4B model.  Here autograd hands the engine the gradient tensors themselves; when a bucket's last one
// has arrived, this kernel moves (or accumulates, for micro-steps after the first) all of them into their places in the
// flat buffer: blockIdx.y = tensor, blockIdx.x strides over its 16-byte pieces.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxCopy = 128;
struct CopyTable {
    const void *src[kMaxCopy];
    long long dst_off[kMaxCopy];       // elements into the flat buffer (multiple of 8)
    int n[kMaxCopy];                   // elements; negative = accumulate
};
template <typename T>
__global__ void __launch_bounds__(256) multi_copy_kernel(const CopyTable tab, T *__restrict__ flat) {
    constexpr int V = 16 / sizeof(T);
    const int e = blockIdx.y;
    const bool accum = tab.n[e] < 0;
    const long long n = accum ? -(long long)tab.n[e] : tab.n[e];
    const T *src = static_cast<const T *>(tab.src[e]);
    T *dst = flat + tab.dst_off[e];
    const long long nv = n / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        uint4 v = ldg_nc_v4(reinterpret_cast<const uint4 *>(src) + i);
        if (accum) {
            const uint4 o = reinterpret_cast<const uint4 *>(dst)[i];
            if constexpr (sizeof(T) == 2) {
                float a[8], b[8];
                unpack8(v, a); unpack8(o, b);
                v.x = pack2(a[0] + b[0], a[1] + b[1]); v.y = pack2(a[2] + b[2], a[3] + b[3]);
                v.z = pack2(a[4] + b[4], a[5] + b[5]); v.w = pack2(a[6] + b[6], a[7] + b[7]);
            } else {
                v.x = __float_as_uint(__uint_as_float(v.x) + __uint_as_float(o.x)); v.y = __float_as_uint(__uint_as_float(v.y) + __uint_as_float(o.y));
                v.z = __float_as_uint(__uint_as_float(v.z) + __uint_as_float(o.z)); v.w = __float_as_uint(__uint_as_float(v.w) + __uint_as_float(o.w));
            }
        }
        reinterpret_cast<uint4 *>(dst)[i] = v;
    }
    if (blockIdx.x == 0) {                                   // tail of a tensor whose size is not a multiple of 16 bytes
        for (long long i = nv * V + threadIdx.x; i < n; i += blockDim.x) {
            if constexpr (sizeof(T) == 2) dst[i] = accum ? __float2bfloat16_rn(__bfloat162float(dst[i]) + __bfloat162float(src[i])) : src[i];
            else dst[i] = accum ? dst[i] + src[i] : src[i];
        }
    }
}

cudaError_t launch_multi_copy(const void *const *srcs, const long long *dst_off, const long long *n, const int *accumulate,
                              int count, void *flat, int elem_bytes, cudaStream_t st) {
    for (int base = 0; base < count; base += kMaxCopy) {
        CopyTable t{};
        const int m = count - base < kMaxCopy ? count - base : kMaxCopy;
        long long biggest = 0;
        for (int i = 0; i < m; i++) {
            t.src[i] = srcs[base + i];
            t.dst_off[i] = dst_off[base + i];
            t.n[i] = (int)(accumulate != nullptr && accumulate[base + i] ? -n[base + i] : n[base + i]);
            if (n[base + i] > biggest) biggest = n[base + i];
        }
        long long gx = (biggest / (16 / elem_bytes) + 255) / 256;
        if (gx > 64) gx = 64;
        if (gx < 1) gx = 1;
        count_launch();
        if (elem_bytes == 2) multi_copy_kernel<bf16><<<dim3((unsigned)gx, (unsigned)m), 256, 0, st>>>(t, static_cast<bf16 *>(flat));
        else multi_copy_kernel<float><<<dim3((unsigned)gx, (unsigned)m), 256, 0, st>>>(t, static_cast<float *>(flat));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace rwkvtts
