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

}  // namespace rwkvtts
