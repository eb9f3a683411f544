// Device radix sort of 32-bit global genome coordinates (index build of rcp_reads_load).
//
// Round-1 implementation: the LSD radix sort of CUB (header templates shipped with the CUDA
// toolkit, instantiated here for sm_100a).  The sort is index-build overhead, NOT part of the
// algorithmic bytes of the coverage path; DESIGN.md lists its replacement by a
// hand-written onesweep pass as follow-up work.
#include <cub/device/device_radix_sort.cuh>

#include "rcp_internal.cuh"

namespace rcp {

int sort_keys_u32(uint32_t* keys, int64_t n, int end_bit) {
    if (n <= 1) return RCP_OK;
    if (n > 0x7fffffff) return fail(RCP_ERR_UNSUPPORTED, "more than 2^31-1 reads in one sort");
    uint32_t* alt = nullptr;
    RCP_TRY(dalloc(&alt, (size_t)n));
    cub::DoubleBuffer<uint32_t> buf(keys, alt);
    size_t tmp_bytes = 0;
    RCP_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, buf, (int)n, 0, end_bit,
                                            g_ctx.stream));
    uint8_t* tmp = nullptr;
    RCP_TRY(dalloc(&tmp, tmp_bytes));
    RCP_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, buf, (int)n, 0, end_bit,
                                            g_ctx.stream));
    g_ctx.launches += 1 + (end_bit + 7) / 8;
    if (buf.Current() != keys) {
        RCP_CUDA(cudaMemcpyAsync(keys, buf.Current(), (size_t)n * sizeof(uint32_t),
                                 cudaMemcpyDeviceToDevice, g_ctx.stream));
    }
    dfree(tmp);
    dfree(alt);
    return RCP_OK;
}

int sort_pairs_u32(uint32_t* keys, uint32_t* vals, int64_t n, int end_bit) {
    if (n <= 1) return RCP_OK;
    if (n > 0x7fffffff) return fail(RCP_ERR_UNSUPPORTED, "more than 2^31-1 reads in one sort");
    uint32_t *kalt = nullptr, *valt = nullptr;
    RCP_TRY(dalloc(&kalt, (size_t)n));
    RCP_TRY(dalloc(&valt, (size_t)n));
    cub::DoubleBuffer<uint32_t> kb(keys, kalt), vb(vals, valt);
    size_t tmp_bytes = 0;
    RCP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, end_bit,
                                             g_ctx.stream));
    uint8_t* tmp = nullptr;
    RCP_TRY(dalloc(&tmp, tmp_bytes));
    RCP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int)n, 0, end_bit,
                                             g_ctx.stream));
    g_ctx.launches += 1 + (end_bit + 7) / 8;
    if (kb.Current() != keys)
        RCP_CUDA(cudaMemcpyAsync(keys, kb.Current(), (size_t)n * sizeof(uint32_t),
                                 cudaMemcpyDeviceToDevice, g_ctx.stream));
    if (vb.Current() != vals)
        RCP_CUDA(cudaMemcpyAsync(vals, vb.Current(), (size_t)n * sizeof(uint32_t),
                                 cudaMemcpyDeviceToDevice, g_ctx.stream));
    dfree(tmp);
    dfree(kalt);
    dfree(valt);
    return RCP_OK;
}

}  // namespace rcp
