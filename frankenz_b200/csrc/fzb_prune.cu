// Ordering of the pass-2 object list of the tensor-core sweep (fzb_sweep_tc.cuh).
//
// Pass 2 skips an 8-model sub-batch when none of the 32 objects of a warp had a weight above the running cut there in
// pass 1 (bit map `live`, [row = model tile x half][object], 16 bits per entry).  The skip is only as good as the
// objects of a warp are alike, so the list is sorted by a key made of the object's live fraction (16 levels) and a
// 24-bit signature (which 24ths of the model sequence - models are ordered by redshift bin - hold live sub-batches).
// Measured on the C3 workload: 51 % of the sub-batches are live per object, 66-69 % per warp after this sort, 100 %
// in arrival order.
#include <cub/device/device_radix_sort.cuh>

#include "fzb_common.cuh"

namespace {

__global__ void k_live_key(const unsigned short* __restrict__ live, int64_t nrows, int64_t No_pad,
                           const int32_t* __restrict__ list, int64_t n, uint32_t* __restrict__ keys, int bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = list[i];
    uint32_t sig = 0;
    int64_t total = 0;
    for (int seg = 0; seg < 24; ++seg) {
        const int64_t r0 = nrows * seg / 24, r1 = nrows * (seg + 1) / 24;
        int c = 0;
        for (int64_t r = r0; r < r1; ++r) c += __popc((unsigned)live[(size_t)r * No_pad + o]);
        total += c;
        if ((int64_t)c * 50 > (r1 - r0) * bits) sig |= 1u << seg;     // more than 2 % of the segment's bits
    }
    const int64_t level = min((int64_t)15, total * 16 / max((int64_t)1, nrows * bits));
    keys[i] = ((uint32_t)level << 24) | sig;
}

// live bits of every (model tile, pass-2 CTA): OR over the 128 list objects of each M-tile and both halves; one warp per entry
__global__ void k_tile_masks(const unsigned short* __restrict__ live, int64_t ntiles, int64_t No_pad, const int32_t* __restrict__ list,
                             int64_t n, int64_t ncta, int tile_objs, unsigned int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= ntiles * ncta) return;
    const int64_t t = w / ncta, c = w - t * ncta;
    const unsigned short* r0 = live + (size_t)(2 * t) * No_pad;
    const unsigned short* r1 = r0 + No_pad;
    unsigned int m0 = 0, m1 = 0;
    for (int i = lane; i < tile_objs; i += 32) {
        const int64_t idx = c * tile_objs + i;
        if (idx < n) {
            const int64_t o = list[idx];
            const unsigned int v = (unsigned int)r0[o] | (unsigned int)r1[o];
            if (i < tile_objs / 2) m0 |= v; else m1 |= v;
        }
    }
    m0 = __reduce_or_sync(0xffffffffu, m0);
    m1 = __reduce_or_sync(0xffffffffu, m1);
    if (lane == 0) out[w] = m0 | (m1 << 16);
}

}  // namespace

// masks for the pass-2 launch over `list` (n objects, tile_objs per CTA, two M-tiles per CTA, two halves per tile row pair)
int fzb_tile_masks(fzb_context* h, const unsigned short* live, int64_t ntiles, int64_t No_pad, const int32_t* list, int64_t n,
                   int tile_objs, unsigned int** out) {
    const int64_t ncta = (n + tile_objs - 1) / tile_objs;
    if (h->fast.tmask.reserve((size_t)ntiles * ncta * 4 + 64)) return 1;
    const int64_t warps = ntiles * ncta;
    k_tile_masks<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, h->stream>>>(live, ntiles, No_pad, list, n, ncta, tile_objs,
                                                                           h->fast.tmask.as<unsigned int>());
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    *out = h->fast.tmask.as<unsigned int>();
    return 0;
}

int fzb_sort_by_live_bits(fzb_context* h, const unsigned short* live, int64_t nrows, int64_t No_pad, int32_t* list,
                          int64_t n, int bits_per_row) {
    if (n <= 32) return 0;
    DevBuf& sb = h->fast.sortbuf;
    size_t tmp_bytes = 0;
    uint32_t* keys_in = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_in, list, list, (int)n, 0, 28, h->stream);
    const size_t need = (size_t)n * 4 * 3 + tmp_bytes + 1024;
    if (sb.reserve(need)) return 1;
    keys_in = sb.as<uint32_t>();
    uint32_t* keys_out = keys_in + n;
    int32_t* vals_out = reinterpret_cast<int32_t*>(keys_out + n);
    void* tmp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(vals_out + n) + 255) & ~(uintptr_t)255);
    k_live_key<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(live, nrows, No_pad, list, n, keys_in, bits_per_row);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    FZB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, list, vals_out, (int)n, 0, 28, h->stream));
    fzb_count_launch(h, 4);
    FZB_CUDA(cudaMemcpyAsync(list, vals_out, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}
