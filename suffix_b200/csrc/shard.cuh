// shard.cuh -- kernels of the multi-GPU LMS-suffix sort (BASELINE config 5):
// the text is cut into one contiguous shard per GPU; classification runs per shard
// (classify2.cuh with halo chars), then the LMS suffixes of ALL shards are sorted by a
// 64-bit window key with one exchange step:
//   window keys -> sampled splitters (all-gather) -> stable partition by destination rank
//   -> all-to-all (grouped ncclSend / ncclRecv) -> local radix sort -> groups / names.
// Replaces, for a sharded text, the reference's LMS placement + first induce + naming
// (src/table.rs:411-416, :421-482, wstring_equal :802-820): the product is the order of
// the LMS suffixes by their first KC characters and dense names of the distinct windows.
// Equal windows (ties) are counted; with ties == 0 the order is the exact suffix order.
#pragma once
#include "lms_sort.cuh"

namespace b200sa {

struct ShardWin {
    const uint8_t *text;     // this rank's shard
    const uint8_t *halo;     // the bytes that follow the shard in the whole text (hl of them)
    const uint32_t *code_of; // [256] global dense codes
    uint64_t n_local, hl;
    uint32_t sigma, kc;
};

// 64-bit window key of the suffix at local position p: kc chars, base sigma, first char most
// significant, zeros past the end of the WHOLE text (= past shard + halo)
__device__ __forceinline__ uint64_t shard_window(const ShardWin &W, uint32_t p) {
    uint64_t key = 0;
    for (uint32_t j = 0; j < W.kc; j++) {
        uint64_t q = (uint64_t)p + j;
        uint32_t c = 0;
        if (q < W.n_local) c = __ldg(W.code_of + __ldg(W.text + q));
        else if (q - W.n_local < W.hl) c = __ldg(W.code_of + __ldg(W.halo + (q - W.n_local)));
        key = key * W.sigma + c;
    }
    return key;
}
__global__ void __launch_bounds__(BLK) k_shard_keys(ShardWin W, const uint32_t *__restrict__ lmsdesc, uint32_t m,
                                                    uint64_t *keys, uint32_t *vals) {
    uint32_t i = blockIdx.x * BLK + threadIdx.x;
    if (i >= m) return;
    uint32_t p = lmsdesc[i];
    keys[i] = shard_window(W, p);
    vals[i] = p;
}
__global__ void __launch_bounds__(BLK) k_hist_to_u64(const uint32_t *h32, unsigned long long *h64, uint32_t k) {
    uint32_t i = blockIdx.x * BLK + threadIdx.x;
    if (i < k) h64[i] = h32[i];
}
// dense order-preserving codes from the all-reduced 768-bin histogram (L | S | LMS per byte)
__global__ void __launch_bounds__(BLK) k_alpha_from_hist64(const unsigned long long *h64, uint32_t *code_of, uint32_t *sigma) {
    __shared__ uint32_t s_w[NWARP + 1];
    uint32_t c = threadIdx.x, total;
    uint32_t present = (h64[c] + h64[256 + c] + h64[512 + c]) > 0 ? 1u : 0u;
    uint32_t inc = block_incl_scan<OpSum>(present, s_w, &total);
    code_of[c] = inc - present;
    if (c == 0) *sigma = total;
}
// every rank takes `per` evenly spaced keys of its (unsorted) list; missing ones are ~0
__global__ void __launch_bounds__(BLK) k_shard_sample(const uint64_t *__restrict__ keys, uint32_t m, uint32_t per, uint64_t *out) {
    uint32_t i = blockIdx.x * BLK + threadIdx.x;
    if (i >= per) return;
    out[i] = m ? keys[(uint64_t)i * m / per] : ~0ull;
}
// splitters: N-1 evenly spaced elements of the sorted sample
__global__ void k_shard_splitters(const uint64_t *__restrict__ sorted, uint32_t total, uint32_t nranks, uint64_t *split) {
    uint32_t i = threadIdx.x;
    if (i + 1 < nranks) split[i] = sorted[(uint64_t)(i + 1) * total / nranks];
}
// destination rank = number of splitters <= key (equal keys always share a destination)
__global__ void __launch_bounds__(BLK) k_shard_dest(const uint64_t *__restrict__ keys, uint32_t m, const uint64_t *__restrict__ split,
                                                    uint32_t nranks, uint8_t *dest, unsigned long long *counts) {
    __shared__ uint32_t s_cnt[16];
    if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t i = blockIdx.x * BLK + threadIdx.x;
    if (i < m) {
        uint64_t k = keys[i];
        uint32_t d = 0;
        for (uint32_t j = 0; j + 1 < nranks; j++) d += (split[j] <= k) ? 1u : 0u;
        dest[i] = (uint8_t)d;
        atomicAdd(&s_cnt[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < nranks && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}
struct DigU8 {
    const uint8_t *d;
    __device__ uint32_t operator()(uint64_t i) const { return d[i]; }
};
struct MoveKV64 {
    const uint64_t *kin; const uint32_t *vin; uint64_t *kout; uint32_t *vout;
    __device__ void operator()(uint64_t i, uint32_t dst) const { kout[dst] = kin[i]; vout[dst] = vin[i]; }
};
// global position of received item j: chunk (= source rank) by its offset, then lo[src] + local position
__global__ void __launch_bounds__(BLK) k_shard_gpos(const uint32_t *__restrict__ idx, const uint32_t *__restrict__ rpos,
                                                    uint32_t cnt, const unsigned long long *__restrict__ chunk_off,
                                                    const unsigned long long *__restrict__ chunk_lo, uint32_t nranks,
                                                    unsigned long long *gpos) {
    uint32_t i = blockIdx.x * BLK + threadIdx.x;
    if (i >= cnt) return;
    uint32_t j = idx[i];
    uint32_t c = 0;
    while (c + 1 < nranks && chunk_off[c + 1] <= j) c++;
    gpos[i] = chunk_lo[c] + rpos[j];
}
// group heads in the sorted slice: key differs, or either neighbour's window runs past the
// end of the whole text (lms_sort.cuh: truncated members are final where the stable sort
// leaves them)
struct InShardHead {
    const uint64_t *K; const unsigned long long *gpos; uint32_t cnt; uint64_t n_total; uint32_t kc;
    __device__ __forceinline__ bool trunc(uint32_t i) const { return gpos[i] + kc > n_total; }
    __device__ __forceinline__ bool head(uint32_t i) const {
        uint32_t h = i > 0 ? i - 1 : 0;
        return (i == 0) | (K[i] != K[h]) | trunc(h) | trunc(i);
    }
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return head((uint32_t)i) ? 1u : 0u; }
};
struct OutShardName {
    InShardHead in; uint32_t name_base; uint32_t *names; unsigned long long *ties;
    __device__ void operator()(uint64_t ii, uint32_t exc, uint32_t v) const {
        uint32_t i = (uint32_t)ii;
        names[i] = name_base + exc + v - 1u;
        bool tail = (i + 1 == in.cnt) || in.head(i + 1);
        if (!(v && tail)) atomicAdd(ties, 1ull);
    }
};

__global__ void __launch_bounds__(BLK) k_add_u32(uint32_t *a, uint32_t n, uint32_t add) {
    uint32_t i = blockIdx.x * BLK + threadIdx.x;
    if (i < n) a[i] += add;
}

}  // namespace b200sa
