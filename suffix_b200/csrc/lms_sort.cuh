// lms_sort.cuh -- direct sort of the LMS suffixes ("LMS substring bucket sort" +
// "rank/rename" of the north-star pipeline, fused into one radix sort with
// refinement): replaces, for texts whose LMS suffixes are told apart by a few
// windows of characters, the reference's stage-1 machinery
//   P5  LMS into bucket tails            src/table.rs:411-416
//   P7/P9 first L/S induce               src/table.rs:421-448
//   P10 compaction, P11 naming, P12      src/table.rs:450-492 (wstring_equal :802-820)
//   P13 recursion on the reduced string  src/table.rs:494-506
//   P15/P16 un-rename                    src/table.rs:512-530
// whose only product is "the LMS suffixes in suffix order" (the seed of the final
// induce, P18, src/table.rs:536-541).
//
// Formulation.  A window key packs the next KC characters of a suffix, first
// character most significant, as a base-sigma number of dense order-preserving
// codes (sigma^KC <= 2^32; for 2-bit packed text simply the 16 two-bit codes), with
// zeros past the end of the text.
//   round 1: one-sweep LSD radix sort of (window(p), p) over all LMS positions;
//   round r: the members of groups that are still tied after (r-1)*KC characters are
//            ordered inside their group by the next window (MSD refinement; tiny
//            groups by counting, otherwise radix on (group, window)).
// End of text (src/table.rs:374: a proper prefix sorts first, there is no sentinel):
// the first sort is fed in DESCENDING text position and every later sort is stable,
// so inside a group of equal padded keys the order is always descending position.
// A member whose window runs past n ("truncated", p + h + KC > n) is a proper prefix
// of every non-truncated member with the same padded key and of every truncated
// member before which it stands, so truncated members are final exactly where the
// stable sort leaves them and each forms a group of its own.
// Texts that stay tied (long repeats: LCP >> KC * rounds) are handed to the robust
// path (stage-1 induce + naming + rank doubling); see lms_direct_sort() in b200sa.cu.
#pragma once
#include "classify.cuh"

namespace b200sa {

struct LmsWin {
    const void *ptext;            // packed words (BITS 2/4) or the byte text (BITS 8)
    const uint32_t *code_of;      // [256] dense codes (BITS 8, sigma < 256)
    uint32_t n, sigma, kc;        // kc = characters per window
};

// pair-reversal of 16 two-bit groups: char 0 (low bits) becomes the most significant
__device__ __forceinline__ uint32_t rev_pairs(uint32_t x) {
    uint32_t y = __brev(x);
    return ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);
}

// window key of text[p .. p+kc), zero-padded past n;  p <= n
template <int BITS>
__device__ __forceinline__ uint32_t lms_window(const LmsWin &W, uint32_t p) {
    const uint32_t left = W.n - p;                 // valid characters from p on
    if (BITS == 2) {
        if (left == 0) return 0u;
        uint32_t x = text_bits<2>(W.ptext, p);
        if (left < 16u) x &= (1u << (2u * left)) - 1u;
        return rev_pairs(x);
    } else if (BITS == 4) {
        if (left == 0) return 0u;
        uint32_t lo = text_bits<4>(W.ptext, p);
        uint32_t hi = (W.kc > 8u && left > 8u) ? text_bits<4>(W.ptext, p + 8u) : 0u;
        uint32_t key = 0;
        for (uint32_t i = 0; i < W.kc; i++) {
            uint32_t c = (i < 8u ? (lo >> (4u * i)) : (hi >> (4u * (i - 8u)))) & 15u;
            key = key * W.sigma + (i < left ? c : 0u);
        }
        return key;
    } else {
        const uint8_t *t = reinterpret_cast<const uint8_t *>(W.ptext);
        uint32_t key = 0;
        if (W.sigma == 256u) {
#pragma unroll
            for (uint32_t i = 0; i < 4u; i++) key = (key << 8) | (i < left ? (uint32_t)__ldg(t + p + i) : 0u);
            return key;
        }
        for (uint32_t i = 0; i < W.kc; i++) {
            uint32_t c = i < left ? __ldg(W.code_of + __ldg(t + p + i)) : 0u;
            key = key * W.sigma + c;
        }
        return key;
    }
}

// first-pass functors of the one-sweep sort: item i = i-th LMS position from the END of the
// text (the fused classifier emits them in that order)
template <int BITS>
struct LmsKeyDesc {
    LmsWin W; const uint32_t *lmsdesc;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return lms_window<BITS>(W, __ldg(lmsdesc + i)); }
    __device__ __forceinline__ uint32_t at(uint64_t, uint32_t pos) const { return lms_window<BITS>(W, pos); }   // value = position
};
struct LmsValDesc {
    const uint32_t *lmsdesc;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return __ldg(lmsdesc + i); }
};

// ---- round 1: groups of equal window keys in the sorted list (slot = index).
// Only LMS positions inside the last `span` characters can be truncated (at most span/2
// of them, the first entries of the descending list).  One small kernel finds their slots -- they
// stand at the very start of their run of equal keys, in descending position -- and
// sets "forced head" bits for the slot and its successor, so that the scan over all m
// elements reads the keys only.
template <int BITS>
__global__ void __launch_bounds__(BLK) k_lms_mark_trunc(LmsWin W, const uint32_t *__restrict__ lmsdesc, uint32_t m,
                                                        const uint32_t *__restrict__ K, const uint32_t *__restrict__ P,
                                                        uint32_t span, uint32_t *forced) {
    uint32_t t = threadIdx.x;
    if (t >= m || t >= span) return;
    uint32_t p = lmsdesc[t];
    if ((uint64_t)p + span <= W.n) return;                // not truncated
    uint32_t key = lms_window<BITS>(W, p);
    uint32_t lo = 0, hi = m;
    while (lo < hi) {                                     // first slot with K >= key
        uint32_t mid = lo + (hi - lo) / 2;
        if (K[mid] < key) lo = mid + 1; else hi = mid;
    }
    for (uint32_t j = lo; j < m && K[j] == key; j++) {
        if (P[j] == p) {
            atomicOr(&forced[j >> 5], 1u << (j & 31));
            if (j + 1 < m) atomicOr(&forced[(j + 1) >> 5], 1u << ((j + 1) & 31));
            break;
        }
    }
}
struct InLmsActive1 {
    const uint32_t *K, *forced; uint32_t m;
    // branch-free: all five loads of an element are independent (a short-circuit chain would
    // serialise them behind branches; measured 3x slower)
    __device__ __forceinline__ bool head(uint32_t i) const {
        uint32_t a = K[i], b = K[i > 0 ? i - 1 : 0], f = forced[i >> 5];
        return (i == 0) | (a != b) | ((f >> (i & 31)) & 1u);
    }
    __device__ __forceinline__ uint32_t operator()(uint64_t ii) const {
        uint32_t i = (uint32_t)ii, j = i + 1 < m ? i + 1 : i;
        uint32_t a = K[i], b = K[i > 0 ? i - 1 : 0], c = K[j], f = forced[i >> 5], f2 = forced[j >> 5];
        bool hd = (i == 0) | (a != b) | ((f >> (i & 31)) & 1u);
        bool tl = (i + 1 == m) | (c != a) | ((f2 >> (j & 31)) & 1u);
        return (hd & tl) ? 0u : 1u;
    }
};
// compaction of the tied elements; ahead[k] = slot if the element starts its group, else 0
// (a max-scan over the compacted list turns it into the group id)
struct OutLmsCompact1 {
    InLmsActive1 in; const uint32_t *P; uint32_t *aslot, *apos, *ahead;
    __device__ void operator()(uint64_t i, uint32_t exc, uint32_t v) const {
        if (v) { aslot[exc] = (uint32_t)i; apos[exc] = P[i]; ahead[exc] = in.head((uint32_t)i) ? (uint32_t)i : 0u; }
    }
};
// The same compaction as ONE specialised single-pass kernel (the generic scan spends ~90
// instructions per element on functor calls and 16 warp scans per thread; here a thread owns 16
// consecutive slots, reads its 18 keys with four 16-byte loads + 2, derives heads / tied flags in
// registers and takes part in one block scan; tile offsets by the same look-back as k_scan_lb).
constexpr int LG_IPT = 16;
constexpr int LG_TILE = BLK * LG_IPT;       // 4096 slots per tile
__global__ void __launch_bounds__(BLK) k_lms_groups1(const uint32_t *__restrict__ K, const uint32_t *__restrict__ P,
                                                     const uint32_t *__restrict__ forced, uint32_t m, uint32_t ntiles,
                                                     ScanState S, uint32_t *aslot, uint32_t *apos, uint32_t *ahead,
                                                     uint32_t *d_total) {
    __shared__ uint32_t s_w[NWARP + 1];
    __shared__ uint32_t s_tile, s_prefix;
    if (threadIdx.x == 0) {
        uint32_t t = atomicAdd(S.ticket, 1u);
        if (t + 1 == ntiles) *S.ticket = 0u;
        s_tile = t;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t i0 = tile * LG_TILE + threadIdx.x * LG_IPT;
    uint32_t k[LG_IPT + 2];                      // k[j+1] = K[i0 + j]; k[0] = K[i0-1]; k[17] = K[i0+16]
    if (i0 + LG_IPT <= m) {
        const uint4 *q = reinterpret_cast<const uint4 *>(K + i0);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            uint4 a = __ldg(q + v);
            k[1 + 4 * v] = a.x; k[2 + 4 * v] = a.y; k[3 + 4 * v] = a.z; k[4 + 4 * v] = a.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < LG_IPT; j++) k[1 + j] = (i0 + j < m) ? __ldg(K + i0 + j) : 0u;
    }
    k[0] = (i0 > 0 && i0 <= m) ? __ldg(K + i0 - 1) : 0u;
    k[LG_IPT + 1] = (i0 + LG_IPT < m) ? __ldg(K + i0 + LG_IPT) : 0u;
    // forced-head bits of slots i0 .. i0+16 (i0 is a multiple of 16)
    uint32_t fw = (i0 < m) ? __ldg(forced + (i0 >> 5)) : 0u;
    uint32_t fbits = (fw >> (i0 & 31)) & 0xffffu;
    if (i0 + LG_IPT < m) {
        uint32_t nxt = ((i0 & 31) == 16) ? __ldg(forced + (i0 >> 5) + 1) : (fw >> 16);
        fbits |= (nxt & 1u) << 16;
    }
    uint32_t head = 0;                           // bit j: slot i0+j starts a group (bit 16: the slot after the chunk)
#pragma unroll
    for (int j = 0; j <= LG_IPT; j++) {
        uint32_t i = i0 + j;
        bool h = (i == 0) | (i >= m) | (k[j + 1] != k[j]) | ((fbits >> j) & 1u);
        head |= (h ? 1u : 0u) << j;
    }
    uint32_t act = 0;                            // bit j: slot i0+j is tied with a neighbour
#pragma unroll
    for (int j = 0; j < LG_IPT; j++) {
        bool a = (i0 + j < m) && !(((head >> j) & 1u) && ((head >> (j + 1)) & 1u));
        act |= (a ? 1u : 0u) << j;
    }
    uint32_t cnt = __popc(act), btot;
    uint32_t inc = block_incl_scan<OpSum>(cnt, s_w, &btot);
    if (warp_id() == 0) {
        if (lane_id() == 0) tile_publish_u32(S, tile, btot);
        uint32_t prefix = tile_walk_u32(S, tile, btot, tile + 1 == ntiles, d_total);
        if (lane_id() == 0) s_prefix = prefix;
    }
    __syncthreads();
    uint32_t at = s_prefix + inc - cnt;
    while (act) {
        uint32_t j = __ffs(act) - 1;
        act &= act - 1;
        uint32_t i = i0 + j;
        aslot[at] = i; apos[at] = __ldg(P + i); ahead[at] = ((head >> j) & 1u) ? i : 0u;
        at++;
    }
}

struct OutMaxInPlace {
    uint32_t *a;
    __device__ void operator()(uint64_t i, uint32_t exc, uint32_t v) const { a[i] = exc > v ? exc : v; }
};

// ---- round r >= 2 over the compacted active list (slot order; groups contiguous)
template <int BITS>
__global__ void __launch_bounds__(BLK) k_lms_refine_keys(LmsWin W, const uint32_t *__restrict__ apos,
                                                         const uint32_t *__restrict__ agrp, uint32_t na, uint32_t h,
                                                         uint64_t *keys) {
    uint32_t j = blockIdx.x * BLK + threadIdx.x;
    if (j >= na) return;
    keys[j] = ((uint64_t)agrp[j] << 32) | lms_window<BITS>(W, apos[j] + h);    // apos + h <= n (not truncated before)
}
struct InLmsGroupR {
    const uint64_t *K; const uint32_t *P, *slot; uint32_t na, n, span;
    __device__ __forceinline__ bool trunc(uint32_t p) const { return (uint64_t)p + span > n; }
    __device__ __forceinline__ bool head(uint32_t j) const {
        return j == 0 || K[j] != K[j - 1] || trunc(P[j - 1]) || trunc(P[j]);
    }
    __device__ unsigned long long operator()(uint64_t jj) const {
        uint32_t j = (uint32_t)jj;
        bool hd = head(j), tl = (j + 1 == na) || head(j + 1);
        return ((unsigned long long)(hd ? slot[j] : 0u) << 32) | ((hd && tl) ? 0u : 1u);
    }
};
struct OutLmsCompactR {
    const uint32_t *P, *slot; uint32_t *list; uint32_t *oslot, *opos, *ogrp;
    __device__ void operator()(uint64_t j, unsigned long long exc, unsigned long long v) const {
        uint32_t p = P[j], sl = slot[j];
        list[sl] = p;                                   // position inside the group is final for this depth
        if ((uint32_t)v) {
            uint32_t eh = (uint32_t)(exc >> 32), vh = (uint32_t)(v >> 32), k = (uint32_t)exc;
            oslot[k] = sl; opos[k] = p; ogrp[k] = eh > vh ? eh : vh;
        }
    }
};

}  // namespace b200sa
