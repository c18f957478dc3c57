// induce3.cuh -- K4/K5 for 2-bit packed text (sigma <= 4, the DNA headline config): the
// one-round bucket steps of induce.cuh (reference src/table.rs:421-448, :543-573,
// head_insert/tail_insert :723-736) with a ranking that needs no MATCH and no
// shared-memory counters:
//   * tiles are cut in PHYSICAL index space at multiples of TILE, so thread t owns 8
//     consecutive, 32-byte aligned entries (two 16-byte loads, one 8-byte pred access);
//     the logical (scan) order is thread-major, ascending or descending;
//   * with at most 4 destination buckets the per-thread counts of all buckets fit ONE
//     64-bit word (4 fields of 16 bits; a tile has 2048 entries), so one block-wide
//     add-scan of that word ranks the whole tile stably for every bucket at once;
//   * the count matrix of a step is G x 4 words instead of G x 256.
// Control flow (peek, big / small steps, run skipping, fill accounting, invariants) is the
// one of induce.cuh.
#pragma once
#include "induce.cuh"

namespace b200sa {

struct Ind3Shared {
    unsigned long long wtot[2][NWARP];   // per-warp packed totals, double buffered over tiles
    uint32_t cbase[4];                   // running destination position per code (inside the bucket part)
    uint32_t ctot[4];                    // step totals per code
    uint32_t code_of[256];
    uint32_t nsig;
};

// Physical chunk of thread t in tile T of a segment: 8 consecutive entries in PHYSICAL order
// (no data movement after the load: the registers are consumed an iteration later, so the
// loads stay in flight under the work on the previous tile).  Logical (scan) order is
// ascending physical index in the L pass and descending in the S pass.
__device__ __forceinline__ void ind3_load8(const uint32_t *__restrict__ src, uint32_t chunk, uint32_t limit,
                                           uint32_t (&raw)[ITEMS]) {
    if (chunk + ITEMS <= limit) {
        const uint4 *q = reinterpret_cast<const uint4 *>(src + chunk);
        uint4 a = __ldcg(q), b = __ldcg(q + 1);
        raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
        raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) raw[i] = (chunk + i < limit) ? __ldcg(src + chunk + i) : 0u;
    }
}

struct Seg3 {
    const uint32_t *src;
    uint8_t *pred;
    uint32_t pa, pb;      // physical range [pa, pb)
    uint32_t limit;       // entries in the underlying array (no access at or beyond)
    uint32_t lo, hi;      // valid destination CODES (inclusive)
    int rev;
};

template <bool SPASS>
__global__ void __launch_bounds__(BLK, 3) k_induce3(InduceArgs A) {
    __shared__ IndShared sh;
    __shared__ Ind3Shared s3;
    cg::grid_group grid = cg::this_grid();
    const uint32_t G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
    const uint32_t w = warp_id(), l = lane_id();

    // ---- init: tables, fill counters, seed (suffix n-1 is L: src/table.rs:422-425)
    sh.bstart[tid] = A.bstart[tid];
    if (tid == 0) sh.bstart[256] = A.bstart[256];
    sh.Lcnt[tid] = A.Lcnt[tid];
    if (SPASS) sh.S_or_lmsoff[tid] = A.Scnt[tid];
    else { sh.S_or_lmsoff[tid] = A.lms_off[tid]; if (tid == 0) sh.S_or_lmsoff[256] = A.lms_off[256]; }
    if (tid < 16) sh.alpha[tid] = A.alpha[tid];
    uint32_t lastc = A.text[A.n - 1];
    sh.fill[tid] = (!SPASS && tid == lastc) ? 1u : 0u;
    if (tid == 0) { sh.st_c = SPASS ? 255 : 0; sh.st_phase = 0; sh.st_begin = 0; sh.streak = 0; sh.streak_c = -1; }
    {
        uint32_t present = (A.Lcnt[tid] + A.Scnt[tid]) > 0 ? 1u : 0u, total;
        uint32_t inc = block_incl_scan<OpSum>(present, sh.sw, &total);
        s3.code_of[tid] = inc - present;
        if (tid == 0) s3.nsig = total;
    }
    __syncthreads();
    if (!SPASS && bid == 0 && tid == 0) A.sa[sh.bstart[lastc]] = A.n - 1u;
    uint32_t bigcount = 0, smallcount = 0, bigtiles = 0;

    while (true) {
        if (tid == 0) induce_peek<SPASS>(A, sh);
        __syncthreads();
        if (A.steplog && bid == 0 && tid == 0) {          // diagnostics: one record per step of this launch
            unsigned long long now, k = A.steplog[0];
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (k < 2000) { A.steplog[1 + 2 * k] = now; A.steplog[2 + 2 * k] = sh.has ? sh.seg.len : 0u; A.steplog[0] = k + 1; }
        }
        if (!sh.has) break;
        if (sh.seg.len <= (uint32_t)TILE) {
            induce_small_episode<SPASS, 2>(A, sh, grid, smallcount);
            continue;
        }
        // -------------------- big step: all blocks
        Seg3 g;
        {
            const Seg &o = sh.seg;
            g.src = o.src; g.pred = o.pred; g.rev = o.rev;     // o.rev == SPASS for every list
            if (o.rev) { g.pb = o.base + 1u; g.pa = g.pb - o.len; }
            else { g.pa = o.base; g.pb = o.base + o.len; }
            g.limit = (o.src == A.sa) ? A.n : sh.S_or_lmsoff[256];     // L pass second list: the LMS list (m entries)
            // destination range in codes (the byte range [lo, hi] holds only bytes that occur at its ends)
            uint32_t lo = 0, hi = 0;
            const uint32_t ns = s3.nsig;
            // smallest code whose byte >= o.lo; largest code whose byte <= o.hi
            while (lo < ns && sh.alpha[lo] < o.lo) lo++;
            hi = ns;
            while (hi > 0 && sh.alpha[hi - 1] > o.hi) hi--;
            g.lo = lo; g.hi = hi;            // valid codes: lo <= d < hi
        }
        uint32_t *cntbuf = A.blk_cnt + (size_t)(bigcount & 1u) * G * 4u;
        bigcount++;
        const uint32_t T0 = g.pa / TILE, T1 = (g.pb - 1u) / TILE;
        const uint32_t tiles = T1 - T0 + 1u;
        bigtiles += tiles;
        const uint32_t tpb = (tiles + G - 1) / G;
        const uint32_t nact = (tiles + tpb - 1) / tpb;
        uint32_t tb0 = bid * tpb, tb1 = tb0 + tpb;
        if (tb1 > tiles) tb1 = tiles;
        constexpr bool REV = SPASS;                 // every S-pass list is scanned downwards, every L-pass list upwards
        auto chunk_of = [&](uint32_t k) -> uint32_t {           // physical start of this thread's chunk in logical tile k
            uint32_t T = REV ? (T1 - k) : (T0 + k);
            return REV ? (T * TILE + TILE - 8u * (tid + 1u)) : (T * TILE + 8u * tid);
        };
        // ---- phase A: count + remember the predecessor chars (three tiles in flight per thread:
        // entries of tile k+2 loading, chars of tile k+1 being gathered, tile k counted)
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (tb0 < tb1) {
            uint32_t s1[ITEMS], s2[ITEMS], d0[ITEMS], d1[ITEMS];
            auto gather = [&](const uint32_t (&sv)[ITEMS], uint32_t chunk, uint32_t (&dv)[ITEMS]) {
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    bool in = p >= g.pa && p < g.pb && sv[i] > 0;
                    dv[i] = in ? text_get<2>(A.ptext, sv[i] - 1u) : 4u;          // 4 = nothing to induce
                }
            };
            ind3_load8(g.src, chunk_of(tb0), g.limit, s1);
            gather(s1, chunk_of(tb0), d0);
            if (tb0 + 1 < tb1) ind3_load8(g.src, chunk_of(tb0 + 1), g.limit, s1);
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 2 < tb1) ind3_load8(g.src, chunk_of(k + 2), g.limit, s2);
                if (k + 1 < tb1) gather(s1, chunk_of(k + 1), d1);
                const uint32_t chunk = chunk_of(k);
                unsigned long long pk = 0;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t d = d0[i];
                    if (!(d >= g.lo && d < g.hi)) d = 4u;
                    c0 += (d == 0u); c1 += (d == 1u); c2 += (d == 2u); c3 += (d == 3u);
                    pk |= (unsigned long long)d << (8 * i);
                }
                if (chunk + ITEMS <= g.limit) *reinterpret_cast<unsigned long long *>(g.pred + chunk) = pk;
                else {
                    for (int i = 0; i < ITEMS; i++) if (chunk + i < g.limit) g.pred[chunk + i] = (uint8_t)(pk >> (8 * i));
                }
#pragma unroll
                for (int i = 0; i < ITEMS; i++) { s1[i] = s2[i]; d0[i] = d1[i]; }
            }
        }
        {   // block totals of the 4 codes
            c0 = __reduce_add_sync(FULL, c0); c1 = __reduce_add_sync(FULL, c1);
            c2 = __reduce_add_sync(FULL, c2); c3 = __reduce_add_sync(FULL, c3);
            __syncthreads();
            if (l == 0) { sh.wcnt[w][0] = c0; sh.wcnt[w][1] = c1; sh.wcnt[w][2] = c2; sh.wcnt[w][3] = c3; }
            __syncthreads();
            if (tid < 4 && bid < nact) {
                uint32_t v = 0;
#pragma unroll
                for (int ww = 0; ww < NWARP; ww++) v += sh.wcnt[ww][tid];
                cntbuf[(size_t)bid * 4u + tid] = v;
            }
        }
        __syncthreads();
        grid.sync();
        // ---- phase B: offsets (every thread reads whole rows of the G x 4 count matrix), scatter
        {
            uint32_t be[4] = {0, 0, 0, 0}, to[4] = {0, 0, 0, 0};
            for (uint32_t b = tid; b < nact; b += BLK) {
                uint4 v = __ldcg(reinterpret_cast<const uint4 *>(cntbuf) + b);
                to[0] += v.x; to[1] += v.y; to[2] += v.z; to[3] += v.w;
                if (b < bid) { be[0] += v.x; be[1] += v.y; be[2] += v.z; be[3] += v.w; }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) { be[q] = __reduce_add_sync(FULL, be[q]); to[q] = __reduce_add_sync(FULL, to[q]); }
            if (l == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) { sh.wcnt[w][8 + q] = be[q]; sh.wcnt[w][12 + q] = to[q]; }
            }
            __syncthreads();
            if (tid < 4) {
                uint32_t bsum = 0, tsum = 0;
#pragma unroll
                for (int ww = 0; ww < NWARP; ww++) { bsum += sh.wcnt[ww][8 + tid]; tsum += sh.wcnt[ww][12 + tid]; }
                s3.cbase[tid] = sh.fill[sh.alpha[tid]] + bsum;
                s3.ctot[tid] = tsum;
            }
        }
        __syncthreads();
        if (tb0 < tb1) {
            // running destination positions of the 4 codes (registers, identical in all threads)
            uint32_t rb0 = s3.cbase[0], rb1 = s3.cbase[1], rb2 = s3.cbase[2], rb3 = s3.cbase[3];
            const uint32_t bs0 = SPASS ? sh.bstart[sh.alpha[0] + 1] - 1u : sh.bstart[sh.alpha[0]];
            const uint32_t bs1 = SPASS ? sh.bstart[sh.alpha[1] + 1] - 1u : sh.bstart[sh.alpha[1]];
            const uint32_t bs2 = SPASS ? sh.bstart[sh.alpha[2] + 1] - 1u : sh.bstart[sh.alpha[2]];
            const uint32_t bs3 = SPASS ? sh.bstart[sh.alpha[3] + 1] - 1u : sh.bstart[sh.alpha[3]];
            uint32_t s_cur[ITEMS], s_nxt[ITEMS];
            unsigned long long p_cur = 0, p_nxt = 0;
            auto load_pred = [&](uint32_t chunk) -> unsigned long long {
                if (chunk + ITEMS <= g.limit) return __ldcg(reinterpret_cast<const unsigned long long *>(g.pred + chunk));
                unsigned long long v = 0;
                for (int i = 0; i < ITEMS; i++) v |= (unsigned long long)((chunk + i < g.limit) ? g.pred[chunk + i] : 4u) << (8 * i);
                return v;
            };
            ind3_load8(g.src, chunk_of(tb0), g.limit, s_cur);
            p_cur = load_pred(chunk_of(tb0));
            uint32_t par = 0;
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 1 < tb1) {
                    ind3_load8(g.src, chunk_of(k + 1), g.limit, s_nxt);
                    p_nxt = load_pred(chunk_of(k + 1));
                }
                // per-thread packed counts (16-bit fields); bytes >= 4 mean "nothing to induce"
                unsigned long long mine = 0;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t d = (uint32_t)(p_cur >> (8 * i)) & 0xffu;
                    if (d < 4u) mine += 1ull << (16 * d);
                }
                unsigned long long inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned long long t = __shfl_up_sync(FULL, inc, o);
                    if ((int)l >= o) inc += t;
                }
                if (l == 31) s3.wtot[par][w] = inc;
                __syncthreads();
                unsigned long long wpre = 0, ttot = 0;
#pragma unroll
                for (int ww = 0; ww < NWARP; ww++) {
                    unsigned long long t = s3.wtot[par][ww];
                    if (ww < (int)w) wpre += t;
                    ttot += t;
                }
                unsigned long long exc = wpre + inc - mine;        // packed exclusive ranks of this thread's first item
                uint32_t e0 = (uint32_t)(exc & 0xffffu), e1 = (uint32_t)((exc >> 16) & 0xffffu);
                uint32_t e2 = (uint32_t)((exc >> 32) & 0xffffu), e3 = (uint32_t)((exc >> 48) & 0xffffu);
#pragma unroll
                for (int ii = 0; ii < ITEMS; ii++) {
                    const int i = REV ? (ITEMS - 1 - ii) : ii;           // logical order inside the chunk
                    uint32_t d = (uint32_t)(p_cur >> (8 * i)) & 0xffu;
                    if (d < 4u) {
                        uint32_t pos, bs;
                        if (d == 0u) { pos = rb0 + e0++; bs = bs0; }
                        else if (d == 1u) { pos = rb1 + e1++; bs = bs1; }
                        else if (d == 2u) { pos = rb2 + e2++; bs = bs2; }
                        else { pos = rb3 + e3++; bs = bs3; }
                        uint32_t slot = SPASS ? bs - pos : bs + pos;
                        A.sa[slot] = s_cur[i] - 1u;
                    }
                }
                rb0 += (uint32_t)(ttot & 0xffffu); rb1 += (uint32_t)((ttot >> 16) & 0xffffu);
                rb2 += (uint32_t)((ttot >> 32) & 0xffffu); rb3 += (uint32_t)((ttot >> 48) & 0xffffu);
                par ^= 1u;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) s_cur[i] = s_nxt[i];
                p_cur = p_nxt;
            }
        }
        __syncthreads();
        if (tid < 4 && tid < s3.nsig) sh.fill[sh.alpha[tid]] += s3.ctot[tid];
        if (tid == 0) { sh.st_c = sh.ns_c; sh.st_phase = sh.ns_phase; sh.st_begin = sh.ns_begin; }
        grid.sync();
    }
    if (bid == 0 && tid == 0) {     // step statistics of this launch (diagnostics, tools/induce_steps.py)
        A.err[4 + (SPASS ? 3 : 0)] = bigcount; A.err[5 + (SPASS ? 3 : 0)] = smallcount; A.err[6 + (SPASS ? 3 : 0)] = bigtiles;
    }
}

}  // namespace b200sa
