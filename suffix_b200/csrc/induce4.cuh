// induce4.cuh -- K4/K5 for 2-bit packed text (sigma <= 4): induce3.cuh (packed-counter ranking on
// physically aligned tiles; reference src/table.rs:421-448, :543-573, :723-736) plus CARRIED
// PREDECESSOR CHARS.  Measured on induce3: the pass is bound by the divergent gathers T[s-1]
// (one 32-byte L2 sector per entry, ~2 LSU wavefront cycles each).  Here the byte that sits
// next to every SA slot carries the next three text chars to the LEFT of the entry,
//     b = c1 | c2 << 2 | c3 << 4 | cnt << 6      (c1 = T[s-1], ...; cnt = how many are valid),
// written by the step that PRODUCES the entry: the child s-1 of an entry with byte b gets
// (b >> 2 with cnt - 1) -- no memory access.  Only the head of an induction path (an LMS
// suffix in the L pass, or an entry whose three chars are used up: cnt == 0) gathers from the
// packed text, one aligned word per gather.  100 MB DNA: ~30 M gathers per pass instead of
// ~96 M.  Steps, ranking, fill accounting and the small-step / run-skipping path are those
// of induce3.cuh / induce.cuh (products of the shared small-step code are marked "unknown").
#pragma once
#include "induce.cuh"

namespace b200sa {

struct Ind4Shared {
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
__device__ __forceinline__ void ind4_load8(const uint32_t *__restrict__ src, uint32_t chunk, uint32_t limit,
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

struct Seg4 {
    const uint32_t *src;
    uint8_t *pred;
    uint32_t pa, pb;      // physical range [pa, pb)
    uint32_t limit;       // entries in the underlying array (no access at or beyond)
    uint32_t lo, hi;      // valid destination CODES (inclusive)
    int rev;
};

template <bool SPASS>
__global__ void __launch_bounds__(BLK, 3) k_induce4(InduceArgs A) {
    __shared__ IndShared sh;
    __shared__ Ind4Shared s4;
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
        s4.code_of[tid] = inc - present;
        if (tid == 0) s4.nsig = total;
    }
    __syncthreads();
    if (!SPASS && bid == 0 && tid == 0) { A.sa[sh.bstart[lastc]] = A.n - 1u; A.pred[sh.bstart[lastc]] = 0; }
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
        Seg4 g;
        {
            const Seg &o = sh.seg;
            g.src = o.src; g.pred = o.pred; g.rev = o.rev;     // o.rev == SPASS for every list
            if (o.rev) { g.pb = o.base + 1u; g.pa = g.pb - o.len; }
            else { g.pa = o.base; g.pb = o.base + o.len; }
            g.limit = (o.src == A.sa) ? A.n : sh.S_or_lmsoff[256];     // L pass second list: the LMS list (m entries)
            // destination range in codes (the byte range [lo, hi] holds only bytes that occur at its ends)
            uint32_t lo = 0, hi = 0;
            const uint32_t ns = s4.nsig;
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
        // ---- phase A: count.  The chars come from the carried bytes (one 8-byte load per thread);
        // only entries with cnt == 0 (heads of induction paths; every entry of an LMS list) read
        // their position and gather ONE aligned word of packed text, then refresh their byte.
        const bool gather_all = (g.src != A.sa);          // LMS list (L pass second list): nothing carried yet
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        auto load_pred = [&](uint32_t chunk) -> unsigned long long {
            if (chunk + ITEMS <= g.limit) return __ldcg(reinterpret_cast<const unsigned long long *>(g.pred + chunk));
            unsigned long long v = 0;
            for (int i = 0; i < ITEMS; i++) v |= (unsigned long long)((chunk + i < g.limit) ? g.pred[chunk + i] : 0u) << (8 * i);
            return v;
        };
        auto byte_from_word = [&](uint32_t sp, uint32_t wd) -> uint32_t {     // carried byte of entry sp from its text word
            uint32_t q = sp - 1u;                     // position of c1 (sp > 0)
            uint32_t av = (q & 15u) + 1u;             // chars of this word at or below q
            uint32_t x = wd << (2u * (15u - (q & 15u)));          // c1 in the top pair
            uint32_t cnt = av < 3u ? av : 3u;
            return ((x >> 30) & 3u) | (((x >> 28) & 3u) << 2) | (((x >> 26) & 3u) << 4) | (cnt << 6);
        };
        if (tb0 < tb1 && gather_all) {
            // every entry gathers one aligned text word (the phase is bound by the LSU wavefronts of
            // these divergent loads, not by their latency: positions are prefetched one tile ahead only)
            uint32_t s0[ITEMS], s1[ITEMS], w0[ITEMS];
            ind4_load8(g.src, chunk_of(tb0), g.limit, s0);
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 1 < tb1) ind4_load8(g.src, chunk_of(k + 1), g.limit, s1);
#pragma unroll
                for (int i = 0; i < ITEMS; i++)
                    w0[i] = __ldg(reinterpret_cast<const uint32_t *>(A.ptext) + ((s0[i] > 0 ? s0[i] - 1u : 0u) >> 4));
                const uint32_t chunk = chunk_of(k);
                unsigned long long pk = 0;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    uint32_t bt = (s0[i] > 0) ? byte_from_word(s0[i], w0[i]) : 0u;
                    uint32_t d = bt & 3u;
                    bool ok = p >= g.pa && p < g.pb && (bt >> 6) != 0u && d >= g.lo && d < g.hi;
                    c0 += (ok && d == 0u); c1 += (ok && d == 1u); c2 += (ok && d == 2u); c3 += (ok && d == 3u);
                    pk |= (unsigned long long)bt << (8 * i);
                }
                // the bytes of a whole chunk go out at once; neighbours outside the segment belong to other
                // LMS lists, which refresh theirs when their turn comes
                if (chunk + ITEMS <= g.limit) *reinterpret_cast<unsigned long long *>(g.pred + chunk) = pk;
                else {
                    for (int i = 0; i < ITEMS; i++) if (chunk + i < g.limit) g.pred[chunk + i] = (uint8_t)(pk >> (8 * i));
                }
#pragma unroll
                for (int i = 0; i < ITEMS; i++) s0[i] = s1[i];
            }
        } else if (tb0 < tb1) {
            unsigned long long p_cur = load_pred(chunk_of(tb0)), p_nxt = 0;
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 1 < tb1) p_nxt = load_pred(chunk_of(k + 1));
                const uint32_t chunk = chunk_of(k);
                // which entries of the chunk lie in the segment and still need their chars
                uint32_t need = 0;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    bool in = p >= g.pa && p < g.pb;
                    if (in && ((p_cur >> (8 * i + 6)) & 3u) == 0u) need |= 1u << i;
                }
                if (need) {                                       // refill: position -> aligned text word -> up to 3 chars
                    uint32_t sv[ITEMS];
                    ind4_load8(g.src, chunk, g.limit, sv);
#pragma unroll
                    for (int i = 0; i < ITEMS; i++) {
                        if ((need >> i) & 1u) {
                            uint32_t sp = sv[i];
                            uint32_t nb = 0;                      // s == 0: nothing to the left, cnt stays 0
                            if (sp > 0) nb = byte_from_word(sp, __ldg(reinterpret_cast<const uint32_t *>(A.ptext) + ((sp - 1u) >> 4)));
                            p_cur = (p_cur & ~(0xffull << (8 * i))) | ((unsigned long long)nb << (8 * i));
                            g.pred[chunk + i] = (uint8_t)nb;      // the scatter phase (and later passes) read the refreshed byte
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    uint32_t bt = (uint32_t)(p_cur >> (8 * i)) & 0xffu;
                    uint32_t d = bt & 3u;
                    bool ok = p >= g.pa && p < g.pb && (bt >> 6) != 0u && d >= g.lo && d < g.hi;
                    c0 += (ok && d == 0u); c1 += (ok && d == 1u); c2 += (ok && d == 2u); c3 += (ok && d == 3u);
                }
                p_cur = p_nxt;
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
        IND_MARK(1)       // count loop + block totals done
        grid.sync();
        IND_MARK(2)       // first grid.sync passed
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
                s4.cbase[tid] = sh.fill[sh.alpha[tid]] + bsum;
                s4.ctot[tid] = tsum;
            }
        }
        __syncthreads();
        IND_MARK(3)       // count matrix summed
        if (tb0 < tb1) {
            // running destination positions of the 4 codes (registers, identical in all threads)
            uint32_t rb0 = s4.cbase[0], rb1 = s4.cbase[1], rb2 = s4.cbase[2], rb3 = s4.cbase[3];
            const uint32_t bs0 = SPASS ? sh.bstart[sh.alpha[0] + 1] - 1u : sh.bstart[sh.alpha[0]];
            const uint32_t bs1 = SPASS ? sh.bstart[sh.alpha[1] + 1] - 1u : sh.bstart[sh.alpha[1]];
            const uint32_t bs2 = SPASS ? sh.bstart[sh.alpha[2] + 1] - 1u : sh.bstart[sh.alpha[2]];
            const uint32_t bs3 = SPASS ? sh.bstart[sh.alpha[3] + 1] - 1u : sh.bstart[sh.alpha[3]];
            uint32_t s_cur[ITEMS], s_nxt[ITEMS];
            unsigned long long p_cur = 0, p_nxt = 0;
            ind4_load8(g.src, chunk_of(tb0), g.limit, s_cur);
            p_cur = load_pred(chunk_of(tb0));
            uint32_t par = 0;
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 1 < tb1) {
                    ind4_load8(g.src, chunk_of(k + 1), g.limit, s_nxt);
                    p_nxt = load_pred(chunk_of(k + 1));
                }
                const uint32_t chunk = chunk_of(k);
                // per-thread packed counts (16-bit fields)
                uint32_t okm = 0;
                unsigned long long mine = 0;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    uint32_t bt = (uint32_t)(p_cur >> (8 * i)) & 0xffu;
                    uint32_t d = bt & 3u;
                    bool ok = p >= g.pa && p < g.pb && (bt >> 6) != 0u && d >= g.lo && d < g.hi;
                    if (ok) { okm |= 1u << i; mine += 1ull << (16 * d); }
                }
                unsigned long long inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned long long t = __shfl_up_sync(FULL, inc, o);
                    if ((int)l >= o) inc += t;
                }
                if (l == 31) s4.wtot[par][w] = inc;
                __syncthreads();
                unsigned long long wpre = 0, ttot = 0;
#pragma unroll
                for (int ww = 0; ww < NWARP; ww++) {
                    unsigned long long t = s4.wtot[par][ww];
                    if (ww < (int)w) wpre += t;
                    ttot += t;
                }
                unsigned long long exc = wpre + inc - mine;        // packed exclusive ranks of this thread's first item
                // stage the tile's products in shared memory, bucket-major in output order, then write
                // them out with consecutive threads on consecutive slots (a direct scatter costs one
                // LSU wavefront per entry, twice with the carried byte)
                const uint32_t t0c = (uint32_t)(ttot & 0xffffu), t1c = (uint32_t)((ttot >> 16) & 0xffffu);
                const uint32_t t2c = (uint32_t)((ttot >> 32) & 0xffffu), t3c = (uint32_t)((ttot >> 48) & 0xffffu);
                const uint32_t o1 = t0c, o2 = t0c + t1c, o3 = o2 + t2c, oall = o3 + t3c;
                uint32_t e0 = (uint32_t)(exc & 0xffffu), e1 = o1 + (uint32_t)((exc >> 16) & 0xffffu);
                uint32_t e2 = o2 + (uint32_t)((exc >> 32) & 0xffffu), e3 = o3 + (uint32_t)((exc >> 48) & 0xffffu);
                uint8_t *stage_b = reinterpret_cast<uint8_t *>(sh.rl);
#pragma unroll
                for (int ii = 0; ii < ITEMS; ii++) {
                    const int i = REV ? (ITEMS - 1 - ii) : ii;           // logical order inside the chunk
                    if ((okm >> i) & 1u) {
                        uint32_t bt = (uint32_t)(p_cur >> (8 * i)) & 0xffu;
                        uint32_t d = bt & 3u;
                        uint32_t at = (d == 0u) ? e0++ : (d == 1u) ? e1++ : (d == 2u) ? e2++ : e3++;
                        sh.ent[at] = s_cur[i] - 1u;
                        stage_b[at] = (uint8_t)(((bt >> 2) & 0x0fu) | (((bt >> 6) - 1u) << 6));      // the child's chars
                    }
                }
                __syncthreads();
                for (uint32_t j = tid; j < oall; j += BLK) {
                    uint32_t pos, bs;
                    if (j < o1) { pos = rb0 + j; bs = bs0; }
                    else if (j < o2) { pos = rb1 + (j - o1); bs = bs1; }
                    else if (j < o3) { pos = rb2 + (j - o2); bs = bs2; }
                    else { pos = rb3 + (j - o3); bs = bs3; }
                    uint32_t slot = SPASS ? bs - pos : bs + pos;
                    A.sa[slot] = sh.ent[j];
                    A.pred[slot] = stage_b[j];
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
        if (tid < 4 && tid < s4.nsig) sh.fill[sh.alpha[tid]] += s4.ctot[tid];
        if (tid == 0) { sh.st_c = sh.ns_c; sh.st_phase = sh.ns_phase; sh.st_begin = sh.ns_begin; }
        IND_MARK(4)       // scatter loop done
        grid.sync();
    }
    if (bid == 0 && tid == 0) {     // step statistics of this launch (diagnostics, tools/induce_steps.py)
        A.err[4 + (SPASS ? 3 : 0)] = bigcount; A.err[5 + (SPASS ? 3 : 0)] = smallcount; A.err[6 + (SPASS ? 3 : 0)] = bigtiles;
    }
}

}  // namespace b200sa
