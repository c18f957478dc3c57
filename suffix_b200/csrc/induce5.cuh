// induce5.cuh -- K4/K5 for 2-bit packed text (sigma <= 4): the one-round bucket steps of
// induce.cuh (reference src/table.rs:421-448, :543-573, :723-736) built from what the step
// timeline of induce3 / induce4 showed (profiles/r02_induce_step_timeline*.txt):
//   * WARP-PRIVATE TILE STREAMS.  A warp owns a contiguous run of 256-entry tiles (8 consecutive,
//     32-byte aligned entries per lane) and walks it alone: packed 4 x 16-bit counters, one warp
//     scan per tile, no block barrier inside the count and scatter loops (the barriers and the
//     exposed load latencies between them, not bytes, bounded the block-wide tiles).
//   * CARRIED CHARS.  A 16-bit word next to every SA slot holds the next SIX text chars to the left
//     of the entry, b = c1 | c2 << 2 | ... | c6 << 10 | cnt << 12, written by the step that PRODUCES
//     the entry; a child inherits (b >> 2, cnt - 1) without touching memory.  The producer
//     refreshes a child that would run dry (cnt == 1 -> one gather for the child; with three chars
//     in a byte a third of all entries needed that, measured), so a consumer only gathers for list
//     heads (LMS suffixes) and for the rare entries made by the shared small-step code (marked 0).
//     The divergent T[s-1] gathers, one LSU wavefront per entry, were the limiter of every count
//     phase.
//   * COALESCED STORES.  A tile's products are staged per warp in shared memory, bucket-major,
//     and written with consecutive lanes on consecutive slots.
// Control flow (peek, big / small steps, run skipping, fill accounting, invariants) is the one
// of induce.cuh.
#pragma once
#include "induce4.cuh"

namespace b200sa {

constexpr uint32_t WT = 256;            // entries per warp tile

struct Ind5Shared {
    uint32_t wrow[NWARP][4];            // per-warp counts of the current step
    uint32_t cbase[4], ctot[4];
    uint32_t code_of[256];
    uint32_t nsig;
};

// pair-reversal of 16 two-bit groups (same as lms_sort.cuh::rev_pairs, which is included later)
__device__ __forceinline__ uint32_t rev_pairs16(uint32_t x) {
    uint32_t y = __brev(x);
    return ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);
}
// carried word of the entry at text position e from the ONE aligned text word that holds T[e-1]
// (up to six chars; fewer when the word starts within six chars)
__device__ __forceinline__ uint32_t carry_from_word(uint32_t e, uint32_t wd) {
    uint32_t q = e - 1u;                          // position of c1 (e > 0)
    uint32_t av = (q & 15u) + 1u;                 // chars of this word at or below q
    uint32_t x = wd << (2u * (15u - (q & 15u)));  // c1 in the top pair
    uint32_t cnt = av < 6u ? av : 6u;
    return (rev_pairs16(x) & 0xfffu) | (cnt << 12);
}
__device__ __forceinline__ uint32_t fresh_carry(const void *__restrict__ ptext, uint32_t e) {
    if (e == 0) return 0u;
    return carry_from_word(e, __ldg(reinterpret_cast<const uint32_t *>(ptext) + ((e - 1u) >> 4)));
}
// 8 carried words of a lane's chunk (16 bytes)
__device__ __forceinline__ void load_carry8(const uint16_t *__restrict__ pc, uint32_t chunk, uint32_t limit, uint32_t (&v)[4]) {
    if (chunk + ITEMS <= limit) {
        uint4 a = __ldcg(reinterpret_cast<const uint4 *>(pc + chunk));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t lo = (chunk + 2 * i < limit) ? pc[chunk + 2 * i] : 0u, hi = (chunk + 2 * i + 1 < limit) ? pc[chunk + 2 * i + 1] : 0u;
            v[i] = lo | (hi << 16);
        }
    }
}
__device__ __forceinline__ uint32_t carry_at(const uint32_t (&v)[4], int i) { return (v[i >> 1] >> ((i & 1) * 16)) & 0xffffu; }

template <bool SPASS>
__global__ void __launch_bounds__(BLK, 3) k_induce5(InduceArgs A) {
    __shared__ IndShared sh;
    __shared__ Ind5Shared s5;
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
        s5.code_of[tid] = inc - present;
        if (tid == 0) s5.nsig = total;
    }
    __syncthreads();
    if (!SPASS && bid == 0 && tid == 0) { A.sa[sh.bstart[lastc]] = A.n - 1u; reinterpret_cast<uint16_t *>(A.pred)[sh.bstart[lastc]] = 0; }
    uint32_t bigcount = 0, smallcount = 0, bigtiles = 0;
    // per-warp staging area (aliases the run-skipping arrays of block 0's small episodes, which are
    // idle during big steps): 256 entries + 256 carried words
    uint32_t *stage_v = sh.ent + w * WT;

    while (true) {
        if (tid == 0) induce_peek<SPASS>(A, sh);
        __syncthreads();
        if (A.steplog && bid == 0 && tid == 0) {
            unsigned long long now, k = A.steplog[0];
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (k < 2000) { A.steplog[1 + 2 * k] = now; A.steplog[2 + 2 * k] = sh.has ? sh.seg.len : 0u; A.steplog[0] = k + 1; }
        }
        if (!sh.has) break;
        if (sh.seg.len <= (uint32_t)TILE) {
            induce_small_episode<SPASS, 2>(A, sh, grid, smallcount);
            continue;
        }
        // -------------------- big step: all blocks, every warp its own run of tiles
        const Seg &o = sh.seg;
        const uint32_t *src = o.src;
        uint8_t *pred = o.pred;
        uint32_t pa, pb;
        if (o.rev) { pb = o.base + 1u; pa = pb - o.len; } else { pa = o.base; pb = o.base + o.len; }
        const uint32_t limit = (o.src == A.sa) ? A.n : sh.S_or_lmsoff[256];
        uint32_t lo = 0, hi = s5.nsig;
        while (lo < s5.nsig && sh.alpha[lo] < o.lo) lo++;
        while (hi > 0 && sh.alpha[hi - 1] > o.hi) hi--;
        const bool gather_all = (src != A.sa);             // LMS list: nothing carried yet
        uint32_t *cntbuf = A.blk_cnt + (size_t)(bigcount & 1u) * G * 4u;
        bigcount++;
        constexpr bool REV = SPASS;
        const uint32_t T0 = pa / WT, T1 = (pb - 1u) / WT;
        const uint32_t tiles = T1 - T0 + 1u;
        bigtiles += (tiles + 7u) / 8u;
        const uint32_t gw = bid * NWARP + w, GW = G * NWARP;
        const uint32_t tpw = (tiles + GW - 1) / GW;
        uint32_t tb0 = gw * tpw, tb1 = tb0 + tpw;
        if (tb0 > tiles) tb0 = tiles;
        if (tb1 > tiles) tb1 = tiles;
        auto chunk_of = [&](uint32_t k) -> uint32_t {
            uint32_t T = REV ? (T1 - k) : (T0 + k);
            return REV ? (T * WT + WT - 8u * (l + 1u)) : (T * WT + 8u * l);
        };
        uint16_t *pc = reinterpret_cast<uint16_t *>(pred);          // carried words of this list
        uint16_t *pc_sa = reinterpret_cast<uint16_t *>(A.pred);     // carried words next to the SA slots
        // ---- phase A: count
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (gather_all) {
            uint32_t s0[ITEMS], s1[ITEMS], wv[ITEMS];
            if (tb0 < tb1) ind4_load8(src, chunk_of(tb0), limit, s0);
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 1 < tb1) ind4_load8(src, chunk_of(k + 1), limit, s1);
                const uint32_t chunk = chunk_of(k);
#pragma unroll
                for (int i = 0; i < ITEMS; i++)           // one aligned text word per entry, all eight in flight
                    wv[i] = __ldg(reinterpret_cast<const uint32_t *>(A.ptext) + ((s0[i] > 0 ? s0[i] - 1u : 0u) >> 4));
                uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    bool in = p >= pa && p < pb;
                    uint32_t bt = (in && s0[i] > 0) ? carry_from_word(s0[i], wv[i]) : 0u;
                    uint32_t d = bt & 3u;
                    bool ok = in && (bt >> 12) != 0u && d >= lo && d < hi;
                    c0 += (ok && d == 0u); c1 += (ok && d == 1u); c2 += (ok && d == 2u); c3 += (ok && d == 3u);
                    out[i >> 1] |= bt << ((i & 1) * 16);
                }
                // neighbours outside the segment belong to other LMS lists, which write theirs when their turn comes
                if (chunk + ITEMS <= limit) *reinterpret_cast<uint4 *>(pc + chunk) = make_uint4(out[0], out[1], out[2], out[3]);
                else {
                    for (int i = 0; i < ITEMS; i++) if (chunk + i < limit) pc[chunk + i] = (uint16_t)(out[i >> 1] >> ((i & 1) * 16));
                }
#pragma unroll
                for (int i = 0; i < ITEMS; i++) s0[i] = s1[i];
            }
        } else {
            uint32_t p_cur[4] = {0, 0, 0, 0}, p_nxt[4] = {0, 0, 0, 0};
            if (tb0 < tb1) load_carry8(pc, chunk_of(tb0), limit, p_cur);
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 1 < tb1) load_carry8(pc, chunk_of(k + 1), limit, p_nxt);
                const uint32_t chunk = chunk_of(k);
                uint32_t need = 0;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    if (p >= pa && p < pb && (carry_at(p_cur, i) >> 12) == 0u) need |= 1u << i;
                }
                if (need) {                               // products of small steps / the seed: chars not carried
                    uint32_t sv[ITEMS];
                    ind4_load8(src, chunk, limit, sv);
#pragma unroll
                    for (int i = 0; i < ITEMS; i++) {
                        if ((need >> i) & 1u) {
                            uint32_t nb = fresh_carry(A.ptext, sv[i]);
                            p_cur[i >> 1] = (p_cur[i >> 1] & ~(0xffffu << ((i & 1) * 16))) | (nb << ((i & 1) * 16));
                            pc[chunk + i] = (uint16_t)nb;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    uint32_t bt = carry_at(p_cur, i);
                    uint32_t d = bt & 3u;
                    bool ok = p >= pa && p < pb && (bt >> 12) != 0u && d >= lo && d < hi;
                    c0 += (ok && d == 0u); c1 += (ok && d == 1u); c2 += (ok && d == 2u); c3 += (ok && d == 3u);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) p_cur[i] = p_nxt[i];
            }
        }
        c0 = __reduce_add_sync(FULL, c0); c1 = __reduce_add_sync(FULL, c1);
        c2 = __reduce_add_sync(FULL, c2); c3 = __reduce_add_sync(FULL, c3);
        if (l == 0) { s5.wrow[w][0] = c0; s5.wrow[w][1] = c1; s5.wrow[w][2] = c2; s5.wrow[w][3] = c3; }
        __syncthreads();
        if (tid < 4) {
            uint32_t v = 0;
#pragma unroll
            for (int ww = 0; ww < NWARP; ww++) v += s5.wrow[ww][tid];
            cntbuf[(size_t)bid * 4u + tid] = v;
        }
        __syncthreads();
        IND_MARK(1)
        grid.sync();
        IND_MARK(2)
        // ---- phase B: offsets from the G x 4 count matrix, then every warp scatters its run
        {
            uint32_t be[4] = {0, 0, 0, 0}, to[4] = {0, 0, 0, 0};
            for (uint32_t b = tid; b < G; b += BLK) {
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
                s5.cbase[tid] = sh.fill[sh.alpha[tid]] + bsum;
                s5.ctot[tid] = tsum;
            }
            __syncthreads();
        }
        IND_MARK(3)
        if (tb0 < tb1) {
            uint32_t rb0 = s5.cbase[0], rb1 = s5.cbase[1], rb2 = s5.cbase[2], rb3 = s5.cbase[3];
            for (uint32_t ww = 0; ww < w; ww++) { rb0 += s5.wrow[ww][0]; rb1 += s5.wrow[ww][1]; rb2 += s5.wrow[ww][2]; rb3 += s5.wrow[ww][3]; }
            const uint32_t bs0 = SPASS ? sh.bstart[sh.alpha[0] + 1] - 1u : sh.bstart[sh.alpha[0]];
            const uint32_t bs1 = SPASS ? sh.bstart[sh.alpha[1] + 1] - 1u : sh.bstart[sh.alpha[1]];
            const uint32_t bs2 = SPASS ? sh.bstart[sh.alpha[2] + 1] - 1u : sh.bstart[sh.alpha[2]];
            const uint32_t bs3 = SPASS ? sh.bstart[sh.alpha[3] + 1] - 1u : sh.bstart[sh.alpha[3]];
            uint32_t s_cur[ITEMS], s_nxt[ITEMS], p_cur[4], p_nxt[4] = {0, 0, 0, 0};
            uint16_t *stage_c = reinterpret_cast<uint16_t *>(sh.rl) + w * WT;
            ind4_load8(src, chunk_of(tb0), limit, s_cur);
            load_carry8(pc, chunk_of(tb0), limit, p_cur);
            for (uint32_t k = tb0; k < tb1; k++) {
                if (k + 1 < tb1) {
                    ind4_load8(src, chunk_of(k + 1), limit, s_nxt);
                    load_carry8(pc, chunk_of(k + 1), limit, p_nxt);
                }
                const uint32_t chunk = chunk_of(k);
                uint32_t okm = 0, dry = 0;
                unsigned long long mine = 0;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    uint32_t p = chunk + i;
                    uint32_t bt = carry_at(p_cur, i);
                    uint32_t d = bt & 3u, cnt = bt >> 12;
                    bool ok = p >= pa && p < pb && cnt != 0u && d >= lo && d < hi;
                    if (ok) { okm |= 1u << i; mine += 1ull << (16 * d); if (cnt == 1u) dry |= 1u << i; }
                }
                // children that would run dry get fresh chars now (rare: every sixth generation of a path);
                // the loads are issued together, not one per branch
                uint32_t wv[ITEMS];
#pragma unroll
                for (int i = 0; i < ITEMS; i++)
                    wv[i] = ((dry >> i) & 1u) && s_cur[i] > 1u ? __ldg(reinterpret_cast<const uint32_t *>(A.ptext) + ((s_cur[i] - 2u) >> 4)) : 0u;
                unsigned long long inc = mine;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    unsigned long long t = __shfl_up_sync(FULL, inc, s);
                    if ((int)l >= s) inc += t;
                }
                const unsigned long long ttot = __shfl_sync(FULL, inc, 31);
                const unsigned long long exc = inc - mine;
                const uint32_t t0c = (uint32_t)(ttot & 0xffffu), t1c = (uint32_t)((ttot >> 16) & 0xffffu);
                const uint32_t t2c = (uint32_t)((ttot >> 32) & 0xffffu), t3c = (uint32_t)((ttot >> 48) & 0xffffu);
                const uint32_t o1 = t0c, o2 = t0c + t1c, o3 = o2 + t2c, oall = o3 + t3c;
                uint32_t e0 = (uint32_t)(exc & 0xffffu), e1 = o1 + (uint32_t)((exc >> 16) & 0xffffu);
                uint32_t e2 = o2 + (uint32_t)((exc >> 32) & 0xffffu), e3 = o3 + (uint32_t)((exc >> 48) & 0xffffu);
                __syncwarp();                                   // the previous tile's copy-out is done
#pragma unroll
                for (int ii = 0; ii < ITEMS; ii++) {
                    const int i = REV ? (ITEMS - 1 - ii) : ii;   // logical order inside the chunk
                    if ((okm >> i) & 1u) {
                        uint32_t bt = carry_at(p_cur, i);
                        uint32_t d = bt & 3u, cnt = bt >> 12;
                        uint32_t cb;
                        if (cnt >= 2u) cb = ((bt >> 2) & 0x3ffu) | ((cnt - 1u) << 12);
                        else cb = (s_cur[i] > 1u) ? carry_from_word(s_cur[i] - 1u, wv[i]) : 0u;
                        uint32_t at = (d == 0u) ? e0++ : (d == 1u) ? e1++ : (d == 2u) ? e2++ : e3++;
                        stage_v[at] = s_cur[i] - 1u;
                        stage_c[at] = (uint16_t)cb;
                    }
                }
                __syncwarp();
                for (uint32_t j = l; j < oall; j += 32) {
                    uint32_t pos, bs;
                    if (j < o1) { pos = rb0 + j; bs = bs0; }
                    else if (j < o2) { pos = rb1 + (j - o1); bs = bs1; }
                    else if (j < o3) { pos = rb2 + (j - o2); bs = bs2; }
                    else { pos = rb3 + (j - o3); bs = bs3; }
                    uint32_t slot = SPASS ? bs - pos : bs + pos;
                    A.sa[slot] = stage_v[j];
                    pc_sa[slot] = stage_c[j];
                }
                rb0 += t0c; rb1 += t1c; rb2 += t2c; rb3 += t3c;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) s_cur[i] = s_nxt[i];
#pragma unroll
                for (int i = 0; i < 4; i++) p_cur[i] = p_nxt[i];
            }
        }
        __syncthreads();
        if (tid < 4 && tid < s5.nsig) sh.fill[sh.alpha[tid]] += s5.ctot[tid];
        if (tid == 0) { sh.st_c = sh.ns_c; sh.st_phase = sh.ns_phase; sh.st_begin = sh.ns_begin; }
        IND_MARK(4)
        grid.sync();
    }
    if (bid == 0 && tid == 0) {
        A.err[4 + (SPASS ? 3 : 0)] = bigcount; A.err[5 + (SPASS ? 3 : 0)] = smallcount; A.err[6 + (SPASS ? 3 : 0)] = bigtiles;
    }
}

}  // namespace b200sa
