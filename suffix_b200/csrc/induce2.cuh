// induce2.cuh -- K4/K5 for packed text (sigma <= 16): the induced fill with MULTI-ROUND
// bucket steps.  Same job and same serial semantics as induce.cuh (reference
// src/table.rs:421-448, :543-573, head_insert/tail_insert :723-736), but one big step
// now plays a whole bucket:
//
//   segment A = the entries present in bucket c when its scan starts (E0), each with the
//     run of c to its left read from ONE window of the packed text (CPW = 16 chars at
//     2 bits, 8 at 4 bits): run length r = number of leading chars equal to c.
//       * chain rounds 1..r emit s-1 .. s-r into bucket c itself; round j holds, in list
//         order, the entries with r >= j, and the rounds lie one after the other
//         (the serial scan visits E0, then what E0 induced into c, ...);
//       * the terminal s-r-1 (char d' != c) goes to bucket d' iff the type test
//         passes (L pass: d' > c, S pass: d' < c); the serial scan reaches it while it
//         visits round r, so inside bucket d' terminals are ordered by (r, list order).
//     Runs of R = CPW-1 or more ("long") emit R rounds; round R then is the list of the
//     next step of the same bucket (continuation).
//   segment B = the bucket's second list (L pass: its LMS suffixes; S pass: its L part),
//     whose products follow every terminal of A in their destination buckets; it is
//     played in the same step unless A has long runs (then it waits for the continuation).
//
// A step is count (classify every entry into a bin (r, d'), remember the bin in pred[],
// per-block histogram) -> grid.sync -> every block derives all positions from the count
// matrix -> stable scatter -> grid.sync: two grid syncs per BUCKET instead of two per
// chain round (100 MB DNA: 4 big steps per pass instead of 23 big + 19 small).
// Lists of at most TILE entries and very long runs take the small-episode path of
// induce.cuh (block 0, run skipping) unchanged.
// Executable model: tests/model_pipeline.py::induce_multiround (tests/test_model.py).
#pragma once
#include "induce.cuh"

namespace b200sa {

template <int BITS>
struct MR {
    static constexpr uint32_t CPW = 32 / BITS;         // chars per window
    static constexpr uint32_t R = CPW - 1;             // chain rounds per step
    static constexpr uint32_t DS = 1u << BITS;         // char slots per class
    static constexpr uint32_t CLS_LONG = R;            // run >= R
    static constexpr uint32_t CLS_B = R + 1;           // second-list entry
    static constexpr uint32_t NBIN = (R + 2) * DS;     // 68 / 144
    static constexpr uint32_t REP = BITS == 2 ? 0x55555555u : 0x11111111u;
};

struct IndShared2 {
    uint32_t tot[256];        // step totals per bin
    uint32_t bexc[256];       // entries of earlier blocks per bin
    uint32_t tb[256];         // running destination position per bin (position inside the bucket part)
    uint32_t code_of[256];    // byte -> dense code
    uint32_t cb[16];          // running position of chain round j inside bucket c
    uint32_t wch[NWARP][16];  // per-warp counts / exclusive offsets of chain round j in the tile
    uint32_t tch[16];         // tile totals per chain round
    uint32_t newfill[16];     // per destination code: fill after the step
    uint32_t nlong, incl_b, chain_add, cont_begin, nsig;
};

// bin of a chain-list entry s of bucket code c: (run length r, terminal char d'); d' == c
// stands for "no predecessor to induce" (text start).
template <int BITS>
__device__ __forceinline__ uint32_t mr_bin_chain(const void *__restrict__ ptext, uint32_t s, uint32_t c) {
    typedef MR<BITS> M;
    if (s == 0) return c;                                        // r = 0, nothing before
    const uint32_t avail = s < M::CPW ? s : M::CPW;
    uint32_t x = (s >= M::CPW) ? text_bits<BITS>(ptext, s - M::CPW)
                               : (text_bits<BITS>(ptext, 0) << (BITS * (M::CPW - s)));   // T[s-1] in the top group
    uint32_t y = x ^ (c * M::REP);
    uint32_t r = y ? (uint32_t)__clz(y) / BITS : M::CPW;
    if (r > avail) r = avail;
    if (r >= M::R) return M::CLS_LONG * M::DS;
    uint32_t d = (r < s) ? ((x >> (32u - BITS * (r + 1u))) & (M::DS - 1u)) : c;
    return r * M::DS + d;
}

template <bool SPASS, int BITS>
__global__ void __launch_bounds__(BLK, INDUCE_MINB) k_induce2(InduceArgs A) {
    typedef MR<BITS> M;
    __shared__ IndShared sh;
    __shared__ IndShared2 s2;
    cg::grid_group grid = cg::this_grid();
    const uint32_t G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
    const uint32_t w = warp_id(), l = lane_id(), lt = lanemask_lt();

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
    {   // dense code of every byte that occurs (order preserving)
        uint32_t present = (A.Lcnt[tid] + A.Scnt[tid]) > 0 ? 1u : 0u, total;
        uint32_t inc = block_incl_scan<OpSum>(present, sh.sw, &total);
        s2.code_of[tid] = inc - present;
        if (tid == 0) s2.nsig = total;
    }
    __syncthreads();
    if (!SPASS && bid == 0 && tid == 0) A.sa[sh.bstart[lastc]] = A.n - 1u;
    uint32_t bigcount = 0, smallcount = 0, bigtiles = 0;

    while (true) {
        if (tid == 0) induce_peek<SPASS>(A, sh);
        __syncthreads();
        if (!sh.has) break;
        if (sh.seg.len <= (uint32_t)TILE) {
            induce_small_episode<SPASS, BITS>(A, sh, grid, smallcount);
            continue;
        }
        // -------------------- big multi-round step: all blocks
        const bool chain = sh.is_chain != 0;
        const int32_t cbyte = chain ? sh.ns_c : (SPASS ? sh.ns_c + 1 : sh.ns_c - 1);   // bucket being scanned
        const uint32_t cc = s2.code_of[cbyte];
        Seg ga, gb;
        ga.len = 0; gb.len = 0;
        if (chain) {
            ga = sh.seg;
            if (SPASS) {
                uint32_t L = sh.Lcnt[cbyte];
                if (L > 0 && cbyte > 0) {
                    gb.src = A.sa; gb.pred = A.pred; gb.len = L; gb.base = sh.bstart[cbyte] + L - 1u; gb.rev = 1;
                    gb.lo = 0; gb.hi = 0;
                }
            } else {
                uint32_t a = sh.S_or_lmsoff[cbyte], b = sh.S_or_lmsoff[cbyte + 1];
                if (b > a && cbyte < 255) {
                    gb.src = A.lms; gb.pred = A.lms_pred; gb.len = b - a; gb.base = a; gb.rev = 0; gb.lo = 0; gb.hi = 0;
                }
            }
        } else {
            gb = sh.seg;
        }
        uint32_t *cntbuf = A.blk_cnt + (size_t)(bigcount & 1u) * G * 256u;
        bigcount++;
        const uint32_t tilesA = (ga.len + TILE - 1) / TILE, tilesB = (gb.len + TILE - 1) / TILE;
        const uint32_t tiles = tilesA + tilesB;
        bigtiles += tiles;
        const uint32_t tpb = (tiles + G - 1) / G;
        const uint32_t nact = (tiles + tpb - 1) / tpb;
        uint32_t tb0 = bid * tpb, tb1 = tb0 + tpb;
        if (tb1 > tiles) tb1 = tiles;
        // ---- phase A: classify, remember the bin, histogram
#pragma unroll
        for (int ww = 0; ww < NWARP; ww++) sh.wcnt[ww][tid] = 0;
        __syncthreads();
        if (tb0 < tb1) {
            uint32_t s_cur[ITEMS], s_nxt[ITEMS];
            {
                const bool isA = tb0 < tilesA;
                tile_load_s(isA ? ga : gb, (isA ? tb0 : tb0 - tilesA) * TILE, s_cur);
            }
            for (uint32_t t = tb0; t < tb1; t++) {
                if (t + 1 < tb1) {
                    const bool nA = t + 1 < tilesA;
                    tile_load_s(nA ? ga : gb, (nA ? t + 1 : t + 1 - tilesA) * TILE, s_nxt);     // prefetch under the gather
                }
                const bool isA = t < tilesA;
                const Seg &g = isA ? ga : gb;
                const uint32_t t0 = (isA ? t : t - tilesA) * TILE;
                uint32_t bin[ITEMS];
#pragma unroll
                for (int r = 0; r < ITEMS; r++) {
                    if (isA) bin[r] = mr_bin_chain<BITS>(A.ptext, s_cur[r], cc);
                    else bin[r] = M::CLS_B * M::DS + (s_cur[r] > 0 ? text_get<BITS>(A.ptext, s_cur[r] - 1u) : cc);
                }
#pragma unroll
                for (int r = 0; r < ITEMS; r++) {
                    uint32_t k = t0 + w * (ITEMS * 32) + r * 32 + l;
                    bool live = k < g.len;
                    if (live) g.pred[g.rev ? g.base - k : g.base + k] = (uint8_t)bin[r];
                    hist_add_private(sh.wcnt[w], bin[r], live);
                }
#pragma unroll
                for (int r = 0; r < ITEMS; r++) s_cur[r] = s_nxt[r];
            }
        }
        __syncthreads();
        if (bid < nact) {
            uint32_t v = 0;
#pragma unroll
            for (int ww = 0; ww < NWARP; ww++) v += sh.wcnt[ww][tid];
            cntbuf[(size_t)bid * 256u + tid] = v;
        }
        __syncthreads();
        grid.sync();
        // ---- phase B: positions from the count matrix
        {
            uint32_t be = 0, to = 0;
            if (tid < M::NBIN) {
                for (uint32_t b = 0; b < nact; b++) {
                    uint32_t v = __ldcg(cntbuf + (size_t)b * 256u + tid);
                    if (b < bid) be += v;
                    to += v;
                }
            }
            s2.bexc[tid] = be;
            s2.tot[tid] = to;
        }
        __syncthreads();
        if (tid < M::DS) {                      // destinations: terminals by round, then the second list
            const uint32_t d = tid;
            const bool ok = d < s2.nsig && (SPASS ? (d < cc) : (d > cc));
            uint32_t run = sh.fill[sh.alpha[d]];
            for (uint32_t r = 0; r < M::R; r++) {
                uint32_t b = r * M::DS + d;
                s2.tb[b] = run + s2.bexc[b];
                run += s2.tot[b];
            }
            uint32_t bB = M::CLS_B * M::DS + d;
            s2.tb[bB] = run + s2.bexc[bB];
            s2.newfill[d] = ok ? run : 0xffffffffu;           // fill after A's terminals (B added below)
        }
        if (tid == 32) {                        // chain rounds (another warp than the loop above)
            uint32_t clsT[M::R + 1], clsB[M::R + 1];
            for (uint32_t r = 0; r <= M::R; r++) {
                uint32_t a = 0, b = 0;
                for (uint32_t d = 0; d < M::DS; d++) { a += s2.tot[r * M::DS + d]; b += s2.bexc[r * M::DS + d]; }
                clsT[r] = a; clsB[r] = b;
            }
            uint32_t geT[M::R + 2], geB[M::R + 2];
            geT[M::R + 1] = 0; geB[M::R + 1] = 0;
            for (int r = (int)M::R; r >= 1; r--) { geT[r] = geT[r + 1] + clsT[r]; geB[r] = geB[r + 1] + clsB[r]; }
            uint32_t run = chain ? sh.fill[cbyte] : 0u, add = 0;
            for (uint32_t j = 1; j <= M::R; j++) {
                if (j == M::R) s2.cont_begin = run;
                s2.cb[j] = run + geB[j];
                run += geT[j];
                add += geT[j];
            }
            s2.chain_add = add;
            s2.nlong = clsT[M::R];
            s2.incl_b = (clsT[M::R] == 0) ? 1u : 0u;
        }
        __syncthreads();
        const bool incl_b = s2.incl_b != 0 || !chain;
        // ---- scatter
        if (tb0 < tb1) {
            uint32_t s_cur[ITEMS], b_cur[ITEMS], s_nxt[ITEMS], b_nxt[ITEMS];
            {
                const bool isA = tb0 < tilesA;
                tile_load_s(isA ? ga : gb, (isA ? tb0 : tb0 - tilesA) * TILE, s_cur);
                tile_load_pred(isA ? ga : gb, (isA ? tb0 : tb0 - tilesA) * TILE, b_cur);
            }
            for (uint32_t t = tb0; t < tb1; t++) {
                if (t + 1 < tb1) {
                    const bool nA = t + 1 < tilesA;
                    tile_load_s(nA ? ga : gb, (nA ? t + 1 : t + 1 - tilesA) * TILE, s_nxt);
                    tile_load_pred(nA ? ga : gb, (nA ? t + 1 : t + 1 - tilesA) * TILE, b_nxt);
                }
                const bool isA = t < tilesA;
                if (isA || incl_b) {
                    const Seg &g = isA ? ga : gb;
                    const uint32_t t0 = (isA ? t : t - tilesA) * TILE;
                    // ---- terminals / second-list products: stable ranking by bin
                    uint32_t vm = 0, rr[ITEMS], rank[ITEMS];
#pragma unroll
                    for (int r = 0; r < ITEMS; r++) {
                        uint32_t k = t0 + w * (ITEMS * 32) + r * 32 + l;
                        uint32_t cls = b_cur[r] / M::DS, d = b_cur[r] % M::DS;
                        bool live = k < g.len;
                        bool ok = live && cls != M::CLS_LONG && (SPASS ? (d < cc) : (d > cc));
                        vm |= (ok ? 1u : 0u) << r;
                        rr[r] = (live && isA) ? cls : 0u;            // chain rounds of this entry (cls <= R for A)
                    }
#pragma unroll
                    for (int ww = 0; ww < NWARP; ww++) sh.wcnt[ww][tid] = 0;
                    // per-warp totals of the chain rounds
                    uint32_t maxr = 0;
#pragma unroll
                    for (int r = 0; r < ITEMS; r++) maxr = rr[r] > maxr ? rr[r] : maxr;
                    maxr = __reduce_max_sync(FULL, maxr);
                    uint32_t mine = 0;
                    for (uint32_t j = 1; j <= maxr; j++) {
                        uint32_t cj = 0;
#pragma unroll
                        for (int r = 0; r < ITEMS; r++) cj += (rr[r] >= j) ? 1u : 0u;
                        uint32_t tj = __reduce_add_sync(FULL, cj);
                        if (l == j) mine = tj;
                    }
                    if (l < 16) s2.wch[w][l] = mine;
                    __syncthreads();
                    {   // tile_rank body (common.cuh) with the chain-round prefix folded into its barriers
#pragma unroll
                        for (int r = 0; r < ITEMS; r++) {
                            bool valid = (vm >> r) & 1u;
                            uint32_t peers = peer_mask<8>(b_cur[r], valid);
                            uint32_t below = __popc(peers & lt);
                            uint32_t base = valid ? sh.wcnt[w][b_cur[r]] : 0u;
                            __syncwarp();
                            if (valid && below == 0) sh.wcnt[w][b_cur[r]] = base + __popc(peers);
                            __syncwarp();
                            rank[r] = base + below;
                        }
                        __syncthreads();
                        {
                            uint32_t run = 0;
#pragma unroll
                            for (int ww = 0; ww < NWARP; ww++) {
                                uint32_t tt = sh.wcnt[ww][tid];
                                sh.wcnt[ww][tid] = run;
                                run += tt;
                            }
                            sh.tcnt[tid] = run;
                            if (tid < 16) {
                                uint32_t crun = 0;
#pragma unroll
                                for (int ww = 0; ww < NWARP; ww++) {
                                    uint32_t tt = s2.wch[ww][tid];
                                    s2.wch[ww][tid] = crun;
                                    crun += tt;
                                }
                                s2.tch[tid] = crun;
                            }
                        }
                        __syncthreads();
                    }
#pragma unroll
                    for (int r = 0; r < ITEMS; r++) {
                        if ((vm >> r) & 1u) {
                            uint32_t b = b_cur[r], cls = b / M::DS, d = b % M::DS;
                            uint32_t pos = s2.tb[b] + sh.wcnt[w][b] + rank[r];
                            uint32_t db = sh.alpha[d];
                            uint32_t slot = SPASS ? (sh.bstart[db + 1] - 1u - pos) : (sh.bstart[db] + pos);
                            uint32_t back = (cls == M::CLS_B) ? 1u : cls + 1u;
                            A.sa[slot] = s_cur[r] - back;
                        }
                    }
                    // ---- chain rounds: round j holds the entries with run >= j, in list order
                    if (isA) {
                        uint32_t wrun = (l < 16) ? s2.wch[w][l] : 0u;      // lane j: offset of this warp in round j
#pragma unroll
                        for (int r = 0; r < ITEMS; r++) {
                            uint32_t rmax = __reduce_max_sync(FULL, rr[r]);
                            for (uint32_t j = 1; j <= rmax; j++) {
                                bool in = rr[r] >= j;
                                uint32_t bal = __ballot_sync(FULL, in);
                                uint32_t off = __shfl_sync(FULL, wrun, j);
                                if (in) {
                                    uint32_t pos = s2.cb[j] + off + __popc(bal & lt);
                                    uint32_t slot = SPASS ? (sh.bstart[cbyte + 1] - 1u - pos) : (sh.bstart[cbyte] + pos);
                                    A.sa[slot] = s_cur[r] - j;
                                }
                                if (l == j) wrun += __popc(bal);
                            }
                        }
                    }
                    __syncthreads();
                    if (tid < M::NBIN) s2.tb[tid] += sh.tcnt[tid];
                    if (isA && tid < 16) s2.cb[tid] += s2.tch[tid];
                    __syncthreads();
                }
#pragma unroll
                for (int r = 0; r < ITEMS; r++) { s_cur[r] = s_nxt[r]; b_cur[r] = b_nxt[r]; }
            }
        }
        __syncthreads();
        // ---- every block advances the fill counters and the scan state identically
        if (tid < M::DS && s2.newfill[tid] != 0xffffffffu)
            sh.fill[sh.alpha[tid]] = s2.newfill[tid] + (incl_b ? s2.tot[M::CLS_B * M::DS + tid] : 0u);
        __syncthreads();
        if (tid == 0) {
            if (chain) {
                sh.fill[cbyte] += s2.chain_add;
                if (s2.nlong > 0) {            // round R is the list of the next step of this bucket
                    sh.st_c = cbyte; sh.st_phase = 0; sh.st_begin = s2.cont_begin;
                } else if (gb.len > 0) {       // second list done in this step: on to the next bucket
                    uint32_t want = SPASS ? sh.S_or_lmsoff[cbyte] : sh.Lcnt[cbyte];
                    if (sh.fill[cbyte] != want && bid == 0) { A.err[0] = 1; A.err[1] = (uint32_t)cbyte; A.err[2] = sh.fill[cbyte]; A.err[3] = want; }
                    sh.st_c = SPASS ? cbyte - 1 : cbyte + 1; sh.st_phase = 0; sh.st_begin = 0;
                } else {                       // no second list: let the peek close the bucket
                    sh.st_c = cbyte; sh.st_phase = 0; sh.st_begin = sh.fill[cbyte];
                }
            } else {
                sh.st_c = sh.ns_c; sh.st_phase = sh.ns_phase; sh.st_begin = sh.ns_begin;
            }
        }
        grid.sync();
    }
    if (bid == 0 && tid == 0) {     // step statistics of this launch (diagnostics, tools/induce_steps.py)
        A.err[4 + (SPASS ? 3 : 0)] = bigcount; A.err[5 + (SPASS ? 3 : 0)] = smallcount; A.err[6 + (SPASS ? 3 : 0)] = bigtiles;
    }
}

}  // namespace b200sa
