/*
 * oracle/sais_oracle.c -- CPU restatement of the BurntSushi/suffix hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs may load this library.
 * The product (libb200sa.so) never links, loads or calls anything in oracle/.
 *
 * Why a restatement: the reference is Rust and no rustc/cargo exists in this
 * image or on the GPU box, so the reference itself cannot be
 * compiled into oracle/_ref.  All arithmetic of the path lives in ONE file,
 * /root/reference/src/table.rs (no third-party dependency), restated here in
 * plain C, phase by phase, same layouts (u32 SA, 1-byte type array, Bins
 * tables) so that it doubles as the timed single-core CPU baseline.
 *
 * Parity pinning: the suffix array of a string is unique, so this oracle is
 * pinned by (1) the reference's own KATs in tests/tests.rs:22-70,152-212 and
 * the doc tests (src/lib.rs:17-18, src/table.rs:220-221), all reproduced in
 * tests/golden/kat.json, (2) the reference's own test oracle `naive_table`
 * (src/table.rs:367-376), restated as oracle_naive_sa below and compared with
 * oracle_sais on every vector, and (3) SHA-256 goldens of the two reference
 * fixtures.  See tests/test_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- src/table.rs:580-585 SuffixType ---- */
enum { TY_ASC = 0, TY_DESC = 1, TY_VALLEY = 2 };

/* table.rs:641-643 is_asc: Ascending | Valley */
static inline int ty_is_asc(uint8_t t) { return t != TY_DESC; }
/* table.rs:663-669 PartialEq: Valley == Ascending */
static inline int ty_equal(uint8_t a, uint8_t b) { return ty_is_asc(a) == ty_is_asc(b); }

/* ---- src/table.rs:671-750 Bins ---- */
typedef struct {
    uint32_t *alphas; size_t nalphas, cap_alphas;
    uint32_t *sizes;  size_t nsizes, cap_sizes;
    uint32_t *ptrs;   size_t nptrs;
} bins_t;

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/* table.rs:678-684 Bins::new (capacity 10,000) */
static void bins_init(bins_t *b)
{
    memset(b, 0, sizeof(*b));
    b->cap_alphas = 10000; b->alphas = (uint32_t *)malloc(b->cap_alphas * sizeof(uint32_t));
    b->cap_sizes = 10000;  b->sizes = (uint32_t *)calloc(b->cap_sizes, sizeof(uint32_t));
}
static void bins_free(bins_t *b) { free(b->alphas); free(b->sizes); free(b->ptrs); }

/* table.rs:739-744 inc_size (resizes on demand) */
static inline void bins_inc_size(bins_t *b, uint32_t c)
{
    if ((size_t)c >= b->nsizes) {
        size_t want = (size_t)c + 1;
        if (want > b->cap_sizes) {
            size_t cap = b->cap_sizes * 2; if (cap < want) cap = want;
            b->sizes = (uint32_t *)realloc(b->sizes, cap * sizeof(uint32_t));
            b->cap_sizes = cap;
        }
        memset(b->sizes + b->nsizes, 0, (want - b->nsizes) * sizeof(uint32_t));
        b->nsizes = want;
    }
    b->sizes[c] += 1;
}
static inline void bins_push_alpha(bins_t *b, uint32_t c)
{
    if (b->nalphas == b->cap_alphas) {
        b->cap_alphas *= 2;
        b->alphas = (uint32_t *)realloc(b->alphas, b->cap_alphas * sizeof(uint32_t));
    }
    b->alphas[b->nalphas++] = c;
}
/* table.rs:706-712 */
static void bins_find_head_pointers(bins_t *b)
{
    uint32_t sum = 0;
    for (size_t k = 0; k < b->nalphas; k++) {
        uint32_t c = b->alphas[k];
        b->ptrs[c] = sum;
        sum += b->sizes[c];
    }
}
/* table.rs:714-720 */
static void bins_find_tail_pointers(bins_t *b)
{
    uint32_t sum = 0;
    for (size_t k = 0; k < b->nalphas; k++) {
        uint32_t c = b->alphas[k];
        sum += b->sizes[c];
        b->ptrs[c] = sum - 1;
    }
}
/* table.rs:723-727 */
static inline void bins_head_insert(bins_t *b, uint32_t *sa, uint32_t i, uint32_t c)
{
    uint32_t *p = &b->ptrs[c];
    sa[*p] = i;
    *p += 1;
}
/* table.rs:730-736 (saturates at 0) */
static inline void bins_tail_insert(bins_t *b, uint32_t *sa, uint32_t i, uint32_t c)
{
    uint32_t *p = &b->ptrs[c];
    sa[*p] = i;
    if (*p > 0) *p -= 1;
}

/* ---- the level routine, instantiated for LexNames (u32) then Utf8 (bytes) ---- */
static void names_sais(uint32_t *sa, uint8_t *types, bins_t *bins, const uint32_t *t, uint32_t n);

#define TEXT_T uint32_t
#define FN(x) names_##x##_impl
#include "sais_level.inc"
#undef TEXT_T
#undef FN
static void names_sais(uint32_t *sa, uint8_t *types, bins_t *bins, const uint32_t *t, uint32_t n)
{
    names_sais_impl(sa, types, bins, t, n);
}

#define TEXT_T uint8_t
#define FN(x) bytes_##x##_impl
#include "sais_level.inc"
#undef TEXT_T
#undef FN

/* ---- exported entry points ---- */

/* src/table.rs:378-386 sais_table.  Returns 0, or -1 if n > u32::MAX (the
 * reference asserts, :380) or on allocation failure. */
int oracle_sais(const uint8_t *text, uint64_t n, uint32_t *sa)
{
    if (n > 0xFFFFFFFFull) return -1;
    if (n == 0) return 0;
    uint8_t *types = (uint8_t *)malloc((size_t)n);   /* SuffixTypes::new, :588-590 */
    if (!types) return -1;
    memset(types, TY_ASC, (size_t)n);
    bins_t bins; bins_init(&bins);
    bytes_sais_impl(sa, types, &bins, text, (uint32_t)n);
    bins_free(&bins);
    free(types);
    return 0;
}

/* Suffix types of the level-0 text as the reference computes them
 * (src/table.rs:592-615): out[i] in {0 Ascending, 1 Descending, 2 Valley}. */
int oracle_types(const uint8_t *text, uint64_t n, uint8_t *out)
{
    if (n > 0xFFFFFFFFull) return -1;
    bytes_types_compute_impl(out, text, (uint32_t)n);
    return 0;
}

/* src/table.rs:367-376 naive_table: sort suffix slices; Rust slice `cmp` is
 * unsigned-byte lexicographic with a proper prefix ordering first. */
static const uint8_t *g_text; static uint64_t g_n;
static int cmp_suffix(const void *pa, const void *pb)
{
    uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb;
    uint64_t la = g_n - a, lb = g_n - b, l = la < lb ? la : lb;
    int r = memcmp(g_text + a, g_text + b, (size_t)l);
    if (r) return r;
    return (la > lb) - (la < lb);
}
int oracle_naive_sa(const uint8_t *text, uint64_t n, uint32_t *sa)
{
    if (n > 0xFFFFFFFFull) return -1;
    for (uint64_t i = 0; i < n; i++) sa[i] = (uint32_t)i;
    g_text = text; g_n = n;
    qsort(sa, (size_t)n, sizeof(uint32_t), cmp_suffix);
    return 0;
}

/* src/table.rs:363-365 lcp_len */
static inline uint32_t lcp_len(const uint8_t *a, uint64_t la, const uint8_t *b, uint64_t lb)
{
    uint64_t l = la < lb ? la : lb, k = 0;
    while (k < l && a[k] == b[k]) k++;
    return (uint32_t)k;
}
/* src/table.rs:348-361 lcp_lens_quadratic: lcp[0]=0, lcp[i]=lcp_len(suf[i-1],suf[i]).
 * (`lcp_lens`, :130-138, also fills a never-read inverse array; not restated
 * in the output but timed by oracle_lcp_lens below.) */
int oracle_lcp_quadratic(const uint8_t *text, uint64_t n, const uint32_t *sa, uint32_t *lcp)
{
    if (n == 0) return 0;
    lcp[0] = 0;
    for (uint64_t i = 0; i + 1 < n; i++)
        lcp[i + 1] = lcp_len(text + sa[i], n - sa[i], text + sa[i + 1], n - sa[i + 1]);
    return 0;
}
/* src/table.rs:130-138 lcp_lens, including the wasted inverse fill (for the
 * timed CPU baseline). */
int oracle_lcp_lens(const uint8_t *text, uint64_t n, const uint32_t *sa, uint32_t *lcp)
{
    uint32_t *inverse = (uint32_t *)calloc((size_t)(n ? n : 1), sizeof(uint32_t));
    if (!inverse) return -1;
    for (uint64_t r = 0; r < n; r++) inverse[sa[r]] = (uint32_t)r;
    volatile uint32_t sink = n ? inverse[0] : 0; (void)sink;
    int rc = oracle_lcp_quadratic(text, n, sa, lcp);
    free(inverse);
    return rc;
}

/* Byte-level Kasai (the algorithm of the commented-out lcp_lens_linear,
 * src/table.rs:314-346, at byte granularity).  Used only to cross-check that
 * it equals the quadratic definition. */
int oracle_lcp_kasai(const uint8_t *text, uint64_t n, const uint32_t *sa, uint32_t *lcp)
{
    if (n == 0) return 0;
    uint32_t *inv = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    if (!inv) return -1;
    for (uint64_t r = 0; r < n; r++) inv[sa[r]] = (uint32_t)r;
    uint64_t h = 0;
    lcp[0] = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t r = inv[i];
        if (r == 0) { h = 0; continue; }
        uint64_t j = sa[r - 1];
        while (i + h < n && j + h < n && text[i + h] == text[j + h]) h++;
        lcp[r] = (uint32_t)h;
        if (h > 0) h--;
    }
    free(inv);
    return 0;
}

/* O(n) suffix-array verifier for sizes where oracle_sais is too slow to be a
 * per-test oracle (BASELINE config 4, 1 GB).  `sa` is the suffix array of
 * `text` -- i.e. equals the order of src/table.rs:374 `text[a..].cmp(&text[b..])`,
 * and by uniqueness the output of the reference's sais() -- iff
 *   (1) sa is a permutation of 0..n-1, and
 *   (2) for every i > 0 with a = sa[i-1], b = sa[i]:  T[a] < T[b], or T[a] == T[b]
 *       and suffix a+1 precedes suffix b+1 (the empty suffix precedes all),
 *       decided by the inverse permutation.
 * Returns 0 if valid, 1 + (index of the first violation) otherwise, -1 on OOM. */
int64_t oracle_verify_sa(const uint8_t *text, uint64_t n, const uint32_t *sa)
{
    if (n == 0) return 0;
    uint32_t *inv = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    if (!inv) return -1;
    memset(inv, 0xff, (size_t)n * sizeof(uint32_t));
    for (uint64_t r = 0; r < n; r++) {
        if (sa[r] >= n || inv[sa[r]] != 0xffffffffu) { free(inv); return 1 + (int64_t)r; }
        inv[sa[r]] = (uint32_t)r;
    }
    for (uint64_t i = 1; i < n; i++) {
        uint64_t a = sa[i - 1], b = sa[i];
        int ok;
        if (text[a] != text[b]) ok = text[a] < text[b];
        else if (a + 1 == n) ok = 1;             /* "c" precedes "c..." */
        else if (b + 1 == n) ok = 0;
        else ok = inv[a + 1] < inv[b + 1];
        if (!ok) { free(inv); return 1 + (int64_t)i; }
    }
    free(inv);
    return 0;
}

/* ---- queries: src/table.rs:197-293, 900-914 ---- */

/* table.rs:900-914 binary_search: first index whose predicate is true */
static int suffix_ge_query(const uint8_t *text, uint64_t n, uint32_t s,
                           const uint8_t *q, uint64_t m)
{   /* query <= text[s..] */
    uint64_t ls = n - s, l = ls < m ? ls : m;
    int r = memcmp(q, text + s, (size_t)l);
    if (r) return r < 0;
    return m <= ls;
}
static int suffix_starts_with(const uint8_t *text, uint64_t n, uint32_t s,
                              const uint8_t *q, uint64_t m)
{
    return (n - s) >= m && memcmp(text + s, q, (size_t)m) == 0;
}
static int bytes_cmp(const uint8_t *a, uint64_t la, const uint8_t *b, uint64_t lb)
{
    uint64_t l = la < lb ? la : lb;
    int r = memcmp(a, b, (size_t)l);
    if (r) return r;
    return (la > lb) - (la < lb);
}

/* table.rs:223-259 positions: writes [*start,*end) into the SA */
int oracle_positions(const uint8_t *text, uint64_t n, const uint32_t *sa,
                     const uint8_t *q, uint64_t m, uint64_t *start, uint64_t *end)
{
    *start = *end = 0;
    if (n == 0 || m == 0) return 0;
    /* :230-232 early outs */
    if (bytes_cmp(q, m, text + sa[0], n - sa[0]) < 0 &&
        !suffix_starts_with(text, n, sa[0], q, m)) return 0;
    if (bytes_cmp(q, m, text + sa[n - 1], n - sa[n - 1]) > 0) return 0;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {                       /* :244-246 */
        uint64_t mid = (lo + hi) / 2;
        if (suffix_ge_query(text, n, sa[mid], q, m)) hi = mid; else lo = mid + 1;
    }
    uint64_t s = lo;
    lo = 0; hi = n - s;
    while (lo < hi) {                       /* :247-250 */
        uint64_t mid = (lo + hi) / 2;
        if (!suffix_starts_with(text, n, sa[s + mid], q, m)) hi = mid; else lo = mid + 1;
    }
    *start = s; *end = s + lo;
    return 0;
}

/* table.rs:279-293 any_position: slice::binary_search_by on the first
 * min(|suffix|, m) bytes; returns 1 and *pos on a hit, 0 otherwise.  (Which
 * of several matches is returned is unspecified in the reference.) */
int oracle_any_position(const uint8_t *text, uint64_t n, const uint32_t *sa,
                        const uint8_t *q, uint64_t m, uint32_t *pos)
{
    if (m == 0) return 0;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        uint32_t s = sa[mid];
        uint64_t ls = n - s, l = ls < m ? ls : m;
        int r = bytes_cmp(text + s, l, q, m);
        if (r == 0) { *pos = s; return 1; }
        if (r < 0) lo = mid + 1; else hi = mid;
    }
    return 0;
}
