// C++ mirror of the reference's tests/tests.rs KATs against
// include/b200sa_table.hpp (runs on the GPU box; compile-checked on CPU).
#include <cassert>
#include <cstdio>
#include <string>
#include <vector>

#include "b200sa_table.hpp"

using b200sa::SuffixTable;

static std::vector<uint32_t> pos(const SuffixTable &t, const char *q) {
    auto r = t.positions(q);
    return std::vector<uint32_t>(r.first, r.second);
}
#define CHECK(x) do { if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } } while (0)

int main() {
    using V = std::vector<uint32_t>;
    // tests/tests.rs:22-70 (values from SURVEY.md Appendix B)
    CHECK(SuffixTable("apple").table() == (V{0, 4, 3, 2, 1}));
    CHECK(SuffixTable("banana").table() == (V{5, 3, 1, 0, 4, 2}));
    CHECK(SuffixTable("mississippi").table() == (V{10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2}));
    CHECK(SuffixTable("tgtgtgtgcaccg").table() == (V{9, 8, 10, 11, 12, 7, 5, 3, 1, 6, 4, 2, 0}));
    CHECK(SuffixTable("").table().empty());
    CHECK(SuffixTable("a").table() == (V{0}));
    CHECK(SuffixTable("ab").table() == (V{0, 1}));
    CHECK(SuffixTable("aa").table() == (V{1, 0}));
    CHECK(SuffixTable(std::string(1, '\0')).table() == (V{0}));
    CHECK(SuffixTable("\xE2\x98\x83" "abc" "\xE2\x98\x83").table() == (V{3, 4, 5, 8, 2, 7, 1, 6, 0}));
    CHECK(SuffixTable("banana").lcp_lens() == (V{0, 1, 3, 0, 0, 2}));
    CHECK(SuffixTable("mississippi").lcp_lens() == (V{0, 1, 1, 4, 0, 0, 1, 0, 2, 1, 3}));
    // tests/tests.rs:100-213
    SuffixTable e("");
    CHECK(pos(e, "").empty() && !e.contains("") && pos(e, "a").empty() && !e.contains("ab"));
    SuffixTable a("a");
    CHECK(pos(a, "").empty() && pos(a, "b").empty() && pos(a, "a") == (V{0}) && a.contains("a") && !a.contains("b"));
    CHECK(pos(SuffixTable("ab"), "b") == (V{1}));
    CHECK(pos(SuffixTable("aa"), "a") == (V{1, 0}));
    CHECK(pos(SuffixTable("zzzzzaazzzzz"), "a") == (V{5, 6}));
    CHECK(pos(SuffixTable("zzzzabczzzzzabczzzzzz"), "abc") == (V{4, 12}));
    CHECK(pos(SuffixTable("az"), "mnomnomnomnomnomnomno").empty());
    CHECK(pos(SuffixTable("zz"), "mnomnomnomnomnomnomno").empty());
    CHECK(pos(SuffixTable("aa"), "mnomnomnomnomnomnomno").empty());
    SuffixTable q("The quick brown fox was very quick.");
    CHECK(pos(q, "quick") == (V{4, 29}));
    auto ap = q.any_position("quick");
    CHECK(ap && (*ap == 4 || *ap == 29));
    SuffixTable sn("\xE2\x98\x83" "abc" "\xE2\x98\x83");
    CHECK(sn.contains("\xE2\x98\x83") && pos(sn, "\xE2\x98\x83") == (V{6, 0}));
    // parts round trip, tests/tests.rs:171-179
    SuffixTable p("po\xC3\xABzie");
    SuffixTable p2 = p;
    auto parts = std::move(p2).into_parts();
    CHECK(p == SuffixTable::from_parts(parts.first, parts.second));
    CHECK(p.len() == 7 && !p.is_empty() && p.suffix(0) == "e");
    std::printf("cpp mirror ok\n");
    return 0;
}
