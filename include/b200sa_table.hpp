// b200sa_table.hpp -- C++ host-side mirror of the reference's `SuffixTable`
// (/root/reference/src/table.rs:54-294) above the C-ABI of b200sa.h.
//
// The reference is compiled (Rust) code and no Rust toolchain exists in this
// image, so the host side above the C-ABI is C++: same method names, argument
// meaning and error behaviour as the reference API.  Construction and
// lcp_lens() run on the GPU through libb200sa.so; the O(m log n) queries stay
// on the host exactly as in the reference.  Where the reference panics
// (text > 2^32-1 bytes, src/table.rs:380; from_parts length mismatch, :117)
// this class throws.  There is no CPU construction path: a failing CUDA call
// throws std::runtime_error.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "b200sa.h"

namespace b200sa {

class Context {
  public:
    explicit Context(int device = 0) {
        int rc = b200sa_ctx_create(device, &ctx_);
        if (rc != B200SA_OK) throw std::runtime_error(std::string("b200sa_ctx_create: ") + b200sa_strerror(rc));
    }
    ~Context() { b200sa_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    b200sa_ctx *get() const { return ctx_; }
    std::mutex &mutex() { return mu_; }
    // The lazily created default context (a context is not thread-safe; the
    // wrapper serialises calls on it, keeping `SuffixTable::new` callable from
    // many threads like the reference).
    static Context &default_context() {
        static Context c(0);
        return c;
    }

  private:
    b200sa_ctx *ctx_ = nullptr;
    std::mutex mu_;
};

class SuffixTable {
  public:
    // SuffixTable::new (src/table.rs:78-85)
    explicit SuffixTable(std::string text) : text_(std::move(text)) {
        table_.resize(text_.size());
        Context &c = Context::default_context();
        std::lock_guard<std::mutex> lk(c.mutex());
        check(c, b200sa_build(c.get(), bytes(), text_.size(), table_.data()));
    }
    // SuffixTable::from_parts (src/table.rs:111-119)
    static SuffixTable from_parts(std::string text, std::vector<uint32_t> table) {
        if (text.size() != table.size()) throw std::invalid_argument("from_parts: text.len() != table.len()");
        return SuffixTable(std::move(text), std::move(table));
    }
    // SuffixTable::into_parts (src/table.rs:125-127)
    std::pair<std::string, std::vector<uint32_t>> into_parts() && { return {std::move(text_), std::move(table_)}; }

    // SuffixTable::lcp_lens (src/table.rs:130-138; values of lcp_lens_quadratic, :348-361)
    std::vector<uint32_t> lcp_lens() const {
        std::vector<uint32_t> lcp(table_.size());
        Context &c = Context::default_context();
        std::lock_guard<std::mutex> lk(c.mutex());
        check(c, b200sa_lcp(c.get(), bytes(), text_.size(), table_.data(), lcp.data()));
        return lcp;
    }

    const std::vector<uint32_t> &table() const { return table_; }   // :142-144
    const std::string &text() const { return text_; }               // :148-150
    size_t len() const { return table_.size(); }                    // :156-158
    bool is_empty() const { return table_.empty(); }                // :162-164
    std::string_view suffix(size_t i) const { return std::string_view(text_).substr(table_[i]); }        // :168-170
    std::string_view suffix_bytes(size_t i) const { return suffix(i); }                                  // :174-176
    bool operator==(const SuffixTable &o) const { return text_ == o.text_ && table_ == o.table_; }       // derive(PartialEq), :54

    // contains (src/table.rs:197-199)
    bool contains(std::string_view query) const { return any_position(query).has_value(); }

    // positions (src/table.rs:223-259): [first,last) into table(), SA order.
    std::pair<const uint32_t *, const uint32_t *> positions(std::string_view query) const {
        const uint32_t *base = table_.data();
        std::string_view text(text_);
        if (text.empty() || query.empty()) return {base, base};
        std::string_view s0 = suffix(0), sl = suffix(len() - 1);
        if ((query < s0 && s0.substr(0, query.size()) != query) || query > sl) return {base, base};
        size_t start = binary_search(0, len(), [&](uint32_t s) { return query <= text.substr(s); });
        size_t cnt = binary_search(start, len(), [&](uint32_t s) { return text.substr(s).substr(0, query.size()) != query; });
        return {base + start, base + start + cnt};
    }

    // any_position (src/table.rs:279-293)
    std::optional<uint32_t> any_position(std::string_view query) const {
        if (query.empty()) return std::nullopt;
        std::string_view text(text_);
        size_t lo = 0, hi = len();
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            std::string_view head = text.substr(table_[mid]).substr(0, query.size());
            int c = head.compare(query);
            if (c == 0) return table_[mid];
            if (c < 0) lo = mid + 1; else hi = mid;
        }
        return std::nullopt;
    }

  private:
    SuffixTable(std::string text, std::vector<uint32_t> table) : text_(std::move(text)), table_(std::move(table)) {}
    const uint8_t *bytes() const { return reinterpret_cast<const uint8_t *>(text_.data()); }
    static void check(Context &c, int rc) {
        if (rc != B200SA_OK)
            throw std::runtime_error(std::string("b200sa: ") + b200sa_strerror(rc) + ": " + b200sa_last_error(c.get()));
    }
    // binary_search (src/table.rs:900-914): number of leading elements of
    // table[from..to) for which pred is false (pred is monotone).
    template <class P>
    size_t binary_search(size_t from, size_t to, P pred) const {
        size_t left = 0, right = to - from;
        while (left < right) {
            size_t mid = (left + right) / 2;
            if (pred(table_[from + mid])) right = mid; else left = mid + 1;
        }
        return left;
    }
    std::string text_;
    std::vector<uint32_t> table_;
};

}  // namespace b200sa
