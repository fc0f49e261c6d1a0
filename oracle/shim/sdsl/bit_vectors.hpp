// Minimal stand-in for the two sdsl-lite types the reference's index builder uses
// (sdsl::bit_vector, sdsl::rank_support_v<>).  TEST INFRASTRUCTURE ONLY: it lets the
// unmodified reference sources under /root/reference compile into oracle/_ref/ without
// the un-vendored sdsl-lite dependency (reference CMakeLists.txt:27-114 fetches it over
// the network).  Nothing in the product path includes this file.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstddef>
#include <istream>
#include <ostream>
#include <vector>

namespace sdsl {

class bit_vector {
  public:
    typedef uint64_t size_type;
    class reference {
        uint64_t* w_; uint64_t m_;
      public:
        reference(uint64_t* w, uint64_t m) : w_(w), m_(m) {}
        reference& operator=(bool b) { if (b) *w_ |= m_; else *w_ &= ~m_; return *this; }
        reference& operator=(const reference& o) { return *this = static_cast<bool>(o); }
        operator bool() const { return (*w_ & m_) != 0; }
    };
    class const_iterator {
        const bit_vector* v_; size_type i_;
      public:
        const_iterator(const bit_vector* v, size_type i) : v_(v), i_(i) {}
        bool operator*() const { return (*v_)[i_]; }
        const_iterator& operator++() { ++i_; return *this; }
        bool operator!=(const const_iterator& o) const { return i_ != o.i_; }
        bool operator==(const const_iterator& o) const { return i_ == o.i_; }
    };
    bit_vector() : n_(0) {}
    bit_vector(size_type n, bool v = false) : n_(n), w_((n + 63) / 64 + 1, v ? ~0ULL : 0ULL) {}
    size_type size() const { return n_; }
    bool empty() const { return n_ == 0; }
    reference operator[](size_type i) { return reference(&w_[i >> 6], 1ULL << (i & 63)); }
    bool operator[](size_type i) const { return (w_[i >> 6] >> (i & 63)) & 1ULL; }
    const_iterator begin() const { return const_iterator(this, 0); }
    const_iterator end() const { return const_iterator(this, n_); }
    const uint64_t* data() const { return w_.data(); }
    void resize(size_type n) { n_ = n; w_.resize((n + 63) / 64 + 1, 0); }
    // sdsl on-disk format for int_vector<1>: u64 bit length, then the 64-bit words.
    void load(std::istream& in) {
        uint64_t n = 0; in.read(reinterpret_cast<char*>(&n), 8);
        n_ = n; w_.assign((n + 63) / 64 + 1, 0);
        in.read(reinterpret_cast<char*>(w_.data()), ((n + 63) / 64) * 8);
    }
    size_type serialize(std::ostream& out) const {
        out.write(reinterpret_cast<const char*>(&n_), 8);
        out.write(reinterpret_cast<const char*>(w_.data()), ((n_ + 63) / 64) * 8);
        return 8 + ((n_ + 63) / 64) * 8;
    }
  private:
    size_type n_;
    std::vector<uint64_t> w_;
};

template <uint8_t t_b = 1, uint8_t t_pat_len = 1>
class rank_support_v {
  public:
    typedef uint64_t size_type;
    rank_support_v() : v_(nullptr) {}
    explicit rank_support_v(const bit_vector* v) : v_(v) {
        if (!v) return;
        size_type nw = (v->size() + 63) / 64 + 1;
        c_.resize(nw + 1);
        size_type s = 0;
        for (size_type i = 0; i < nw; i++) { c_[i] = s; s += __builtin_popcountll(v->data()[i]); }
        c_[nw] = s;
    }
    // number of set bits in [0, i)
    size_type operator()(size_type i) const {
        size_type w = i >> 6, b = i & 63;
        size_type r = c_[w];
        if (b) r += __builtin_popcountll(v_->data()[w] & ((1ULL << b) - 1));
        return r;
    }
    size_type rank(size_type i) const { return (*this)(i); }
    void set_vector(const bit_vector* v) { v_ = v; }
  private:
    const bit_vector* v_;
    std::vector<size_type> c_;
};

}  // namespace sdsl
