// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// RHeap: restatement of Rust's std::collections::BinaryHeap (max-heap) *including its
// tie behaviour*, because the reference keeps search_layer's two queues in BinaryHeaps whose
// Ord compares the distance only (/root/reference/src/hnsw.rs:273-297, used at 940-1053,
// 1364-1409, 1544).  Which of several equal-distance items survives a pop is decided by the
// sift rules below, so the oracle's `std` mode replays them literally.
//
// Source of the rules: Rust std `alloc::collections::binary_heap` (push -> sift_up stops on
// `elem <= parent`; pop -> swap last into root, sift_down_to_bottom picking the right child
// when `left <= right`, then sift_up; into_sorted_vec -> swap(0,end) + sift_down_range;
// retain -> Vec::retain + rebuild_tail).  Restated from the published algorithm; the Rust
// toolchain is absent here so this restatement is pinned only by the unit checks in
// tests/test_oracle.py (test_rheap_*) ("parity unpinned" w.r.t. the real std).
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace oracle {

template <class Item, class Cmp>
class RHeap {
 public:
  explicit RHeap(Cmp c) : cmp_(c) {}
  size_t size() const { return v_.size(); }
  bool empty() const { return v_.empty(); }
  const Item& peek() const { return v_[0]; }
  const std::vector<Item>& items() const { return v_; }
  void clear() { v_.clear(); }

  void push(const Item& it) {
    size_t old = v_.size();
    v_.push_back(it);
    sift_up(0, old);
  }

  Item pop() {
    Item item = v_.back();
    v_.pop_back();
    if (!v_.empty()) {
      std::swap(item, v_[0]);
      sift_down_to_bottom(0);
    }
    return item;
  }

  // ascending order, like BinaryHeap::into_sorted_vec
  std::vector<Item> into_sorted_vec() {
    size_t end = v_.size();
    while (end > 1) {
      end -= 1;
      std::swap(v_[0], v_[end]);
      sift_down_range(0, end);
    }
    return std::move(v_);
  }

  template <class Pred>
  void retain(Pred keep) {
    size_t first_removed = v_.size();
    size_t w = 0;
    for (size_t i = 0; i < v_.size(); ++i) {
      if (keep(v_[i])) {
        if (w != i) v_[w] = v_[i];
        ++w;
      } else if (i < first_removed) {
        first_removed = i;
      }
    }
    v_.resize(w);
    rebuild_tail(first_removed);
  }

 private:
  bool le(const Item& a, const Item& b) const { return cmp_(a, b) <= 0; }
  bool lt(const Item& a, const Item& b) const { return cmp_(a, b) < 0; }
  bool ge(const Item& a, const Item& b) const { return cmp_(a, b) >= 0; }

  size_t sift_up(size_t start, size_t pos) {
    Item elt = v_[pos];
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (le(elt, v_[parent])) break;
      v_[pos] = v_[parent];
      pos = parent;
    }
    v_[pos] = elt;
    return pos;
  }

  void sift_down_range(size_t pos, size_t end) {
    Item elt = v_[pos];
    size_t child = 2 * pos + 1;
    size_t lim = end >= 2 ? end - 2 : 0;  // end.saturating_sub(2)
    while (child <= lim && end >= 2) {
      if (le(v_[child], v_[child + 1])) child += 1;
      if (ge(elt, v_[child])) {
        v_[pos] = elt;
        return;
      }
      v_[pos] = v_[child];
      pos = child;
      child = 2 * pos + 1;
    }
    if (end >= 1 && child == end - 1 && lt(elt, v_[child])) {
      v_[pos] = v_[child];
      pos = child;
    }
    v_[pos] = elt;
  }

  void sift_down_to_bottom(size_t pos) {
    size_t end = v_.size();
    size_t start = pos;
    Item elt = v_[pos];
    size_t child = 2 * pos + 1;
    size_t lim = end >= 2 ? end - 2 : 0;
    while (child <= lim && end >= 2) {
      if (le(v_[child], v_[child + 1])) child += 1;
      v_[pos] = v_[child];
      pos = child;
      child = 2 * pos + 1;
    }
    if (end >= 1 && child == end - 1) {
      v_[pos] = v_[child];
      pos = child;
    }
    v_[pos] = elt;
    sift_up(start, pos);
  }

  void rebuild() {
    size_t n = v_.size() / 2;
    while (n > 0) {
      n -= 1;
      sift_down_range(n, v_.size());
    }
  }

  static size_t log2_fast(size_t x) {
    size_t r = 0;
    while (x >>= 1) ++r;
    return r;
  }

  void rebuild_tail(size_t start) {
    if (start == v_.size()) return;
    size_t tail_len = v_.size() - start;
    bool better_to_rebuild;
    if (start < tail_len) {
      better_to_rebuild = true;
    } else if (v_.size() <= 2048) {
      better_to_rebuild = 2 * v_.size() < tail_len * log2_fast(start);
    } else {
      better_to_rebuild = 2 * v_.size() < tail_len * 11;
    }
    if (better_to_rebuild) {
      rebuild();
    } else {
      for (size_t i = start; i < v_.size(); ++i) sift_up(0, i);
    }
  }

  Cmp cmp_;
  std::vector<Item> v_;
};

}  // namespace oracle
