// rp_dbbuild_merge.h -- host merge of the per-batch results of the phylo-k-mer generation.  A build whose tuples
// do not fit one pass runs in batches of consecutive nodes; every batch comes back as its sorted, max-reduced
// (code << 16 | node, score) pairs.  Two batches can hold the same pair key (several ancestral nodes map to one
// original node id), so the merge applies addTuple's rule again: the maximum per (k-mer, node)
// (CustomHash_v4_FastUtil81.java:73-90).  Plain C++: tests compile it with g++.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <utility>
#include <vector>

namespace rp {

struct BatchPairs {
  std::vector<uint64_t> key;   // code << 16 | node, ascending, distinct within the batch
  std::vector<float> score;
};

// -> the CSR arrays rp_db_load takes: keys ascending, a key's postings by ascending node id
inline void merge_batches(const std::vector<BatchPairs>& batches, std::vector<uint64_t>& keys, std::vector<uint64_t>& offsets,
                          std::vector<uint16_t>& post_node, std::vector<float>& post_score) {
  size_t total = 0;
  for (const BatchPairs& b : batches) total += b.key.size();
  std::vector<std::pair<uint64_t, float>> all;
  all.reserve(total);
  for (const BatchPairs& b : batches)
    for (size_t i = 0; i < b.key.size(); i++) all.emplace_back(b.key[i], b.score[i]);
  // every batch is sorted: merge them pairwise (stable, linear) instead of sorting the concatenation
  std::vector<size_t> bounds(1, 0);
  for (const BatchPairs& b : batches) bounds.push_back(bounds.back() + b.key.size());
  auto by_key = [](const std::pair<uint64_t, float>& a, const std::pair<uint64_t, float>& b) { return a.first < b.first; };
  while (bounds.size() > 2) {
    std::vector<size_t> next(1, 0);
    for (size_t i = 0; i + 1 < bounds.size(); i += 2) {
      if (i + 2 < bounds.size()) {
        std::inplace_merge(all.begin() + bounds[i], all.begin() + bounds[i + 1], all.begin() + bounds[i + 2], by_key);
        next.push_back(bounds[i + 2]);
      } else {
        next.push_back(bounds[i + 1]);
      }
    }
    bounds.swap(next);
  }
  keys.clear(); offsets.clear(); post_node.clear(); post_score.clear();
  for (size_t i = 0; i < all.size(); i++) {
    const uint64_t pk = all[i].first, code = pk >> 16;
    if (i && all[i - 1].first == pk) {  // the same (k-mer, node) from another batch: keep the maximum
      if (all[i].second > post_score.back()) post_score.back() = all[i].second;
      continue;
    }
    if (keys.empty() || keys.back() != code) { keys.push_back(code); offsets.push_back(post_node.size()); }
    post_node.push_back((uint16_t)(pk & 0xFFFFu));
    post_score.push_back(all[i].second);
  }
  offsets.push_back(post_node.size());
}

}  // namespace rp
