// rp_xchg_plan.h -- the host-side bookkeeping of the exchange form (rp_xchg.cu): who sends how much to whom, and
// where it lands.  Pure C++ (no CUDA), so that the N > 1 logic is testable on a CPU: tests/test_sharding_gloo.py
// all_gathers random count matrices over gloo and checks that every rank's plan agrees with its peers'.
//
//   probes[(w * J + j) * W + o]   probes of rank w's sub-batch j whose key belongs to owner o  (all-gathered)
//   units [(o * W + p) * J + j]   32 B units of posting blocks owner o sends home p for p's sub-batch j  (all-gathered)
#pragma once
#include <stdint.h>

#include <vector>

namespace rp {

struct XKeyPlan {
  std::vector<uint64_t> send_cnt, recv_cnt;   // [W] keys this rank sends to / receives from each peer
  std::vector<uint64_t> send_off, recv_off;   // [W+1] start of the peer's segment in the send / receive order
  std::vector<uint64_t> seg_first;            // [W*J+1] first received probe of (source p, sub-batch j); last = total
  std::vector<uint64_t> home_first;           // [W*J] first SENT probe of (owner o, sub-batch j) inside o's stream
};
inline XKeyPlan plan_keys(int W, int J, int me, const uint64_t* probes) {
  auto S = [&](int w, int j, int o) { return probes[((size_t)w * J + j) * W + o]; };
  XKeyPlan k;
  k.send_cnt.assign(W, 0); k.recv_cnt.assign(W, 0); k.send_off.assign(W + 1, 0); k.recv_off.assign(W + 1, 0);
  for (int p = 0; p < W; p++)
    for (int j = 0; j < J; j++) { k.send_cnt[p] += S(me, j, p); k.recv_cnt[p] += S(p, j, me); }
  for (int p = 0; p < W; p++) { k.send_off[p + 1] = k.send_off[p] + k.send_cnt[p]; k.recv_off[p + 1] = k.recv_off[p] + k.recv_cnt[p]; }
  k.seg_first.clear();
  for (int p = 0; p < W; p++) {
    uint64_t i = k.recv_off[p];
    for (int j = 0; j < J; j++) { k.seg_first.push_back(i); i += S(p, j, me); }
  }
  k.seg_first.push_back(k.recv_off[W]);
  k.home_first.assign((size_t)W * J, 0);
  for (int o = 0; o < W; o++) {
    uint64_t q = 0;
    for (int j = 0; j < J; j++) { k.home_first[(size_t)o * J + j] = q; q += S(me, j, o); }
  }
  return k;
}

struct XPayPlan {   // everything in 32 B units, per sub-batch j and peer p: index j * W + p
  std::vector<uint64_t> send_off, send_cnt;   // in this rank's send buffer (as owner), what goes to home p
  std::vector<uint64_t> recv_off, recv_cnt;   // in this rank's receive buffer (as home), what comes from owner p
  uint64_t send_cap = 0, recv_cap = 0;        // largest sub-batch
  uint64_t recv_total = 0;                    // over all sub-batches
};
// direct_local: a rank reads the blocks of its own partition in place (nothing is packed or copied for p == me)
inline XPayPlan plan_payload(int W, int J, int me, bool direct_local, const uint64_t* units) {
  auto U = [&](int o, int p, int j) { return units[((size_t)o * W + p) * J + j]; };
  XPayPlan y;
  y.send_off.assign((size_t)J * W, 0); y.send_cnt = y.send_off; y.recv_off = y.send_off; y.recv_cnt = y.send_off;
  for (int j = 0; j < J; j++) {
    uint64_t out = 0, in = 0;
    for (int p = 0; p < W; p++) {
      const bool skip = direct_local && p == me;
      y.send_off[(size_t)j * W + p] = out;
      y.recv_off[(size_t)j * W + p] = in;
      y.send_cnt[(size_t)j * W + p] = skip ? 0 : U(me, p, j);
      y.recv_cnt[(size_t)j * W + p] = skip ? 0 : U(p, me, j);
      out += y.send_cnt[(size_t)j * W + p];
      in += y.recv_cnt[(size_t)j * W + p];
    }
    if (out > y.send_cap) y.send_cap = out;
    if (in > y.recv_cap) y.recv_cap = in;
    y.recv_total += in;
  }
  return y;
}

}  // namespace rp
