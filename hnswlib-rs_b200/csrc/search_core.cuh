// search_layer on a warp: the ef-bounded best-first expansion of one layer.
// Restates /root/reference/src/hnsw.rs:922-1064 (search_layer) for one warp that owns the whole
// queue state of one query; used by the query kernel (search.cu) and the insert kernel (build.cu).
#pragma once
#include "common.cuh"

namespace hb {

struct WarpSmem {
  uint4* q4;          // query row, d4 16-byte chunks (zero padded)
  uint64_t* wbuf;     // queue keys, capacity >= ef
  uint32_t* cand_id;  // 64 slots reserved, 32 used (one chunk of neighbours)
  float* cand_d;      // 64
};

struct Stats {
  unsigned evals, expansions, adj;
};

// Unfiltered search_layer.  On return Q (in s.wbuf) holds min(ef, reached) keys ascending by
// (dist, id).  Equivalence with the two-heap reference loop (no filter):
//   * accept rule `d < d(f) || |W| < ef` + bounded W  == keep the ef smallest keys seen (SortedQueue)
//   * C.pop() nearest-first + stop when d(c) > d(f)   == expand the nearest unexpanded entry of W until none is left:
//     a candidate evicted from W has key > every key of the (full) W, and f only decreases afterwards,
//     so it would trip the stop rule the moment it is popped.
// Ties on distance are ordered by id (oracle MODE_DET).
template <class Op, int CH, int U, class Queue>
__device__ __forceinline__ void search_layer(const GraphView& g, const WarpSmem& s, Stage& stg, Visited& vis,
                                             Queue& Q, uint32_t ep, int ef, int layer, Stats& st, bool& overflow) {
  const int lane = lane_id();
  const uint4* vec4 = reinterpret_cast<const uint4*>(g.vec);
  vis.begin();
  __syncwarp();  // earlier reads of cand_id (descent, previous layer) happen-before the write below
  if (lane == 0) s.cand_id[0] = ep;
  __syncwarp();
  warp_dists<Op, CH, U>(vec4, g.d4, g.dim, s.q4, s.cand_id, 1, s.cand_d);  // hnsw.rs:952
  __syncwarp();
  st.evals += 1;
  const float d0 = Op::post(s.cand_d[0]);
  vis.test_and_set(ep, lane == 0);  // hnsw.rs:955-956
  Q.reset(s.wbuf, ef);
  Q.push_first(make_key(d0, ep));  // hnsw.rs:958-967 (ep enters W and C)
  for (;;) {
    const int idx = Q.first_unexpanded();  // C.pop(): nearest candidate (hnsw.rs:971)
    if (idx < 0) break;                    // C empty (969) or stop rule (981-993), see header
    const uint32_t c = key_id(Q.get(idx));
    int cap;
    const uint32_t* ids = list_ids(g, c, layer, cap);  // hnsw.rs:1006
    {  // pull the adjacency rows of the two most likely next candidates towards L2 while this one is expanded
      int n1, n2, n3;
      Q.next3(idx + 1, n1, n2, n3);
      const uint32_t c1 = n1 >= 0 ? key_id(Q.get(n1)) : INVALID_ID;
      const uint32_t c2 = n2 >= 0 ? key_id(Q.get(n2)) : INVALID_ID;
      const uint32_t pc = lane == 0 ? c1 : (lane == 1 ? c2 : INVALID_ID);
      if (pc != INVALID_ID) {
        int pcap;
        const uint32_t* pids = list_ids(g, pc, layer, pcap);
        if (pids) asm volatile("prefetch.global.L2 [%0];" ::"l"(pids));
      }
    }
    Q.mark_expanded(idx);
    st.expansions += 1;
    for (int base = 0; base < cap; base += 32) {  // hnsw.rs:1013, 32 neighbours at a time
      const uint32_t nid = (base + lane < cap) ? ids[base + lane] : INVALID_ID;
      const unsigned valid = __ballot_sync(FULL, nid != INVALID_ID);
      st.adj += __popc(valid);
      const bool fresh = vis.test_and_set(nid, nid != INVALID_ID);  // hnsw.rs:1016-1017
      const unsigned m = __ballot_sync(FULL, fresh);
      const int cnt = __popc(m);
      if (cnt) {
        const int pos = __popc(m & ((1u << lane) - 1u));
        if (fresh) s.cand_id[pos] = nid;
        __syncwarp();
        warp_dists_staged<Op, CH, U>(vec4, g.d4, g.dim, s.q4, s.cand_id, cnt, s.cand_d, stg);  // hnsw.rs:1026
        __syncwarp();
        st.evals += cnt;
        const uint64_t key = lane < cnt ? make_key(Op::post(s.cand_d[lane]), s.cand_id[lane]) : ~0ull;
        unsigned acc = __ballot_sync(FULL, lane < cnt && Q.accepts(key));  // hnsw.rs:1028
        while (acc) {
          const int j = __ffs(acc) - 1;
          acc &= acc - 1;
          const uint64_t kj = __shfl_sync(FULL, key, j);
          if (Q.accepts(kj)) Q.insert(kj);  // hnsw.rs:1035-1053
        }
      }
      if (valid != FULL) break;  // lists are dense prefixes terminated by INVALID_ID
    }
    if (vis.overflowing()) {
      overflow = true;
      break;
    }
  }
}

}  // namespace hb
