// Batched insert kernels.  One warp owns one new point.
//   phase A (insert_search_kernel): upper-layer ef=1 descent, per-layer search_layer(ef_construction),
//            heuristic neighbour selection, write of the new point's own lists
//            == /root/reference/src/hnsw.rs:1110-1205 (insert_slice) + 1299-1421 (select_neighbours)
//   phase B (insert_link_kernel): reverse links under a per-point lock
//            == /root/reference/src/hnsw.rs:1241-1289 (reverse_update_neighborhood_simple),
//            including its quirk that every back-link is filed under the NEW point's level (1257).
// The reference races inserts under parking_lot locks on a rayon pool (hnsw.rs:1224-1238); here a
// batch of inserts searches the graph as it stood at the start of the batch (phase A is read-only
// on other points' lists) and links afterwards; see DESIGN.md "batched insert".
#include "kernels.h"
#include "search_core.cuh"

namespace hb {

// dists from the vector of point `e` to kept[0..cnt): stage e's row as the "query"
template <class Op, int CH, int U>
__device__ __forceinline__ void dists_from_point(const GraphView& g, uint4* qe4, uint32_t e, const uint32_t* kept,
                                                 int cnt, float* out) {
  const uint4* vec4 = reinterpret_cast<const uint4*>(g.vec);
  const int lane = lane_id();
  __syncwarp();
  for (int i = lane; i < g.d4; i += 32) qe4[i] = __ldg(vec4 + (size_t)e * g.d4 + i);
  __syncwarp();
  warp_dists<Op, CH, U>(vec4, g.d4, g.dim, qe4, kept, cnt, out);
  __syncwarp();
}

template <class Op, int CH, int U, int NS>
__global__ void __launch_bounds__(BUILD_THREADS) insert_search_kernel(InsertParams p) {
  using Queue = typename QueueSel<NS>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GraphView& g = p.g;
  unsigned char* base = smem_raw + (size_t)warp * p.smem_per_warp;
  // layout: q4 | qe4 | wbuf[ef_c] | cand_id[32] cand_d[32] | sel_id[nbmax] sel_d[nbmax] tmp[nbmax] | disc[ef_c] (u16)
  WarpSmem s;
  size_t off = stage_bytes(g.d4);
  Stage stg;
  stg.buf = off ? reinterpret_cast<uint4*>(base) : nullptr;
  s.q4 = reinterpret_cast<uint4*>(base + off);
  off += (size_t)g.d4 * 16;
  uint4* qe4 = reinterpret_cast<uint4*>(base + off);
  off += (size_t)g.d4 * 16;
  s.wbuf = reinterpret_cast<uint64_t*>(base + off);
  off += (size_t)p.q_smem * 8;
  s.cand_id = reinterpret_cast<uint32_t*>(base + off);
  off += 256;
  s.cand_d = reinterpret_cast<float*>(base + off);
  off += 256;
  stg.bar = reinterpret_cast<uint64_t*>(base + off);
  off += 16;
  stg.phase = 0;
  if (lane == 0) mbar_init(stg.bar, 1);
  __syncwarp();
  const int nbmax = g.deg0;
  uint32_t* sel_id = reinterpret_cast<uint32_t*>(base + off);
  off += (size_t)nbmax * 4;
  float* sel_d = reinterpret_cast<float*>(base + off);
  off += (size_t)nbmax * 4;
  float* tmp = reinterpret_cast<float*>(base + off);
  off += (size_t)nbmax * 4;
  uint16_t* disc = reinterpret_cast<uint16_t*>(base + off);

  const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + warp;
  Visited vis;
  vis.init(p.vis, slot);
  Queue Q;
  Q.reset(s.wbuf, p.ef_c);
  Stats st{0, 0, 0};
  const uint4* vec4 = reinterpret_cast<const uint4*>(g.vec);

  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(p.work_counter, 1u);
    wi = __shfl_sync(FULL, wi, 0);
    if (wi >= p.count) break;
    const uint32_t x = p.first + wi;
    const int lv = g.level[x];
    const unsigned mask = p.layer_mask[wi];
    for (int i = lane; i < g.d4; i += 32) s.q4[i] = __ldg(vec4 + (size_t)x * g.d4 + i);
    __syncwarp();

    bool overflow = false;
    uint32_t cur = g.entry;
    // dist_to_entry, hnsw.rs:1110-1112
    if (lane == 0) s.cand_id[0] = cur;
    __syncwarp();
    warp_dists<Op, CH, U>(vec4, g.d4, g.dim, s.q4, s.cand_id, 1, s.cand_d);
    __syncwarp();
    float dist_to_entry = Op::post(s.cand_d[0]);
    // ---- layers above the new point's level: ef = 1 (hnsw.rs:1114-1155).  The reference also pushes
    // the result into new_point.neighbours[l] for l above its level (1140-1144); that list can never
    // be traversed (DESIGN.md "lists above a point's level") and is not materialised.
    for (int l = g.entry_level; l > lv; --l) {
      if (!((mask >> l) & 1u)) continue;  // points_by_layer[l].is_empty() => empty result (942-946)
      search_layer<Op, CH, U, Queue>(g, s, stg, vis, Q, cur, 1, l, st, overflow);
      if (overflow) break;
      const uint64_t k0 = Q.get(0);
      const float t = key_dist(k0);  // == dist(data, ep) recomputed at 1146
      if (t < dist_to_entry) {       // 1147-1150
        cur = key_id(k0);
        dist_to_entry = t;
      }
    }
    // ---- layers level..0: ef_construction search + selection (hnsw.rs:1158-1205)
    for (int l = lv; l >= 0 && !overflow; --l) {
      if (!((mask >> l) & 1u)) continue;
      search_layer<Op, CH, U, Queue>(g, s, stg, vis, Q, cur, p.ef_c, l, st, overflow);
      if (overflow) break;
      const int n = Q.n;
      const int nb = (l == 0) ? g.deg0 : g.M;  // 1177-1183
      int cnt = 0;
      // extend_candidates (1318-1362, layer 0 only): with |cand| <= nb the reference adds the neighbours of the
      // candidates that are not candidates themselves, then runs the heuristic instead of taking everything.
      // When |cand| < ef_construction the search ended because every reachable node was visited, accepted
      // (|W| < ef) and expanded, so those neighbours are all candidates already and the extension set is
      // empty; the host only allows the flag when ef_construction > 2*max_nb_connection >= |cand|.
      const bool heuristic_on_few = p.extend && l == 0;
      if (n <= nb && !heuristic_on_few) {
        // 1318-1327: few candidates, take them all nearest first
        for (int i = lane; i < n; i += 32) {
          const uint64_t k = Q.local(i);
          sel_id[i] = key_id(k);
          sel_d[i] = key_dist(k);
        }
        cnt = n;
        __syncwarp();
      } else {
        int ndisc = 0;
        for (int i = 0; i < n && cnt < nb; ++i) {  // 1365: pop nearest while |out| < nb
          const uint64_t k = Q.get(i);
          const uint32_t e = key_id(k);
          const float de = key_dist(k);
          bool keep = true;
          if (cnt > 0) {  // 1372-1376: reject when some kept d has dist(e,d) <= dist(e,q)
            dists_from_point<Op, CH, U>(g, qe4, e, sel_id, cnt, tmp);
            st.evals += cnt;
            for (int b = 0; b < cnt; b += 32) {
              const bool bad = (b + lane < cnt) && (Op::post(tmp[b + lane]) <= de);
              if (__any_sync(FULL, bad)) {
                keep = false;
                break;
              }
            }
          }
          __syncwarp();
          if (keep) {
            if (lane == 0) {
              sel_id[cnt] = e;
              sel_d[cnt] = de;
            }
            cnt++;
          } else if (p.keep_pruned) {  // 1387-1392
            if (lane == 0) disc[ndisc] = (uint16_t)i;
            ndisc++;
          }
          __syncwarp();
        }
        if (p.keep_pruned && cnt < nb && ndisc > 0) {
          // 1399-1409: back-fill with the nearest discarded ones, then the caller sorts (1195).
          // Kept and discarded are both ascending sub-sequences of Q, so a merge by key restores order.
          const int take = min(ndisc, nb - cnt);
          // serial merge, executed uniformly by the warp (take <= nb, rare option); lane 0 writes
          {
            int a = cnt - 1, b = take - 1, o = cnt + take - 1;
            while (b >= 0) {
              const uint64_t kb = Q.get(disc[b]) & ~1ull;
              const bool from_a = a >= 0 && make_key(sel_d[a], sel_id[a]) > kb;
              __syncwarp();
              if (lane == 0) {
                sel_id[o] = from_a ? sel_id[a] : key_id(kb);
                sel_d[o] = from_a ? sel_d[a] : key_dist(kb);
              }
              __syncwarp();
              if (from_a) --a; else --b;
              --o;
            }
          }
          cnt += take;
          __syncwarp();
        }
      }
      // own list of layer l (hnsw.rs:1197), ascending, INVALID padded
      {
        uint32_t* ids;
        float* ds;
        int cap;
        if (l == 0) {
          ids = g.adj0 + (size_t)x * g.deg0;
          ds = g.adj0_d + (size_t)x * g.deg0;
          cap = g.deg0;
        } else {
          const size_t li = (size_t)g.up_off[x] + (l - 1);
          ids = g.adjU + li * g.M;
          ds = g.adjU_d + li * g.M;
          cap = g.M;
        }
        for (int i = lane; i < cap; i += 32) {
          ids[i] = i < cnt ? sel_id[i] : INVALID_ID;
          ds[i] = i < cnt ? sel_d[i] : 0.f;
        }
      }
      if (cnt > 0) cur = sel_id[0];  // 1201-1203
      __syncwarp();
    }
    if (overflow && lane == 0) atomicExch(p.status, 1);
  }
  vis.save(p.vis, slot);
  if (p.stats && lane == 0) {
    atomicAdd(p.stats + 0, (unsigned long long)st.evals);
    atomicAdd(p.stats + 1, (unsigned long long)st.expansions);
    atomicAdd(p.stats + 2, (unsigned long long)st.adj);
  }
}

// ------------------------------------------------------------------------------------------------
// phase B
__device__ __forceinline__ void lock_point(int* locks, uint32_t q) {
  if (lane_id() == 0) {
    while (atomicCAS(locks + q, 0, 1) != 0) {
      __nanosleep(64);
    }
    __threadfence();
  }
  __syncwarp();
}
__device__ __forceinline__ void unlock_point(int* locks, uint32_t q) {
  __syncwarp();
  if (lane_id() == 0) {
    __threadfence();
    atomicExch(locks + q, 0);
  }
  __syncwarp();
}

// add (x, d) to the sorted list ids/ds of capacity cap; drop the farthest when over capacity
// (push + sort_unstable + pop, hnsw.rs:1268-1284).  Warp-collective, list is locked.
__device__ __forceinline__ void list_add_sorted(uint32_t* ids, float* ds, int cap, uint32_t x, float d) {
  const int lane = lane_id();
  const uint64_t key = make_key(d, x);
  int n = 0, pos = 0;
  bool already = false;
  for (int b = 0; b < cap; b += 32) {
    const int i = b + lane;
    uint32_t id = INVALID_ID;
    float di = 0.f;
    if (i < cap) {
      id = __ldcg(ids + i);
      di = __ldcg(ds + i);
    }
    const bool valid = id != INVALID_ID;
    n += __popc(__ballot_sync(FULL, valid));
    pos += __popc(__ballot_sync(FULL, valid && make_key(di, id) < key));
    already |= __any_sync(FULL, valid && id == x) != 0;
  }
  if (already) return;  // hnsw.rs:1258-1267
  if (n == cap && pos == cap) return;  // pushed then popped again
  const int new_n = n < cap ? n + 1 : cap;
  int top = new_n - 1;
  while (top > pos) {
    const int lo = top - 31 > pos + 1 ? top - 31 : pos + 1;
    const int i = lo + lane;
    uint32_t vi = 0;
    float vd = 0.f;
    if (i <= top) {
      vi = __ldcg(ids + i - 1);
      vd = __ldcg(ds + i - 1);
    }
    __syncwarp();
    if (i <= top) {
      __stcg(ids + i, vi);
      __stcg(ds + i, vd);
    }
    __syncwarp();
    top = lo - 1;
  }
  if (lane == 0) {
    __stcg(ids + pos, x);
    __stcg(ds + pos, d);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(BUILD_THREADS) insert_link_kernel(InsertParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GraphView& g = p.g;
  const uint32_t wstride = gridDim.x * (BUILD_THREADS / 32);
  for (uint32_t wi = blockIdx.x * (BUILD_THREADS / 32) + warp; wi < p.count; wi += wstride) {
    const uint32_t x = p.first + wi;
    const int L = g.level[x];
    for (int l = L; l >= 0; --l) {  // hnsw.rs:1248
      const uint32_t* ids;
      const float* ds;
      int cap;
      if (l == 0) {
        ids = g.adj0 + (size_t)x * g.deg0;
        ds = g.adj0_d + (size_t)x * g.deg0;
        cap = g.deg0;
      } else {
        const size_t li = (size_t)g.up_off[x] + (l - 1);
        ids = g.adjU + li * g.M;
        ds = g.adjU_d + li * g.M;
        cap = g.M;
      }
      for (int j = 0; j < cap; ++j) {  // hnsw.rs:1249
        const uint32_t q = ids[j];
        if (q == INVALID_ID) break;
        if (q == x) continue;  // 1250
        const float d = ds[j];
        // target list: q.neighbours[L] with L = the NEW point's level (1257)
        uint32_t* tids;
        float* tds;
        int tcap;
        if (L == 0) {
          tids = g.adj0 + (size_t)q * g.deg0;
          tds = g.adj0_d + (size_t)q * g.deg0;
          tcap = g.deg0;  // 1272-1276: 2*max_nb_connection at layer 0
        } else {
          if (L > (int)g.plevel[q]) continue;  // a list no search can ever read; not materialised
          const size_t li = (size_t)g.up_off[q] + (L - 1);
          tids = g.adjU + li * g.M;
          tds = g.adjU_d + li * g.M;
          tcap = g.M;
        }
        lock_point(p.locks, q);
        list_add_sorted(tids, tds, tcap, x, d);
        unlock_point(p.locks, q);
      }
    }
  }
  (void)lane;
}

template <class Op, int NS>
static cudaError_t launch_insert_for_op(const InsertParams& p, int grid, size_t smem, cudaStream_t st, bool query_only,
                                        int* blocks_per_sm) {
  const int ch = p.g.d4 / 8;
#define HB_LAUNCH(CHV, UV)                                                                              \
  do {                                                                                                  \
    auto kern = insert_search_kernel<Op, CHV, UV, NS>;                                                  \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                     \
    if (blocks_per_sm) {                                                                                \
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, p.threads, smem);      \
      if (e != cudaSuccess) return e;                                                                   \
    }                                                                                                   \
    if (!query_only) kern<<<grid, p.threads, smem, st>>>(p);                                        \
    return cudaGetLastError();                                                                          \
  } while (0)
  if constexpr (Specialise<Op>::value) {
    if (ch == 1) HB_LAUNCH(1, 4);
    if (ch == 2) HB_LAUNCH(2, 4);
    if (ch == 4) HB_LAUNCH(4, 2);
  }
  HB_LAUNCH(0, 2);
#undef HB_LAUNCH
}

cudaError_t launch_insert_search(const InsertParams& p, int metric, int dtype, int grid, size_t smem, cudaStream_t st,
                                 bool query_only, int* blocks_per_sm) {
  return dispatch_op(metric, dtype, [&](auto tag) -> cudaError_t {
    using Op = typename decltype(tag)::type;
    if constexpr (Specialise<Op>::value) {
      if (p.q_kind == 108) return launch_insert_for_op<Op, 108>(p, grid, smem, st, query_only, blocks_per_sm);
      if (p.q_kind == 104) return launch_insert_for_op<Op, 104>(p, grid, smem, st, query_only, blocks_per_sm);
    }
    return launch_insert_for_op<Op, 0>(p, grid, smem, st, query_only, blocks_per_sm);
  });
}

cudaError_t launch_insert_link(const InsertParams& p, int grid, cudaStream_t st) {
  insert_link_kernel<<<grid, BUILD_THREADS, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace hb
