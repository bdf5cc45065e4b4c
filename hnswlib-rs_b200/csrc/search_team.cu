// Query kernel, team form: EIGHT lanes own one query, four queries ride one warp.
//
// Restates /root/reference/src/hnsw.rs:1487-1580 (search_filter without a filter: entry fetch, one hop per
// upper layer, search_layer on the lowest populated layer, ascending top-k) and hnsw.rs:922-1064 (search_layer),
// with the batch contract of parallel_search (hnsw.rs:1612-1635: one answer per query, in input order).
//
// Why teams.  A point row is read by 8 lanes x 16 bytes whatever the kernel shape, so in the warp-per-query kernel
// (search.cu) 24 of 32 lanes idle through all the per-query bookkeeping: picking the next candidate, marking it,
// queue inserts, the loop control.  Here the four teams of a warp run the SAME instruction stream on four queries:
// a warp instruction that used to serve one query serves four, and a 10 000-query batch is resident at once
// (148 SMs x 16 warps x 4 = 9472 teams) instead of queueing for a second and third wave.
//
// A team is a small state machine; one loop iteration is one STEP = "score up to 32 rows named by one chunk of one
// adjacency list":
//   FETCH   take a query index, stage the query, registers <- the lane's 16-byte chunks of it
//   DESC    layer entry_level+1: score the entry point; layers entry_level..1: score pivot.neighbours[layer],
//           strict '<' first minimum becomes the pivot (hnsw.rs:1511-1529)
//   SEARCH  pop the nearest unexpanded entry of W, read its list, drop visited ids, score the rest, insert
//           (hnsw.rs:969-1057)
// All four teams execute every part of the step together, each under its own predicates: every warp collective
// (vote, shuffle, warp barrier) is issued in converged code with the full mask.
//
// Queue: W is a sorted array of (dist,id) keys in shared memory, 8 lanes wide; which entries are still unexpanded
// (the reference's C, see search_core.cuh) is a bit mask in registers, so "pop nearest candidate" is a find-first-set.
// Visited set: the per-slot epoch-tagged table of common.cuh, probed four ids per lane, claimed with atomicCAS.
// Distances: identical lane/chunk mapping and reduction tree as warp_dists (bit-identical results); the 8-lane
// reduction of four rows is done as one transposed reduction (4+2+1 shuffles instead of 4 x 3).
// Rows are fetched four per team at a time into registers (the query stays in shared memory so that 64 registers can
// hold rows in flight); the lines of the step's later rows are prefetched to L2 right after the list is known, so
// only the first block pays the HBM latency.
// Visited table: 16-byte buckets of four entries, one 128-bit load per probe, linear probing over buckets.
#include "kernels.h"
#include "team_common.cuh"

namespace hb {

enum : int { TS_FETCH = 0, TS_DESC = 1, TS_SEARCH = 2, TS_DONE = 3 };

// Four rows' per-lane partial sums -> the finished value of row (g >> 1) in lane g, over the team's 8 lanes.
// Same pairing as reduce8's xor butterfly (4, 2, 1): a lane adds its own partial and the partner's, so every sum has
// the operands of the butterfly (IEEE addition is commutative): bit-identical.
template <class R>
__device__ __forceinline__ R treduce4(const R& a0, const R& a1, const R& a2, const R& a3, int g) {
  const bool h4 = (g & 4) != 0;
  const R b0 = radd(h4 ? a2 : a0, rshfl(h4 ? a0 : a2, 4));
  const R b1 = radd(h4 ? a3 : a1, rshfl(h4 ? a1 : a3, 4));
  const bool h2 = (g & 2) != 0;
  R c = radd(h2 ? b1 : b0, rshfl(h2 ? b0 : b1, 2));
  c = radd(c, rshfl(c, 1));
  return c;
}

// 4 consecutive ids of a list per lane (ids 4g..4g+3 of the 32-id chunk at `base`), INVALID_ID beyond the capacity
__device__ __forceinline__ void load_ids4(const uint32_t* ids, int lcap, int base, int g, uint32_t (&nid)[4]) {
  const int o = base + 4 * g;
  if (ids != nullptr && (lcap & 3) == 0 && o + 3 < lcap) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(ids + o));
    nid[0] = v.x; nid[1] = v.y; nid[2] = v.z; nid[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) nid[j] = (ids != nullptr && o + j < lcap) ? __ldg(ids + o + j) : INVALID_ID;
  }
}

// CH = 16-byte chunks per lane per row (row = CH * 128 bytes), QC = queue slots (>= ef), STATS = traversal counters
template <class Op, int CH, int QC, bool STATS>
__global__ void __launch_bounds__(TEAM_THREADS, TEAM_MIN_BLOCKS) search_team_kernel(SearchParams p) {
  typedef typename MaskSel<QC>::type MaskT;
  typedef typename Op::red_t red_t;
  constexpr int TEAM_SMEM = QC * 8 + 128 + CH * 128;  // queue keys, the step's row ids, the query
  constexpr int BL = QC / 8;                         // queue entries per lane block
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const GraphView& G = p.g;
  const int g = (int)pin(threadIdx.x & 7), tl0 = (int)pin(threadIdx.x & 24);
  const uint32_t wa = pin(smem_u32(smem_raw) + (uint32_t)(threadIdx.x >> 3) * TEAM_SMEM);  // queue keys
  const uint32_t ca = wa + QC * 8;                                                           // the step's row ids
  const uint32_t qa = ca + 128;                                                              // the query, zero padded
  const char* const vecb = reinterpret_cast<const char*>(G.vec);
  const uint32_t row_bytes = (uint32_t)CH * 128u;
  const uint64_t pol_rows = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();

  // ---- visited table of this team (Visited of common.cuh, 8 lanes wide)
  const uint32_t slot = blockIdx.x * (TEAM_THREADS / 8) + (threadIdx.x >> 3);
  uint32_t* const vtab = p.vis.tables + (size_t)slot * p.vis.cap;
  const uint32_t vmask = (p.vis.cap >> 2) - 1, vlimit = p.vis.cap - (p.vis.cap >> 2);  // vmask over buckets
  const int vshift = p.vis.shift + 2, id_bits = p.vis.id_bits;
  const uint32_t epoch_max = (id_bits >= 32) ? 0u : ((1u << (32 - id_bits)) - 1u);
  uint32_t epoch = p.vis.epochs[slot], vused = 0;

  int state = TS_FETCH, layer = 0, base = 0, n = 0;
  uint32_t qi = 0, cur = INVALID_ID, newpiv = INVALID_ID;
  float best = 0.f;
  MaskT open;
  open.clear();
  uint64_t thr = ~0ull;
  unsigned evals = 0, expans = 0, adjr = 0;
  bool overflow = false;
  const int cap = p.ef;

  for (;;) {
    // ================================================================ completion + FETCH (rare: once per query)
    const bool finished = state == TS_SEARCH && base == 0 && (open.none() || overflow);
    if (__any_sync(FULL, finished || state == TS_FETCH)) {
      if (finished) {  // ascending top-k (hnsw.rs:1544-1579): the queue is sorted
        int count = n < p.k ? n : p.k;  // hnsw.rs:1547 (n <= ef)
        if (overflow) {
          if (g == 0) atomicExch(p.status, 1);
          count = 0;
        }
        const size_t ob = (size_t)qi * p.k;
        for (int j = g; j < p.k; j += 8) {
          if (j < count) {
            const uint64_t key = lds64(wa + 8 * j);
            const uint32_t id = key_id(key);
            p.out_nb[ob + j] = NeighbourOut{G.origin[id], key_dist(key), id};
          } else {
            p.out_nb[ob + j] = NeighbourOut{~0ull, __int_as_float(0x7f800000), INVALID_ID};
          }
        }
        if (g == 0) p.out_count[qi] = count;
        overflow = false;
        state = TS_FETCH;
      }
      __syncwarp();
      uint32_t q = 0;
      if (state == TS_FETCH && g == 0) q = atomicAdd(p.work_counter, 1u);
      q = __shfl_sync(FULL, q, tl0);
      if (state == TS_FETCH) {
        if (q >= p.nq) {
          state = TS_DONE;
        } else {  // the query row, zero padded, into the staging buffer (cf. stage_row_bytes)
          qi = q;
          const char* src = reinterpret_cast<const char*>(p.queries) + (size_t)qi * p.q_stride_bytes;
          const int nw = p.q_bytes >> 2;
          if ((reinterpret_cast<size_t>(src) & 3) == 0) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
            for (int i = g; i < CH * 32; i += 8) {
              uint32_t v = 0;
              if (i < nw) v = s32[i];
              else if ((i << 2) < p.q_bytes)
                for (int b = 0; b < (p.q_bytes & 3); ++b) v |= (uint32_t)(uint8_t)src[(nw << 2) + b] << (8 * b);
              sts32(qa + 4 * i, v);
            }
          } else {
            for (int i = g; i < CH * 32; i += 8) {
              uint32_t v = 0;
              for (int b = 0; b < 4; ++b)
                if ((i << 2) + b < p.q_bytes) v |= (uint32_t)(uint8_t)src[(i << 2) + b] << (8 * b);
              sts32(qa + 4 * i, v);
            }
          }
        }
      }
      __syncwarp();
      if (state == TS_FETCH) {
        cur = newpiv = G.entry;  // hnsw.rs:1498-1506
        best = __int_as_float(0x7f800000);
        layer = G.entry_level + 1;  // the "layer" above the top: the step that scores the entry point itself
        base = 0;
        state = TS_DESC;
      }
      if (__all_sync(FULL, state == TS_DONE)) break;
    }

    // ================================================================ the step's list chunk: 4 ids per lane
    const bool single = state == TS_DESC && layer > G.entry_level;  // score the entry point alone
    const bool listing = (state == TS_DESC && !single) || state == TS_SEARCH;
    const int lyr = state == TS_SEARCH ? p.layer0 : layer;
    if (state == TS_SEARCH && base == 0) {  // C.pop(): the nearest unexpanded entry of W (hnsw.rs:971)
      const int idx = open.first();
      open.drop_first();
      cur = key_id(lds64(wa + 8 * idx));
      if (STATS) expans += 1;
      // pull the adjacency rows of the next two candidates towards L2 while this one is expanded
      MaskT o2 = open;
      if (g == 2) o2.drop_first();
      if ((g == 1 || g == 2) && !o2.none()) {
        const uint32_t pc = key_id(lds64(wa + 8 * o2.first()));
        int pcap;
        const uint32_t* pids = list_ids(G, pc, lyr, pcap);
        if (pids) asm volatile("prefetch.global.L2 [%0];" ::"l"(pids));
      }
    }
    uint32_t nid[4] = {INVALID_ID, INVALID_ID, INVALID_ID, INVALID_ID};
    int lcap = 0;
    if (listing) {
      const uint32_t* ids = list_ids(G, cur, lyr, lcap);  // hnsw.rs:1006 / 1511
      load_ids4(ids, lcap, base, g, nid);
    }
    const int mine = (nid[0] != INVALID_ID) + (nid[1] != INVALID_ID) + (nid[2] != INVALID_ID) + (nid[3] != INVALID_ID);
    const int fl = __popc((__ballot_sync(FULL, mine == 4) >> tl0) & 0xFFu);  // lists are dense prefixes
    const int part = __shfl_sync(FULL, mine, tl0 + (fl & 7));
    const int nvalid = fl == 8 ? 32 : 4 * fl + part;
    const bool more = (nvalid == 32) && (base + 32 < lcap);
    if (STATS) adjr += nvalid;

    // ---- SEARCH: visited test-and-set of the lane's (up to) four ids (hnsw.rs:1016-1017); DESC: every id is scored.
    // One probe = one 16-byte bucket of four entries: the id is there, or the bucket has a free (stale-epoch) entry to
    // claim, or the probe moves to the next bucket.  Ids only ever enter the first bucket of their sequence that has
    // room, so "a bucket with room and without the id" proves absence.
    const uint32_t vtag = epoch << id_bits;
    uint32_t h[4];
    bool pend[4], fresh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = (nid[j] * 2654435761u) >> vshift;  // bucket index: vshift = 32 - log2(cap / 4)
      pend[j] = state == TS_SEARCH && nid[j] != INVALID_ID;
      fresh[j] = state == TS_DESC && nid[j] != INVALID_ID;
    }
    while (__any_sync(FULL, pend[0] | pend[1] | pend[2] | pend[3])) {
      uint4 bk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (pend[j]) bk[j] = ld_keep4(reinterpret_cast<const uint4*>(vtab) + h[j], pol_keep);
      int sl[4];
      uint32_t cv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sl[j] = -1;
        cv[j] = 0;
        if (pend[j]) {
          const uint32_t want = vtag | nid[j];
          if (bk[j].x == want || bk[j].y == want || bk[j].z == want || bk[j].w == want) {
            pend[j] = false;  // already visited
          } else {
            if ((bk[j].w >> id_bits) != epoch) { sl[j] = 3; cv[j] = bk[j].w; }
            if ((bk[j].z >> id_bits) != epoch) { sl[j] = 2; cv[j] = bk[j].z; }
            if ((bk[j].y >> id_bits) != epoch) { sl[j] = 1; cv[j] = bk[j].y; }
            if ((bk[j].x >> id_bits) != epoch) { sl[j] = 0; cv[j] = bk[j].x; }
            if (sl[j] < 0) h[j] = (h[j] + 1) & vmask;  // bucket full: next one
          }
        }
      }
      uint32_t old[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (pend[j] && sl[j] >= 0) old[j] = atomicCAS(vtab + 4 * h[j] + sl[j], cv[j], vtag | nid[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (pend[j] && sl[j] >= 0) {
          if (old[j] == cv[j]) {
            fresh[j] = true;
            pend[j] = false;
          } else if (old[j] == (vtag | nid[j])) {
            pend[j] = false;  // the same id twice in one chunk: its first occurrence recorded it
          }  // else: another id of this chunk took the entry; look at the bucket again
        }
      }
    }
    // ---- the team's row list: SEARCH compacts the fresh ids, DESC keeps list order (its tie rule needs positions)
    int n_t = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned fb = (__ballot_sync(FULL, fresh[j]) >> tl0) & 0xFFu;
      const int at = state == TS_DESC ? 4 * g + j : n_t + __popc(fb & ((1u << g) - 1u));
      if (fresh[j]) sts32(ca + 4 * at, nid[j]);
      n_t += __popc(fb);
    }
    if (single) {
      if (g == 0) sts32(ca, cur);
      n_t = 1;
    }
    if (state == TS_SEARCH) vused += n_t;
    if (STATS) evals += n_t;
    __syncwarp();

    // ================================================================ scoring: rows cand[0..n_t) of every team
    const int maxn = __reduce_max_sync(FULL, n_t);
    // lines of the rows after the first block -> L2, so that the later blocks find them there
    for (int l = 4 * CH + g; l < maxn * CH; l += 8) {
      if (l < n_t * CH) {
        const uint32_t id = lds32(ca + 4 * (l / CH));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(vecb + (size_t)id * row_bytes + (uint32_t)(l % CH) * 128u));
      }
    }
    uint64_t dkey = ~0ull;  // DESC: smallest (distance, list position) seen by this lane
    for (int b = 0; b < maxn; b += 4) {
      uint4 x[4][CH];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (b + u < n_t) {
          const uint4* row = reinterpret_cast<const uint4*>(vecb + (size_t)lds32(ca + 4 * (b + u)) * row_bytes) + g;
#pragma unroll
          for (int i = 0; i < CH; ++i) x[u][i] = ldg_stream(row + 8 * i, pol_rows);
        }
      }
      typename Op::acc_t acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = Op::zero();
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const uint4 q = lds128(qa + 16 * (g + 8 * i));
#pragma unroll
        for (int u = 0; u < 4; ++u) Op::chunk(acc[u], q, x[u][i]);  // rows beyond n_t: garbage in, result unused
      }
      red_t a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = Op::fold(acc[u]);
      const red_t tot = treduce4<red_t>(a[0], a[1], a[2], a[3], g);
      const float dist = Op::post(Op::finish(tot, G.dim));  // hnsw.rs:1026 / 1518
      const int r = b + (g >> 1);
      const bool has = r < n_t;
      uint64_t key = ~0ull;
      if (has) {
        if (state == TS_DESC) {
          key = ((uint64_t)__float_as_uint(dist) << 32) | (uint32_t)r;
          dkey = key < dkey ? key : dkey;
        } else {
          key = make_key(dist, lds32(ca + 4 * r));
        }
      }
      // ---- SEARCH: W and C updates (hnsw.rs:1028-1053), one accepted candidate per team per round
      unsigned tbits = (__ballot_sync(FULL, state == TS_SEARCH && has && (g & 1) == 0 && key < thr) >> tl0) & 0xFFu;
      while (__any_sync(FULL, tbits != 0)) {
        const bool act0 = tbits != 0;
        const int src = tl0 + (act0 ? __ffs(tbits) - 1 : 0);
        tbits &= tbits - 1;
        const uint64_t kj = __shfl_sync(FULL, key, src);
        const bool act = act0 && kj < thr;  // the bound may have tightened since the ballot
        // position = number of keys below kj: the lane blocks wholly below, then inside the first block that is not
        const unsigned b1 = (__ballot_sync(FULL, act && lds64(wa + 8 * (BL * g + BL - 1)) < kj) >> tl0) & 0xFFu;
        const int nb = __popc(b1) & 7;
        int pos = BL * nb;
#pragma unroll
        for (int t = 0; t < BL; t += 8)
          pos += __popc((__ballot_sync(FULL, act && lds64(wa + 8 * (BL * nb + t + g)) < kj) >> tl0) & 0xFFu);
        // shift [pos, top] up by one (the last entry of a full queue drops out), 16 entries per round, top down
        const int lo = act ? pos : 0x7fffffff;
        int top = (n < cap ? n : cap - 1) - 1;
        while (__any_sync(FULL, top >= lo)) {
          const int i0 = top - g, i1 = top - 8 - g;
          uint64_t v0 = 0, v1 = 0;
          if (i0 >= lo) v0 = lds64(wa + 8 * i0);
          if (i1 >= lo) v1 = lds64(wa + 8 * i1);
          __syncwarp();
          if (i0 >= lo) sts64(wa + 8 * i0 + 8, v0);
          if (i1 >= lo) sts64(wa + 8 * i1 + 8, v1);
          top -= 16;
        }
        if (act && g == 0) sts64(wa + 8 * pos, kj);
        __syncwarp();
        if (act) {
          n = n < cap ? n + 1 : cap;
          thr = lds64(wa + 8 * (cap - 1));
          open.insert_at(pos, cap);
        }
      }
    }

    // ================================================================ after the step
    if (state == TS_SEARCH) {
      if (vused >= vlimit) {
        overflow = true;
        base = 0;
      } else {
        base = more ? base + 32 : 0;
      }
    }
    __syncwarp();  // this step's reads of the row ids and the queue precede the next step's writes
    if (__any_sync(FULL, state == TS_DESC)) {
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const uint64_t other = __shfl_xor_sync(FULL, dkey, o);
        dkey = other < dkey ? other : dkey;
      }
      bool start = false;
      if (state == TS_DESC) {
        const float dmin = __uint_as_float((uint32_t)(dkey >> 32));
        if (n_t > 0 && dmin < best) {  // strict '<' in list order == the first minimum, if below `best`
          best = dmin;
          newpiv = lds32(ca + 4 * ((uint32_t)dkey & 31u));
        }
        if (more) {
          base += 32;
        } else {
          base = 0;
          cur = newpiv;  // hnsw.rs:1526-1528
          layer -= 1;
          start = layer < 1;
        }
      }
      __syncwarp();  // the row ids are read before the next step rewrites them
      if (start) {
        // ---- search_layer on the lowest populated layer starts (hnsw.rs:1531-1542, 940-967)
        if (epoch >= epoch_max) {
          for (uint32_t i = g; i < p.vis.cap; i += 8) __stcg(vtab + i, 0u);
          epoch = 0;
        }
        epoch += 1;
#pragma unroll
        for (int t = 0; t < BL; ++t) sts64(wa + 8 * (g + 8 * t), ~0ull);
      }
      __syncwarp();  // table clear and queue fill before the seeds
      if (start) {
        if (g == 0) {
          sts64(wa, make_key(best, cur));  // the entry of the layer enters W and C
          st_keep(vtab + 4 * ((cur * 2654435761u) >> vshift), (epoch << id_bits) | cur, pol_keep);
        }
        if (STATS) evals += 1;  // search_layer's own evaluation of its entry point (hnsw.rs:952): the value is `best`
        vused = 1;
        n = 1;
        open.set_only(0);
        state = TS_SEARCH;
      }
      __syncwarp();
      if (start) thr = lds64(wa + 8 * (cap - 1));
    }
  }

  if (g == 0) {
    p.vis.epochs[slot] = epoch;
    if (STATS && p.stats) {
      atomicAdd(p.stats + 0, (unsigned long long)evals);
      atomicAdd(p.stats + 1, (unsigned long long)expans);
      atomicAdd(p.stats + 2, (unsigned long long)adjr);
    }
  }
}

template <class Op, int QC>
static cudaError_t launch_team_for_op(const SearchParams& p, int grid, size_t smem, cudaStream_t st, bool query_only,
                                      int* blocks_per_sm) {
  const int ch = p.g.d4 / 8;
#define HB_LAUNCH_TEAM2(CHV, STV)                                                                            \
  do {                                                                                                       \
    auto kern = search_team_kernel<Op, CHV, QC, STV>;                                                        \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return e;                                                                          \
    if (blocks_per_sm) {                                                                                     \
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, TEAM_THREADS, smem);            \
      if (e != cudaSuccess) return e;                                                                        \
    }                                                                                                        \
    if (!query_only) kern<<<grid, TEAM_THREADS, smem, st>>>(p);                                              \
    return cudaGetLastError();                                                                               \
  } while (0)
#define HB_LAUNCH_TEAM(CHV)                  \
  do {                                       \
    if (p.stats) HB_LAUNCH_TEAM2(CHV, true); \
    HB_LAUNCH_TEAM2(CHV, false);             \
  } while (0)
  if (ch == 1) HB_LAUNCH_TEAM(1);
  if (ch == 2) HB_LAUNCH_TEAM(2);
  if (ch == 4) HB_LAUNCH_TEAM(4);
#undef HB_LAUNCH_TEAM
#undef HB_LAUNCH_TEAM2
  return cudaErrorInvalidValue;
}

cudaError_t launch_search_team(const SearchParams& p, int metric, int dtype, int grid, size_t smem, cudaStream_t st,
                               bool query_only, int* blocks_per_sm) {
  return dispatch_op(metric, dtype, [&](auto tag) -> cudaError_t {
    using Op = typename decltype(tag)::type;
    if constexpr (TeamOp<Op>::value) {
      if (p.q_smem == 64) return launch_team_for_op<Op, 64>(p, grid, smem, st, query_only, blocks_per_sm);
      if (p.q_smem == 128) return launch_team_for_op<Op, 128>(p, grid, smem, st, query_only, blocks_per_sm);
    }
    return cudaErrorInvalidValue;
  });
}

}  // namespace hb
