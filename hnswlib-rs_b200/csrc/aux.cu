// Stand-alone kernels on the index's point store:
//   dist_batch_kernel  — K1: queries x candidate ids -> distances.  The batched form of
//                        Distance<T>::eval (crate anndists; call sites /root/reference/src/hnsw.rs:1026,1518).
//   bruteforce_kernel  — K5: exact k nearest neighbours by linear scan, the GPU counterpart of
//                        brute_force_neighbours in /root/reference/tests/serpar.rs:42-70 (recall ground truth).
#include "index.h"
#include "kernels.h"

namespace hb {

struct AuxParams {
  GraphView g;
  const void* queries;  // [nq][q_bytes] raw element bytes
  int q_bytes;
  uint32_t nq;
  const uint32_t* cand;  // [nq][m]
  uint32_t m;
  float* out;            // [nq][m]
  int k;
  uint32_t* out_ids;     // [nq][k]
  float* out_dist;       // [nq][k]
  int smem_per_warp;
};

template <class Op, int CH, int U>
__global__ void __launch_bounds__(256) dist_batch_kernel(AuxParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = smem_raw + (size_t)warp * p.smem_per_warp;
  uint4* q4 = reinterpret_cast<uint4*>(base);
  uint32_t* cid = reinterpret_cast<uint32_t*>(base + (size_t)p.g.d4 * 16);
  float* cd = reinterpret_cast<float*>(cid + 32);
  const uint4* vec4 = reinterpret_cast<const uint4*>(p.g.vec);
  const uint32_t wstride = gridDim.x * 8;
  for (uint32_t qi = blockIdx.x * 8 + warp; qi < p.nq; qi += wstride) {
    __syncwarp();
    stage_row_bytes(q4, reinterpret_cast<const char*>(p.queries) + (size_t)qi * p.q_bytes, p.q_bytes, p.g.d4 * 16);
    for (uint32_t b = 0; b < p.m; b += 32) {
      const int cnt = min(32u, p.m - b);
      if (lane < cnt) cid[lane] = p.cand[(size_t)qi * p.m + b + lane];
      __syncwarp();
      warp_dists<Op, CH, U>(vec4, p.g.d4, p.g.dim, q4, cid, cnt, cd);
      __syncwarp();
      if (lane < cnt) p.out[(size_t)qi * p.m + b + lane] = Op::post(cd[lane]);
      __syncwarp();
    }
  }
}

template <class Op, int CH, int U>
__global__ void __launch_bounds__(256) bruteforce_kernel(AuxParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = smem_raw + (size_t)warp * p.smem_per_warp;
  uint4* q4 = reinterpret_cast<uint4*>(base);
  uint32_t* cid = reinterpret_cast<uint32_t*>(base + (size_t)p.g.d4 * 16);
  float* cd = reinterpret_cast<float*>(cid + 32);
  uint64_t* wbuf = reinterpret_cast<uint64_t*>(cd + 32);
  const uint4* vec4 = reinterpret_cast<const uint4*>(p.g.vec);
  const uint32_t wstride = gridDim.x * 8;
  SortedQueue Q;
  for (uint32_t qi = blockIdx.x * 8 + warp; qi < p.nq; qi += wstride) {
    __syncwarp();
    stage_row_bytes(q4, reinterpret_cast<const char*>(p.queries) + (size_t)qi * p.q_bytes, p.q_bytes, p.g.d4 * 16);
    Q.reset(wbuf, p.k);
    for (uint32_t b = 0; b < p.g.n; b += 32) {
      const int cnt = min(32u, p.g.n - b);
      cid[lane] = b + lane;
      __syncwarp();
      warp_dists<Op, CH, U>(vec4, p.g.d4, p.g.dim, q4, cid, cnt, cd);
      __syncwarp();
      const uint64_t key = lane < cnt ? make_key(Op::post(cd[lane]), b + lane) : ~0ull;
      unsigned acc = __ballot_sync(FULL, lane < cnt && Q.accepts(key));
      while (acc) {
        const int j = __ffs(acc) - 1;
        acc &= acc - 1;
        const uint64_t kj = __shfl_sync(FULL, key, j);
        if (Q.accepts(kj)) Q.insert(kj);
      }
      __syncwarp();
    }
    for (int j = lane; j < p.k; j += 32) {
      p.out_ids[(size_t)qi * p.k + j] = j < Q.n ? key_id(Q.w[j]) : INVALID_ID;
      p.out_dist[(size_t)qi * p.k + j] = j < Q.n ? key_dist(Q.w[j]) : __int_as_float(0x7f800000);
    }
  }
}

template <class Op>
static cudaError_t launch_aux_for_op(const AuxParams& p, bool brute, int grid, size_t smem, cudaStream_t st) {
  const int ch = p.g.d4 / 8;
#define HB_LAUNCH(CHV, UV)                                                                                \
  do {                                                                                                    \
    if (brute) {                                                                                          \
      auto kern = bruteforce_kernel<Op, CHV, UV>;                                                         \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                     \
      kern<<<grid, 256, smem, st>>>(p);                                                                   \
    } else {                                                                                              \
      auto kern = dist_batch_kernel<Op, CHV, UV>;                                                         \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                     \
      kern<<<grid, 256, smem, st>>>(p);                                                                   \
    }                                                                                                     \
    return cudaGetLastError();                                                                            \
  } while (0)
  if constexpr (Specialise<Op>::value) {
    if (ch == 1) HB_LAUNCH(1, 4);
    if (ch == 4) HB_LAUNCH(4, 2);
  }
  HB_LAUNCH(0, 2);
#undef HB_LAUNCH
}

static cudaError_t launch_aux(const AuxParams& p, int metric, int dtype, bool brute, int grid, size_t smem, cudaStream_t st) {
  return dispatch_op(metric, dtype, [&](auto tag) -> cudaError_t {
    using Op = typename decltype(tag)::type;
    return launch_aux_for_op<Op>(p, brute, grid, smem, st);
  });
}

#define HB_CUDA(call)                                     \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

int Index::dist_batch(const void* queries, size_t nq, int d, const uint32_t* cand, size_t m, float* out) {
  if (nq == 0 || m == 0) return 0;
  if (d != dim) return fail("query length differs from the index dimension");
  for (size_t i = 0; i < nq * m; ++i)
    if (cand[i] >= n) return fail("candidate id out of range");
  HB_CUDA(cudaSetDevice(device));
  void* dq = nullptr;
  uint32_t* dc = nullptr;
  float* dout = nullptr;
  HB_CUDA(cudaMalloc(&dq, nq * d * es));
  HB_CUDA(cudaMalloc(&dc, nq * m * 4));
  HB_CUDA(cudaMalloc(&dout, nq * m * 4));
  cudaMemcpyAsync(dq, queries, nq * d * es, cudaMemcpyHostToDevice, stream_);
  cudaMemcpyAsync(dc, cand, nq * m * 4, cudaMemcpyHostToDevice, stream_);
  AuxParams p{};
  p.g = view();
  p.queries = dq;
  p.q_bytes = d * es;
  p.nq = (uint32_t)nq;
  p.cand = dc;
  p.m = (uint32_t)m;
  p.out = dout;
  p.smem_per_warp = p.g.d4 * 16 + 256;
  const size_t smem = (size_t)p.smem_per_warp * 8;
  const int grid = (int)std::min<size_t>((size_t)sm_count_ * 8, (nq + 7) / 8);
  cudaError_t e = launch_aux(p, metric, dtype, false, grid, smem, stream_);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, nq * m * 4, cudaMemcpyDeviceToHost, stream_);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
  cudaFree(dq);
  cudaFree(dc);
  cudaFree(dout);
  if (e != cudaSuccess) return cuda_fail(e, "dist_batch");
  return 0;
}

int Index::bruteforce(const void* queries, size_t nq, int d, size_t k, uint32_t* out_ids, float* out_dist) {
  if (nq == 0 || k == 0) return 0;
  if (d != dim) return fail("query length differs from the index dimension");
  HB_CUDA(cudaSetDevice(device));
  void* dq = nullptr;
  uint32_t* dids = nullptr;
  float* dd = nullptr;
  HB_CUDA(cudaMalloc(&dq, nq * d * es));
  HB_CUDA(cudaMalloc(&dids, nq * k * 4));
  HB_CUDA(cudaMalloc(&dd, nq * k * 4));
  cudaMemcpyAsync(dq, queries, nq * d * es, cudaMemcpyHostToDevice, stream_);
  AuxParams p{};
  p.g = view();
  p.queries = dq;
  p.q_bytes = d * es;
  p.nq = (uint32_t)nq;
  p.k = (int)k;
  p.out_ids = dids;
  p.out_dist = dd;
  p.smem_per_warp = (int)(((size_t)p.g.d4 * 16 + 256 + k * 8 + 15) & ~(size_t)15);
  const size_t smem = (size_t)p.smem_per_warp * 8;
  if (smem > 220 * 1024) {
    cudaFree(dq); cudaFree(dids); cudaFree(dd);
    return fail("k / dimension too large for the brute-force kernel");
  }
  const int grid = (int)std::min<size_t>((size_t)sm_count_ * 4, (nq + 7) / 8);
  cudaError_t e = launch_aux(p, metric, dtype, true, grid, smem, stream_);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_ids, dids, nq * k * 4, cudaMemcpyDeviceToHost, stream_);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, dd, nq * k * 4, cudaMemcpyDeviceToHost, stream_);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
  cudaFree(dq);
  cudaFree(dids);
  cudaFree(dd);
  if (e != cudaSuccess) return cuda_fail(e, "bruteforce");
  return 0;
}

}  // namespace hb
