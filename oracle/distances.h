// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// CPU restatement of the distance functions the reference takes from the external crate
// `anndists` (requirement "0.1", /root/reference/Cargo.toml:89-91; NOT vendored under
// /root/reference, no Cargo.lock => exact version unpinned).  Call sites on the path:
// /root/reference/src/hnsw.rs:952,1026,1112,1146,1359,1374,1506,1518.
// Semantics restated from the crate's published behaviour:
//   DistL1  = sum |a-b|                      DistL2 = sqrt(sum (a-b)^2)   (un-squared!)
//   DistDot = max(0, 1 - sum a*b)            (inputs expected unit-norm)
//   DistCosine = max(0, 1 - ab/sqrt(aa*bb))  with f64 accumulation, 0 when a norm is 0
//   DistHamming = #{a_i != b_i}/len          DistJaccard = 1 - sum min / sum max
//   DistHellinger = sqrt(max(0, 1 - sum sqrt(a*b)))
//   DistJeffreys = sum (a-b) ln(max(a,1e-30)/max(b,1e-30))
//   DistJensenShannon = sqrt(0.5 * sum [a ln(a/m) + b ln(b/m)]), m=(a+b)/2
// All return f32.  PARITY UNPINNED: the reference holds no golden distance values and
// neither rustc nor the crate source exist in this container.
//
// Two summation orders are provided for the f32 accumulate-type metrics:
//   ORDER_REF : the reference's CPU shape (simdeez AVX2: 8 independent lanes over
//               chunks_exact(8), mul then add, horizontal add, scalar tail, then finish).
//   ORDER_GPU : the order the CUDA kernels use (documented in DESIGN.md "bit-exact
//               summation order"): 8 lanes g=0..7, lane g owns float4 chunks g, g+8, ...,
//               accumulates with fused multiply-add in element order, then a 4/2/1 xor
//               butterfly.  Used so that GPU-vs-oracle comparisons can be bit-exact; the
//               two orders are compared with the 1e-5 relative tolerance north_star names.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <type_traits>

namespace oracle {

enum Metric : int {
  METRIC_L1 = 0,
  METRIC_L2 = 1,
  METRIC_DOT = 2,
  METRIC_COSINE = 3,
  METRIC_HAMMING = 4,
  METRIC_JACCARD = 5,
  METRIC_HELLINGER = 6,
  METRIC_JEFFREYS = 7,
  METRIC_JENSENSHANNON = 8,
};

enum SumOrder : int { ORDER_REF = 0, ORDER_GPU = 1 };

// ---------------------------------------------------------------- per-element functors
struct AccL1 {
  static constexpr bool packed = true;  // GPU uses packed f32x2 math for this metric on f32 data
  static inline float step(float acc, float a, float b, bool fused) {
    (void)fused;
    return acc + std::fabs(a - b);
  }
  static inline float finish(float s) { return s; }
};
struct AccL2 {
  static constexpr bool packed = true;  // GPU uses packed f32x2 math for this metric on f32 data
  static inline float step(float acc, float a, float b, bool fused) {
    float df = a - b;
    return fused ? std::fmaf(df, df, acc) : acc + df * df;
  }
  static inline float finish(float s) { return std::sqrt(s); }
};
struct AccDot {
  static constexpr bool packed = true;  // GPU uses packed f32x2 math for this metric on f32 data
  static inline float step(float acc, float a, float b, bool fused) {
    return fused ? std::fmaf(a, b, acc) : acc + a * b;
  }
  static inline float finish(float s) {
    float r = 1.0f - s;
    return r > 0.f ? r : 0.f;
  }
};
struct AccHellinger {
  static constexpr bool packed = false;  // GPU uses packed f32x2 math for this metric on f32 data
  static inline float step(float acc, float a, float b, bool fused) {
    (void)fused;
    return acc + std::sqrt(a * b);
  }
  static inline float finish(float s) {
    float r = 1.0f - s;
    return std::sqrt(r > 0.f ? r : 0.f);
  }
};
struct AccJeffreys {
  static constexpr bool packed = false;  // GPU uses packed f32x2 math for this metric on f32 data
  static inline float step(float acc, float a, float b, bool fused) {
    (void)fused;
    float am = a > 1e-30f ? a : 1e-30f, bm = b > 1e-30f ? b : 1e-30f;
    return acc + (a - b) * std::log(am / bm);
  }
  static inline float finish(float s) { return s; }
};
struct AccJS {
  static constexpr bool packed = false;  // GPU uses packed f32x2 math for this metric on f32 data
  static inline float step(float acc, float a, float b, bool fused) {
    (void)fused;
    float m = 0.5f * (a + b);
    float t = 0.f;
    if (a > 0.f) t += a * std::log(a / m);
    if (b > 0.f) t += b * std::log(b / m);
    return acc + t;
  }
  static inline float finish(float s) {
    float r = 0.5f * s;
    return std::sqrt(r > 0.f ? r : 0.f);
  }
};

// ORDER_REF: 8 lanes over chunks of 8, horizontal add ((0+1)+(2+3))+((4+5)+(6+7)), scalar tail
template <class Acc, class T>
static inline float accumulate_ref(const T* a, const T* b, size_t d) {
  float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t n8 = d / 8 * 8;
  for (size_t i = 0; i < n8; i += 8)
    for (int l = 0; l < 8; ++l) lane[l] = Acc::step(lane[l], (float)a[i + l], (float)b[i + l], false);
  float s = ((lane[0] + lane[1]) + (lane[2] + lane[3])) + ((lane[4] + lane[5]) + (lane[6] + lane[7]));
  for (size_t i = n8; i < d; ++i) s = Acc::step(s, (float)a[i], (float)b[i], false);
  return Acc::finish(s);
}

// ORDER_GPU: see header comment.  Mirrors hnswlib-rs_b200/csrc/dist.cuh exactly.
template <class Acc, class T>
static inline float accumulate_gpu(const T* a, const T* b, size_t d) {
  // a 16-byte chunk holds EPC elements (4 x f32/i32/u32, 8 x u16, 16 x u8); lane g owns chunks g, g+8, ...
  // f32 data with a packed-math metric (L1, L2, Dot): TWO partial sums per lane, elements 0,2 of each 4-element
  // chunk in the first, 1,3 in the second, added once at the end (FADD2/FFMA2 on the GPU).  Otherwise one sum.
  const size_t EPC = 16 / sizeof(T);
  const bool two = std::is_same<T, float>::value && Acc::packed;
  float p[8];
  for (int g = 0; g < 8; ++g) {
    float acc0 = 0.f, acc1 = 0.f;
    for (size_t c4 = g; EPC * c4 < d; c4 += 8)
      for (size_t k = 0; k < EPC; ++k) {
        size_t e = EPC * c4 + k;
        if (e >= d) continue;
        if (two && (k & 1)) acc1 = Acc::step(acc1, (float)a[e], (float)b[e], true);
        else acc0 = Acc::step(acc0, (float)a[e], (float)b[e], true);
      }
    p[g] = two ? acc0 + acc1 : acc0;
  }
  for (int g = 0; g < 4; ++g) p[g] = p[g] + p[g + 4];
  for (int g = 0; g < 2; ++g) p[g] = p[g] + p[g + 2];
  return Acc::finish(p[0] + p[1]);
}

template <class T>
static inline float cosine_ref(const T* a, const T* b, size_t d) {
  double ab = 0, aa = 0, bb = 0;
  for (size_t i = 0; i < d; ++i) {
    double x = (double)a[i], y = (double)b[i];
    ab += x * y;
    aa += x * x;
    bb += y * y;
  }
  if (aa > 0. && bb > 0.) {
    double r = 1. - ab / std::sqrt(aa * bb);
    return (float)(r > 0. ? r : 0.);
  }
  return 0.f;
}

template <class T>
static inline float cosine_gpu(const T* a, const T* b, size_t d) {
  double pab[8], paa[8], pbb[8];
  for (int g = 0; g < 8; ++g) {
    double ab = 0, aa = 0, bb = 0;
    for (size_t c4 = g; 4 * c4 < d; c4 += 8)
      for (size_t k = 0; k < 4; ++k) {
        size_t e = 4 * c4 + k;
        if (e < d) {
          double x = (double)a[e], y = (double)b[e];
          ab = std::fma(x, y, ab);
          aa = std::fma(x, x, aa);
          bb = std::fma(y, y, bb);
        }
      }
    pab[g] = ab; paa[g] = aa; pbb[g] = bb;
  }
  for (int g = 0; g < 4; ++g) { pab[g] += pab[g + 4]; paa[g] += paa[g + 4]; pbb[g] += pbb[g + 4]; }
  for (int g = 0; g < 2; ++g) { pab[g] += pab[g + 2]; paa[g] += paa[g + 2]; pbb[g] += pbb[g + 2]; }
  double ab = pab[0] + pab[1], aa = paa[0] + paa[1], bb = pbb[0] + pbb[1];
  if (aa > 0. && bb > 0.) {
    double r = 1. - ab / std::sqrt(aa * bb);
    return (float)(r > 0. ? r : 0.);
  }
  return 0.f;
}

template <class T>
static inline float hamming(const T* a, const T* b, size_t d) {
  size_t n = 0;
  for (size_t i = 0; i < d; ++i) n += (a[i] != b[i]);
  return (float)n / (float)d;
}

template <class T>
static inline float jaccard(const T* a, const T* b, size_t d) {
  // weighted Jaccard with integer sums (u8/u16/u32); for f32 data sums are taken in f64
  if constexpr (std::is_integral<T>::value) {
    uint64_t mn = 0, mx = 0;
    for (size_t i = 0; i < d; ++i) {
      mn += (uint64_t)(a[i] < b[i] ? a[i] : b[i]);
      mx += (uint64_t)(a[i] < b[i] ? b[i] : a[i]);
    }
    if (mx == 0) return 0.f;
    return (float)(1.0 - (double)mn / (double)mx);
  } else {
    double mn = 0, mx = 0;
    for (size_t i = 0; i < d; ++i) {
      mn += (double)(a[i] < b[i] ? a[i] : b[i]);
      mx += (double)(a[i] < b[i] ? b[i] : a[i]);
    }
    if (mx <= 0.) return 0.f;
    return (float)(1.0 - mn / mx);
  }
}

template <class T>
static inline float eval_dist(int metric, int order, const T* a, const T* b, size_t d) {
  const bool g = (order == ORDER_GPU);
  switch (metric) {
    case METRIC_L1: return g ? accumulate_gpu<AccL1>(a, b, d) : accumulate_ref<AccL1>(a, b, d);
    case METRIC_L2: return g ? accumulate_gpu<AccL2>(a, b, d) : accumulate_ref<AccL2>(a, b, d);
    case METRIC_DOT: return g ? accumulate_gpu<AccDot>(a, b, d) : accumulate_ref<AccDot>(a, b, d);
    case METRIC_COSINE: return g ? cosine_gpu(a, b, d) : cosine_ref(a, b, d);
    case METRIC_HAMMING: return hamming(a, b, d);
    case METRIC_JACCARD: return jaccard(a, b, d);
    case METRIC_HELLINGER: return g ? accumulate_gpu<AccHellinger>(a, b, d) : accumulate_ref<AccHellinger>(a, b, d);
    case METRIC_JEFFREYS: return g ? accumulate_gpu<AccJeffreys>(a, b, d) : accumulate_ref<AccJeffreys>(a, b, d);
    case METRIC_JENSENSHANNON: return g ? accumulate_gpu<AccJS>(a, b, d) : accumulate_ref<AccJS>(a, b, d);
  }
  return NAN;
}

}  // namespace oracle
