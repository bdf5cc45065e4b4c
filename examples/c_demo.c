/* Minimal C caller of libhnsw_b200.so through the reference's own entry points (libext.rs names):
 * build an index, search it, release everything.  Also proves include/hnsw_b200.h is plain C.
 *   gcc -std=c99 -Iinclude examples/c_demo.c -Lhnswlib-rs_b200/lib -lhnsw_b200 -Wl,-rpath,$PWD/hnswlib-rs_b200/lib -o c_demo
 * Without a usable CUDA device the constructor returns NULL and the demo prints the library's error (no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>

#include "hnsw_b200.h"

int main(void) {
  const char* name = "DistL2";
  const HnswApif32* h = init_hnsw_f32(16, 200, 6, (const uint8_t*)name);
  if (!h) {
    printf("init_hnsw_f32 failed: %s\n", hnsw_b200_last_error());
    return 2;
  }
  enum { N = 2000, D = 16 };
  float* data = (float*)malloc(sizeof(float) * N * D);
  const float** rows = (const float**)malloc(sizeof(float*) * N);
  size_t* ids = (size_t*)malloc(sizeof(size_t) * N);
  unsigned s = 12345u;
  for (int i = 0; i < N * D; ++i) {
    s = s * 1664525u + 1013904223u;
    data[i] = (float)(s >> 8) / 16777216.0f;
  }
  for (int i = 0; i < N; ++i) {
    rows[i] = data + (size_t)i * D;
    ids[i] = 1000 + (size_t)i;
  }
  parallel_insert_f32((HnswApif32*)h, N, D, rows, ids);
  const Neighbourhood_api* r = search_neighbours_f32(h, D, rows[7], 5, 32);
  if (!r) {
    printf("search failed: %s\n", hnsw_b200_last_error());
    return 3;
  }
  printf("query = stored point 7: %lld neighbours, nearest id %zu at distance %g\n", (long long)r->nbgh,
         r->neighbours[0].id, (double)r->neighbours[0].d);
  int ok = r->nbgh == 5 && r->neighbours[0].id == 1007 && r->neighbours[0].d == 0.0f;
  hnsw_b200_free_neighbourhood(r);
  const Vec_api_Neighbourhood_api* v = parallel_search_neighbours_f32(h, 3, D, rows, 4, 32);
  ok = ok && v && v->len == 3 && v->ptr[2].neighbours[0].id == 1002;
  hnsw_b200_free_vec_api(v);
  /* extensions: two batches in flight from this one thread (submit / wait), the replicas call (a no-op on one device) */
  {
    int dev0 = 0;
    uint64_t out_ids[2][8 * 4];
    float out_d[2][8 * 4];
    int32_t cnt[2][8];
    int64_t t0, t1;
    ok = ok && hnsw_b200_replicate((void*)h, 1, &dev0) == 0 && hnsw_b200_replica_count(h) == 0;
    t0 = hnsw_b200_search_flat_submit(h, data, 8, D, 4, 32, out_ids[0], out_d[0], NULL, NULL, cnt[0]);
    t1 = hnsw_b200_search_flat_submit(h, data, 8, D, 4, 32, out_ids[1], out_d[1], NULL, NULL, cnt[1]);
    ok = ok && t0 >= 0 && t1 >= 0 && hnsw_b200_search_flat_wait(h, t0) == 0 && hnsw_b200_search_flat_wait(h, t1) == 0;
    ok = ok && cnt[0][7] == 4 && out_ids[0][7 * 4] == 1007 && out_d[0][7 * 4] == 0.0f && out_ids[1][2 * 4] == 1002;
  }
  drop_hnsw_f32(h);
  free(data);
  free(rows);
  free(ids);
  printf(ok ? "c_demo ok\n" : "c_demo FAILED\n");
  return ok ? 0 : 1;
}
