/* c_api_example.c -- the boundary from plain C99: what a cgo / JNI / N-API binding of libgpc_b200.so would do.
 *
 *   c_api_example <forest.txt> <width> <height> <left.raw> <right.raw> [naive]
 *
 * left.raw / right.raw are headerless 8-bit grayscale images of width x height (width a multiple of 16).  Runs the
 * window sparsematch.cpp:45-52 times (preprocessImage x2 + rectifiedMatch) with the settings of sparsematch.cpp:29-34
 * and prints the candidate counts, the number of supports and an FNV-1a digest of the ordered support list
 * (SURVEY.md 8c).  "naive" selects the results of the reference's SSE=OFF build. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpc_b200.h"

static uint8_t* read_raw(const char* path, size_t n) {
  FILE* f = fopen(path, "rb");
  uint8_t* p = (uint8_t*)malloc(n);
  if (!f || !p || fread(p, 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: %s forest.txt width height left.raw right.raw [naive]\n", argv[0]); return 2; }
  const int w = atoi(argv[2]), h = atoi(argv[3]);
  const size_t n = (size_t)w * (size_t)h;
  uint8_t* left = read_raw(argv[4], n);
  uint8_t* right = read_raw(argv[5], n);

  gpc_ctx* ctx = NULL;
  int rc = gpc_create(&ctx, 0, w, h, 1);
  if (rc != GPC_OK) { fprintf(stderr, "gpc_create: %s\n", gpc_last_error(NULL)); return 3; }
  if (argc > 6 && strcmp(argv[6], "naive") == 0) rc = gpc_set_result_mode(ctx, GPC_RESULTS_NAIVE);
  gpc_forest forest;
  if (rc == GPC_OK) rc = gpc_read_forest(argv[1], &forest);
  if (rc == GPC_OK) rc = gpc_set_forest(ctx, &forest);

  gpc_settings s;
  memset(&s, 0, sizeof s);
  s.gradient_threshold = 5; s.disp_high = 128; s.vertical_tolerance = 0; s.epipolar_mode = 1; s.num_threads = 1;

  const int cap = (w > 26 && h > 26) ? (w - 26) * (h - 26) : 1;
  gpc_support* out = (gpc_support*)malloc((size_t)cap * sizeof(gpc_support));
  int n_out = 0, ncl = 0, ncr = 0;
  if (rc == GPC_OK) rc = gpc_match_pair(ctx, left, right, w, h, w, &s, out, cap, &n_out, &ncl, &ncr);
  if (rc != GPC_OK) { fprintf(stderr, "error %d (%s): %s\n", rc, gpc_status_string(rc), gpc_last_error(ctx)); return 4; }

  uint64_t dg = 1469598103934665603ull;
  for (int i = 0; i < n_out; i++) {
    const int32_t v[3] = {out[i].x, out[i].y, (int32_t)out[i].d};
    for (int k = 0; k < 3; k++) { dg ^= (uint32_t)v[k]; dg *= 1099511628211ull; }
  }
  printf("#candidatesL:%d, #candidatesR:%d, num matches:%d, digest:%016llx\n", ncl, ncr, n_out, (unsigned long long)dg);
  gpc_destroy(ctx);
  free(out); free(left); free(right);
  return 0;
}
