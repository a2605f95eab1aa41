/* pool_example.c -- every GPU of the box on one batch of stereo pairs, from plain C99 (include/gpc_b200.h).
 *
 *   pool_example <forest.txt> <w> <h> <n_pairs> [n_gpus]
 *
 * Makes n_pairs copies of a deterministic test pattern (left = pattern, right = pattern shifted by 7 pixels), runs
 * gpc_pool_match_batch over n_gpus devices (default: device 0 twice -- two contexts on one GPU, which exercises the same
 * code on a single-GPU box) and checks that every pair returned the same supports.  Prints "ok <supports per pair>".
 * There is no reference counterpart: the reference processes one pair per call on one CPU core (sparsematch.cpp:45-52). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpc_b200.h"

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: pool_example forest w h n_pairs [n_gpus]\n"); return 2; }
  const int w = atoi(argv[2]), h = atoi(argv[3]), n_pairs = atoi(argv[4]);
  int n_dev = argc > 5 ? atoi(argv[5]) : 0;
  int devices[64];
  if (n_dev <= 0) { n_dev = 2; devices[0] = 0; devices[1] = 0; }
  else for (int i = 0; i < n_dev && i < 64; i++) devices[i] = i;
  const size_t P = (size_t)w * h;
  gpc_forest forest;
  if (gpc_read_forest(argv[1], &forest) != GPC_OK) { fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }

  gpc_pool* pool = NULL;
  if (gpc_pool_create(&pool, devices, n_dev, w, h, n_pairs) != GPC_OK) { fprintf(stderr, "gpc_pool_create: %s\n", gpc_last_error(NULL)); return 4; }
  if (gpc_pool_set_forest(pool, &forest) != GPC_OK) { fprintf(stderr, "set_forest: %s\n", gpc_pool_last_error(pool)); return 5; }

  /* page-locked, portable host buffers: every device DMAs from / into them directly */
  uint8_t* images = (uint8_t*)gpc_host_alloc(2 * P * (size_t)n_pairs);
  const int64_t cap = (int64_t)n_pairs * (w - 26) * (h - 26);
  gpc_support* out = (gpc_support*)gpc_host_alloc((size_t)cap * sizeof(gpc_support));
  int64_t* offsets = (int64_t*)malloc(((size_t)n_pairs + 1) * sizeof(int64_t));
  if (!images || !out || !offsets) return 6;
  uint32_t s = 12345u;
  for (size_t i = 0; i < P; i++) { s = s * 1664525u + 1013904223u; images[i] = (uint8_t)(s >> 24); }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) images[P + (size_t)y * w + x] = images[(size_t)y * w + (x + 7) % w];
  for (int p = 1; p < n_pairs; p++) memcpy(images + 2 * P * (size_t)p, images, 2 * P);

  gpc_settings st = {5, 128, 0, 1, 0, 1};              /* sparsematch.cpp:29-34 */
  const int rc = gpc_pool_match_batch(pool, images, n_pairs, w, h, &st, out, cap, offsets, NULL);
  if (rc != GPC_OK) { fprintf(stderr, "gpc_pool_match_batch: %s\n", gpc_pool_last_error(pool)); return 7; }
  const int64_t n0 = offsets[1] - offsets[0];
  for (int p = 1; p < n_pairs; p++) {
    if (offsets[p + 1] - offsets[p] != n0 || memcmp(out + offsets[p], out, (size_t)n0 * sizeof(gpc_support)) != 0) {
      fprintf(stderr, "pair %d differs from pair 0\n", p);
      return 8;
    }
  }
  printf("ok %lld\n", (long long)n0);
  gpc_host_free(images); gpc_host_free(out); free(offsets);
  gpc_pool_destroy(pool);
  return 0;
}
