// Standalone harness: runs the real match_rows_kernel on dumped hash rows many times and
// compares the emitted (x) sets with a CPU evaluation.  Build: nvcc ... match_harness.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>
#include <algorithm>
#include "../../opengpc_b200/csrc/match_rows.cu"
using namespace gpc;
int main(int argc, char** argv) {
  const int W = 1024, H = (argc > 1 ? atoi(argv[1]) : 64), Y = 20;
  std::vector<uint32_t> rows(4 * W);
  FILE* f = fopen("scripts/micro/rows.bin", "rb"); if (!f) { printf("no rows.bin\n"); return 1; }
  if (fread(rows.data(), 4, rows.size(), f) != rows.size()) return 1; fclose(f);
  int smem_max = 200 * 1024; configure_match_rows(smem_max);
  for (int which = 0; which < 2; which++) {
    std::vector<uint32_t> h((size_t)2 * H * W, 0);
    for (int x = 0; x < W; x++) { h[(size_t)Y * W + x] = rows[(2 * which) * W + x]; h[(size_t)(H + Y) * W + x] = rows[(2 * which + 1) * W + x]; }
    h[(size_t)(H + Y + 1) * W + 500] = 0x80000000u | 0x1234567u;
    std::map<uint32_t, std::pair<int,int>> cl, cr;   // key -> (count, x)
    for (int x = 0; x < W; x++) {
      uint32_t a = h[(size_t)Y * W + x], b = h[(size_t)(H + Y) * W + x];
      if (a >> 31) { auto& e = cl[a & 0x7fffffffu]; e.first++; e.second = x; }
      if (b >> 31) { auto& e = cr[b & 0x7fffffffu]; e.first++; e.second = x; }
    }
    std::vector<int> expx;
    for (auto& kv : cl) if (kv.second.first == 1) { auto it = cr.find(kv.first); if (it != cr.end() && it->second.first == 1) expx.push_back(kv.second.second); }
    uint32_t *d_hash, *d_stage; int32_t *d_last, *d_rowmatch;
    cudaMalloc(&d_hash, h.size() * 4); cudaMalloc(&d_stage, (size_t)H * W * 4); cudaMalloc(&d_last, 8); cudaMalloc(&d_rowmatch, H * 4);
    cudaMemcpy(d_hash, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int last[2] = {Y, Y + 1}; cudaMemcpy(d_last, last, 8, cudaMemcpyHostToDevice);
    MatchArgs a{}; a.hash = d_hash; a.lastrow = d_last; a.stage = d_stage; a.rowmatch = d_rowmatch; a.W = W; a.H = H;
    a.disp_high = 5000; a.vertical_tolerance = 0; a.wcap = W - 26; a.table_log2 = 12; a.x_bits = 10; a.pow2cap = 1024; a.key_bits = 31;
    int nbad = 0;
    for (int rep = 0; rep < 200; rep++) {
      cudaMemset(d_stage, 0xee, (size_t)H * W * 4);
      launch_match_rows(a, 1, 0);
      std::vector<int32_t> rm(H); std::vector<uint32_t> st(W);
      cudaMemcpy(rm.data(), d_rowmatch, H * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(st.data(), d_stage + (size_t)Y * W, W * 4, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<int> gx; for (int i = 0; i < rm[Y] && i < W; i++) gx.push_back((int)(st[i] >> 16));
      std::vector<int> ex = expx; std::sort(ex.begin(), ex.end()); std::vector<int> g2 = gx; std::sort(g2.begin(), g2.end());
      if (ex != g2) { nbad++; if (nbad <= 3) printf("  row set %d rep %d: expected %zu got %d\n", which, rep, ex.size(), rm[Y]); }
    }
    printf("row set %d: %d bad runs of 200 (expected %zu matches)\n", which, nbad, expx.size());
  }
  return 0;
}
