// nn_cells.cu — fused bidirectional 1-NN (Chamfer / Hausdorff / normal borrowing front end) over the cell-grid blobs of
// geoa3_cell_sort: the same (min distance, lowest index) answers as nn_pair.cu, but a query only looks at the cells
// within reach of its current best instead of streaming (a pruned part of) the whole cloud.
//
// One launch covers both directions (blockIdx.x < tiles of direction 0: queries = adv, candidates = ori; the rest:
// swapped).  A CTA stages ONE candidate blob (single TMA bulk copy) and owns 256 consecutive queries of the query
// cloud's cell order, so a warp's 32 queries are neighbours in space.  Every query starts from a seed candidate
// (caller hint, default: the point with the same index — in the attack adv_i is a perturbed ori_i) whose PINNED
// distance initialises `best`; the cell rows within sqrt(best) are walked (row offsets relative to the lane's first row
// are warp-uniform, the x interval per row comes from what is left of best after the row's y/z slab distance, and
// shrinks as best improves) and every candidate met is folded in with the exact lexicographic rule
// (d < best) or (d == best and index < argbest) on the pinned fma chain — bit-identical to the plain ascending scan
// for ANY seed and ANY grid; a bad seed only costs speed.  In an attack step the perturbation is a small fraction of
// the point spacing, so a query meets ~10 candidates instead of the hundreds the box-pruned scan of nn_pair.cu visits.
//
// Exactness of the cell walk: see knn_cells.cu (monotone cell mapping used identically by writer and reader, radii
// inflated over every rounding involved); the query may lie outside the candidate cloud's bounding box (cells clamp).
#include "cells.cuh"

namespace geoa3 {

constexpr int NC_THREADS = 256;

struct NcDir {
  const unsigned char* qblobs;  // blobs of the query cloud (queries are taken in ITS cell order)
  const unsigned char* cblobs;  // blobs of the candidate cloud
  int nq, gq, nc, gc;           // points / cells per axis of the two clouds
  const int32_t* hint;          // [b][nq] seed candidate per ORIGINAL query index (nullable; may alias idx)
  float* dist;                  // [b][nq]
  int32_t* idx;                 // [b][nq]
  int tiles;
};

__global__ void __launch_bounds__(NC_THREADS)
nn_cells_kernel(const NcDir d0, const NcDir d1) {
  extern __shared__ __align__(16) unsigned char nc_smem[];
  __shared__ __align__(8) unsigned long long nc_bar;
  const bool dir1 = (int)blockIdx.x >= d0.tiles;
  const NcDir& D = dir1 ? d1 : d0;
  const int tile = (int)blockIdx.x - (dir1 ? d0.tiles : 0);
  const int cloud = blockIdx.y, tid = threadIdx.x;
  const int nq = D.nq, m = D.nc, G = D.gc;
  const int ncell = G * G * G;
  const size_t cbytes = kc_blob_bytes(m, ncell);
  kc_stage_issue(&nc_bar, nc_smem, D.cblobs + (size_t)cloud * cbytes, (unsigned)cbytes);

  const unsigned char* qblob = D.qblobs + (size_t)cloud * kc_blob_bytes(nq, D.gq * D.gq * D.gq);
  const int slot = tile * NC_THREADS + tid;  // position in the query cloud's cell order
  const bool live = slot < nq;
  const float4 q = reinterpret_cast<const float4*>(qblob + KC_HDR)[min(slot, nq - 1)];
  const int qo = __float_as_int(q.w);        // ORIGINAL index of the query
  int seed = D.hint ? D.hint[(size_t)cloud * nq + qo] : qo;
  seed = (int)min((unsigned)seed, (unsigned)(m - 1));
  kc_stage_wait(&nc_bar);

  const float* sgp = reinterpret_cast<const float*>(nc_smem);
  const unsigned a4 = (unsigned)__cvta_generic_to_shared(nc_smem + KC_HDR);
  const unsigned acs = (unsigned)__cvta_generic_to_shared(nc_smem + kc_cs_off(m));
  const unsigned aip = (unsigned)__cvta_generic_to_shared(nc_smem + kc_ip_off(m, ncell));
  const float lox = sgp[0], loy = sgp[1], loz = sgp[2], ihx = sgp[3], ihy = sgp[4], ihz = sgp[5];
  const float hy = sgp[7], hz = sgp[8], gm1 = sgp[10];
  // slack covers the rounding of q +- r as well: the query need not lie inside the candidate cloud's box
  const float slack = fmaf(8e-6f, fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fabsf(q.z)), sgp[9]);

  float best;
  int bidx;
  {
    const float4 c = kc_lds128(a4 + kc_lds16(aip + (unsigned)seed * 2u) * 16u);
    best = dist2(c.x, c.y, c.z, q.x, q.y, q.z);
    bidx = seed;
  }
  int z0 = 0, y0 = 0, nz = -1, ny = -1;
  if (live) {
    const float r = (best < 3e38f ? kc_sqrt(best * KC_REL) * KC_REL : KC_INF) + slack;
    z0 = kc_cell(q.z - r, loz, ihz, gm1);
    nz = kc_cell(q.z + r, loz, ihz, gm1) - z0;
    y0 = kc_cell(q.y - r, loy, ihy, gm1);
    ny = kc_cell(q.y + r, loy, ihy, gm1) - y0;
  }
  const int wz = __reduce_max_sync(0xffffffffu, nz), wy = __reduce_max_sync(0xffffffffu, ny);
  for (int oz = 0; oz <= wz; ++oz) {
    const int rz = z0 + min(oz, max(nz, 0));
    const float zl = fmaf((float)rz, hz, loz);
    const float ez = fmaxf(fmaxf(zl - q.z, q.z - (zl + hz)) - slack, 0.f);  // lower bound of |c.z - q.z| in this slab
    const float ez2 = ez * ez;
    for (int oy = 0; oy <= wy; ++oy) {
      const int ry = y0 + min(oy, max(ny, 0));
      const float yl = fmaf((float)ry, hy, loy);
      const float ey = fmaxf(fmaxf(yl - q.y, q.y - (yl + hy)) - slack, 0.f);
      // what the CURRENT best leaves for (c.x - q.x)^2 in this row
      const float rem = (oz <= nz && oy <= ny) ? fmaf(best, KC_REL, 1e-37f) - ez2 - ey * ey : -1.f;
      const float rx = (rem < 3e38f ? kc_sqrt(fmaxf(rem, 0.f)) * KC_REL : KC_INF) + slack;
      const unsigned ab = acs + (unsigned)((rz * G + ry) * G) * 2u;
      const int x0 = kc_cell(q.x - rx, lox, ihx, gm1), x1 = kc_cell(q.x + rx, lox, ihx, gm1);
      const int s = (int)kc_lds16(ab + x0 * 2);
      const int len = rem >= 0.f ? (int)kc_lds16(ab + x1 * 2 + 2) - s : 0;
      const int wl = __reduce_max_sync(0xffffffffu, len);
      const unsigned ca = a4 + (unsigned)s * 16u;
      const unsigned cl = ca + (unsigned)max(len - 1, 0) * 16u;  // reads past the lane's own range are clamped to it
      auto fold = [&](int i, const float4 c) {
        const float d = dist2(c.x, c.y, c.z, q.x, q.y, q.z);  // the pinned arithmetic decides
        const int ci = __float_as_int(c.w);
        const bool better = i < len && (d < best || (d == best && ci < bidx));
        best = better ? d : best;
        bidx = better ? ci : bidx;
      };
      int i = 0;
      for (; i + 1 < wl; i += 2) {
        const float4 c0 = kc_lds128(min(ca + (unsigned)i * 16u, cl));
        const float4 c1 = kc_lds128(min(ca + (unsigned)i * 16u + 16u, cl));
        fold(i, c0);
        fold(i + 1, c1);
      }
      if (i < wl) fold(i, kc_lds128(min(ca + (unsigned)i * 16u, cl)));
    }
  }
  if (live) {
    D.dist[(size_t)cloud * nq + qo] = best;
    D.idx[(size_t)cloud * nq + qo] = bidx;
  }
}

}  // namespace geoa3

extern "C" int geoa3_nn_pair_cells(const void* blobs_adv, const void* blobs_ori, int b, int n, int m, int g_adv, int g_ori,
                                   const int32_t* hint_a2o, const int32_t* hint_o2a, float* d_a2o, int32_t* jstar,
                                   float* d_o2a, int32_t* istar, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(blobs_adv && blobs_ori && d_a2o && jstar && b > 0 && n > 0 && m > 0 && g_adv >= 1 && g_ori >= 1);
  GEOA3_CHECK_ARG((d_o2a == nullptr) == (istar == nullptr));
  GEOA3_CHECK_ARG(((reinterpret_cast<uintptr_t>(blobs_adv) | reinterpret_cast<uintptr_t>(blobs_ori)) & 15) == 0);
  if (b > 65535 || n > 65535 || m > 65535 || g_adv > 32 || g_ori > 32) return GEOA3_EUNSUPPORTED;
  const bool both = d_o2a != nullptr;
  const size_t smem = max(kc_blob_bytes(m, g_ori * g_ori * g_ori), both ? kc_blob_bytes(n, g_adv * g_adv * g_adv) : (size_t)0);
  if (smem > 226 * 1024) return GEOA3_EUNSUPPORTED;  // the candidate blob has to fit one CTA's shared memory
  NcDir d0, d1;
  d0.qblobs = reinterpret_cast<const unsigned char*>(blobs_adv);
  d0.cblobs = reinterpret_cast<const unsigned char*>(blobs_ori);
  d0.nq = n; d0.gq = g_adv; d0.nc = m; d0.gc = g_ori;
  d0.hint = hint_a2o; d0.dist = d_a2o; d0.idx = jstar; d0.tiles = ceil_div(n, NC_THREADS);
  d1.qblobs = d0.cblobs; d1.cblobs = d0.qblobs;
  d1.nq = m; d1.gq = g_ori; d1.nc = n; d1.gc = g_adv;
  d1.hint = hint_o2a; d1.dist = d_o2a; d1.idx = istar; d1.tiles = both ? ceil_div(m, NC_THREADS) : 0;
  static PerDeviceOnce once;
  if (once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(nn_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return (int)e;
    once.done();
  }
  nn_cells_kernel<<<dim3(d0.tiles + d1.tiles, b), NC_THREADS, smem, (cudaStream_t)stream>>>(d0, d1);
  return GEOA3_LAUNCH_RESULT();
}
