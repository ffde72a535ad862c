/*
 * geoa3_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement of the GeoA3 hot path, used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the checker for
 * the CUDA kernels in geoa3_b200/csrc.  Index-producing routines reproduce the
 * reference arithmetic bit for bit (fp32, explicit fmaf chain, compile with
 * -ffp-contract=off); value-producing routines (losses, kappa, gradients) are
 * evaluated in fp64 from the fp32 inputs and compared at 1e-5 relative.
 *
 * Two distance chains are pinned: the kNN loop of pytorch3d (`dist += diff*diff`, x then y then z) and the
 * contracted pointnet2_ops expression (y then x then z, see dist2f_pn2).  They differ in the last ulp.
 *
 * Reference statements followed (paths relative to /root/reference):
 *   Lib/loss_utils.py:28-35   chamfer_loss        -> orc_loss_fwd (cd)
 *   Lib/loss_utils.py:45-50   hausdorff_loss      -> orc_loss_fwd (hd)
 *   Lib/loss_utils.py:52-62   _get_kappa_ori      -> orc_knn + orc_kappa
 *   Lib/loss_utils.py:64-82   _get_kappa_adv      -> orc_nn1 + orc_knn + orc_kappa
 *   Lib/loss_utils.py:84-97   curvature_loss      -> orc_loss_fwd (curv)
 *   Lib/utility.py:30-31      _normalize (eps 1e-12 clamp)
 *   .../_ext-src/src/sampling_gpu.cu:59-173   furthest point sampling (tie order, origin skip)
 *   .../_ext-src/src/sampling_gpu.cu:8-47     gather_points(+grad)
 *   .../_ext-src/src/ball_query_gpu.cu:9-44   ball query (first nsample hits, first-hit fill)
 *   .../_ext-src/src/group_points_gpu.cu:8-64 group_points(+grad)
 *   .../_ext-src/src/interpolate_gpu.cu:9-143 three_nn / three_interpolate(+grad)
 *   .../_ext-src/include/cuda_utils.h:13-19   opt_n_threads (FPS block size => tie order)
 * The kNN arithmetic itself lives in un-vendored, un-pinned pytorch3d
 * (knn_points); its published per-pair loop `diff=p1[d]-p2[d]; dist+=diff*diff`
 * compiles to the fma chain below under nvcc's default -fmad=true.  PARITY OF
 * THE kNN BOUNDARY IS PINNED BY THIS FILE + the dense formulation the reference
 * keeps in comments (loss_utils.py:30-31,54-56,67-69), not by a reference test.
 *
 * Layouts: clouds are channel-first [b][3][n] fp32 (as the losses receive them)
 * unless a routine says AoS [b][n][3] (the pointnet2_ops convention).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* squared distance, fp32, exactly: t=dx*dx; t=fma(dy,dy,t); t=fma(dz,dz,t) */
static inline float dist2f(float px, float py, float pz, float qx, float qy, float qz) {
  float dx = px - qx, dy = py - qy, dz = pz - qz;
  float t = dx * dx;
  t = fmaf(dy, dy, t);
  t = fmaf(dz, dz, t);
  return t;
}

/* pointnet2_ops kernels: the source expression (dx*dx)+(dy*dy)+(dz*dz) is contracted by nvcc into
 * t=dy*dy; t=fma(dx,dx,t); t=fma(dz,dz,t)  (read off the SASS of the reference compiled for sm_100:
 * FPS, ball_query and three_nn all start from the y term). */
static inline float dist2f_pn2(float px, float py, float pz, float qx, float qy, float qz) {
  float dx = px - qx, dy = py - qy, dz = pz - qz;
  float t = dy * dy;
  t = fmaf(dx, dx, t);
  t = fmaf(dz, dz, t);
  return t;
}

ORC_API void orc_pairdist(const float *Q, const float *R, int b, int n, int m, float *out) {
  /* out[b][n][m] — the dense matrix the reference comments describe (loss_utils.py:30) */
  for (int c = 0; c < b; ++c) {
    const float *q = Q + (size_t)c * 3 * n, *r = R + (size_t)c * 3 * m;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
        out[((size_t)c * n + i) * m + j] = dist2f(q[i], q[n + i], q[2 * n + i], r[j], r[m + j], r[2 * m + j]);
  }
}

/* 1-NN of every query in R: ties -> lowest index (scan ascending, strict <). loss_utils.py:32,33,48,70,92 */
ORC_API void orc_nn1(const float *Q, const float *R, int b, int n, int m, float *dmin, int32_t *arg) {
  for (int c = 0; c < b; ++c) {
    const float *q = Q + (size_t)c * 3 * n, *r = R + (size_t)c * 3 * m;
    for (int i = 0; i < n; ++i) {
      float best = INFINITY;
      int bi = 0;
      for (int j = 0; j < m; ++j) {
        float d = dist2f(q[i], q[n + i], q[2 * n + i], r[j], r[m + j], r[2 * m + j]);
        if (d < best) { best = d; bi = j; }
      }
      dmin[(size_t)c * n + i] = best;
      arg[(size_t)c * n + i] = bi;
    }
  }
}

/* K smallest (dist, idx) lexicographic, ascending. idx/dist are [b][n][K]. loss_utils.py:57,77 (K=k+1).
 * If m < K the tail is filled with idx -1 / dist +inf. */
ORC_API void orc_knn(const float *Q, const float *R, int b, int n, int m, int K, int32_t *idx, float *dist) {
  float *bd = (float *)malloc(sizeof(float) * K);
  int32_t *bj = (int32_t *)malloc(sizeof(int32_t) * K);
  for (int c = 0; c < b; ++c) {
    const float *q = Q + (size_t)c * 3 * n, *r = R + (size_t)c * 3 * m;
    for (int i = 0; i < n; ++i) {
      int cnt = 0;
      for (int j = 0; j < m; ++j) {
        float d = dist2f(q[i], q[n + i], q[2 * n + i], r[j], r[m + j], r[2 * m + j]);
        if (cnt == K && !(d < bd[K - 1])) continue; /* equal distance, higher index loses */
        int p = cnt < K ? cnt++ : K - 1;
        while (p > 0 && d < bd[p - 1]) { bd[p] = bd[p - 1]; bj[p] = bj[p - 1]; --p; }
        bd[p] = d; bj[p] = j;
      }
      for (int t = 0; t < K; ++t) {
        size_t o = ((size_t)c * n + i) * K + t;
        idx[o] = t < cnt ? bj[t] : -1;
        if (dist) dist[o] = t < cnt ? bd[t] : INFINITY;
      }
    }
  }
  free(bd); free(bj);
}

/* kappa_i = (1/k) sum_m |<n_i, v/max(|v|,1e-12)>|, v = P[nbr[i][m]] - P[i]. fp64 evaluation.
 * nbr is [b][n][k] (already without the dropped column 0). nrm is the per-point normal [b][3][n]
 * (ori: its own normals, adv: normals borrowed through jstar). loss_utils.py:59-62,79-82; utility.py:30-31 */
ORC_API void orc_kappa(const float *P, const float *nrm, const int32_t *nbr, int b, int n, int k, double *kappa) {
  for (int c = 0; c < b; ++c) {
    const float *p = P + (size_t)c * 3 * n, *nn = nrm + (size_t)c * 3 * n;
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int t = 0; t < k; ++t) {
        int j = nbr[((size_t)c * n + i) * k + t];
        /* the subtraction happens in fp32 in the reference (nn_pts - pc), keep it */
        double vx = (double)(float)(p[j] - p[i]), vy = (double)(float)(p[n + j] - p[n + i]),
               vz = (double)(float)(p[2 * n + j] - p[2 * n + i]);
        double L = sqrt(vx * vx + vy * vy + vz * vz);
        if (L < 1e-12) L = 1e-12;
        acc += fabs((vx * nn[i] + vy * nn[n + i] + vz * nn[2 * n + i]) / L);
      }
      kappa[(size_t)c * n + i] = acc / k;
    }
  }
}

/* gather normals through jstar: out[c][:,i] = nrm[c][:, jstar[i]]  (loss_utils.py:71) */
ORC_API void orc_gather3(const float *src, const int32_t *idx, int b, int n_src, int n_out, float *out) {
  for (int c = 0; c < b; ++c)
    for (int ch = 0; ch < 3; ++ch)
      for (int i = 0; i < n_out; ++i)
        out[((size_t)c * 3 + ch) * n_out + i] = src[((size_t)c * 3 + ch) * n_src + idx[(size_t)c * n_out + i]];
}

/* Forward of the fused loss group, fp64 values, fp32-exact index decisions.
 * outputs per cloud: cd, hd, curv; hd_arg = lowest i attaining max dmin (fp32 compare). */
ORC_API void orc_loss_fwd(const float *d_a2o, const float *d_o2a, const int32_t *jstar, const double *kappa_adv,
                          const float *kappa_ori, int b, int n, double *cd, double *hd, int32_t *hd_arg,
                          double *curv) {
  for (int c = 0; c < b; ++c) {
    double s1 = 0, s2 = 0, sc = 0;
    float mx = -1.f;
    int am = 0;
    for (int i = 0; i < n; ++i) {
      size_t o = (size_t)c * n + i;
      s1 += d_a2o[o];
      s2 += d_o2a[o];
      if (d_a2o[o] > mx) { mx = d_a2o[o]; am = i; }
      if (kappa_adv) {
        double e = kappa_adv[o] - (double)kappa_ori[(size_t)c * n + jstar[o]];
        sc += e * e;
      }
    }
    cd[c] = s1 / n + s2 / n;
    hd[c] = mx;
    hd_arg[c] = am;
    curv[c] = sc / n;
  }
}

/* Closed-form gradient of  g_cd*CD + g_hd*HD + g_cu*CUR  w.r.t. adv (matches the
 * reference's autograd through loss_utils.py:28-97 to ~1e-15 in fp64).  G is [b][3][n] fp64.
 * Contributions are accumulated in ascending source order. */
ORC_API void orc_loss_bwd(const float *adv, const float *ori, const float *nrm_adv, const float *kappa_ori,
                          const int32_t *jstar, const int32_t *istar, const int32_t *nbr, const float *d_a2o,
                          const double *g_cd, const double *g_hd, const double *g_cu, int b, int n, int k,
                          double *G) {
  double *kap = (double *)malloc(sizeof(double) * (size_t)b * n);
  if (k > 0) orc_kappa(adv, nrm_adv, nbr, b, n, k, kap);
  for (int c = 0; c < b; ++c) {
    const float *a = adv + (size_t)c * 3 * n, *o = ori + (size_t)c * 3 * n, *nn = nrm_adv + (size_t)c * 3 * n;
    double *g = G + (size_t)c * 3 * n;
    const int32_t *js = jstar + (size_t)c * n, *is = istar + (size_t)c * n;
    memset(g, 0, sizeof(double) * 3 * n);
    double w = g_cd[c] * 2.0 / n;
    for (int i = 0; i < n; ++i)
      for (int ch = 0; ch < 3; ++ch) g[ch * n + i] += w * ((double)a[ch * n + i] - (double)o[ch * n + js[i]]);
    for (int j = 0; j < n; ++j)
      for (int ch = 0; ch < 3; ++ch) g[ch * n + is[j]] += w * ((double)a[ch * n + is[j]] - (double)o[ch * n + j]);
    /* Hausdorff: gradient only to the arg-max row (lowest index on ties) */
    {
      float mx = -1.f;
      int am = 0;
      for (int i = 0; i < n; ++i)
        if (d_a2o[(size_t)c * n + i] > mx) { mx = d_a2o[(size_t)c * n + i]; am = i; }
      for (int ch = 0; ch < 3; ++ch)
        g[ch * n + am] += g_hd[c] * 2.0 * ((double)a[ch * n + am] - (double)o[ch * n + js[am]]);
    }
    if (k <= 0) continue;
    for (int i = 0; i < n; ++i) {
      double gk = g_cu[c] * 2.0 / n * (kap[(size_t)c * n + i] - (double)kappa_ori[(size_t)c * n + js[i]]);
      double nx = nn[i], ny = nn[n + i], nz = nn[2 * n + i];
      for (int t = 0; t < k; ++t) {
        int j = nbr[((size_t)c * n + i) * k + t];
        double vx = (double)(float)(a[j] - a[i]), vy = (double)(float)(a[n + j] - a[n + i]),
               vz = (double)(float)(a[2 * n + j] - a[2 * n + i]);
        double L = sqrt(vx * vx + vy * vy + vz * vz);
        double dx, dy, dz;
        if (L >= 1e-12) {
          double ux = vx / L, uy = vy / L, uz = vz / L;
          double s = ux * nx + uy * ny + uz * nz;
          double sg = (s > 0) - (s < 0);
          double f = gk / k * sg / L;
          dx = f * (nx - ux * s); dy = f * (ny - uy * s); dz = f * (nz - uz * s);
        } else { /* clamp active: no gradient through the norm */
          double s = (vx * nx + vy * ny + vz * nz) / 1e-12;
          double sg = (s > 0) - (s < 0);
          double f = gk / k * sg / 1e-12;
          dx = f * nx; dy = f * ny; dz = f * nz;
        }
        g[j] += dx; g[n + j] += dy; g[2 * n + j] += dz;
        g[i] -= dx; g[n + i] -= dy; g[2 * n + i] -= dz;
      }
    }
  }
  free(kap);
}

/* ---------------------------------------------------------------- pointnet2_ops */

/* cuda_utils.h:13-19 */
ORC_API int orc_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  if (v < 1) v = 1;
  return v;
}

/* sampling_gpu.cu:69-173 (+ sampling.cpp:66-87: temp init 1e10, idxs zero-init). xyz AoS [b][n][3]. */
ORC_API void orc_fps(const float *xyz, int b, int n, int m, int32_t *idxs) {
  if (m <= 0) return;
  int BS = orc_opt_n_threads(n);
  float *temp = (float *)malloc(sizeof(float) * n);
  float *dists = (float *)malloc(sizeof(float) * BS);
  int *dists_i = (int *)malloc(sizeof(int) * BS);
  for (int c = 0; c < b; ++c) {
    const float *p = xyz + (size_t)c * n * 3;
    int32_t *out = idxs + (size_t)c * m;
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    for (int j = 0; j < m; ++j) out[j] = 0;
    int old = 0;
    for (int j = 1; j < m; ++j) {
      float x1 = p[old * 3], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int tid = 0; tid < BS; ++tid) {
        int besti = 0;
        float best = -1.f;
        for (int k = tid; k < n; k += BS) {
          float x2 = p[k * 3], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue;
          float d = dist2f_pn2(x2, y2, z2, x1, y1, z1);
          float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int s = BS / 2; s >= 1; s >>= 1)
        for (int tid = 0; tid < s; ++tid) {
          float v1 = dists[tid], v2 = dists[tid + s];
          int i1 = dists_i[tid], i2 = dists_i[tid + s];
          dists[tid] = v1 > v2 ? v1 : v2;
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      old = dists_i[0];
      out[j] = old;
    }
  }
  free(temp); free(dists); free(dists_i);
}

/* Lib/utility.py:175-187 farthest_points_sample: plain FPS from a given start index per cloud — no frozen points,
 * arg-max ties -> lowest index (torch.argmax), running minimum initialised to +inf, distance = the Euclidean NORM
 * (`torch.norm(diff, dim=1)`, :183): sqrt of (dx*dx + dy*dy) + dz*dz with every product and sum rounded separately —
 * the arithmetic of torch's CUDA norm reduction (checked on a B200: 250 clouds x 1024 picks identical to the torch
 * loop; an fma chain diverges) — so that rounding ties fall the way the reference's do.  xyz AoS [b][n][3]. */
ORC_API void orc_fps_from(const float *xyz, int b, int n, int m, const int32_t *start, int32_t *idxs) {
  if (m <= 0) return;
  float *temp = (float *)malloc(sizeof(float) * n);
  for (int c = 0; c < b; ++c) {
    const float *p = xyz + (size_t)c * n * 3;
    int32_t *out = idxs + (size_t)c * m;
    for (int k = 0; k < n; ++k) temp[k] = INFINITY;
    int old = start[c];
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      float x1 = p[old * 3], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      float best = -1.f;
      int besti = 0;
      for (int k = 0; k < n; ++k) {
        float dx = p[k * 3] - x1, dy = p[k * 3 + 1] - y1, dz = p[k * 3 + 2] - z1;
        float d = sqrtf((dx * dx + dy * dy) + dz * dz);  /* separately rounded squares (-ffp-contract=off) */
        float d2 = fminf(d, temp[k]);
        temp[k] = d2;
        if (d2 > best) { best = d2; besti = k; }
      }
      old = besti;
      out[j] = old;
    }
  }
  free(temp);
}

/* ball_query_gpu.cu:9-44 (+ ball_query.cpp:19-21 zero init). AoS inputs. */
ORC_API void orc_ball_query(const float *new_xyz, const float *xyz, int b, int n, int m, float radius,
                            int nsample, int32_t *idx) {
  float radius2 = radius * radius;
  for (int c = 0; c < b; ++c) {
    const float *p = xyz + (size_t)c * n * 3, *q = new_xyz + (size_t)c * m * 3;
    int32_t *out = idx + (size_t)c * m * nsample;
    for (int j = 0; j < m; ++j) {
      for (int l = 0; l < nsample; ++l) out[j * nsample + l] = 0;
      int cnt = 0;
      for (int k = 0; k < n && cnt < nsample; ++k) {
        float d2 = dist2f_pn2(q[j * 3], q[j * 3 + 1], q[j * 3 + 2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) out[j * nsample + l] = k;
          out[j * nsample + cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* group_points_gpu.cu:8-28: points [b][c][n], idx [b][np][ns] -> out [b][c][np][ns] */
ORC_API void orc_group_points(const float *points, const int32_t *idx, int b, int c, int n, int np, int ns,
                              float *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < np; ++j)
        for (int k = 0; k < ns; ++k)
          out[(((size_t)i * c + l) * np + j) * ns + k] =
              points[((size_t)i * c + l) * n + idx[((size_t)i * np + j) * ns + k]];
}

/* group_points_gpu.cu:43-64, accumulated in ascending (j,k) order, fp64 accumulators */
ORC_API void orc_group_points_grad(const float *grad_out, const int32_t *idx, int b, int c, int n, int np, int ns,
                                   double *grad_points) {
  memset(grad_points, 0, sizeof(double) * (size_t)b * c * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < np; ++j)
        for (int k = 0; k < ns; ++k)
          grad_points[((size_t)i * c + l) * n + idx[((size_t)i * np + j) * ns + k]] +=
              grad_out[(((size_t)i * c + l) * np + j) * ns + k];
}

/* sampling_gpu.cu:8-20 */
ORC_API void orc_gather_points(const float *points, const int32_t *idx, int b, int c, int n, int m, float *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
}

/* sampling_gpu.cu:34-47 */
ORC_API void orc_gather_points_grad(const float *grad_out, const int32_t *idx, int b, int c, int n, int m,
                                    double *grad_points) {
  memset(grad_points, 0, sizeof(double) * (size_t)b * c * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        grad_points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] += grad_out[((size_t)i * c + l) * m + j];
}

/* interpolate_gpu.cu:9-59: unknown [b][n][3], known [b][m][3] -> dist2 [b][n][3] f32, idx [b][n][3].
 * Running best1..3 are doubles initialised to 1e40; d is the fp32 fma chain promoted on compare. */
ORC_API void orc_three_nn(const float *unknown, const float *known, int b, int n, int m, float *dist2,
                          int32_t *idx) {
  for (int c = 0; c < b; ++c) {
    const float *u = unknown + (size_t)c * n * 3, *kn = known + (size_t)c * m * 3;
    for (int j = 0; j < n; ++j) {
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int b1 = 0, b2 = 0, b3 = 0;
      for (int k = 0; k < m; ++k) {
        float d = dist2f_pn2(u[j * 3], u[j * 3 + 1], u[j * 3 + 2], kn[k * 3], kn[k * 3 + 1], kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
        } else if (d < best2) {
          best3 = best2; b3 = b2; best2 = d; b2 = k;
        } else if (d < best3) {
          best3 = d; b3 = k;
        }
      }
      size_t o = ((size_t)c * n + j) * 3;
      dist2[o] = (float)best1; dist2[o + 1] = (float)best2; dist2[o + 2] = (float)best3;
      idx[o] = b1; idx[o + 1] = b2; idx[o + 2] = b3;
    }
  }
}

/* interpolate_gpu.cu:72-101: points [b][c][m], idx/weight [b][n][3] -> out [b][c][n] (fp64 eval) */
ORC_API void orc_three_interpolate(const float *points, const int32_t *idx, const float *weight, int b, int c,
                                   int m, int n, double *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        size_t o = ((size_t)i * n + j) * 3;
        const float *p = points + ((size_t)i * c + l) * m;
        out[((size_t)i * c + l) * n + j] = (double)p[idx[o]] * weight[o] + (double)p[idx[o + 1]] * weight[o + 1] +
                                           (double)p[idx[o + 2]] * weight[o + 2];
      }
}

/* interpolate_gpu.cu:116-143 */
ORC_API void orc_three_interpolate_grad(const float *grad_out, const int32_t *idx, const float *weight, int b,
                                        int c, int n, int m, double *grad_points) {
  memset(grad_points, 0, sizeof(double) * (size_t)b * c * m);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        size_t o = ((size_t)i * n + j) * 3;
        double g = grad_out[((size_t)i * c + l) * n + j];
        double *gp = grad_points + ((size_t)i * c + l) * m;
        gp[idx[o]] += g * weight[o];
        gp[idx[o + 1]] += g * weight[o + 1];
        gp[idx[o + 2]] += g * weight[o + 2];
      }
}
