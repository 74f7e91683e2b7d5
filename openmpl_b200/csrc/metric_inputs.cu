// K6: MPJPE accumulators (evaluate.py:91-125 + function_mpl.py:674-687), P-MPJPE accumulators (pose_utils.py:61-143)
// and the batched input builder
// (joints_dataset_mpl.py:615-648,701-715,762-772,817-820,872-904).  Both are HBM-bound streaming kernels.
#include "kernels.cuh"

namespace mpl {

// acc layout (doubles), L = MPL_METRIC_ACC_LEN(J) = 11 J + 1:
//   [0,J)      sum_b sqrt(nansum_k d^2)  absolute
//   [J,2J)     same, root-relative (a masked root coordinate blanks that coordinate of the whole pose, as the double
//              root-centring of evaluate() + calc_mpjpe(mode='relative') does with NaNs)
//   [2J,5J)    sum_b |d| per joint-dim over unmasked entries, absolute
//   [5J,8J)    same, root-relative
//   [8J,11J)   unmasked count per joint-dim
//   [11J]      pose count
// room: the un-scaling validate() applies to predictions and targets of room-normalised datasets before storing them
// (function_mpl.py:476-488): v * scale + centre per coordinate, in fp32 like the numpy arrays it works on (identity when unused)
struct RoomAffine {
  float scale[3], offset[3];
  int live;
};

__global__ void __launch_bounds__(256) mpjpe_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                    const float* __restrict__ conf, int64_t B, int J, float unit,
                                                    const RoomAffine room, double* __restrict__ acc) {
  extern __shared__ double sacc[];  // 11 J + 1
  const int L = 11 * J + 1;
  for (int i = threadIdx.x; i < L; i += blockDim.x) sacc[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * warps_per_block;
  for (int j0 = 0; j0 < J; j0 += 32) {
    const int j = j0 + lane;
    const bool live = j < J;
    double a_abs = 0, a_rel = 0, d_abs[3] = {0, 0, 0}, d_rel[3] = {0, 0, 0}, cnt[3] = {0, 0, 0};
    for (int64_t b = warp_global; b < B; b += warp_stride) {
      if (!live) continue;
      const float* p = pred + (b * J + j) * 3;
      const float* g = gt + (b * J + j) * 3;
      const float* p0 = pred + b * J * 3;
      const float* g0 = gt + b * J * 3;
      double sa = 0.0, sr = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const bool ok = conf == nullptr || conf[(b * J + j) * 3 + k] > 0.f;
        const bool root_ok = conf == nullptr || conf[b * J * 3 + k] > 0.f;
        float pf = p[k], gf = g[k], p0f = p0[k], g0f = g0[k];
        if (room.live) {
          pf = __fadd_rn(__fmul_rn(pf, room.scale[k]), room.offset[k]);
          gf = __fadd_rn(__fmul_rn(gf, room.scale[k]), room.offset[k]);
          p0f = __fadd_rn(__fmul_rn(p0f, room.scale[k]), room.offset[k]);
          g0f = __fadd_rn(__fmul_rn(g0f, room.scale[k]), room.offset[k]);
        }
        const double pk = (double)pf * unit, gk = (double)gf * unit;
        const double da = pk - gk;
        const double dr = (pk - (double)p0f * unit) - (gk - (double)g0f * unit);
        if (ok) {
          sa = fma(da, da, sa);
          d_abs[k] += fabs(da);
          d_rel[k] += fabs(dr);
          cnt[k] += 1.0;
          if (root_ok) sr = fma(dr, dr, sr);
        }
      }
      a_abs += sqrt(sa);
      a_rel += sqrt(sr);
    }
    if (live) {
      atomicAdd(&sacc[j], a_abs);
      atomicAdd(&sacc[J + j], a_rel);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        atomicAdd(&sacc[2 * J + j * 3 + k], d_abs[k]);
        atomicAdd(&sacc[5 * J + j * 3 + k], d_rel[k]);
        atomicAdd(&sacc[8 * J + j * 3 + k], cnt[k]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L - 1; i += blockDim.x)
    if (sacc[i] != 0.0) atomicAdd(&acc[i], sacc[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&acc[L - 1], (double)B);
}

int launch_mpjpe_accumulate(const float* pred, const float* gt, const float* conf3d, int64_t B, int J, float unit_scale,
                            const float* room_affine, double* acc, cudaStream_t s) {
  if (B == 0) return MPL_OK;
  RoomAffine room{};
  if (room_affine != nullptr) {
    for (int k = 0; k < 3; ++k) { room.scale[k] = room_affine[k]; room.offset[k] = room_affine[3 + k]; }
    room.live = 1;
  }
  const int64_t want = ceil_div(B, 8 * 16);  // 8 warps per block, ~16 poses per warp
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 8 * kNumSMs ? 8 * kNumSMs : want));
  mpjpe_kernel<<<grid, 256, (size_t)(11 * J + 1) * sizeof(double), s>>>(pred, gt, conf3d, B, J, unit_scale, room, acc);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

// ---- P-MPJPE: Procrustes-aligned error (pose_utils.py:61-143) -----------------------------------------------------
// One thread per pose, fp64 throughout; a tile of poses is staged through shared memory with coalesced loads (a pose
// is 3 J consecutive floats, so per-thread global reads would be strided).  Per pose, with A = gt, B = pred (x unit):
// centre both, normalise by their Frobenius norms, M = A0^T B0 (3x3) = U S V^T, R = V U^T, optional reflection rule
// on the smallest singular value, then Z = a_norm * tr(S) * B0 R + mean(A) (scaling) or b_norm * B0 R + mean(A).
// acc layout (doubles), L = MPL_PMETRIC_ACC_LEN(J) = J + 3:
//   [0,J) sum_b ||Z_j - A_j||;  [J] sum_b d (normalised residual);  [J+1] sum_b scale;  [J+2] pose count

// One-sided Jacobi on the columns of G (3x3): on return G = U diag(sigma) with orthogonal columns and the rotations
// accumulated in V, i.e. M = U diag(sigma) V^T.  Unordered sigma >= 0.
__device__ __forceinline__ void svd3_one_sided(double (&G)[3][3], double (&V)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) V[i][k] = i == k ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {              // converges in <= 4 sweeps in practice
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        alpha = fma(G[i][p], G[i][p], alpha);
        beta = fma(G[i][q], G[i][q], beta);
        gamma = fma(G[i][p], G[i][q], gamma);
      }
      if (gamma * gamma <= 1e-30 * alpha * beta) continue;   // columns orthogonal to 1e-15 relative
      rotated = true;
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = rsqrt(1.0 + t * t), sn = c * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double gp = G[i][p], gq = G[i][q];
        G[i][p] = c * gp - sn * gq;
        G[i][q] = sn * gp + c * gq;
        const double vp = V[i][p], vq = V[i][q];
        V[i][p] = c * vp - sn * vq;
        V[i][q] = sn * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
}

__global__ void __launch_bounds__(64) pmpjpe_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                    int64_t B, int J, float unit, int scaling, int reflection,
                                                    double* __restrict__ acc) {
  extern __shared__ __align__(16) unsigned char psm[];
  const int tile = blockDim.x, row = 3 * J;
  double* sacc = reinterpret_cast<double*>(psm);                       // J + 3
  float* sp = reinterpret_cast<float*>(sacc + (J + 3 + 1) / 2 * 2);    // [tile][3 J]
  float* sg = sp + (size_t)tile * row;
  for (int i = threadIdx.x; i < J + 3; i += blockDim.x) sacc[i] = 0.0;
  const int64_t n_tiles = (B + tile - 1) / tile;
  for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
    const int64_t b0 = tl * tile;
    const int n = (int)(B - b0 < tile ? B - b0 : tile);
    __syncthreads();
    for (int i = threadIdx.x; i < n * row; i += blockDim.x) {
      sp[i] = pred[b0 * row + i];
      sg[i] = gt[b0 * row + i];
    }
    __syncthreads();
    const bool live = (int)threadIdx.x < n;
    const float* p = sp + (size_t)threadIdx.x * row;
    const float* g = sg + (size_t)threadIdx.x * row;
    const double u = (double)unit;
    double am[3] = {0, 0, 0}, bm[3] = {0, 0, 0}, R[3][3], zs = 0.0, d = 0.0, scale = 0.0;
    if (live) {
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          am[k] += (double)g[j * 3 + k] * u;
          bm[k] += (double)p[j * 3 + k] * u;
        }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        am[k] /= (double)J;
        bm[k] /= (double)J;
      }
      double ssx = 0, ssy = 0, G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, V[3][3];
      for (int j = 0; j < J; ++j) {
        double a0[3], c0[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          a0[k] = (double)g[j * 3 + k] * u - am[k];
          c0[k] = (double)p[j * 3 + k] * u - bm[k];
          ssx = fma(a0[k], a0[k], ssx);
          ssy = fma(c0[k], c0[k], ssy);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int k = 0; k < 3; ++k) G[i][k] = fma(a0[i], c0[k], G[i][k]);
      }
      const double a_norm = sqrt(ssx), b_norm = sqrt(ssy), inv = 1.0 / (a_norm * b_norm);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) G[i][k] *= inv;
      svd3_one_sided(G, V);
      double sig[3], U[3][3];
      int imin = 0, imax = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sig[c] = sqrt(G[0][c] * G[0][c] + G[1][c] * G[1][c] + G[2][c] * G[2][c]);
        if (sig[c] < sig[imin]) imin = c;
        if (sig[c] > sig[imax]) imax = c;
      }
      // left vectors; a vanishing singular value (planar / collinear prediction) gets the completion of the others
      const double tiny = 1e-14 * sig[imax];
      int n_ok = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const bool ok = sig[c] > tiny;
        n_ok += ok;
#pragma unroll
        for (int i = 0; i < 3; ++i) U[i][c] = ok ? G[i][c] / sig[c] : 0.0;
      }
      if (n_ok == 2) {
        const int c = imin, a = (c + 1) % 3, b = (c + 2) % 3;
        U[0][c] = U[1][a] * U[2][b] - U[2][a] * U[1][b];
        U[1][c] = U[2][a] * U[0][b] - U[0][a] * U[2][b];
        U[2][c] = U[0][a] * U[1][b] - U[1][a] * U[0][b];
      }
      if (reflection >= 0) {                                  // :112-120, det(R) = det(V) det(U)
        double detR;
        {
          double Rt[3][3];
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) Rt[a][b] = V[a][0] * U[b][0] + V[a][1] * U[b][1] + V[a][2] * U[b][2];
          detR = Rt[0][0] * (Rt[1][1] * Rt[2][2] - Rt[1][2] * Rt[2][1]) - Rt[0][1] * (Rt[1][0] * Rt[2][2] - Rt[1][2] * Rt[2][0]) +
                 Rt[0][2] * (Rt[1][0] * Rt[2][1] - Rt[1][1] * Rt[2][0]);
        }
        if ((reflection != 0) != (detR < 0.0)) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c == imin) {
              V[0][c] = -V[0][c];
              V[1][c] = -V[1][c];
              V[2][c] = -V[2][c];
              sig[c] = -sig[c];
            }
        }
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) R[a][b] = V[a][0] * U[b][0] + V[a][1] * U[b][1] + V[a][2] * U[b][2];
      const double s_trace = sig[0] + sig[1] + sig[2];
      if (scaling) {
        scale = s_trace * a_norm / b_norm;
        d = 1.0 - s_trace * s_trace;
        zs = scale;                                           // a_norm * s_trace * (B0 / b_norm)
      } else {
        scale = 1.0;
        d = 1.0 + ssy / ssx - 2.0 * s_trace * b_norm / a_norm;
        zs = 1.0;                                             // b_norm * (B0 / b_norm)
      }
    }
    const int lane = threadIdx.x & 31;
    for (int j = 0; j < J; ++j) {
      double e = 0.0;
      if (live) {
        double c0[3], acc2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) c0[k] = (double)p[j * 3 + k] * u - bm[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double z = zs * (c0[0] * R[0][k] + c0[1] * R[1][k] + c0[2] * R[2][k]) + am[k];
          const double df = z - (double)g[j * 3 + k] * u;
          acc2 = fma(df, df, acc2);
        }
        e = sqrt(acc2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if (lane == 0) atomicAdd(&sacc[j], e);
    }
    double sd = live ? d : 0.0, ssc = live ? scale : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
      ssc += __shfl_xor_sync(0xffffffffu, ssc, o);
    }
    if (lane == 0) {
      atomicAdd(&sacc[J], sd);
      atomicAdd(&sacc[J + 1], ssc);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < J + 2; i += blockDim.x) atomicAdd(&acc[i], sacc[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&acc[J + 2], (double)B);
}

int launch_pmpjpe_accumulate(const float* pred, const float* gt, int64_t B, int J, float unit_scale, int scaling,
                             int reflection, double* acc, cudaStream_t s) {
  if (B == 0) return MPL_OK;
  int tile = 64;
  auto smem = [&](int t) { return (size_t)((J + 3 + 1) / 2 * 2) * sizeof(double) + (size_t)2 * t * 3 * J * sizeof(float); };
  if (smem(tile) > 48 * 1024) tile = 32;
  if (smem(tile) > 48 * 1024) {
    set_error("mpl_pmpjpe_accumulate: num_joints too large for the shared-memory pose tile");
    return MPL_ERR_INVALID_ARGUMENT;
  }
  const int64_t n_tiles = ceil_div(B, (int64_t)tile);
  const unsigned grid = (unsigned)(n_tiles > 16 * kNumSMs ? 16 * kNumSMs : n_tiles);
  pmpjpe_kernel<<<grid, tile, smem(tile), s>>>(pred, gt, B, J, unit_scale, scaling, reflection, acc);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

// One thread per (pose, view, joint).  Double arithmetic in the order numpy applies it (the dataset code works in
// float64 and casts to float32 at the end), so results are bit-identical to the reference's per-sample path.
__global__ void __launch_bounds__(256) build_inputs_kernel(const float* __restrict__ pix, const double* __restrict__ calib,
                                                           int64_t B, int V, int J, float* __restrict__ poses,
                                                           float* __restrict__ rays, float* __restrict__ centers) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = B * V * J;
  if (idx >= total) return;
  const int j = (int)(idx % J);
  const int v = (int)((idx / J) % V);
  const double* c = calib + v * 18;
  const double w = c[16], h = c[17];
  double u = pix[idx * 3], vv = pix[idx * 3 + 1], conf = pix[idx * 3 + 2];
  // clip + confidence zeroing (joints_dataset_mpl.py:709-715)
  if (!(0.0 < u)) conf = 0.0;
  if (!(u < w - 1.0)) conf = 0.0;
  if (!(0.0 < vv)) conf = 0.0;
  if (!(vv < h - 1.0)) conf = 0.0;
  u = fmin(fmax(u, 0.0), w - 1.0);
  vv = fmin(fmax(vv, 0.0), h - 1.0);
  // screen normalisation (:817-820) of the joint and of the intrinsics (:615-623)
  const double x = (u / w) * 2.0 - 1.0, y = (vv / w) * 2.0 - h / w;
  const double cx = ((double)c[14] / w) * 2.0 - 1.0, cy = ((double)c[15] / w) * 2.0 - h / w;
  const double fx = (double)c[12] / w * 2.0, fy = (double)c[13] / w * 2.0;
  poses[idx * 3] = (float)x;
  poses[idx * 3 + 1] = (float)y;
  poses[idx * 3 + 2] = (float)conf;
  // rays = R^T [ (x - cx) / fx, (y - cy) / fy, 1 ] + t   (:872-898, USE_T)
  const double dx = (x - cx) / fx, dy = (y - cy) / fy, dz = 1.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double r = (double)c[0 + k] * dx + (double)c[3 + k] * dy + (double)c[6 + k] * dz + (double)c[9 + k];
    rays[idx * 3 + k] = (float)r;
  }
  if (j == 0) {
    const int64_t bv = idx / J;
#pragma unroll
    for (int k = 0; k < 3; ++k) centers[bv * 3 + k] = (float)c[9 + k];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// N3: MHP-style synthetic projector.  3D poses from a counter-based generator keyed by (seed, GLOBAL pose index) --
// any sharding of the index range over ranks / micro-batches yields the same data -- pushed through V calibrations:
// x_cam = R (X - t), u = f x_cam / z + c (MPL/lib/utils/calib.py:42-77).  Emits raw detector-style pixels (u, v, conf)
// for mpl_build_inputs and the 3D target.  The generator is numpy's Philox4x64-10 stream, bit for bit (openmpl_b200/
// synth.py draws the same uniforms on the host), so the device data can be checked against the host generator.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x64_10(uint64_t n, uint64_t key0, uint64_t (&out)[4]) {
  uint64_t c0 = n, c1 = 0, c2 = 0, c3 = 0, k0 = key0, k1 = 0;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) { k0 += 0x9E3779B97F4A7C15ull; k1 += 0xBB67AE8584CAA73Bull; }
    const uint64_t hi0 = __umul64hi(0xD2E7470EE14C6C93ull, c0), lo0 = 0xD2E7470EE14C6C93ull * c0;
    const uint64_t hi1 = __umul64hi(0xCA5A826395121157ull, c2), lo1 = 0xCA5A826395121157ull * c2;
    const uint64_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct UniformCursor {  // sequential reader of one pose's uniform stream with a one-block cache
  uint64_t v[4];
  int64_t blk;
  uint64_t base, key;
  __device__ UniformCursor(uint64_t base_, uint64_t key_) : blk(-1), base(base_), key(key_) {}
  __device__ double get(int i) {
    const int64_t b = i >> 2;
    if (b != blk) { philox4x64_10(base + (uint64_t)b + 1ull, key, v); blk = b; }  // numpy increments the counter before generating
    const int k = i & 3;
    const uint64_t raw = k == 0 ? v[0] : (k == 1 ? v[1] : (k == 2 ? v[2] : v[3]));
    return (double)(raw >> 11) * (1.0 / 9007199254740992.0);
  }
};

__global__ void __launch_bounds__(128) synth_project_kernel(uint64_t seed, int64_t start, int64_t B, int V, int J,
                                                            const double* __restrict__ calib, const double* __restrict__ room,
                                                            int conf_ones, float* __restrict__ pix, float* __restrict__ target) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = 3 * (J - 1);
  const int per_pose = 3 + 2 * n + V * J;
  const uint64_t blocks = (uint64_t)((per_pose + 3) / 4);
  const uint64_t base = (uint64_t)(start + b) * blocks;
  UniformCursor ca(base, seed), cb(base, seed);
  double root[3];
  root[0] = room[0] + ca.get(0) * (room[1] - room[0]);
  root[1] = room[2] + ca.get(1) * (room[3] - room[2]);
  root[2] = 0.8 + 0.2 * ca.get(2);
  float* tg = target + b * J * 3;
  float* px = pix + b * V * J * 3;
  UniformCursor cc(base, seed);
  for (int j = 0; j < J; ++j) {
    double X[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      X[k] = root[k];
      if (j > 0) {
        const int i = (j - 1) * 3 + k;
        const double u1 = fmax(ca.get(3 + i), 1e-12), u2 = cb.get(3 + n + i);
        X[k] += 0.25 * (sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2));  // Box-Muller, fixed draw count
      }
      tg[j * 3 + k] = (float)X[k];
    }
    for (int v = 0; v < V; ++v) {
      const double* c = calib + v * 18;
      double xc[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        xc[r] = c[3 * r] * (X[0] - c[9]) + c[3 * r + 1] * (X[1] - c[10]) + c[3 * r + 2] * (X[2] - c[11]);
      const double u = xc[0] / xc[2] * c[12] + c[14], w = xc[1] / xc[2] * c[13] + c[15];
      double conf = conf_ones ? 1.0 : 0.3 + 0.7 * cc.get(3 + 2 * n + v * J + j);
      if (!(xc[2] > 0.0)) conf = 0.0;  // behind the camera: never a detection
      float* o = px + (v * J + j) * 3;
      o[0] = (float)u; o[1] = (float)w; o[2] = (float)conf;
    }
  }
}

int launch_synth_project(uint64_t seed, int64_t start, int64_t B, int V, int J, const double* calib, const double* room,
                         int conf_ones, float* pix, float* target, cudaStream_t s) {
  if (B == 0) return MPL_OK;
  synth_project_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(seed, start, B, V, J, calib, room, conf_ones, pix, target);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

int launch_build_inputs(const float* pix, const double* calib, int64_t B, int V, int J, float* poses, float* rays,
                        float* centers, cudaStream_t s) {
  const int64_t total = B * V * J;
  if (total == 0) return MPL_OK;
  build_inputs_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(pix, calib, B, V, J, poses, rays, centers);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

}  // namespace mpl
