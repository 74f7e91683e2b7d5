// K6: MPJPE accumulators (evaluate.py:91-125 + function_mpl.py:674-687) and the batched input builder
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
__global__ void __launch_bounds__(256) mpjpe_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                    const float* __restrict__ conf, int64_t B, int J, float unit,
                                                    double* __restrict__ acc) {
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
        const double pk = (double)p[k] * unit, gk = (double)g[k] * unit;
        const double da = pk - gk;
        const double dr = (pk - (double)p0[k] * unit) - (gk - (double)g0[k] * unit);
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
                            double* acc, cudaStream_t s) {
  if (B == 0) return MPL_OK;
  const int64_t want = ceil_div(B, 8 * 16);  // 8 warps per block, ~16 poses per warp
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 8 * kNumSMs ? 8 * kNumSMs : want));
  mpjpe_kernel<<<grid, 256, (size_t)(11 * J + 1) * sizeof(double), s>>>(pred, gt, conf3d, B, J, unit_scale, acc);
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
