// The HBM-bound edge kernels of the forward for the shipped widths (d % 4 == 0): K1 joint embedding, FPT token build and
// the K5 fused head, written so that every global access is a 16-byte (or a coalesced 128-byte-per-warp) transaction
// and the per-element index arithmetic of the generic kernels (kernels_generic.cu) disappears.  The generic kernels stay
// as the path for odd widths / unaligned pointers; launch_* in kernels_generic.cu dispatch here first.
#include <cuda_fp16.h>

#include <algorithm>

#include <atomic>

#include "kernels.cuh"

namespace mpl {

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------------------------------------
// K1 joint embedding (multiview_mpl.py:349-398).  grid.y = view; a thread owns 4 consecutive channels for the whole
// launch (its rows of W_e / b_e / W_c live in registers) and walks over token rows; a warp writes 512 contiguous bytes.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_vec_kernel(const EmbedArgs a) {
  const int v = blockIdx.y;
  const int d4 = a.d >> 2;                 // float4 chunks per token row
  const int c = (threadIdx.x % d4) * 4;    // first channel of this thread
  const int rows_per_pass = 256 / d4;
  const float* We = a.We[v] + c * a.in_ch;
  float w[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    w[i][0] = __ldg(We + i * a.in_ch);
    w[i][1] = __ldg(We + i * a.in_ch + 1);
    w[i][2] = (a.in_ch == 3) ? __ldg(We + i * a.in_ch + 2) : 0.f;
  }
  const float4 be = ld4(a.be[v] + c);
  const bool conf_emb = a.add_conf || a.mult_conf;
  float4 wc = make_float4(0.f, 0.f, 0.f, 0.f), bc = wc;
  if (conf_emb) { wc = ld4(a.Wc[v] + c); bc = ld4(a.bc[v] + c); }
  float wl[4][3];
  float4 bl = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.spatial_pos_mode == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { wl[i][0] = __ldg(a.Wl + (c + i) * 3); wl[i][1] = __ldg(a.Wl + (c + i) * 3 + 1); wl[i][2] = __ldg(a.Wl + (c + i) * 3 + 2); }
    bl = ld4(a.bl + c);
  }
  const int64_t rows = a.B * a.J;
  const float* poses = a.poses[v];
  float* xv = a.x + (int64_t)v * rows * a.d;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_pass + threadIdx.x / d4; row < rows; row += (int64_t)gridDim.x * rows_per_pass) {
    const int64_t b = row / a.J;
    const int j = (int)(row - b * a.J);
    const float* p = poses + b * a.pose_stride + j * 3;
    const float px = __ldg(p), py = __ldg(p + 1), pc = __ldg(p + 2);
    float o[4];
    const float bev[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = fmaf(w[i][2], pc, fmaf(w[i][1], py, fmaf(w[i][0], px, bev[i])));  // w[i][2] = 0 when in_ch == 2
    if (conf_emb) {
      const float ce[4] = {fmaf(wc.x, pc, bc.x), fmaf(wc.y, pc, bc.y), fmaf(wc.z, pc, bc.z), fmaf(wc.w, pc, bc.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (a.add_conf) o[i] += ce[i];
        if (a.mult_conf) o[i] *= ce[i];
      }
    }
    const float4 ps = ld4(a.Ps[v] + j * a.d + c);
    o[0] += ps.x; o[1] += ps.y; o[2] += ps.z; o[3] += ps.w;
    if (a.spatial_pos_mode == 1) {
      const float4 t = ld4(a.pos3d + j * a.pos3d_ld + c);
      o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
    } else if (a.spatial_pos_mode == 2) {
      const float* r = a.rays[v] + b * a.pose_stride + j * 3;
      const float* ce = a.centers[v] + b * a.center_stride;
      const float dx = __ldg(r) - __ldg(ce), dy = __ldg(r + 1) - __ldg(ce + 1), dz = __ldg(r + 2) - __ldg(ce + 2);
      const float inv = 1.0f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);  // F.normalize eps
      const float blv[4] = {bl.x, bl.y, bl.z, bl.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] += fmaf(wl[i][2], dz * inv, fmaf(wl[i][1], dy * inv, fmaf(wl[i][0], dx * inv, blv[i])));
    }
    *reinterpret_cast<float4*>(xv + row * a.d + c) = make_float4(o[0], o[1], o[2], o[3]);
    if (a.conf != nullptr && c == 0) a.conf[(int64_t)v * rows + row] = pc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// FPT token build (multiview_mpl.py:463-499): one thread per 4 consecutive channels of tok [B, V, tok_w]; the three
// layouts (no ray token / [x | ray] per joint / J pose tokens then J ray tokens) differ only in how a 4-channel chunk
// maps to (joint, segment), and d % 4 == 0 keeps every chunk inside one segment.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) token_build_vec_kernel(const TokenArgs a) {
  const int w4 = a.tok_w >> 2;
  const int64_t total = a.B * a.V * (int64_t)w4;
  const int d = a.d, J = a.J;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % w4) * 4;
    const int64_t bv = idx / w4;
    const int v = (int)(bv % a.V);
    const int64_t b = bv / a.V;
    int j, c, pos_c;
    bool is_ray = false, add_pos = true;
    if (a.ray_layout == 1) {
      j = e / (2 * d);
      c = e - j * 2 * d;
      pos_c = c;
      if (c >= d) { is_ray = true; c -= d; }
    } else if (a.ray_layout == 2) {
      const int t = e / d;
      c = e - t * d;
      pos_c = c;
      if (t >= J) { is_ray = true; j = t - J; add_pos = false; } else { j = t; }
    } else {
      j = e / d;
      c = e - j * d;
      pos_c = c;
    }
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (is_ray || (add_pos && a.pos_table == nullptr)) {
      const float* r = a.rays[v] + b * a.pose_stride + j * 3;
      const float* ce = a.centers[v] + b * a.center_stride;
      dx = __ldg(r) - __ldg(ce);
      dy = __ldg(r + 1) - __ldg(ce + 1);
      dz = __ldg(r + 2) - __ldg(ce + 2);
    }
    float o[4];
    if (is_ray) {
      const float4 br = ld4(a.br + c);
      const float brv[4] = {br.x, br.y, br.z, br.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* wr = a.Wr + (c + i) * 3;
        o[i] = fmaf(__ldg(wr + 2), dz, fmaf(__ldg(wr + 1), dy, fmaf(__ldg(wr), dx, brv[i])));
      }
    } else {
      const float4 xv = *reinterpret_cast<const float4*>(a.xn + (((int64_t)v * a.B + b) * J + j) * d + c);
      o[0] = xv.x; o[1] = xv.y; o[2] = xv.z; o[3] = xv.w;
      if (a.Wcf != nullptr) {
        const float pc = __ldg(a.poses[v] + b * a.pose_stride + j * 3 + 2);
        const float4 wc = ld4(a.Wcf + c), bc = ld4(a.bcf + c);
        o[0] += fmaf(wc.x, pc, bc.x); o[1] += fmaf(wc.y, pc, bc.y); o[2] += fmaf(wc.z, pc, bc.z); o[3] += fmaf(wc.w, pc, bc.w);
      }
    }
    if (add_pos) {
      if (a.pos_table != nullptr) {
        const float4 t = ld4(a.pos_table + j * a.pos_w + pos_c);
        o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
      } else {
        const float inv = 1.0f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
        const float4 bl = ld4(a.bl + pos_c);
        const float blv[4] = {bl.x, bl.y, bl.z, bl.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* wl = a.Wl + (pos_c + i) * 3;
          o[i] += fmaf(__ldg(wl + 2), dz * inv, fmaf(__ldg(wl + 1), dy * inv, fmaf(__ldg(wl), dx * inv, blv[i])));
        }
      }
    }
    *reinterpret_cast<float4*>(a.tok + idx * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K5 fused head (multiview_mpl.py:425-446, 517-523).  Pooling of ONE pose by one warp: the stripped row is S = E / 32
// segments of 32 channels; a lane owns two consecutive channels of two segments at a time (lanes 0-15: segment 2k,
// lanes 16-31: segment 2k + 1), so every access is an 8-byte load of a fully used 128-byte line and the arithmetic runs
// on packed f32x2 instructions.  View_norm and the head LayerNorm reduce with shuffles only — no shared memory, no block
// barriers.  The views are taken four at a time: all 4 x NK loads of a group are in flight together.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }

// NK: slots of two segments (ceil(S / 2) <= NK); SS: segment stride in floats, 0 = a.seg_stride; ST: S, 0 = a.E / 32
// PL: the token rows come as two bf16 planes (hi + lo, LayerNorm-fused bf16 mode): two 4-byte loads instead of one 8-byte one
template <int NK, int SS, int ST, bool PL>
__device__ __forceinline__ void head_pool_pose(const HeadArgs& a, int64_t b, int lane, float2 (&p)[NK]) {
  const int S = ST ? ST : (a.E >> 5), hi = lane >> 4;
  const int stride = SS ? SS : a.seg_stride;
  const int col0 = hi * stride + (lane & 15) * 2;  // slot k: + 2 k stride
  const int ch0 = hi * 32 + (lane & 15) * 2;       // slot k: + 64 k
  const float invE = 1.0f / (float)a.E;
  bool ok[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) ok[k] = (ST && 2 * k + 1 < ST) || 2 * k + hi < S;
  float2 acc[NK];  // sum_v (w_v rstd_v) (x_v - m_v)
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = f2(0.f);
  float wsum = 0.f;
  for (int v0 = 0; v0 < a.V; v0 += 4) {
    float2 x[4][NK];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // a view past V re-reads view V - 1 and is dropped by its zero weight below
      const int64_t roff = (b * a.V + min(v0 + u, a.V - 1)) * (int64_t)a.tok_w + col0;
      if constexpr (PL) {
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          x[u][k] = f2(0.f);
          if (ok[k]) {
            const uint32_t h = *reinterpret_cast<const uint32_t*>(a.tok_hi + roff + 2 * k * stride);
            const uint32_t l = *reinterpret_cast<const uint32_t*>(a.tok_lo + roff + 2 * k * stride);
            x[u][k] = make_float2(__uint_as_float(h << 16) + __uint_as_float(l << 16),
                                  __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u));
          }
        }
      } else {
        const float* r = a.tok + roff;
#pragma unroll
        for (int k = 0; k < NK; ++k) x[u][k] = ok[k] ? *reinterpret_cast<const float2*>(r + 2 * k * stride) : f2(0.f);
      }
    }
    float m[4], rs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 s = x[u][0];
#pragma unroll
      for (int k = 1; k < NK; ++k) s = __fadd2_rn(s, x[u][k]);
      m[u] = s.x + s.y;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) m[u] = warp_sum(m[u]) * invE;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 nm = f2(-m[u]);
      float2 q = f2(0.f);
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        x[u][k] = ok[k] ? __fadd2_rn(x[u][k], nm) : f2(0.f);
        q = __ffma2_rn(x[u][k], x[u][k], q);
      }
      rs[u] = q.x + q.y;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float w = (v0 + u < a.V) ? __ldg(a.wm_w + min(v0 + u, a.V - 1)) : 0.f;
      rs[u] = (v0 + u < a.V) ? rsqrtf(warp_sum(rs[u]) * invE + 1e-6f) * w : 0.f;
      wsum += w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 r2 = f2(rs[u]);
#pragma unroll
      for (int k = 0; k < NK; ++k) acc[k] = __ffma2_rn(r2, x[u][k], acc[k]);
    }
  }
  // sum_v w_v ((x - m_v) rstd_v gamma + beta) + b = gamma * acc + beta * sum_v w_v + b
  const float2 ws2 = f2(wsum), wmb2 = f2(__ldg(a.wm_b));
  float2 s = f2(0.f);
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int e = ch0 + 64 * k;
    p[k] = ok[k] ? __ffma2_rn(acc[k], __ldg(reinterpret_cast<const float2*>(a.vn_w + e)),
                              __ffma2_rn(ws2, __ldg(reinterpret_cast<const float2*>(a.vn_b + e)), wmb2))
                 : f2(0.f);
    s = __fadd2_rn(s, p[k]);
  }
  const float2 nmean = f2(-warp_sum(s.x + s.y) * invE);
  float2 q = f2(0.f);
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    p[k] = ok[k] ? __fadd2_rn(p[k], nmean) : f2(0.f);
    q = __ffma2_rn(p[k], p[k], q);
  }
  const float2 rstd = f2(rsqrtf(warp_sum(q.x + q.y) * invE + 1e-5f));  // head LayerNorm: default eps (multiview_mpl.py:284)
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int e = ch0 + 64 * k;
    if (ok[k])
      p[k] = __ffma2_rn(__fmul2_rn(p[k], rstd), __ldg(reinterpret_cast<const float2*>(a.hn_w + e)),
                        __ldg(reinterpret_cast<const float2*>(a.hn_b + e)));
  }
}

// One CTA (8 warps) per group of 16 poses (= one MMA row tile).
//   Phase 1: each warp pools two poses, one after the other, four views in flight at a time (ray-half strip ->
//            View_norm -> view-weighted sum -> head LayerNorm; registers + shuffles only, packed f32x2 arithmetic) and
//            parks the result in shared memory as pool[pose][channel].
//   Phase 2: the E -> 3J Linear on the tensor cores at fp32-grade accuracy: both operands are split into fp16 hi + lo
//            parts (x = hi + lo to ~2^-22; the weights are scaled by 2^10 first so their lo parts stay normal) and every
//            product is three mma.sync m16n8k16 (hi*hi + hi*lo + lo*hi, fp32 accumulate).  The k tiles are dealt round-
//            robin to the warps; the pre-split weights sit in fragment order (one 16-byte load per lane, k tile and n tile).
//   Phase 3: the per-warp partial tiles are added in a fixed order through shared memory, the bias is added, 3J floats
//            per pose are written.
// The head serves every precision mode, hence the split: a single bf16 / fp16 product would cost ~2e-3 of output scale.
constexpr int HEAD_PB = 16;
constexpr int HEAD_WARPS = 8;
constexpr float HEAD_WSCALE = 1024.0f;

__device__ __forceinline__ void split_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma_f16_acc(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NK, int SS, int ST, bool PL>
__global__ void __launch_bounds__(HEAD_WARPS * 32, 2) head_block_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float hsm[];  // pool [HEAD_PB][pitch]; reused as partial [HEAD_WARPS][HEAD_PB][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int E = a.E;
  const int pitch = E + 8;          // (E + 8) % 32 == 8 for E % 32 == 0: conflict-free 8-byte A-fragment loads
  const int ktiles = E >> 4, ntiles = (a.out_dim + 7) >> 3;
  const int64_t groups = (a.B + HEAD_PB - 1) / HEAD_PB;
  const int S = E >> 5;
  const int ch0 = (lane >> 4) * 32 + (lane & 15) * 2;  // the lane's channel pair of slot 0 (head_pool_pose)
  for (int64_t gi = blockIdx.x; gi < groups; gi += gridDim.x) {
    const int64_t b0 = gi * HEAD_PB;
    // ---- phase 1 ----
#pragma unroll 1
    for (int pi = 0; pi < 2; ++pi) {
      const int p0 = warp * 2 + pi;
      float2 pp[NK];
      head_pool_pose<NK, SS, ST, PL>(a, min(b0 + p0, a.B - 1), lane, pp);
#pragma unroll
      for (int k = 0; k < NK; ++k)
        if (2 * k + (lane >> 4) < S) *reinterpret_cast<float2*>(hsm + p0 * pitch + ch0 + 64 * k) = pp[k];
    }
    __syncthreads();
    // ---- phase 2 ----
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
    for (int kt = warp; kt < ktiles; kt += HEAD_WARPS) {
      const float* r0 = hsm + g * pitch + 16 * kt + 2 * t;
      const float* r1 = r0 + 8 * pitch;
      const float2 x0 = *reinterpret_cast<const float2*>(r0), x1 = *reinterpret_cast<const float2*>(r1);
      const float2 x2 = *reinterpret_cast<const float2*>(r0 + 8), x3 = *reinterpret_cast<const float2*>(r1 + 8);
      uint32_t ah[4], al[4];
      split_f16(x0.x, x0.y, ah[0], al[0]);
      split_f16(x1.x, x1.y, ah[1], al[1]);
      split_f16(x2.x, x2.y, ah[2], al[2]);
      split_f16(x3.x, x3.y, ah[3], al[3]);
      const uint4* wb = reinterpret_cast<const uint4*>(a.hwT) + ((size_t)kt * ntiles) * 32 + lane;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < ntiles) {
          const uint4 w = __ldg(wb + nt * 32);  // (hi b0, hi b1, lo b0, lo b1)
          mma_f16_acc(acc[nt], ah, w.x, w.y);
          mma_f16_acc(acc[nt], ah, w.z, w.w);
          mma_f16_acc(acc[nt], al, w.x, w.y);
        }
      }
    }
    __syncthreads();  // every warp is done reading the pooled values: the buffer becomes the partial-sum exchange
    // ---- phase 3 ----
    float* part = hsm + (size_t)warp * HEAD_PB * 64;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<float2*>(part + g * 64 + 8 * nt + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(part + (g + 8) * 64 + 8 * nt + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < HEAD_PB * 64; idx += HEAD_WARPS * 32) {
      const int p = idx >> 6, o = idx & 63;
      if (o < a.out_dim && b0 + p < a.B) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < HEAD_WARPS; ++w) sum += hsm[(size_t)w * HEAD_PB * 64 + idx];
        a.out[(b0 + p) * a.out_dim + o] = fmaf(sum, 1.0f / HEAD_WSCALE, __ldg(a.hb + o));
      }
    }
    __syncthreads();  // the buffer is reused by the next group
  }
}

// head.1.weight [out_dim, E] -> fp16 hi / lo parts of 2^10 * W in mma B-fragment order: [k tile][n tile][lane] x uint4
// (hi b0, hi b1, lo b0, lo b1), b0 = W[8nt+g][16kt+2t, +1], b1 = W[8nt+g][16kt+2t+8, +9]   (pack time)
__global__ void head_transpose_kernel(const float* __restrict__ W, uint4* __restrict__ WB, int out_dim, int E) {
  const int ntiles = (out_dim + 7) >> 3, ktiles = E >> 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ktiles * ntiles * 32) return;
  const int lane = idx & 31, nt = (idx >> 5) % ntiles, kt = (idx >> 5) / ntiles;
  const int g = lane >> 2, t = lane & 3;
  const int n = 8 * nt + g, k = 16 * kt + 2 * t;
  auto w = [&](int kk) { return n < out_dim ? W[(int64_t)n * E + kk] * HEAD_WSCALE : 0.f; };
  uint4 o;
  split_f16(w(k), w(k + 1), o.x, o.z);
  split_f16(w(k + 8), w(k + 9), o.y, o.w);
  WB[idx] = o;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// returns MPL_OK if it launched, 1 if the shapes / alignment need the generic kernel
int try_launch_embed_vec(const EmbedArgs& a, cudaStream_t s) {
  if (a.d % 4 != 0 || a.d > 1024 || 256 % (a.d / 4) != 0 || !aligned16(a.x)) return 1;
  for (int v = 0; v < a.V; ++v) {
    if (!aligned16(a.be[v]) || !aligned16(a.Ps[v])) return 1;
    if ((a.add_conf || a.mult_conf) && (!aligned16(a.Wc[v]) || !aligned16(a.bc[v]))) return 1;
  }
  if (a.spatial_pos_mode == 1 && (!aligned16(a.pos3d) || a.pos3d_ld % 4 != 0)) return 1;
  if (a.spatial_pos_mode == 2 && !aligned16(a.bl)) return 1;
  const int64_t rows = a.B * a.J;
  if (rows == 0) return MPL_OK;
  const int rows_per_pass = 256 / (a.d / 4);
  const int64_t blocks = ceil_div(rows, rows_per_pass);
  const int64_t per_view = std::max<int64_t>(1, (int64_t)kNumSMs * 16 / a.V);
  dim3 grid((unsigned)std::min<int64_t>(blocks, per_view), (unsigned)a.V);
  embed_vec_kernel<<<grid, 256, 0, s>>>(a);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

int try_launch_token_build_vec(const TokenArgs& a, cudaStream_t s) {
  if (a.d % 4 != 0 || a.tok_w % 4 != 0 || !aligned16(a.tok) || !aligned16(a.xn)) return 1;
  if (a.Wcf != nullptr && (!aligned16(a.Wcf) || !aligned16(a.bcf))) return 1;
  if (a.br != nullptr && !aligned16(a.br)) return 1;
  if (a.pos_table != nullptr && (!aligned16(a.pos_table) || a.pos_w % 4 != 0)) return 1;
  if (a.pos_table == nullptr && !aligned16(a.bl)) return 1;
  const int64_t total = a.B * a.V * (int64_t)(a.tok_w / 4);
  if (total == 0) return MPL_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 32);
  token_build_vec_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

bool head_warp_supports(const HeadArgs& a) {
  if (a.E > 32 * 18 || a.E % 32 != 0 || a.out_dim > 64 || !(a.seg_len == 32 || a.seg_len == a.E)) return false;
  const int stride = a.seg_len == a.E ? 32 : a.seg_stride;
  return a.tok_w % 2 == 0 && stride % 2 == 0;
}

int try_launch_head_warp(const HeadArgs& a, cudaStream_t s) {
  if (!head_warp_supports(a) || a.hwT == nullptr) return 1;
  const int stride = a.seg_len == a.E ? 32 : a.seg_stride;  // contiguous row: segment i starts at column 32 i
  auto aligned8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
  const bool planes = a.tok_hi != nullptr;
  if ((planes ? (!aligned8(a.tok_hi) || !aligned8(a.tok_lo)) : !aligned8(a.tok)) || a.tok_w % 2 != 0 || stride % 2 != 0 ||
      !aligned8(a.vn_w) || !aligned8(a.vn_b) || !aligned8(a.hn_w) || !aligned8(a.hn_b))
    return 1;
  if (a.B == 0) return MPL_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(a.B, HEAD_PB), (int64_t)kNumSMs * 2);
  const size_t smem = std::max((size_t)(a.E + 8) * HEAD_PB, (size_t)HEAD_WARPS * HEAD_PB * 64) * sizeof(float);
  static std::atomic<unsigned char> attr_set[64][8];  // per device and instantiation; setting an attribute twice is harmless
  int dev = 0;
  MPL_CUDA(cudaGetDevice(&dev));
  // the shipped width (17 segments; ray-stripped or contiguous rows) with everything compiled in, else runtime
  // stride / segment count with 9 slots (up to 18 segments) or 5 (up to 10); x2: fp32 rows or two bf16 planes
  const int which = (a.E == 32 * 17 && stride == 64 ? 0 : (a.E == 32 * 17 && stride == 32 ? 1 : (a.E > 32 * 10 ? 2 : 3))) + (planes ? 4 : 0);
  void (*const kerns[8])(const HeadArgs) = {head_block_kernel<9, 64, 17, false>, head_block_kernel<9, 32, 17, false>,
                                            head_block_kernel<9, 0, 0, false>,   head_block_kernel<5, 0, 0, false>,
                                            head_block_kernel<9, 64, 17, true>,  head_block_kernel<9, 32, 17, true>,
                                            head_block_kernel<9, 0, 0, true>,    head_block_kernel<5, 0, 0, true>};
  auto kern = kerns[which];
  if (dev < 0 || dev >= 64 || !attr_set[dev][which].load(std::memory_order_acquire)) {
    MPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev][which].store(1, std::memory_order_release);
  }
  HeadArgs b = a;
  b.seg_stride = stride;
  kern<<<(unsigned)blocks, HEAD_WARPS * 32, smem, s>>>(b);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

int launch_head_transpose(const float* W, float* WT, int out_dim, int E, cudaStream_t s) {
  const int64_t n = (int64_t)(E >> 4) * ((out_dim + 7) >> 3) * 32;
  head_transpose_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(W, reinterpret_cast<uint4*>(WT), out_dim, E);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

}  // namespace mpl
