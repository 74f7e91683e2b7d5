// The HBM-bound edge kernels of the forward for the shipped widths (d % 4 == 0): K1 joint embedding, FPT token build and
// the K5 fused head, written so that every global access is a 16-byte (or a coalesced 128-byte-per-warp) transaction
// and the per-element index arithmetic of the generic kernels (kernels_generic.cu) disappears.  The generic kernels stay
// as the path for odd widths / unaligned pointers; launch_* in kernels_generic.cu dispatch here first.
#include <algorithm>

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
// K5 fused head (multiview_mpl.py:425-446, 517-523): one WARP per PAIR of poses.  Channel e of the stripped row lives in
// lane e % 32, register e / 32 (segments are multiples of 32 wide for the shipped d = 32, so every load is a coalesced
// 128-byte line); View_norm and the head LayerNorm reduce with shuffles only — no shared memory, no block barriers.
// The E -> 3J Linear reads each weight row once per pose pair (L1-resident, 3J*E*4 = 111 KB) and reduces across lanes.
// ---------------------------------------------------------------------------------------------------------------------
// View_norm -> view-weighted sum -> head LayerNorm for TWO poses at once (independent load / reduction chains interleave)
template <int NV, bool FULL>  // FULL: E == 32 * NV, no channel predicates
__device__ __forceinline__ void head_pool_two(const HeadArgs& a, int64_t ba, int64_t bb, int lane, const int (&colv)[NV],
                                              float (&pa)[NV], float (&pb)[NV]) {
  const int E = a.E;
  const float invE = 1.0f / (float)E;
  const float wmb = __ldg(a.wm_b);
#pragma unroll
  for (int i = 0; i < NV; ++i) { pa[i] = wmb; pb[i] = wmb; }
  for (int v = 0; v < a.V; ++v) {
    const float* ra = a.tok + (ba * a.V + v) * (int64_t)a.tok_w + lane;
    const float* rb = a.tok + (bb * a.V + v) * (int64_t)a.tok_w + lane;
    float xa[NV], xb[NV];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const bool ok = FULL || 32 * i + lane < E;
      xa[i] = ok ? ra[colv[i]] : 0.f;
      xb[i] = ok ? rb[colv[i]] : 0.f;
      sa += xa[i];
      sb += xb[i];
    }
    const float ma = warp_sum(sa) * invE, mb = warp_sum(sb) * invE;
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const bool ok = FULL || 32 * i + lane < E;
      const float ta = ok ? xa[i] - ma : 0.f, tb = ok ? xb[i] - mb : 0.f;
      qa = fmaf(ta, ta, qa);
      qb = fmaf(tb, tb, qb);
    }
    const float rsa = rsqrtf(warp_sum(qa) * invE + 1e-6f), rsb = rsqrtf(warp_sum(qb) * invE + 1e-6f);
    const float wv = __ldg(a.wm_w + v);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = 32 * i + lane;
      if (FULL || e < E) {
        const float g = __ldg(a.vn_w + e), bt = __ldg(a.vn_b + e);
        pa[i] = fmaf(wv, fmaf((xa[i] - ma) * rsa, g, bt), pa[i]);
        pb[i] = fmaf(wv, fmaf((xb[i] - mb) * rsb, g, bt), pb[i]);
      }
    }
  }
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { const bool ok = FULL || 32 * i + lane < E; sa += ok ? pa[i] : 0.f; sb += ok ? pb[i] : 0.f; }
  const float ma = warp_sum(sa) * invE, mb = warp_sum(sb) * invE;
  float qa = 0.f, qb = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const bool ok = FULL || 32 * i + lane < E;
    const float ta = ok ? pa[i] - ma : 0.f, tb = ok ? pb[i] - mb : 0.f;
    qa = fmaf(ta, ta, qa);
    qb = fmaf(tb, tb, qb);
  }
  const float rsa = rsqrtf(warp_sum(qa) * invE + 1e-5f), rsb = rsqrtf(warp_sum(qb) * invE + 1e-5f);  // head LN: eps 1e-5 (multiview_mpl.py:284)
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = 32 * i + lane;
    const float g = (FULL || e < E) ? __ldg(a.hn_w + e) : 0.f, bt = (FULL || e < E) ? __ldg(a.hn_b + e) : 0.f;
    pa[i] = fmaf((pa[i] - ma) * rsa, g, bt);
    pb[i] = fmaf((pb[i] - mb) * rsb, g, bt);
  }
}

// One CTA (8 warps) per group of HEAD_PB poses.
//   Phase 1: each warp pools its 4 poses (View_norm -> view-weighted sum -> head LayerNorm, registers + shuffles only,
//            lane = channel % 32) and parks the result in shared memory as pool[channel][pose].
//   Phase 2: the E channels are split over the warps; lane l owns outputs l and l + 32 of ALL 32 poses, so every element
//            of the transposed, 64-padded head weight is read once per 32 poses (coalesced) and feeds 64 FMAs; the pooled
//            values of a channel arrive as eight broadcast 16-byte shared-memory loads.
//   Phase 3: the 8 per-warp partial sums of every (pose, output) are added in a fixed order through shared memory.
constexpr int HEAD_PB = 16;       // poses per CTA iteration (16: pooled values + exchange buffer stay at 35 KB per CTA, so the
                                  // 139 KB transposed head weight remains L1-resident next to two CTAs per SM)
constexpr int HEAD_WARPS = 8;

template <int NV, bool FULL>
__global__ void __launch_bounds__(HEAD_WARPS * 32, 2) head_block_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float hsm[];  // pool [E][HEAD_PB]; reused as partial [HEAD_WARPS][HEAD_PB][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int E = a.E;
  const int64_t groups = (a.B + HEAD_PB - 1) / HEAD_PB;
  const int e_per_warp = (E + HEAD_WARPS - 1) / HEAD_WARPS;
  // column of channel group i in the token row: 32-channel segments seg_stride apart (ray halves stripped, d = 32) or
  // one contiguous row (seg_len == E, seg_stride == 32 passed by the launcher): one multiply, no division
  int colv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) colv[i] = i * a.seg_stride;
  for (int64_t gi = blockIdx.x; gi < groups; gi += gridDim.x) {
    const int64_t b0 = gi * HEAD_PB;
    // ---- phase 1 ----
#pragma unroll 1
    for (int pi = 0; pi < HEAD_PB / HEAD_WARPS; pi += 2) {
      const int p0 = warp * (HEAD_PB / HEAD_WARPS) + pi;
      float pa[NV], pb[NV];
      head_pool_two<NV, FULL>(a, min(b0 + p0, a.B - 1), min(b0 + p0 + 1, a.B - 1), lane, colv, pa, pb);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (FULL || 32 * i + lane < E) *reinterpret_cast<float2*>(hsm + (32 * i + lane) * HEAD_PB + p0) = make_float2(pa[i], pb[i]);
    }
    __syncthreads();
    // ---- phase 2 ----
    float acc0[HEAD_PB], acc1[HEAD_PB];
#pragma unroll
    for (int p = 0; p < HEAD_PB; ++p) { acc0[p] = 0.f; acc1[p] = 0.f; }
    const int e_lo = warp * e_per_warp, e_hi = min(E, e_lo + e_per_warp);
    const float* wt = a.hwT + lane;
    // weights of 4 channels are fetched (L2 / L1) one 4-channel step ahead of their use
    float wn[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = min(e_lo + u, E - 1);
      wn[u][0] = __ldg(wt + e * 64); wn[u][1] = __ldg(wt + e * 64 + 32);
    }
#pragma unroll 1
    for (int e4 = e_lo; e4 < e_hi; e4 += 4) {
      float wc[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) { wc[u][0] = wn[u][0]; wc[u][1] = wn[u][1]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = min(e4 + 4 + u, E - 1);
        wn[u][0] = __ldg(wt + e * 64); wn[u][1] = __ldg(wt + e * 64 + 32);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (e4 + u >= e_hi) break;
        const float w0 = wc[u][0], w1 = wc[u][1];
        const float4* pv = reinterpret_cast<const float4*>(hsm + (e4 + u) * HEAD_PB);
#pragma unroll
        for (int p4 = 0; p4 < HEAD_PB / 4; ++p4) {
          const float4 v = pv[p4];
          acc0[4 * p4] = fmaf(v.x, w0, acc0[4 * p4]);         acc1[4 * p4] = fmaf(v.x, w1, acc1[4 * p4]);
          acc0[4 * p4 + 1] = fmaf(v.y, w0, acc0[4 * p4 + 1]); acc1[4 * p4 + 1] = fmaf(v.y, w1, acc1[4 * p4 + 1]);
          acc0[4 * p4 + 2] = fmaf(v.z, w0, acc0[4 * p4 + 2]); acc1[4 * p4 + 2] = fmaf(v.z, w1, acc1[4 * p4 + 2]);
          acc0[4 * p4 + 3] = fmaf(v.w, w0, acc0[4 * p4 + 3]); acc1[4 * p4 + 3] = fmaf(v.w, w1, acc1[4 * p4 + 3]);
        }
      }
    }
    __syncthreads();  // every warp is done reading the pooled values: the buffer becomes the partial-sum exchange
    // ---- phase 3 ----
    float* part = hsm + (size_t)warp * HEAD_PB * 64;
#pragma unroll
    for (int p = 0; p < HEAD_PB; ++p) {
      part[p * 64 + lane] = acc0[p];
      part[p * 64 + 32 + lane] = acc1[p];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < HEAD_PB * 64; idx += HEAD_WARPS * 32) {
      const int p = idx >> 6, o = idx & 63;
      if (o < a.out_dim && b0 + p < a.B) {
        float sum = __ldg(a.hb + o);
#pragma unroll
        for (int w = 0; w < HEAD_WARPS; ++w) sum += hsm[(size_t)w * HEAD_PB * 64 + idx];
        a.out[(b0 + p) * a.out_dim + o] = sum;
      }
    }
    __syncthreads();  // the buffer is reused by the next group
  }
}

// head.1.weight [out_dim, E] -> transposed and padded [E, 64] (pack time)
__global__ void head_transpose_kernel(const float* __restrict__ W, float* __restrict__ WT, int out_dim, int E) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E * 64) return;
  const int e = idx >> 6, o = idx & 63;
  WT[idx] = (o < out_dim) ? W[(int64_t)o * E + e] : 0.f;
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

int try_launch_head_warp(const HeadArgs& a, cudaStream_t s) {
  if (a.E > 32 * 17 || a.out_dim > 64 || a.hwT == nullptr || !(a.seg_len == 32 || a.seg_len == a.E)) return 1;
  if (a.B == 0) return MPL_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(a.B, HEAD_PB), (int64_t)kNumSMs * 2);
  const size_t smem = std::max((size_t)a.E * HEAD_PB, (size_t)HEAD_WARPS * HEAD_PB * 64) * sizeof(float);
  static bool attr_set[64][3] = {};
  int dev = 0;
  MPL_CUDA(cudaGetDevice(&dev));
  const int which = (a.E == 32 * 17) ? 2 : (a.E > 32 * 9 ? 1 : 0);
  auto kern = which == 2 ? head_block_kernel<17, true> : (which == 1 ? head_block_kernel<17, false> : head_block_kernel<9, false>);
  if (dev < 0 || dev >= 64 || !attr_set[dev][which]) {
    MPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev][which] = true;
  }
  HeadArgs b = a;
  if (a.seg_len == a.E) b.seg_stride = 32;  // contiguous row: channel group i starts at column 32 i
  kern<<<(unsigned)blocks, HEAD_WARPS * 32, smem, s>>>(b);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

int launch_head_transpose(const float* W, float* WT, int out_dim, int E, cudaStream_t s) {
  head_transpose_kernel<<<(unsigned)ceil_div((int64_t)E * 64, 256), 256, 0, s>>>(W, WT, out_dim, E);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

}  // namespace mpl
