// K2: the Spatial Pose Transformer as ONE kernel (multiview_mpl.py:400-412 with Block/Attention/Mlp :21-92).
//
// The SPT is latency work, not FLOPs: 17 tokens x 32 channels per (pose, view) set, head_dim 4, depth+1 block
// applications.  Launched layer by layer it is ~14 small kernels per application and re-reads the residual stream
// from HBM every time; here a CTA keeps 16 sets resident for the whole stack:
//   * 16 sets x 17 tokens = 272 rows = 17 MMA row tiles of 16 -> 17 warps, zero padding (16*17 == 17*16);
//   * the residual stream lives in registers as mma.sync C fragments (16 fp32 per lane) across all applications;
//   * LayerNorm is computed on the fragments (quad shuffles), the four Linears are bf16 mma.sync m16n8k16 with fp32
//     accumulation, their weights pre-packed in fragment order and double-buffered in shared memory via cp.async;
//   * q|k|v of the 16 sets are staged once in shared memory (fp32), the 17x17 softmax per head is done by one thread
//     per (row, half of the heads) in registers — the sets are far too short for flash-style tiling;
//   * erf-form GELU (A&S 7.1.25, |erf error| < 2.5e-5, result rounded to bf16 for the next mma), residual adds and
//     the final Spatial_norm stay in fp32 registers.
// HBM traffic per set: 17*32*4 B in + out, nothing in between.  (tcgen05 needs 128-row operand tiles staged through
// shared memory for every one of the 52 tiny GEMMs per set; the warp-level mma keeps the operands in registers.)
#include "kernels.cuh"

namespace mpl {

namespace {

constexpr int J = 17, D = 32, HID = 64, HEADS = 8;
constexpr int SETS = 16, ROWS = SETS * J, WARPS = 17, THREADS = WARPS * 32;
constexpr int QS = 100;  // fp32 row pitch of the q|k|v staging buffer (96 + 4: spreads rows over banks, keeps 16B alignment)
constexpr int AS = 40;   // bf16 row pitch of the attention-output buffer (32 + 8: conflict-free A-fragment loads)

// fragment-packed layer blob (32-bit words): B fragments of the four weight matrices, then the fp32 vectors
constexpr int OFF_QKV = 0, OFF_PROJ = 1536, OFF_FC1 = 2048, OFF_FC2 = 3072, FRAG_WORDS = 4096;
constexpr int F_N1W = 0, F_N1B = 32, F_QKVB = 64, F_PROJB = 160, F_N2W = 192, F_N2B = 224, F_FC1B = 256, F_FC2B = 320;
constexpr int VEC_WORDS = 352, LAYER_WORDS = FRAG_WORDS + VEC_WORDS;  // 4448 words = 17792 bytes

constexpr int SMEM_W = 2 * LAYER_WORDS * 4;
constexpr int SMEM_QKV = ROWS * QS * 4;
constexpr int SMEM_AO = ROWS * AS * 2;
constexpr int SMEM_TOTAL = SMEM_W + SMEM_QKV + SMEM_AO;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// LayerNorm (eps 1e-6, biased variance) of the two rows a lane co-owns, straight to bf16 A fragments (2 k-tiles of 16)
__device__ __forceinline__ void ln_to_afrag(const float (&x)[4][4], const float* __restrict__ gw, const float* __restrict__ gb,
                                            int t, uint32_t (&a)[2][4]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { s0 += x[nt][0] + x[nt][1]; s1 += x[nt][2] + x[nt][3]; }
  const float m0 = quad_sum(s0) * (1.0f / D), m1 = quad_sum(s1) * (1.0f / D);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float d0 = x[nt][0] - m0, d1 = x[nt][1] - m0, d2 = x[nt][2] - m1, d3 = x[nt][3] - m1;
    q0 += d0 * d0 + d1 * d1;
    q1 += d2 * d2 + d3 * d3;
  }
  const float r0 = rsqrtf(quad_sum(q0) * (1.0f / D) + 1e-6f), r1 = rsqrtf(quad_sum(q1) * (1.0f / D) + 1e-6f);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 w = *reinterpret_cast<const float2*>(gw + 8 * nt + 2 * t);
    const float2 b = *reinterpret_cast<const float2*>(gb + 8 * nt + 2 * t);
    const float y0 = (x[nt][0] - m0) * r0 * w.x + b.x, y1 = (x[nt][1] - m0) * r0 * w.y + b.y;
    const float y2 = (x[nt][2] - m1) * r1 * w.x + b.x, y3 = (x[nt][3] - m1) * r1 * w.y + b.y;
    a[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(y0, y1);
    a[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(y2, y3);
  }
}

// one output tile (16 rows x 8 features): acc = bias, then KT k-steps with B fragments read from the packed blob
template <int KT>
__device__ __forceinline__ void gemm_tile(float (&c)[4], const uint32_t (&a)[KT][4], const uint32_t* __restrict__ wfrag, int nt,
                                          const float* __restrict__ bias, int lane, int t) {
  const float2 b2 = *reinterpret_cast<const float2*>(bias + 8 * nt + 2 * t);
  c[0] = b2.x; c[1] = b2.y; c[2] = b2.x; c[3] = b2.y;
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    const uint2 b = *reinterpret_cast<const uint2*>(wfrag + ((nt * KT + kt) * 32 + lane) * 2);
    mma_bf16_16816(c, a[kt], b.x, b.y);
  }
}

struct SptArgs {
  const float* x_in;                  // [V, B, J, 32] joint embeddings
  float* x_out;                       // [V, B, J, 32] after the stack and Spatial_norm
  const uint32_t* wpack[kMaxViews];   // per view stack: [depth][LAYER_WORDS]
  const float* sn_w;
  const float* sn_b;
  const float* conf;                  // [V, B, J] or null: confidence_as_attention_uncertainty_weight
  int64_t B;
  int depth;
  float scale_log2e;                  // head_dim^-0.5 (or qk_scale) * log2(e)
};

__device__ __forceinline__ void load_layer_async(uint32_t* dst, const uint32_t* src) {
  constexpr int CHUNKS = LAYER_WORDS * 4 / 16;  // 1112 x 16 bytes
  for (int i = threadIdx.x; i < CHUNKS; i += THREADS) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + i * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i * 4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void wait_async_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(THREADS, 1) spt_fused_kernel(const SptArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* wbuf = reinterpret_cast<uint32_t*>(smem);
  float* qkv_s = reinterpret_cast<float*>(smem + SMEM_W);
  __nv_bfloat16* ao_s = reinterpret_cast<__nv_bfloat16*>(smem + SMEM_W + SMEM_QKV);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int view = blockIdx.y;
  const int64_t tile = blockIdx.x;
  const int64_t rows_in_view = args.B * J;
  const int64_t view_row0 = (int64_t)view * rows_in_view;
  const int r0 = warp * 16 + g, r1 = r0 + 8;                    // CTA-local rows of this lane's fragments
  const int64_t gr0 = tile * ROWS + r0, gr1 = tile * ROWS + r1;  // rows inside the view
  const bool ok0 = gr0 < rows_in_view, ok1 = gr1 < rows_in_view;
  const uint32_t* wsrc = args.wpack[view];

  load_layer_async(wbuf, wsrc);

  // residual stream: C-fragment layout, x[nt][0..1] = row r0 cols 8nt+2t,+1 ; x[nt][2..3] = row r1
  float x[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
    if (ok0) v0 = *reinterpret_cast<const float2*>(args.x_in + (view_row0 + gr0) * D + 8 * nt + 2 * t);
    if (ok1) v1 = *reinterpret_cast<const float2*>(args.x_in + (view_row0 + gr1) * D + 8 * nt + 2 * t);
    x[nt][0] = v0.x; x[nt][1] = v0.y; x[nt][2] = v1.x; x[nt][3] = v1.y;
  }

  // attention role of this thread: one row, four heads
  const int arow = threadIdx.x >> 1, ahh = threadIdx.x & 1;
  const int aset0 = (arow / J) * J;  // first row of the row's set
  const int64_t agr = tile * ROWS + arow;
  float aconf = 1.0f;
  if (args.conf != nullptr && agr < rows_in_view) aconf = __ldg(args.conf + view_row0 + agr);

  for (int layer = 0; layer < args.depth; ++layer) {
    wait_async_all();
    __syncthreads();  // this layer's weights have landed; every warp is done with the other buffer
    const uint32_t* w = wbuf + (layer & 1) * LAYER_WORDS;
    const float* wv = reinterpret_cast<const float*>(w + FRAG_WORDS);
    if (layer + 1 < args.depth) load_layer_async(wbuf + ((layer + 1) & 1) * LAYER_WORDS, wsrc + (size_t)(layer + 1) * LAYER_WORDS);

    // block applications of this layer (multiview_mpl.py:405-410): [confidence-weighted], [last layer: once more], plain
    const int n_apps = (args.conf != nullptr ? 1 : 0) + (layer == args.depth - 1 ? 2 : 1);
    for (int app = 0; app < n_apps; ++app) {
      const bool weighted = (args.conf != nullptr) && app == 0;
      // ---- LN1 + QKV ----
      {
        uint32_t a[2][4];
        ln_to_afrag(x, wv + F_N1W, wv + F_N1B, t, a);
#pragma unroll
        for (int nt = 0; nt < 12; ++nt) {
          float c[4];
          gemm_tile<2>(c, a, w + OFF_QKV, nt, wv + F_QKVB, lane, t);
          *reinterpret_cast<float2*>(qkv_s + r0 * QS + 8 * nt + 2 * t) = make_float2(c[0], c[1]);
          *reinterpret_cast<float2*>(qkv_s + r1 * QS + 8 * nt + 2 * t) = make_float2(c[2], c[3]);
        }
      }
      __syncthreads();
      // ---- attention: softmax(q k^T * scale) v over the 17 tokens of the row's set, 4 heads per thread ----
      {
        const float rowscale = weighted ? aconf : 1.0f;
        uint32_t packed[8];
#pragma unroll
        for (int hl = 0; hl < 4; ++hl) {
          const int h = ahh * 4 + hl;
          float4 q = *reinterpret_cast<const float4*>(qkv_s + arow * QS + 4 * h);
          q.x *= args.scale_log2e; q.y *= args.scale_log2e; q.z *= args.scale_log2e; q.w *= args.scale_log2e;
          float sc[J];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const float4 k = *reinterpret_cast<const float4*>(qkv_s + (aset0 + j) * QS + D + 4 * h);
            sc[j] = q.x * k.x + q.y * k.y + q.z * k.z + q.w * k.w;
            mx = fmaxf(mx, sc[j]);
          }
          float sum = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const float p = exp2f(sc[j] - mx);
            const float4 v = *reinterpret_cast<const float4*>(qkv_s + (aset0 + j) * QS + 2 * D + 4 * h);
            sum += p;
            o0 = fmaf(p, v.x, o0); o1 = fmaf(p, v.y, o1); o2 = fmaf(p, v.z, o2); o3 = fmaf(p, v.w, o3);
          }
          const float inv = rowscale / sum;
          packed[2 * hl] = pack_bf16(o0 * inv, o1 * inv);
          packed[2 * hl + 1] = pack_bf16(o2 * inv, o3 * inv);
        }
        uint4* dst = reinterpret_cast<uint4*>(ao_s + arow * AS + 16 * ahh);
        dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      }
      __syncthreads();
      // ---- proj + residual ----
      {
        uint32_t a[2][4];
        const uint32_t* ao32 = reinterpret_cast<const uint32_t*>(ao_s);
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          a[kt][0] = ao32[r0 * (AS / 2) + 8 * kt + t];
          a[kt][1] = ao32[r1 * (AS / 2) + 8 * kt + t];
          a[kt][2] = ao32[r0 * (AS / 2) + 8 * kt + 4 + t];
          a[kt][3] = ao32[r1 * (AS / 2) + 8 * kt + 4 + t];
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float c[4];
          gemm_tile<2>(c, a, w + OFF_PROJ, nt, wv + F_PROJB, lane, t);
#pragma unroll
          for (int i = 0; i < 4; ++i) x[nt][i] += c[i];
        }
      }
      // ---- LN2 + fc1 + GELU + fc2 + residual ----
      {
        uint32_t a[2][4];
        ln_to_afrag(x, wv + F_N2W, wv + F_N2B, t, a);
        uint32_t a2[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          float c[4];
          gemm_tile<2>(c, a, w + OFF_FC1, nt, wv + F_FC1B, lane, t);
          a2[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(gelu_erf_fast(c[0]), gelu_erf_fast(c[1]));
          a2[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(gelu_erf_fast(c[2]), gelu_erf_fast(c[3]));
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float c[4];
          gemm_tile<4>(c, a2, w + OFF_FC2, nt, wv + F_FC2B, lane, t);
#pragma unroll
          for (int i = 0; i < 4; ++i) x[nt][i] += c[i];
        }
      }
    }
  }

  // ---- Spatial_norm (multiview_mpl.py:412), fp32 out ----
  {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { s0 += x[nt][0] + x[nt][1]; s1 += x[nt][2] + x[nt][3]; }
    const float m0 = quad_sum(s0) * (1.0f / D), m1 = quad_sum(s1) * (1.0f / D);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float d0 = x[nt][0] - m0, d1 = x[nt][1] - m0, d2 = x[nt][2] - m1, d3 = x[nt][3] - m1;
      q0 += d0 * d0 + d1 * d1;
      q1 += d2 * d2 + d3 * d3;
    }
    const float rs0 = rsqrtf(quad_sum(q0) * (1.0f / D) + 1e-6f), rs1 = rsqrtf(quad_sum(q1) * (1.0f / D) + 1e-6f);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float2 wg = __ldg(reinterpret_cast<const float2*>(args.sn_w + 8 * nt + 2 * t));
      const float2 bg = __ldg(reinterpret_cast<const float2*>(args.sn_b + 8 * nt + 2 * t));
      if (ok0)
        *reinterpret_cast<float2*>(args.x_out + (view_row0 + gr0) * D + 8 * nt + 2 * t) =
            make_float2((x[nt][0] - m0) * rs0 * wg.x + bg.x, (x[nt][1] - m0) * rs0 * wg.y + bg.y);
      if (ok1)
        *reinterpret_cast<float2*>(args.x_out + (view_row0 + gr1) * D + 8 * nt + 2 * t) =
            make_float2((x[nt][2] - m1) * rs1 * wg.x + bg.x, (x[nt][3] - m1) * rs1 * wg.y + bg.y);
    }
  }
}

// fp32 parameters of one Block -> the fragment-packed layer blob
struct SptPackArgs {
  const float *n1w, *n1b, *qkvw, *qkvb, *projw, *projb, *n2w, *n2b, *fc1w, *fc1b, *fc2w, *fc2b;
  uint32_t* dst;
};

__global__ void spt_pack_kernel(const SptPackArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= LAYER_WORDS) return;
  if (i < FRAG_WORDS) {
    const float* W;
    int K, KT, base;
    if (i < OFF_PROJ) { W = a.qkvw; K = 32; KT = 2; base = OFF_QKV; }
    else if (i < OFF_FC1) { W = a.projw; K = 32; KT = 2; base = OFF_PROJ; }
    else if (i < OFF_FC2) { W = a.fc1w; K = 32; KT = 2; base = OFF_FC1; }
    else { W = a.fc2w; K = 64; KT = 4; base = OFF_FC2; }
    const int j = i - base;
    const int reg = j & 1, lane = (j >> 1) & 31, tl = j >> 6;
    const int nt = tl / KT, kt = tl % KT;
    const int g = lane >> 2, t = lane & 3;
    const int n = 8 * nt + g, k = 16 * kt + 2 * t + 8 * reg;
    a.dst[i] = pack_bf16(W[n * K + k], W[n * K + k + 1]);
  } else {
    const int f = i - FRAG_WORDS;
    float v;
    if (f < F_N1B) v = a.n1w[f - F_N1W];
    else if (f < F_QKVB) v = a.n1b[f - F_N1B];
    else if (f < F_PROJB) v = a.qkvb ? a.qkvb[f - F_QKVB] : 0.f;
    else if (f < F_N2W) v = a.projb[f - F_PROJB];
    else if (f < F_N2B) v = a.n2w[f - F_N2W];
    else if (f < F_FC1B) v = a.n2b[f - F_N2B];
    else if (f < F_FC2B) v = a.fc1b[f - F_FC1B];
    else v = a.fc2b[f - F_FC2B];
    a.dst[i] = __float_as_uint(v);
  }
}

}  // namespace

bool spt_fused_supports(int J_, int d, int H, int hidden) { return J_ == J && d == D && H == HEADS && hidden == HID; }
size_t spt_fused_layer_bytes() { return (size_t)LAYER_WORDS * 4; }

int launch_spt_pack_layer(const float* n1w, const float* n1b, const float* qkvw, const float* qkvb, const float* projw,
                          const float* projb, const float* n2w, const float* n2b, const float* fc1w, const float* fc1b,
                          const float* fc2w, const float* fc2b, void* dst, cudaStream_t s) {
  SptPackArgs a{n1w, n1b, qkvw, qkvb, projw, projb, n2w, n2b, fc1w, fc1b, fc2w, fc2b, reinterpret_cast<uint32_t*>(dst)};
  spt_pack_kernel<<<(LAYER_WORDS + 255) / 256, 256, 0, s>>>(a);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

int launch_spt_fused(const float* x_in, float* x_out, const void* const* wpack_per_view, int V, int64_t B, int depth,
                     float scale, const float* sn_w, const float* sn_b, const float* conf, cudaStream_t s) {
  if (B == 0 || V == 0) return MPL_OK;
  SptArgs a{};
  a.x_in = x_in;
  a.x_out = x_out;
  for (int v = 0; v < V; ++v) a.wpack[v] = reinterpret_cast<const uint32_t*>(wpack_per_view[v]);
  a.sn_w = sn_w;
  a.sn_b = sn_b;
  a.conf = conf;
  a.B = B;
  a.depth = depth;
  a.scale_log2e = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    MPL_CUDA(cudaFuncSetAttribute(spt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(B, SETS), (unsigned)V);
  spt_fused_kernel<<<grid, THREADS, SMEM_TOTAL, s>>>(a);
  MPL_LAUNCH_CHECK();
  return MPL_OK;
}

}  // namespace mpl
