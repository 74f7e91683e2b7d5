// K3: the FPT projection GEMMs (QKV, proj, fc1, fc2 of multiview_mpl.py:26-36,48-66 batched over B*V rows) on the
// 5th-generation tensor cores: Y[M,N] = epilogue(A[M,K] . W[N,K]^T + bias).
//
// Persistent, warp-specialised, hand-written tcgen05 kernel:
//   warp 0     TMA producer   cp.async.bulk.tensor 2D tiles (128B swizzle) into a ring of shared-memory stages
//   warp 1     MMA issuer     one thread issues tcgen05.mma (kind::f16 bf16 or kind::tf32), accumulators in TMEM
//   warp 2     TMEM allocator 512 columns = two 256-column accumulator stages (MMA of tile i+1 overlaps epilogue of i)
//   warps 4-11 epilogue       tcgen05.ld -> registers -> bias / GELU / fp32 residual add -> global (two warps per
//                             TMEM lane quarter, two 32-column chunks in flight per warp)
// CG = 1: one CTA per 128 x 256 output tile.  CG = 2: a CTA pair (cluster 2x1x1, cta_group::2) per 256 x 256 tile, each
// CTA staging half of A and half of W, which halves the shared-memory and L2 operand traffic per MMA.
// Every FPT width is a multiple of 17 (D = 1088 = 17*64): ragged N tiles use a narrower UMMA N (multiple of 16) and
// the K tail is zero-filled by TMA, so no padding copies are needed.
#include <cuda.h>

#include <mutex>

#include "kernels.cuh"
#include "ptx.cuh"

namespace mpl {

namespace {

constexpr int BM = 128;          // accumulator rows per CTA (TMEM lanes)
constexpr int BN = 256;          // accumulator columns per tile
constexpr int KB_BYTES = 128;    // bytes of K per stage row = one 128B swizzle span
constexpr int UMMA_K_BYTES = 32; // bytes of K per tcgen05.mma
constexpr int NUM_EPI_WARPS = 8;   // two per TMEM lane quarter: even / odd 32-column chunks
constexpr int NUM_THREADS = (4 + NUM_EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;

template <int CG>
struct Cfg {
  static constexpr int A_BYTES = BM * KB_BYTES;             // 16 KB
  static constexpr int B_ROWS = BN / CG;                    // W rows staged per CTA
  static constexpr int B_BYTES = B_ROWS * KB_BYTES;         // 32 KB / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;     // 48 KB / 32 KB
  static constexpr int STAGES = (CG == 1) ? 4 : 6;          // 192 KB of operand ring either way
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES;
};

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

// kind::f16 / kind::tf32 instruction descriptor: fp32 accumulate, K-major A and B
__device__ __forceinline__ uint32_t make_idesc(int kind, int m, int n) {
  const uint32_t fmt = (kind == 0) ? 1u : 2u;  // BF16 : TF32
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float round_tf32_dev(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// EPI: 0 bias -> operand dtype, 1 bias + GELU -> operand dtype, 2 Y(fp32) += acc + bias, 3 bias -> fp32
template <int KIND, int EPI, int NV>
__device__ __forceinline__ void epilogue_store(const uint32_t (&v)[NV], const float* __restrict__ bias, void* Y, int64_t row,
                                               int N, int n, bool row_ok) {
  float f[NV];
#pragma unroll
  for (int i = 0; i < NV; i += 4) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n + i));
    f[i] = __uint_as_float(v[i]) + b4.x;
    f[i + 1] = __uint_as_float(v[i + 1]) + b4.y;
    f[i + 2] = __uint_as_float(v[i + 2]) + b4.z;
    f[i + 3] = __uint_as_float(v[i + 3]) + b4.w;
  }
  if constexpr (EPI == 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) f[i] = (KIND == 0) ? gelu_erf_fast(f[i]) : gelu_erf(f[i]);
  }
  if (!row_ok) return;
  if constexpr (EPI == 2) {
    float4* y = reinterpret_cast<float4*>(reinterpret_cast<float*>(Y) + row * N + n);
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      float4 r = y[i];
      r.x += f[4 * i]; r.y += f[4 * i + 1]; r.z += f[4 * i + 2]; r.w += f[4 * i + 3];
      y[i] = r;
    }
  } else if constexpr (EPI == 3 || KIND == 1) {
    float4* y = reinterpret_cast<float4*>(reinterpret_cast<float*>(Y) + row * N + n);
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      float4 r = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
      if constexpr (KIND == 1 && EPI == 1) {  // feeds the next kind::tf32 GEMM: round-to-nearest once, here
        r.x = round_tf32_dev(r.x); r.y = round_tf32_dev(r.y); r.z = round_tf32_dev(r.z); r.w = round_tf32_dev(r.w);
      }
      y[i] = r;
    }
  } else {
    uint4* y = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(Y) + row * N + n);
#pragma unroll
    for (int i = 0; i < NV / 8; ++i) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(f[8 * i], f[8 * i + 1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(f[8 * i + 2], f[8 * i + 3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(f[8 * i + 4], f[8 * i + 5]);
      __nv_bfloat162 p3 = __floats2bfloat162_rn(f[8 * i + 6], f[8 * i + 7]);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      pk.z = *reinterpret_cast<uint32_t*>(&p2);
      pk.w = *reinterpret_cast<uint32_t*>(&p3);
      y[i] = pk;
    }
  }
}

template <int CG, int KIND, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const float* __restrict__ bias, void* Y, int64_t M, int N, int K) {
  using C = Cfg<CG>;
  constexpr int ESZ = (KIND == 0) ? 2 : 4;
  constexpr int BK = KB_BYTES / ESZ;  // K elements per stage
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte aligned bases
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);   // the leader's arrive.expect_tx covers the bytes of both CTAs of a pair
      ptx::mbar_init(empty_bar(s), 1);  // one tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);        // one tcgen05.commit
      ptx::mbar_init(tempty_bar(s), NUM_EPI_WARPS * CG);  // one arrive per epilogue warp of every CTA of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<CG>(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_tiles = (N + BN - 1) / BN;
  const int64_t m_tiles = (M + (int64_t)BM * CG - 1) / ((int64_t)BM * CG);
  const int64_t total_tiles = m_tiles * n_tiles;
  const int64_t first_tile = blockIdx.x / CG;
  const int64_t tile_stride = gridDim.x / CG;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t leader_full0 = (CG == 2) ? ptx::mapa(full_bar(0), 0) : 0u;
      for (int64_t tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int64_t m_blk = tile / n_tiles;
        const int n_blk = (int)(tile % n_tiles);
        const int n_size = min(BN, N - n_blk * BN);
        const int32_t m0 = (int32_t)(m_blk * BM * CG + cta_rank * BM);
        const int32_t n0 = n_blk * BN + (int32_t)cta_rank * (n_size / CG);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + C::A_BYTES;
          if constexpr (CG == 1) {
            ptx::mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
            ptx::tma_load_2d(a_dst, &tmA, full_bar(stage), kb * BK, m0);
            ptx::tma_load_2d(b_dst, &tmB, full_bar(stage), kb * BK, n0);
          } else {
            const uint32_t lbar = leader_full0 + 8u * stage;
            // the peer's bytes may land before this expect_tx: the phase still cannot complete without this arrive
            if (is_leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            ptx::tma_load_2d_pair(a_dst, &tmA, lbar, kb * BK, m0);
            ptx::tma_load_2d_pair(b_dst, &tmB, lbar, kb * BK, n0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    if (lane == 0 && is_leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int n_blk = (int)(tile % n_tiles);
        const int n_size = min(BN, N - n_blk * BN);
        const uint32_t idesc = make_idesc(KIND, BM * CG, n_size);
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator stage
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t a_src = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_src = a_src + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < KB_BYTES / UMMA_K_BYTES; ++k) {
            const uint64_t adesc = make_smem_desc(a_src + k * UMMA_K_BYTES);
            const uint64_t bdesc = make_smem_desc(b_src + k * UMMA_K_BYTES);
            ptx::umma<CG, KIND>(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 1) ptx::umma_commit(empty_bar(stage));
          else ptx::umma_commit_pair(empty_bar(stage), 0x3);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if constexpr (CG == 1) ptx::umma_commit(tfull_bar(acc));
        else ptx::umma_commit_pair(tfull_bar(acc), 0x3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    const int q = warp & 3;          // TMEM lane quarter this warp may read (warp id % 4)
    const int half = (warp - 4) >> 2;  // 0: even 32-column chunks, 1: odd chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t leader_tempty0 = (CG == 2) ? ptx::mapa(tempty_bar(0), 0) : 0u;
    for (int64_t tile = first_tile; tile < total_tiles; tile += tile_stride) {
      const int64_t m_blk = tile / n_tiles;
      const int n_blk = (int)(tile % n_tiles);
      const int n_size = min(BN, N - n_blk * BN);
      const int64_t row = m_blk * BM * CG + cta_rank * BM + q * 32 + lane;
      const bool row_ok = row < M;
      const int ncol0 = n_blk * BN;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      int c = half * 32;
      for (; c + 96 <= n_size; c += 128) {  // two chunks (c, c + 64) per wait
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32(taddr + c, va);
        ptx::tmem_ld_32x32(taddr + c + 64, vb);
        ptx::tmem_ld_wait();
        epilogue_store<KIND, EPI, 32>(va, bias, Y, row, N, ncol0 + c, row_ok);
        epilogue_store<KIND, EPI, 32>(vb, bias, Y, row, N, ncol0 + c + 64, row_ok);
      }
      for (; c + 32 <= n_size; c += 64) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c, v);
        ptx::tmem_ld_wait();
        epilogue_store<KIND, EPI, 32>(v, bias, Y, row, N, ncol0 + c, row_ok);
      }
      if (c < n_size) {  // n_size % 32 == 16: the trailing half chunk belongs to the warp whose turn it is
        uint32_t v[16];
        ptx::tmem_ld_32x16(taddr + c, v);
        ptx::tmem_ld_wait();
        epilogue_store<KIND, EPI, 16>(v, bias, Y, row, N, ncol0 + c, row_ok);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) ptx::mbar_arrive(tempty_bar(acc));
        else ptx::mbar_arrive_cluster(leader_tempty0 + 8u * acc);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// [rows, K] row-major matrix, box = 128 bytes of K x box_rows rows, 128B swizzle, out-of-bounds -> zeros
int make_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int K, int esz, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return MPL_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)K * esz};
  const cuuint32_t box[2] = {(cuuint32_t)(KB_BYTES / esz), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld K=%d esz=%d box_rows=%d ptr=%p)", (int)r,
              (long long)rows, K, esz, box_rows, ptr);
    return MPL_ERR_CUDA;
  }
  return MPL_OK;
}

int g_gemm_cta_group = 2;  // CTA pairs by default: half the operand traffic per SM (measured faster on every FPT shape)

template <int CG, int KIND, int EPI>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, void* Y, int64_t M, int N, int K,
               cudaStream_t s) {
  using C = Cfg<CG>;
  auto kern = gemm_tcgen05_kernel<CG, KIND, EPI>;
  static bool attr_set = false;  // per instantiation; the attribute is per function, valid on every device of the process
  if (!attr_set) {
    MPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int n_tiles = (N + BN - 1) / BN;
  const int64_t m_tiles = ceil_div(M, (int64_t)BM * CG);
  const int64_t total = m_tiles * n_tiles;
  const int64_t max_groups = kNumSMs / CG;
  const unsigned groups = (unsigned)(total < max_groups ? total : max_groups);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MPL_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, bias, Y, M, N, K));
  return MPL_OK;
}

template <int CG, int KIND>
int launch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, void* Y, int64_t M, int N, int K,
               cudaStream_t s) {
  switch (epi) {
    case 0: return launch_one<CG, KIND, 0>(tmA, tmB, bias, Y, M, N, K, s);
    case 1: return launch_one<CG, KIND, 1>(tmA, tmB, bias, Y, M, N, K, s);
    case 2: return launch_one<CG, KIND, 2>(tmA, tmB, bias, Y, M, N, K, s);
    default: return launch_one<CG, KIND, 3>(tmA, tmB, bias, Y, M, N, K, s);
  }
}

}  // namespace

void set_gemm_cta_group(int cg) { g_gemm_cta_group = (cg == 2) ? 2 : 1; }
int get_gemm_cta_group() { return g_gemm_cta_group; }

bool gemm_tcgen05_supports(int N, int K, int dtype) {
  const int esz = (dtype == MPL_PREC_BF16) ? 2 : 4;
  return N >= 16 && N % 16 == 0 && K >= 1 && ((int64_t)K * esz) % 16 == 0;
}

int launch_gemm_tcgen05(const void* A, const void* W, const float* bias, void* Y, int64_t M, int N, int K, int dtype,
                        int epilogue, int out_fp32, cudaStream_t s) {
  if (M == 0) return MPL_OK;
  if (dtype != MPL_PREC_BF16 && dtype != MPL_PREC_TF32) {
    set_error("launch_gemm_tcgen05: dtype must be MPL_PREC_BF16 or MPL_PREC_TF32");
    return MPL_ERR_INVALID_ARGUMENT;
  }
  if (!gemm_tcgen05_supports(N, K, dtype)) {
    set_error("tcgen05 projection needs N %% 16 == 0 and a 16-byte aligned row pitch (N=%d K=%d)", N, K);
    return MPL_ERR_UNSUPPORTED;
  }
  if (bias == nullptr) {
    set_error("launch_gemm_tcgen05: bias must not be null");
    return MPL_ERR_INVALID_ARGUMENT;
  }
  const int esz = (dtype == MPL_PREC_BF16) ? 2 : 4;
  const int cg = g_gemm_cta_group;
  CUtensorMap tmA, tmB;
  MPL_TRY(make_tmap(&tmA, A, M, K, esz, BM));
  MPL_TRY(make_tmap(&tmB, W, N, K, esz, BN / cg));
  int epi = epilogue;
  if (epilogue == EPI_BIAS && out_fp32) epi = 3;
  const int kind = (dtype == MPL_PREC_BF16) ? 0 : 1;
  if (cg == 1) {
    return kind == 0 ? launch_epi<1, 0>(epi, tmA, tmB, bias, Y, M, N, K, s) : launch_epi<1, 1>(epi, tmA, tmB, bias, Y, M, N, K, s);
  }
  return kind == 0 ? launch_epi<2, 0>(epi, tmA, tmB, bias, Y, M, N, K, s) : launch_epi<2, 1>(epi, tmA, tmB, bias, Y, M, N, K, s);
}

}  // namespace mpl
