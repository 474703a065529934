// Implicit-GEMM convolution / linear on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by TMA).
//
// Replaces the cuDNN/cuBLAS call sites nn.Conv2d (unet.py:76,108,155,169,176-180,313,402) and nn.Linear qkv/proj_out
// (rpe.py:111-112) of the reference for bf16 operands.
//
// GEMM view:   Y[m, co] = sum_{tap, ci} A[pixel(m, tap), ci] * W[tap][co][ci]  (+ second K-segment: 1x1 skip conv)
//   m = (frame n, oh, ow) in NHWC order.  One CTA computes a 128-pixel x BN-channel output tile.
//   A tile  : im2col is done by the TMA engine itself — the activation tensor is described as a 4-D tensor
//             (C, W, H, N) and the 128 pixels of an M tile form a box (64 ch, Wbox, Hbox, Nbox); tap (r,s) is the same
//             box shifted by (s-pad, r-pad), out-of-bounds rows/columns are zero-filled by the hardware (= padding).
//             Stride-2 convs use the tensor map's elementStrides (traversal stride 2 in W and H).
//   B tile  : weights pre-packed [tap][co_pad][ci_pad] bf16 (K-major), box (64 ci, BN co, 1 tap).
//   Both land in shared memory in the canonical 128-byte-swizzled K-major layout that tcgen05.mma reads directly.
//   Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..5 = epilogue
//             (TMEM -> registers -> padded smem staging -> coalesced bias/residual/store/GroupNorm-statistics).
//   The staging buffer aliases the operand ring (all MMAs have retired when the epilogue starts; one tile per CTA),
//   so two CTAs fit per SM and one CTA's epilogue overlaps the other's main loop.
#include "tc_common.cuh"
#include <mutex>

namespace fdm {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;          // bf16 elements per K chunk = 128 bytes = one swizzle row
constexpr int TC_THREADS = 192;    // 6 warps
constexpr int A_TILE_BYTES = TC_BM * TC_BK * 2;

struct TcParams {
  const float* bias;
  const float* resid;
  float* y_f32;
  __nv_bfloat16* y_op;
  double* stats;
  int M, HWo, Wo, Ho, Cout;
  int wbox, hbox, nbox;  // output-pixel geometry of an M tile
  int taps, ksize, pad, stride;
  int kchunks0, kchunks1;
  int klast0, klast1;  // 16-wide k-steps actually issued in the last 64-channel chunk of each K segment
  int out_nchw;
  unsigned long long* trace;  // debug: per-CTA phase timestamps (fdm_debug_set_trace), NULL in production
};

template <int BN>
struct TcSmem {
  static constexpr int B_TILE_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = BN >= 128 ? 3 : 4;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int ROW = BN + 4;  // staging row stride (floats): 16-byte aligned, conflict-free float4 access
  static constexpr int STAGING_BYTES = TC_BM * ROW * 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BYTES = (RING_BYTES > STAGING_BYTES ? RING_BYTES : STAGING_BYTES) + 1024;  // + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap ta0,
                                                            const __grid_constant__ CUtensorMap tw0,
                                                            const __grid_constant__ CUtensorMap ta1,
                                                            const __grid_constant__ CUtensorMap tw1, const TcParams p) {
  using S = TcSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[S::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[S::STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, n_off = blockIdx.y * BN;
  unsigned long long* tr = p.trace ? p.trace + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
#define FDM_TRACE(slot) do { if (tr != nullptr) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tr[slot] = t_; } } while (0)
  if (threadIdx.x == 0) FDM_TRACE(0);
  const int m0 = mt * TC_BM;
  const int iters0 = p.taps * p.kchunks0;
  const int iters = iters0 + p.kchunks1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&ta0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tw0) : "memory");
    if (p.kchunks1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&ta1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tw1) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(S::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) FDM_TRACE(1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // first output pixel of the tile -> box origin in (w, h, n)
      int w0, h0, n0;
      if (p.nbox > 1) {
        w0 = 0; h0 = 0; n0 = mt * p.nbox;
      } else {
        n0 = m0 / p.HWo;
        int r = m0 - n0 * p.HWo;
        h0 = r / p.Wo;
        w0 = r - h0 * p.Wo;
      }
      for (int it = 0; it < iters; ++it) {
        const int stage = it % S::STAGES;
        const uint32_t phase = (it / S::STAGES) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_TILE_BYTES;
        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
        if (it < iters0) {
          const int tap = it / p.kchunks0, kc = it - tap * p.kchunks0;
          const int r = tap / p.ksize, s = tap - r * p.ksize;
          tma_load_4d(a_dst, &ta0, &full_bar[stage], kc * TC_BK, w0 * p.stride + s - p.pad, h0 * p.stride + r - p.pad, n0);
          tma_load_3d(b_dst, &tw0, &full_bar[stage], kc * TC_BK, n_off, tap);
        } else {
          const int kc = it - iters0;
          tma_load_4d(a_dst, &ta1, &full_bar[stage], kc * TC_BK, w0, h0, n0);
          tma_load_3d(b_dst, &tw1, &full_bar[stage], kc * TC_BK, n_off, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int it = 0; it < iters; ++it) {
        const int stage = it % S::STAGES;
        const uint32_t phase = (it / S::STAGES) & 1;
        mbar_wait(&full_bar[stage], phase);
        if (it == 0) FDM_TRACE(2);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(a_addr);
        const uint64_t bdesc = make_smem_desc(a_addr + A_TILE_BYTES);
        // channels beyond C0/C1 in the last chunk are TMA zero-fill: skip their k-steps
        int nk = TC_BK / 16;
        if (it < iters0) {
          if ((it % p.kchunks0) == p.kchunks0 - 1) nk = p.klast0;
        } else if (it == iters - 1) {
          nk = p.klast1;
        }
        for (int k = 0; k < nk; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) start-address field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
      }
      umma_commit(&tmem_full_bar);       // accumulator complete
      FDM_TRACE(3);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int g = warp & 3;  // TMEM lane group this warp may access: lanes [32g, 32g+32)
    mbar_wait(&tmem_full_bar, 0);
    if (warp == 2 && lane == 0) FDM_TRACE(4);
    tcgen05_fence_after();
    float* stg = reinterpret_cast<float*>(smem);  // aliases the operand ring: every MMA (hence every smem read) has retired
    float* my_row = stg + (size_t)(g * 32 + lane) * S::ROW;
    if constexpr (BN == 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(g * 32) << 16), v);
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(my_row + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    } else {
#pragma unroll
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(g * 32) << 16) + c, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(my_row + c + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
    __syncwarp();
    if (warp == 2 && lane == 0) FDM_TRACE(5);
    if (p.out_nchw) {
      // narrow head conv (unet.py:402,462-464): eps written as [N][Cout][Ho][Wo] fp32; lane <-> pixel row, few channels
      const int row = g * 32 + lane, m = m0 + row;
      if (m < p.M) {
        const int f = m / p.HWo, r = m - f * p.HWo;
        for (int c = 0; c < BN && n_off + c < p.Cout; ++c) {
          float v = stg[(size_t)row * S::ROW + c] + (p.bias ? p.bias[n_off + c] : 0.f);
          p.y_f32[((size_t)f * p.Cout + n_off + c) * p.HWo + r] = v;
        }
      }
    } else {
    // phase 2: this warp's 32 rows, lanes across channels (float4 each): coalesced global traffic
    constexpr int LPR = BN / 4;         // lanes per row
    constexpr int RPI = 32 / LPR > 0 ? 32 / LPR : 1;  // rows per iteration (BN <= 128)
    const int cl = (lane % LPR) * 4, rsub = lane / LPR;
    const int cg = n_off + cl;
    const bool col_ok = cg < p.Cout;    // Cout % 4 == 0
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && p.bias != nullptr) bias = *reinterpret_cast<const float4*>(p.bias + cg);
    const int seg_rows = p.HWo < 32 ? p.HWo : 32;  // rows of one frame inside this warp's 32 rows
    for (int seg0 = 0; seg0 < 32; seg0 += seg_rows) {
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      const int mseg = m0 + g * 32 + seg0;
      for (int rr = rsub; rr < seg_rows; rr += RPI) {
        const int row = g * 32 + seg0 + rr;
        const int m = m0 + row;
        if (m < p.M && col_ok) {
          float4 a = *reinterpret_cast<const float4*>(stg + (size_t)row * S::ROW + cl);
          float v[4] = {a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w};
          const size_t o = (size_t)m * p.Cout + cg;
          if (p.resid != nullptr) {
            float4 r4 = *reinterpret_cast<const float4*>(p.resid + o);
            v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { s1[j] += v[j]; s2[j] += v[j] * v[j]; }
          if (p.y_f32 != nullptr) *reinterpret_cast<float4*>(p.y_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
          if (p.y_op != nullptr) OpType<__nv_bfloat16>::store4(p.y_op + o, make_float4(v[0], v[1], v[2], v[3]));
        }
      }
      if (p.stats != nullptr) {
        // combine the RPI row-subgroups of the warp, then one atomic per (frame, channel) per warp segment
#pragma unroll
        for (int off = 16; off >= LPR; off >>= 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
          }
        }
        if (rsub == 0 && col_ok && mseg < p.M) {
          double* dst = p.stats + ((size_t)(mseg / p.HWo) * p.Cout + cg) * 2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            atomicAdd(dst + 2 * j, (double)s1[j]);
            atomicAdd(dst + 2 * j + 1, (double)s2[j]);
          }
        }
      }
    }
    }  // !out_nchw
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FDM_TRACE(6);
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::TMEM_COLS));
  }
}

// ---------------------------------------------------------------- host side
static unsigned long long* g_trace = nullptr;
EncodeTiledFn get_tensormap_encoder() {
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

// activations [N][H][W][C] bf16 as a (C, W, H, N) tensor; box (64, wbox*sw, hbox*sh, nbox) traversed with stride (1, s, s, 1)
static bool encode_act(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int wbox, int hbox, int nbox, int stride) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(wbox * stride), (cuuint32_t)(hbox * stride), (cuuint32_t)nbox};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// weights [taps][co_pad][ci_pad] bf16 as a (ci_pad, co_pad, taps) tensor; box (64, BN, 1)
static bool encode_w(CUtensorMap* m, const void* ptr, int taps, int co_pad, int ci_pad, int bn) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)ci_pad, (cuuint64_t)co_pad, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)ci_pad * 2, (cuuint64_t)co_pad * ci_pad * 2};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)bn, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
static int launch_tc(const CUtensorMap& ta0, const CUtensorMap& tw0, const CUtensorMap& ta1, const CUtensorMap& tw1,
                     const TcParams& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::BYTES);
  });
  if (attr_err != cudaSuccess) {
    set_last_error(attr_err);
    return FDM_ERR_CUDA;
  }
  dim3 grid((p.M + TC_BM - 1) / TC_BM, (p.Cout + BN - 1) / BN);
  conv_tc_kernel<BN><<<grid, TC_THREADS, TcSmem<BN>::BYTES, st>>>(ta0, tw0, ta1, tw1, p);
  return check_launch();
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int conv_tc_launch(const fdm_conv_args* a, cudaStream_t st) {
  FDM_REQUIRE(a->a_dtype == FDM_BF16, FDM_ERR_UNSUPPORTED);
  FDM_REQUIRE(!a->upsample, FDM_ERR_UNSUPPORTED);
  FDM_REQUIRE(a->C0 % 8 == 0 && (a->a1 == nullptr || a->C1 % 8 == 0), FDM_ERR_UNSUPPORTED);
  if (a->out_nchw) {
    FDM_REQUIRE(a->y_f32 != nullptr && a->y_op == nullptr && a->resid == nullptr && a->stats == nullptr, FDM_ERR_UNSUPPORTED);
  } else {
    FDM_REQUIRE(a->Cout % 4 == 0, FDM_ERR_UNSUPPORTED);
  }
  FDM_REQUIRE(a->y_op == nullptr || a->op_dtype == FDM_BF16, FDM_ERR_UNSUPPORTED);
  const int pad = a->ksize / 2;
  const int Ho = (a->Hin + 2 * pad - a->ksize) / a->stride + 1, Wo = (a->Win + 2 * pad - a->ksize) / a->stride + 1;
  const int HWo = Ho * Wo;
  TcParams p;
  p.bias = a->bias; p.resid = a->resid; p.y_f32 = a->y_f32; p.y_op = reinterpret_cast<__nv_bfloat16*>(a->y_op); p.stats = reinterpret_cast<double*>(a->stats);
  p.M = a->N * HWo; p.HWo = HWo; p.Wo = Wo; p.Ho = Ho; p.Cout = a->Cout;
  // M-tile geometry: 128 consecutive output pixels must form a box in (w, h, n)
  if (Wo >= TC_BM) {
    FDM_REQUIRE(Wo % TC_BM == 0, FDM_ERR_UNSUPPORTED);
    p.wbox = TC_BM; p.hbox = 1; p.nbox = 1;
  } else if (HWo >= TC_BM) {
    FDM_REQUIRE(TC_BM % Wo == 0 && HWo % TC_BM == 0, FDM_ERR_UNSUPPORTED);
    p.wbox = Wo; p.hbox = TC_BM / Wo; p.nbox = 1;
  } else {
    FDM_REQUIRE(TC_BM % HWo == 0, FDM_ERR_UNSUPPORTED);
    p.wbox = Wo; p.hbox = Ho; p.nbox = TC_BM / HWo;
  }
  // GroupNorm statistics are flushed per warp segment: a warp's 32 rows must not straddle frames unevenly
  FDM_REQUIRE(a->stats == nullptr || HWo % 32 == 0 || 32 % HWo == 0, FDM_ERR_UNSUPPORTED);
  p.taps = a->ksize * a->ksize; p.ksize = a->ksize; p.pad = pad; p.stride = a->stride;
  p.kchunks0 = (a->C0 + TC_BK - 1) / TC_BK;
  p.kchunks1 = a->a1 ? (a->C1 + TC_BK - 1) / TC_BK : 0;
  p.klast0 = (a->C0 - (p.kchunks0 - 1) * TC_BK + 15) / 16;
  p.klast1 = a->a1 ? (a->C1 - (p.kchunks1 - 1) * TC_BK + 15) / 16 : 0;
  p.out_nchw = a->out_nchw;
  p.trace = g_trace;
  const int bn = a->Cout % 128 == 0 ? 128 : (a->Cout >= 64 ? 64 : (a->Cout > 16 ? 32 : 16));
  const int co_pad = round_up(a->Cout, 16);
  CUtensorMap ta0, tw0, ta1, tw1;
  bool ok = encode_act(&ta0, a->a0, a->N, a->Hin, a->Win, a->C0, p.wbox, p.hbox, p.nbox, a->stride) &&
            encode_w(&tw0, a->w0, p.taps, co_pad, round_up(a->C0, TC_BK), bn);
  if (ok && a->a1) {
    ok = encode_act(&ta1, a->a1, a->N, Ho, Wo, a->C1, p.wbox, p.hbox, p.nbox, 1) &&
         encode_w(&tw1, a->w1, 1, co_pad, round_up(a->C1, TC_BK), bn);
  } else {
    ta1 = ta0;
    tw1 = tw0;
  }
  FDM_REQUIRE(ok, FDM_ERR_UNSUPPORTED);
  if (bn == 128) return launch_tc<128>(ta0, tw0, ta1, tw1, p, st);
  if (bn == 64) return launch_tc<64>(ta0, tw0, ta1, tw1, p, st);
  if (bn == 32) return launch_tc<32>(ta0, tw0, ta1, tw1, p, st);
  return launch_tc<16>(ta0, tw0, ta1, tw1, p, st);
}

}  // namespace fdm

// debug hook (not part of the product ABI): per-CTA globaltimer stamps of the next conv_tc launches are written to
// trace[cta][8] = {start, prologue done, first operands landed, last MMA issued, accumulator ready, staged, done}
extern "C" void fdm_debug_set_trace(unsigned long long* trace) { fdm::g_trace = trace; }
