// Attention-block linears with the GroupNorm in the operand path (tcgen05.mma, persistent CTAs, weights resident in shared memory).
//
// Reference call sites (rpe.py:111-113,135-140,170-173): every RPEAttention does
//     x = norm(x);  qkv = self.qkv(x);  ... attention ...;  return x + self.proj_out(h)
// i.e. a GroupNorm, a C -> 3C linear, and a C -> C linear whose residual is the NORMALISED input.  As separate kernels that is a
// normalisation pass (fp32 in, fp32 + bf16 out), a GEMM of 960 one-shot CTAs, and a second GEMM re-reading the fp32 normalised
// tensor.  Here the normalisation never exists in memory:
//   * qkv  : A-operand tiles are built IN shared memory from the raw fp32 activations (TMA box -> 4 transform warps: statistics,
//            scale/shift, bf16, 128-byte-swizzled K-major rows) and multiplied with W_qkv, which stays resident for the CTA's life;
//            a_mode 1 = per-frame GroupNorm from the (sum, sum of squares) pairs the producing kernel's epilogue left behind,
//            a_mode 2 = the temporal GroupNorm (statistics over C/32 channels x T frames per (video, pixel)), computed in the tile
//            itself: a tile is PL pixels x all T frames, fetched as ONE 4-D TMA box of the [B*T][HW][C] tensor.
//   * proj : a_mode 0 (bf16 operand by TMA, double buffered); the epilogue adds the residual GN(x) RECOMPUTED from x and the saved
//            statistics (resid_mode 2: per-(pixel, group) mean / rstd written by the a_mode 2 launch; resid_mode 3: per-frame sums),
//            and accumulates the GroupNorm statistics of its own output like every conv epilogue (fp64 atomics).
// Pipeline: warp 0 = TMA (weights before griddepcontrol.wait: they do not depend on the previous kernel), warp 1 = MMA issuer,
// warps 2-5 = transform producers, warps 6-13 = epilogue (two warps per TMEM lane quarter, 64 columns each).  Accumulators: a ring
// of four 128-column TMEM slots, so the MMAs of the next tile run under the epilogue of this one.
// Shapes: K = C = 128 (two 64-channel chunks, GroupNorm groups of 4 channels = one float4), Cout = 128 / 256 / 384.
#include "tc_common.cuh"
#include <mutex>

namespace fdm {

constexpr int NL_THREADS = 448;   // 14 warps
constexpr int NL_K = 128;
constexpr int NL_A_BYTES = 2 * 16384;       // [2 chunks][128 rows][128 B]
constexpr int NL_RAW_BYTES = 128 * NL_K * 4;
constexpr int NL_STG_ROW = 36;              // floats per staged row (fp32 epilogue)
constexpr int NL_STG_F32 = 32 * NL_STG_ROW * 4;
constexpr int NL_STG_BF16 = 2560;

struct NlParams {
  const float* x;
  const double* stats;
  float* tstats;
  const float* gamma;
  const float* beta;
  const float* bias;
  const float* resid;
  float* y_f32;
  __nv_bfloat16* y_op;
  double* out_stats;
  int M, B, T, HW, Cout, NS;
  int a_mode, resid_mode;
  int PL, tpv, tiles;  // a_mode 2: pixels per tile, tiles per video
  int NA;              // A buffers (2 when TMA writes them, 1 behind the transform warps)
  float eps;
  long long* trace;    // debug: per-CTA phase timestamps (clock64), NULL in production (fdm_debug_nl_trace)
};

#define NL_TRACE(slot)                                                                      \
  do {                                                                                      \
    if (p.trace != nullptr && (slot) < 64) p.trace[blockIdx.x * 64 + (slot)] = clock64();   \
  } while (0)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// tile row -> row of the [M][C] tensors, or -1.  a_mode 2 tiles are ordered (frame, pixel): row = t * PL + pl.
__device__ __forceinline__ int nl_row(const NlParams& p, int tile, int r) {
  if (p.a_mode == 2) {
    const int t = r / p.PL, pl = r - t * p.PL;
    const int vb = tile / p.tpv, px = (tile - vb * p.tpv) * p.PL + pl;
    if (t >= p.T || px >= p.HW) return -1;
    return (vb * p.T + t) * p.HW + px;
  }
  const int m = tile * 128 + r;
  return m < p.M ? m : -1;
}

__global__ void __launch_bounds__(NL_THREADS, 1) norm_linear_kernel(const __grid_constant__ CUtensorMap ta,
                                                                   const __grid_constant__ CUtensorMap tw, const NlParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t w_bar, raw_full, raw_empty, a_full[2], a_empty[2], t_full[4], t_empty[4];
  __shared__ uint32_t tmem_slot;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w_bytes = p.Cout * NL_K * 2;
  uint8_t* w_s = smem;                                 // [2 chunks][Cout rows][128 B]
  uint8_t* a_s = w_s + w_bytes;                        // [NA][2 chunks][128 rows][128 B]
  uint8_t* raw_s = a_s + p.NA * NL_A_BYTES;            // [128 rows][128 fp32]   (a_mode 1, 2)
  uint8_t* stg_base = raw_s + (p.a_mode ? NL_RAW_BYTES : 0);
  const bool f32_path = p.y_f32 != nullptr;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&ta) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tw) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&w_bar, 1);
    mbar_init(&raw_full, 1);
    mbar_init(&raw_empty, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], p.a_mode ? 4 : 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) NL_TRACE(0);

  if (warp == 0) {
    // ===================== TMA: weights once (independent of the previous kernel: issued before griddepcontrol.wait), then tiles
    if (elect_one_sync()) {
      mbar_expect_tx(&w_bar, (uint32_t)w_bytes);
      for (int c = 0; c < 2; ++c)
        for (int n0 = 0; n0 < p.Cout; n0 += 128) tma_load_3d(w_s + c * p.Cout * 128 + n0 * 128, &tw, &w_bar, c * 64, n0, 0);
    }
    __syncwarp();
    pdl_wait();
    if (lane == 0) NL_TRACE(1);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      if (p.a_mode == 0) {
        const int s = it & 1;
        mbar_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&a_full[s], NL_A_BYTES);
          for (int c = 0; c < 2; ++c) tma_load_3d(a_s + s * NL_A_BYTES + c * 16384, &ta, &a_full[s], c * 64, tile * 128, 0);
        }
      } else {
        mbar_wait(&raw_empty, (it & 1) ^ 1);
        if (elect_one_sync()) {
          if (p.a_mode == 1) {
            mbar_expect_tx(&raw_full, NL_RAW_BYTES);
            tma_load_3d(raw_s, &ta, &raw_full, 0, tile * 128, 0);
          } else {
            const int vb = tile / p.tpv, px0 = (tile - vb * p.tpv) * p.PL;
            mbar_expect_tx(&raw_full, (uint32_t)(p.PL * p.T * NL_K * 4));
            tma_load_4d(raw_s, &ta, &raw_full, 0, px0, 0, vb);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer
    pdl_wait();
    constexpr uint32_t idesc = make_idesc(128);
    mbar_wait(&w_bar, 0);
    if (lane == 0) NL_TRACE(2);
    uint32_t seq = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int s = p.NA == 2 ? (it & 1) : 0;
      const uint32_t ph = p.NA == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(&a_full[s], ph);
      if (lane == 0) NL_TRACE(8 + 8 * it + 2);
      tcgen05_fence_after();
      const uint32_t a_lo = smem_desc_lo(smem_u32(a_s + s * NL_A_BYTES));
      for (int j = 0; j < p.NS; ++j, ++seq) {
        const uint32_t slot = seq & 3;
        mbar_wait(&t_empty[slot], ((seq >> 2) & 1) ^ 1);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t d = tmem_base + slot * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t al = a_lo + c * (16384 >> 4);
            const uint32_t bl = smem_desc_lo(smem_u32(w_s + c * p.Cout * 128 + j * 128 * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16_lo(d, al + 2 * kk, bl + 2 * kk, idesc, (c | kk) != 0);
          }
          umma_commit(&t_full[slot]);
          if (j == p.NS - 1) umma_commit(&a_empty[s]);
        }
        __syncwarp();
      }
      if (lane == 0) NL_TRACE(8 + 8 * it + 3);
    }
  } else if (warp < 6) {
    // ===================== transform producers: raw fp32 tile -> GroupNorm -> bf16 A operand (K-major, 128-byte swizzle)
    pdl_wait();
    if (p.a_mode != 0) {
      const int pw = warp - 2;
      const int c4 = lane * 4;  // this lane's 4 channels = one GroupNorm group (C = 128)
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + c4));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + c4));
      // destination of this lane's 8 bytes inside a row: chunk lane/16, 16-byte piece (lane%16)/2 (XOR row%8), half lane%2
      const uint32_t d_chunk = (uint32_t)(lane >> 4) * 16384u, d_piece = (uint32_t)((lane & 15) >> 1), d_half = (uint32_t)(lane & 1) * 8u;
      const float* raw = reinterpret_cast<const float*>(raw_s);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        if (p.a_mode == 1) {
          // per-frame statistics of the rows of this warp: fetched while the tile is still in flight
          const int r0 = pw * 32;
          int cur_n = -1;
          float mul[4] = {0.f, 0.f, 0.f, 0.f}, add[4] = {0.f, 0.f, 0.f, 0.f};
          auto frame_coefs = [&](int n) {
            const double2* st = reinterpret_cast<const double2*>(p.stats + ((size_t)n * NL_K + c4) * 2);
            const double2 s0 = st[0], s1 = st[1], s2 = st[2], s3 = st[3];
            const double cnt = 4.0 * (double)p.HW;
            const double mean = (s0.x + s1.x + s2.x + s3.x) / cnt;
            const double var = fmax((s0.y + s1.y + s2.y + s3.y) / cnt - mean * mean, 0.0);
            const float mf = (float)mean, rs = (float)(1.0 / sqrt(var + (double)p.eps));
            mul[0] = g4.x * rs; mul[1] = g4.y * rs; mul[2] = g4.z * rs; mul[3] = g4.w * rs;
            add[0] = b4.x - mf * mul[0]; add[1] = b4.y - mf * mul[1]; add[2] = b4.z - mf * mul[2]; add[3] = b4.w - mf * mul[3];
          };
          {
            const int m = tile * 128 + r0;
            if (m < p.M) { cur_n = m / p.HW; frame_coefs(cur_n); }
          }
          mbar_wait(&raw_full, it & 1);
          if (threadIdx.x == 64) NL_TRACE(8 + 8 * it + 0);
          mbar_wait(&a_empty[0], (it & 1) ^ 1);
          uint8_t* A = a_s;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int row = r0 + rr, m = tile * 128 + row;
            if (m < p.M) {
              const int n = m / p.HW;
              if (n != cur_n) { cur_n = n; frame_coefs(n); }
            }
            const float4 v = *reinterpret_cast<const float4*>(raw + row * NL_K + c4);
            uint2 o;
            o.x = pack2_bf16(fmaf(v.x, mul[0], add[0]), fmaf(v.y, mul[1], add[1]));
            o.y = pack2_bf16(fmaf(v.z, mul[2], add[2]), fmaf(v.w, mul[3], add[3]));
            *reinterpret_cast<uint2*>(A + d_chunk + row * 128 + ((d_piece ^ (uint32_t)(row & 7)) << 4) + d_half) = o;
          }
        } else {
          mbar_wait(&raw_full, it & 1);
          if (threadIdx.x == 64) NL_TRACE(8 + 8 * it + 0);
          mbar_wait(&a_empty[0], (it & 1) ^ 1);
          uint8_t* A = a_s;
          const int vb = tile / p.tpv, px0 = (tile - vb * p.tpv) * p.PL;
          const float cnt = 4.f * (float)p.T;
          for (int pl = pw; pl < p.PL; pl += 4) {
            // statistics of (pixel pl, group lane) over the T frames, one pass around a pivot (the group's first value)
            const float* col = raw + pl * NL_K + c4;
            const int rstride = p.PL * NL_K;
            const float pivot = col[0];
            float s = 0.f, ss = 0.f;
#pragma unroll 4
            for (int t = 0; t < p.T; ++t) {
              const float4 v = *reinterpret_cast<const float4*>(col + t * rstride);
              const float d0 = v.x - pivot, d1 = v.y - pivot, d2 = v.z - pivot, d3 = v.w - pivot;
              s += (d0 + d1) + (d2 + d3);
              ss = fmaf(d0, d0, ss); ss = fmaf(d1, d1, ss); ss = fmaf(d2, d2, ss); ss = fmaf(d3, d3, ss);
            }
            const float md = s / cnt;
            const float mean = pivot + md;
            const float rstd = rsqrtf(fmaxf(ss / cnt - md * md, 0.f) + p.eps);
            if (p.tstats != nullptr && px0 + pl < p.HW)
              *reinterpret_cast<float2*>(p.tstats + (((size_t)vb * p.HW + px0 + pl) * 32 + lane) * 2) = make_float2(mean, rstd);
#pragma unroll 4
            for (int t = 0; t < p.T; ++t) {
              const float4 v = *reinterpret_cast<const float4*>(col + t * rstride);
              const int row = t * p.PL + pl;
              uint2 o;
              o.x = pack2_bf16((v.x - mean) * rstd * g4.x + b4.x, (v.y - mean) * rstd * g4.y + b4.y);
              o.y = pack2_bf16((v.z - mean) * rstd * g4.z + b4.z, (v.w - mean) * rstd * g4.w + b4.w);
              *reinterpret_cast<uint2*>(A + d_chunk + row * 128 + ((d_piece ^ (uint32_t)(row & 7)) << 4) + d_half) = o;
            }
          }
        }
        fence_proxy_async_smem();  // the A tile is read by the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&raw_empty);
          mbar_arrive(&a_full[0]);
        }
        if (threadIdx.x == 64) NL_TRACE(8 + 8 * it + 1);
      }
    }
  } else {
    // ===================== epilogue (8 warps): TMEM lane quarter = warp % 4, column half of every 128-column slot = (warp - 6) / 4
    pdl_wait();
    const int q = warp & 3, hf = (warp - 6) >> 2;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t seq = 0;
    if (!f32_path) {
      // ---- bf16 outputs (qkv): + bias, pack, transpose 64 bytes per lane through the warp's staging slice, 4 lanes per row
      uint8_t* stg = stg_base + (warp - 6) * NL_STG_BF16;
      const int piece = lane & 3;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        int ro[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ro[i] = nl_row(p, tile, q * 32 + i * 8 + (lane >> 2));
        for (int j = 0; j < p.NS; ++j, ++seq) {
          const uint32_t slot = seq & 3;
          mbar_wait(&t_full[slot], (seq >> 2) & 1);
          if (threadIdx.x == 192 && j == 0) NL_TRACE(8 + 8 * it + 4);
          tcgen05_fence_after();
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int col = hf * 64 + cc * 32, gcol = j * 128 + col;
            uint32_t v[32];
            tmem_ld_32x32b_x32(trow + slot * 128 + col, v);
            if (cc == 1) {  // both halves of this warp's columns are in registers: hand the slot back
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&t_empty[slot]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol + 8 * k));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol + 8 * k + 4));
              uint4 w;
              w.x = pack2_bf16(__uint_as_float(v[8 * k]) + b0.x, __uint_as_float(v[8 * k + 1]) + b0.y);
              w.y = pack2_bf16(__uint_as_float(v[8 * k + 2]) + b0.z, __uint_as_float(v[8 * k + 3]) + b0.w);
              w.z = pack2_bf16(__uint_as_float(v[8 * k + 4]) + b1.x, __uint_as_float(v[8 * k + 5]) + b1.y);
              w.w = pack2_bf16(__uint_as_float(v[8 * k + 6]) + b1.z, __uint_as_float(v[8 * k + 7]) + b1.w);
              *reinterpret_cast<uint4*>(stg + lane * 80 + k * 16) = w;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = i * 8 + (lane >> 2);
              if (ro[i] >= 0)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.y_op + (size_t)ro[i] * p.Cout + gcol) + piece * 16) =
                    *reinterpret_cast<const uint4*>(stg + row * 80 + piece * 16);
            }
            __syncwarp();
          }
        }
        if (threadIdx.x == 192) NL_TRACE(8 + 8 * it + 5);
      }
    } else {
      // ---- fp32 outputs (proj_out): + bias + residual (plain, or GN(x) recomputed), GroupNorm statistics of the result
      float* stg = reinterpret_cast<float*>(stg_base + (warp - 6) * NL_STG_F32);
      const int sub = lane >> 3, cq = (lane & 7) * 4;
      const bool two_frames = p.HW < 32;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int m_w = tile * 128 + q * 32;
        for (int j = 0; j < p.NS; ++j, ++seq) {
          const uint32_t slot = seq & 3;
          mbar_wait(&t_full[slot], (seq >> 2) & 1);
          if (threadIdx.x == 192 && j == 0) NL_TRACE(8 + 8 * it + 4);
          tcgen05_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int col = j * 128 + hf * 64 + cc * 32 + cq;
            {
              uint32_t v[32];
              tmem_ld_32x32b_x32(trow + slot * 128 + hf * 64 + cc * 32, v);
              if (cc == 1) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[slot]);
              }
#pragma unroll
              for (int k = 0; k < 32; k += 4)
                *reinterpret_cast<float4*>(stg + lane * NL_STG_ROW + k) =
                    make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
            }
            __syncwarp();
            const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
            if (p.resid_mode >= 2) {
              g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
              b4 = __ldg(reinterpret_cast<const float4*>(p.beta + col));
            }
            // resid_mode 3: per-frame scale / shift of this lane's group, for the (at most two) frames of the warp's 32 rows
            float mul[2][4], add[2][4];
            if (p.resid_mode == 3) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int mseg = m_w + hh * 16;
                if (hh == 1 && !two_frames) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) { mul[1][k] = mul[0][k]; add[1][k] = add[0][k]; }
                } else if (mseg < p.M) {
                  const double2* st = reinterpret_cast<const double2*>(p.stats + ((size_t)(mseg / p.HW) * p.Cout + col) * 2);
                  const double2 s0 = st[0], s1 = st[1], s2 = st[2], s3 = st[3];
                  const double cnt = 4.0 * (double)p.HW;
                  const double mean = (s0.x + s1.x + s2.x + s3.x) / cnt;
                  const double var = fmax((s0.y + s1.y + s2.y + s3.y) / cnt - mean * mean, 0.0);
                  const float mf = (float)mean, rs = (float)(1.0 / sqrt(var + (double)p.eps));
                  mul[hh][0] = g4.x * rs; mul[hh][1] = g4.y * rs; mul[hh][2] = g4.z * rs; mul[hh][3] = g4.w * rs;
#pragma unroll
                  for (int k = 0; k < 4; ++k) add[hh][k] = (k == 0 ? b4.x : k == 1 ? b4.y : k == 2 ? b4.z : b4.w) - mf * mul[hh][k];
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k) { mul[hh][k] = 0.f; add[hh][k] = 0.f; }
                }
              }
            }
            float s1[2][4], s2[2][4];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
              for (int k = 0; k < 4; ++k) s1[hh][k] = s2[hh][k] = 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              // the residual loads of four rows are issued before any is used
              float4 rx[4];
              float2 ts[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int m = m_w + (half * 4 + u) * 4 + sub;
                rx[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                ts[u] = make_float2(0.f, 0.f);
                if (m < p.M) {
                  if (p.resid_mode == 1) rx[u] = __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)m * p.Cout + col));
                  else if (p.resid_mode >= 2) rx[u] = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)m * p.Cout + col));
                  if (p.resid_mode == 2) {
                    const int n = m / p.HW, px = m - n * p.HW, b = n / p.T;
                    ts[u] = __ldg(reinterpret_cast<const float2*>(p.tstats + (((size_t)b * p.HW + px) * 32 + (col >> 2)) * 2));
                  }
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int row = (half * 4 + u) * 4 + sub;
                const int m = m_w + row;
                const float4 a = *reinterpret_cast<const float4*>(stg + row * NL_STG_ROW + cq);
                float r[4] = {rx[u].x, rx[u].y, rx[u].z, rx[u].w};
                if (p.resid_mode == 2) {
                  r[0] = (r[0] - ts[u].x) * ts[u].y * g4.x + b4.x;
                  r[1] = (r[1] - ts[u].x) * ts[u].y * g4.y + b4.y;
                  r[2] = (r[2] - ts[u].x) * ts[u].y * g4.z + b4.z;
                  r[3] = (r[3] - ts[u].x) * ts[u].y * g4.w + b4.w;
                } else if (p.resid_mode == 3) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) r[k] = fmaf(r[k], mul[half][k], add[half][k]);
                }
                const float o[4] = {a.x + bias.x + r[0], a.y + bias.y + r[1], a.z + bias.z + r[2], a.w + bias.w + r[3]};
                if (m < p.M) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) { s1[half][k] += o[k]; s2[half][k] = fmaf(o[k], o[k], s2[half][k]); }
                  const size_t off = (size_t)m * p.Cout + col;
                  *reinterpret_cast<float4*>(p.y_f32 + off) = make_float4(o[0], o[1], o[2], o[3]);
                  if (p.y_op != nullptr) OpType<__nv_bfloat16>::store4(p.y_op + off, make_float4(o[0], o[1], o[2], o[3]));
                }
              }
            }
            if (p.out_stats != nullptr) {
              // combine the 4 row-subgroups (lanes with equal lane % 8); one fp64 atomic per (frame, channel, moment) per warp
#pragma unroll
              for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  s1[hh][k] += __shfl_xor_sync(0xffffffffu, s1[hh][k], 8);
                  s2[hh][k] += __shfl_xor_sync(0xffffffffu, s2[hh][k], 8);
                  s1[hh][k] += __shfl_xor_sync(0xffffffffu, s1[hh][k], 16);
                  s2[hh][k] += __shfl_xor_sync(0xffffffffu, s2[hh][k], 16);
                }
              if (sub == 0) {
                if (two_frames) {
#pragma unroll
                  for (int hh = 0; hh < 2; ++hh) {
                    const int mseg = m_w + hh * 16;
                    if (mseg < p.M) {
                      double* dst = p.out_stats + ((size_t)(mseg / p.HW) * p.Cout + col) * 2;
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        atomicAdd(dst + 2 * k, (double)s1[hh][k]);
                        atomicAdd(dst + 2 * k + 1, (double)s2[hh][k]);
                      }
                    }
                  }
                } else if (m_w < p.M) {
                  double* dst = p.out_stats + ((size_t)(m_w / p.HW) * p.Cout + col) * 2;
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    atomicAdd(dst + 2 * k, (double)(s1[0][k] + s1[1][k]));
                    atomicAdd(dst + 2 * k + 1, (double)(s2[0][k] + s2[1][k]));
                  }
                }
              }
            }
            __syncwarp();  // the staging slice is overwritten by the next chunk
          }
        }
        if (threadIdx.x == 192) NL_TRACE(8 + 8 * it + 5);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) NL_TRACE(3);
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---------------------------------------------------------------------------------------------------------------- host
static bool nl_encode(CUtensorMap* m, CUtensorMapDataType dt, int esz, const void* ptr, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, bool swizzle) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  (void)esz;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, dt, rank, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static long long* g_nl_trace = nullptr;

int norm_linear_supported(const fdm_norm_linear_args* a) {
  if (a->K != NL_K || a->Cout % 128 != 0 || a->Cout < 128 || a->Cout > 384) return 0;
  if (a->B <= 0 || a->T <= 0 || a->HW <= 0) return 0;
  if (a->a_mode < 0 || a->a_mode > 2 || a->resid_mode < 0 || a->resid_mode > 3) return 0;
  if (a->a_mode == 2 && a->T > 128) return 0;
  const bool f32_path = a->y_f32 != nullptr;
  if (f32_path) {
    if (a->a_mode == 2) return 0;
    // a warp's 32 rows hold one frame or exactly two (statistics and the per-frame residual coefficients are kept per half-warp)
    if ((a->out_stats != nullptr || a->resid_mode == 3) && !(a->HW % 32 == 0 || a->HW == 16)) return 0;
    if (a->resid_mode >= 2 && a->Cout != NL_K) return 0;
  } else {
    if (a->y_op == nullptr || a->resid_mode != 0 || a->out_stats != nullptr) return 0;
  }
  return 1;
}

}  // namespace fdm

// debug hook (not part of the public header): device buffer of 148 x 64 clock64 stamps filled by the next launches, or NULL
extern "C" void fdm_debug_nl_trace(void* device_buf) { fdm::g_nl_trace = reinterpret_cast<long long*>(device_buf); }

extern "C" int fdm_norm_linear_supported(const fdm_norm_linear_args* a) { return a ? fdm::norm_linear_supported(a) : 0; }

extern "C" int fdm_norm_linear(const fdm_norm_linear_args* a, void* stream) {
  using namespace fdm;
  FDM_REQUIRE(a && a->w && a->bias, FDM_ERR_BAD_ARG);
  FDM_REQUIRE(norm_linear_supported(a), FDM_ERR_UNSUPPORTED);
  if (a->a_mode == 0) FDM_REQUIRE(a->a_op != nullptr, FDM_ERR_BAD_ARG);
  else FDM_REQUIRE(a->x != nullptr && a->gamma != nullptr && a->beta != nullptr, FDM_ERR_BAD_ARG);
  if (a->a_mode == 1) FDM_REQUIRE(a->stats != nullptr, FDM_ERR_BAD_ARG);
  if (a->resid_mode == 1) FDM_REQUIRE(a->resid != nullptr, FDM_ERR_BAD_ARG);
  if (a->resid_mode == 2) FDM_REQUIRE(a->x && a->tstats && a->gamma && a->beta, FDM_ERR_BAD_ARG);
  if (a->resid_mode == 3) FDM_REQUIRE(a->x && a->stats && a->gamma && a->beta, FDM_ERR_BAD_ARG);
  NlParams p;
  p.x = a->x; p.stats = a->stats; p.tstats = a->tstats; p.gamma = a->gamma; p.beta = a->beta; p.bias = a->bias; p.resid = a->resid;
  p.y_f32 = a->y_f32; p.y_op = reinterpret_cast<__nv_bfloat16*>(a->y_op); p.out_stats = a->out_stats;
  p.B = a->B; p.T = a->T; p.HW = a->HW; p.Cout = a->Cout; p.NS = a->Cout / 128;
  p.M = a->B * a->T * a->HW;
  p.a_mode = a->a_mode; p.resid_mode = a->resid_mode; p.eps = a->eps;
  p.NA = a->a_mode == 0 ? 2 : 1;
  p.trace = g_nl_trace;
  p.PL = 1; p.tpv = 1;
  if (a->a_mode == 2) {
    p.PL = 128 / a->T;
    if (p.PL > a->HW) p.PL = a->HW;
    p.tpv = (a->HW + p.PL - 1) / p.PL;
    p.tiles = a->B * p.tpv;
  } else {
    p.tiles = (p.M + 127) / 128;
  }
  CUtensorMap ta, tw;
  {
    cuuint64_t dims[3] = {(cuuint64_t)NL_K, (cuuint64_t)a->Cout, 1};
    cuuint64_t strides[2] = {(cuuint64_t)NL_K * 2, (cuuint64_t)a->Cout * NL_K * 2};
    cuuint32_t box[3] = {64, 128, 1};
    FDM_REQUIRE(nl_encode(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->w, 3, dims, strides, box, true), FDM_ERR_UNSUPPORTED);
  }
  if (a->a_mode == 0) {
    cuuint64_t dims[3] = {(cuuint64_t)NL_K, (cuuint64_t)p.M, 1};
    cuuint64_t strides[2] = {(cuuint64_t)NL_K * 2, (cuuint64_t)p.M * NL_K * 2};
    cuuint32_t box[3] = {64, 128, 1};
    FDM_REQUIRE(nl_encode(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->a_op, 3, dims, strides, box, true), FDM_ERR_UNSUPPORTED);
  } else if (a->a_mode == 1) {
    cuuint64_t dims[3] = {(cuuint64_t)NL_K, (cuuint64_t)p.M, 1};
    cuuint64_t strides[2] = {(cuuint64_t)NL_K * 4, (cuuint64_t)p.M * NL_K * 4};
    cuuint32_t box[3] = {NL_K, 128, 1};
    FDM_REQUIRE(nl_encode(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->x, 3, dims, strides, box, false), FDM_ERR_UNSUPPORTED);
  } else {
    // [B*T][HW][C] fp32 as (channel, pixel, frame, video): a tile = PL pixels x all T frames of one video
    cuuint64_t dims[4] = {(cuuint64_t)NL_K, (cuuint64_t)a->HW, (cuuint64_t)a->T, (cuuint64_t)a->B};
    cuuint64_t strides[3] = {(cuuint64_t)NL_K * 4, (cuuint64_t)a->HW * NL_K * 4, (cuuint64_t)a->T * a->HW * NL_K * 4};
    cuuint32_t box[4] = {NL_K, (cuuint32_t)p.PL, (cuuint32_t)a->T, 1};
    FDM_REQUIRE(nl_encode(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->x, 4, dims, strides, box, false), FDM_ERR_UNSUPPORTED);
  }
  const int smem = a->Cout * NL_K * 2 + p.NA * NL_A_BYTES + (a->a_mode ? NL_RAW_BYTES : 0) +
                   8 * (a->y_f32 ? NL_STG_F32 : NL_STG_BF16) + 1024;
  constexpr int SMEM_MAX = 226 * 1024;
  FDM_REQUIRE(smem <= SMEM_MAX, FDM_ERR_UNSUPPORTED);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(norm_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX); });
  if (attr_err != cudaSuccess) {
    set_last_error(attr_err);
    return FDM_ERR_CUDA;
  }
  const int grid = p.tiles < 148 ? p.tiles : 148;
  fdm::launch(norm_linear_kernel, dim3(grid), dim3(NL_THREADS), (size_t)smem, reinterpret_cast<cudaStream_t>(stream), ta, tw, p);
  return check_launch();
}
