// EXPERIMENTAL (off by default, FDM_TEMPORAL_MMA=1): temporal attention with the two contextual relative-position SCORE terms
// on tensor cores (bf16 mode).  Correct (kernel test vs torch), but measured no faster than attn_simt.cu on B200 in three
// variants (fragments from global memory; + cp.async double buffering; + 32-wide chunks with fragments from the staged
// tile): 2.27-2.37 ms per cfg4 step vs 2.25 ms.  ncu: L1/TEX throughput ~78 % in every variant — the per-pixel K/V staging
// traffic, not the FMA count, bounds this op.  Kept as the starting point for a tcgen05/TMA formulation.
//
//   S[t,s] = scale * ( q_t.k_s  +  q_t.Rk[t,s]  +  k_s.Rq[s,t] )        (rpe.py:144-152)
//
// q_t.k_s is a per-pixel T x T x F batch with no shared operand (CUDA cores, as in attn_simt.cu), but the two R terms ARE
// GEMMs over pixels, because R depends on (b, t, s) only:
//   D1[px, s] = Q[px, t, :] . Rk[t, s, :]      for a fixed query frame t :  [32 px x F] x [F x T]
//   D2[px, t] = K[px, s, :] . Rq[s, t, :]      for a fixed key frame s   :  [32 px x F] x [F x nw query frames of the block]
// They are 2/3 of the score FMAs.  Here each warp computes them with mma.sync.m16n8k16 (bf16 in, fp32 accumulate; T <= 40 is
// far below the tcgen05 tile minimum and there is nothing to pipeline), scatters the accumulator fragments into a shared
// [query frame][key frame][pixel] buffer, and the CUDA-core loop that follows starts its scores from that buffer and only adds
// q.k.  The P.(V + Rv) output phase is unchanged.
//
// grid (ceil(HW/32), heads, B * tgroups); block = nw warps; warp w owns query frame t = tg*nw + w; lane <-> pixel.
#include "common.cuh"

namespace fdm {

constexpr int TM_MAX_WARPS = 12;

struct TMParams {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* Rq_op;  // [B][T][T][C] bf16
  const __nv_bfloat16* Rk_op;
  const float* Rv;             // [B][T][T][C] fp32
  const float* mask;
  __nv_bfloat16* out;
  int B, T, HW, C, heads, F, tgroups, nw;
  float scale;
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint32_t ldg_u32(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }

// The first versions of this kernel were bound by L1/TEX sector throughput (ncu: 78 %), not by FMA issue: K/V were staged
// in 8-element slices (16 of every 32-byte sector used, each sector fetched twice) and the mma A fragments came from
// scattered 4-byte global loads.  Now a chunk is FC = 16 or 32 head dims (whole sectors per (frame, pixel)), the K tile is
// staged ONCE per chunk as packed bf16 pairs [s][pair][px], and it serves both the CUDA-core q.k loop (lane = pixel:
// conflict-free) and the A fragments of the D2 GEMM (word (pair, px) IS the fragment register).
template <int TP, int FC>
__global__ void __launch_bounds__(TM_MAX_WARPS * 32) attn_temporal_mma_kernel(TMParams p) {
  extern __shared__ __align__(16) float tm_smem[];
  pdl_launch_dependents();
  pdl_wait();
  constexpr int TN = (TP + 7) / 8;              // key-frame n-tiles of D1
  constexpr int QN = (TM_MAX_WARPS + 7) / 8;    // n-tiles over the block's query frames (D2)
  constexpr int PAIRS = FC / 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nthreads = blockDim.x;
  const int gid = lane >> 2, tig = lane & 3;
  const int T = p.T, C = p.C, F = p.F, HW = p.HW, nw = p.nw;
  uint32_t* kv = reinterpret_cast<uint32_t*>(tm_smem);                  // [T][PAIRS][32] packed bf16 pairs
  float* rv = tm_smem + (size_t)T * PAIRS * 32 + (size_t)w * T * FC;   // this warp's [T][FC] slice of Rv
  float* sbuf = tm_smem + (size_t)T * PAIRS * 32 + (size_t)nw * T * FC;  // [nw][T][32]: R-term scores
  const int px0 = blockIdx.x * 32, h = blockIdx.y;
  const int b = blockIdx.z / p.tgroups, tg = blockIdx.z - b * p.tgroups;
  const int px = min(px0 + lane, HW - 1);
  const bool px_ok = px0 + lane < HW;
  const size_t tok = (size_t)3 * C;
  const float* maskb = p.mask ? p.mask + (size_t)b * T : nullptr;
  const int t = tg * nw + w;
  const bool act = t < T;
  const int tt = act ? t : 0;
  const int NCH = (F + FC - 1) / FC;

  // K (sel = C) or V (sel = 2C) chunk -> kv tile: 16-byte global loads (8 head dims), whole sectors per (frame, pixel)
  auto stage = [&](int sel, int f0, int fc) {
    const int groups = fc / 8;
    for (int i = threadIdx.x; i < T * 32 * groups; i += nthreads) {
      const int g8 = i % groups, pl = (i / groups) % 32, s = i / (32 * groups);
      const int pp = min(px0 + pl, HW - 1);
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.qkv + ((size_t)(b * T + s) * HW + pp) * tok + sel + h * F + f0 + g8 * 8));
      uint32_t* d = kv + ((size_t)s * PAIRS + g8 * 4) * 32 + pl;
      d[0] = u.x; d[32] = u.y; d[64] = u.z; d[96] = u.w;
    }
  };

  stage(C, 0, min(FC, F));
  // ---------------- D1[px, s] = Q[px, t, :] . Rk[t, s, :]   (this warp's query frame; operands straight from global/L2)
  if (act) {
    int prow[2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int r = 0; r < 2; ++r) prow[m][r] = min(px0 + m * 16 + gid + r * 8, HW - 1);
    float acc[2][TN][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < TN; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
    const __nv_bfloat16* qb = p.qkv + (size_t)(b * T + tt) * HW * tok + h * F;
    const __nv_bfloat16* rkb = p.Rk_op + ((size_t)(b * T + tt) * T) * C + h * F;
    for (int k0 = 0; k0 < F; k0 += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const __nv_bfloat16* r0 = qb + (size_t)prow[m][0] * tok + k0 + tig * 2;
        const __nv_bfloat16* r1 = qb + (size_t)prow[m][1] * tok + k0 + tig * 2;
        a[m][0] = ldg_u32(r0); a[m][1] = ldg_u32(r1); a[m][2] = ldg_u32(r0 + 8); a[m][3] = ldg_u32(r1 + 8);
      }
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        const int s = n * 8 + gid;
        uint32_t bb[2] = {0u, 0u};
        if (s < T) {
          const __nv_bfloat16* rr = rkb + (size_t)s * C + k0 + tig * 2;
          bb[0] = ldg_u32(rr); bb[1] = ldg_u32(rr + 8);
        }
        mma_bf16_16816(acc[0][n], a[0], bb);
        mma_bf16_16816(acc[1][n], a[1], bb);
      }
    }
    float* sb = sbuf + (size_t)w * T * 32;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < TN; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int s = n * 8 + tig * 2 + (i & 1), pl = m * 16 + gid + (i >> 1) * 8;
          if (s < T) sb[s * 32 + pl] = acc[m][n][i];
        }
  }

  // ---------------- K chunks: D2 partial (tensor cores, A fragments from the staged tile) + q.k partial (CUDA cores)
  float S[TP];
#pragma unroll
  for (int s = 0; s < TP; ++s) S[s] = 0.f;
  for (int c = 0; c < NCH; ++c) {
    const int f0 = c * FC, fc = min(FC, F - f0);
    if (c > 0) {
      __syncthreads();  // everyone is done with the previous chunk's tile
      stage(C, f0, fc);
    }
    __syncthreads();    // tile staged (and, for c == 0, every D1 row of sbuf written)
    // D2[px, t'] += K[px, s, chunk] . Rq[s, t', chunk]   for the key frames s = w, w + nw, ... of this warp
    for (int s = w; s < T; s += nw) {
      float acc[2][QN][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < QN; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
      const __nv_bfloat16* rqb = p.Rq_op + ((size_t)(b * T + s) * T) * C + h * F + f0;
      const uint32_t* kt = kv + (size_t)s * PAIRS * 32;
      for (int k0 = 0; k0 < fc; k0 += 16) {
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const uint32_t* r = kt + (size_t)(k0 / 2 + tig) * 32 + m * 16 + gid;
          a[m][0] = r[0]; a[m][1] = r[8]; a[m][2] = r[4 * 32]; a[m][3] = r[4 * 32 + 8];
        }
#pragma unroll
        for (int n = 0; n < QN; ++n) {
          const int tl = n * 8 + gid, tq = tg * nw + tl;
          uint32_t bb[2] = {0u, 0u};
          if (tl < nw && tq < T) {
            const __nv_bfloat16* rr = rqb + (size_t)tq * C + k0 + tig * 2;
            bb[0] = ldg_u32(rr); bb[1] = ldg_u32(rr + 8);
          }
          mma_bf16_16816(acc[0][n], a[0], bb);
          mma_bf16_16816(acc[1][n], a[1], bb);
        }
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < QN; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int tl = n * 8 + tig * 2 + (i & 1), pl = m * 16 + gid + (i >> 1) * 8;
            if (tl < nw && tg * nw + tl < T) sbuf[((size_t)tl * T + s) * 32 + pl] += acc[m][n][i];  // one writer per entry
          }
    }
    // q.k partial
    if (act) {
      const __nv_bfloat16* qrow = p.qkv + ((size_t)(b * T + tt) * HW + px) * tok + h * F + f0;
      float q[FC];
#pragma unroll
      for (int f = 0; f < FC; f += 4) {
        if (f < fc) {
          const float4 v = OpType<__nv_bfloat16>::load4(qrow + f);
          q[f] = v.x; q[f + 1] = v.y; q[f + 2] = v.z; q[f + 3] = v.w;
        } else {
          q[f] = q[f + 1] = q[f + 2] = q[f + 3] = 0.f;
        }
      }
#pragma unroll
      for (int s = 0; s < TP; ++s) {
        if (s < T) {
          const uint32_t* wv = kv + (size_t)s * PAIRS * 32 + lane;
          float acc0 = S[s], acc1 = 0.f;
#pragma unroll
          for (int fp = 0; fp < PAIRS; ++fp) {
            if (2 * fp < fc) {
              const uint32_t u = wv[fp * 32];
              acc0 = fmaf(q[2 * fp], __uint_as_float(u << 16), acc0);
              acc1 = fmaf(q[2 * fp + 1], __uint_as_float(u & 0xffff0000u), acc1);
            }
          }
          S[s] = acc0 + acc1;
        }
      }
    }
  }
  __syncthreads();  // all D2 contributions are in sbuf
  // ---------------- masked softmax (fp32)
  if (act) {
    const bool gt = maskb ? maskb[t] > 0.5f : true;
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < TP; ++s) {
      if (s < T) {
        const bool ok = maskb ? ((maskb[s] > 0.5f) == gt) : true;
        S[s] = ok ? (S[s] + sbuf[((size_t)w * T + s) * 32 + lane]) * p.scale : -INFINITY;
        mx = fmaxf(mx, S[s]);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < TP; ++s) {
      if (s < T) {
        S[s] = expf(S[s] - mx);
        sum += S[s];
      }
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int s = 0; s < TP; ++s)
      if (s < T) S[s] *= inv;
  }
  // ---------------- output: O[t] = sum_s P[t,s] * (v_s + Rv[t,s])
  for (int c = 0; c < NCH; ++c) {
    const int f0 = c * FC, fc = min(FC, F - f0);
    __syncthreads();
    stage(2 * C, f0, fc);
    for (int i = lane; i < T * (fc / 4); i += 32) {
      const int s = i / (fc / 4), fq = i - s * (fc / 4);
      *reinterpret_cast<float4*>(rv + s * FC + fq * 4) =
          __ldg(reinterpret_cast<const float4*>(p.Rv + (((size_t)(b * T + tt) * T + s) * C + h * F + f0 + fq * 4)));
    }
    __syncthreads();
    if (act) {
      float o[FC];
#pragma unroll
      for (int f = 0; f < FC; ++f) o[f] = 0.f;
#pragma unroll
      for (int s = 0; s < TP; ++s) {
        if (s < T) {
          const uint32_t* wv = kv + (size_t)s * PAIRS * 32 + lane;
          const float pr = S[s];
#pragma unroll
          for (int fq = 0; fq < FC / 4; ++fq) {
            if (fq * 4 < fc) {
              const float4 r4 = *reinterpret_cast<const float4*>(rv + s * FC + fq * 4);
              const uint32_t u0 = wv[(2 * fq) * 32], u1 = wv[(2 * fq + 1) * 32];
              o[4 * fq] = fmaf(pr, __uint_as_float(u0 << 16) + r4.x, o[4 * fq]);
              o[4 * fq + 1] = fmaf(pr, __uint_as_float(u0 & 0xffff0000u) + r4.y, o[4 * fq + 1]);
              o[4 * fq + 2] = fmaf(pr, __uint_as_float(u1 << 16) + r4.z, o[4 * fq + 2]);
              o[4 * fq + 3] = fmaf(pr, __uint_as_float(u1 & 0xffff0000u) + r4.w, o[4 * fq + 3]);
            }
          }
        }
      }
      if (px_ok) {
        __nv_bfloat16* orow = p.out + ((size_t)(b * T + t) * HW + px) * C + h * F + f0;
#pragma unroll
        for (int f = 0; f < FC; f += 4)
          if (f < fc) OpType<__nv_bfloat16>::store4(orow + f, make_float4(o[f], o[f + 1], o[f + 2], o[f + 3]));
      }
    }
  }
}

template <int TP, int FC>
static int launch_tm(TMParams& p, cudaStream_t st) {
  const int base_blocks = ((p.HW + 31) / 32) * p.heads * p.B;
  int groups = (296 + base_blocks - 1) / base_blocks;
  const int min_groups = (p.T + TM_MAX_WARPS - 1) / TM_MAX_WARPS, max_groups = (p.T + 1) / 2;
  if (groups < min_groups) groups = min_groups;
  if (groups > max_groups) groups = max_groups;
  if (groups < 1) groups = 1;
  p.nw = (p.T + groups - 1) / groups;
  p.tgroups = (p.T + p.nw - 1) / p.nw;
  dim3 grid((p.HW + 31) / 32, p.heads, p.B * p.tgroups);
  const size_t smem = ((size_t)p.T * (FC / 2) * 32 + (size_t)p.nw * p.T * FC + (size_t)p.nw * p.T * 32) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_temporal_mma_kernel<TP, FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fdm::launch(attn_temporal_mma_kernel<TP, FC>, grid, dim3(p.nw * 32), smem, st, p);
  return check_launch();
}

// bf16 qkv / out with bf16 copies of the Rq, Rk tables; FDM_ERR_UNSUPPORTED -> the caller uses the CUDA-core kernel
int attn_temporal_mma_launch(const fdm_attn_temporal_args* a, cudaStream_t st) {
  FDM_REQUIRE(a->qkv_dtype == FDM_BF16 && a->out_dtype == FDM_BF16 && a->Rq_op != nullptr && a->Rk_op != nullptr, FDM_ERR_UNSUPPORTED);
  const int F = a->C / a->heads;
  FDM_REQUIRE(F % 16 == 0 && a->T <= 40 && a->C % 8 == 0, FDM_ERR_UNSUPPORTED);
  TMParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(a->qkv);
  p.Rq_op = reinterpret_cast<const __nv_bfloat16*>(a->Rq_op);
  p.Rk_op = reinterpret_cast<const __nv_bfloat16*>(a->Rk_op);
  p.Rv = a->Rv; p.mask = a->mask; p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.B = a->B; p.T = a->T; p.HW = a->HW; p.C = a->C; p.heads = a->heads; p.F = F;
  p.scale = 1.0f / sqrtf((float)F);
  // 32-wide head-dim chunks while the key tile stays small, 16-wide for long clips (shared memory per block)
  if (p.T <= 8) return launch_tm<8, 32>(p, st);
  if (p.T <= 16) return launch_tm<16, 32>(p, st);
  if (p.T <= 24) return launch_tm<24, 32>(p, st);
  if (p.T <= 32) return launch_tm<32, 16>(p, st);
  return launch_tm<40, 16>(p, st);
}

}  // namespace fdm
