// Attention cores on CUDA cores (fp32 math, fp32 softmax), used by both precision modes:
//   * temporal attention with the contextual relative-position terms and the two-group mask computed
//     in-kernel (rpe.py:139-170): sequences are T <= 40 frames, far below the tcgen05 tile minimum
//     (M = 64/128), so the QK^T part runs on CUDA cores (SURVEY §7 "hard parts").
//   * spatial attention (rpe.py:139-144,163-166 with no RPE / mask): flash-style streaming softmax.
#include "common.cuh"

namespace fdm {

// =====================================================================================================
// temporal attention
//   grid (ceil(HW/32), heads, B); block 256 = 8 warps; lane <-> pixel, warp <-> query frame t
//   S[t,s] = scale*( q_t.k_s + q_t.Rk[t,s] + k_s.Rq[s,t] );  P = softmax over {s : mask_s == mask_t}
//   O[t]   = sum_s P[t,s] * (v_s + Rv[t,s])
// =====================================================================================================
constexpr int TA_FC = 8;   // head-dim chunk staged in shared memory
constexpr int TA_NW = 8;   // warps = query frames handled by one block

struct TAParams {
  const void* qkv; const float* Rq; const float* Rk; const float* Rv; const float* mask; void* out;
  int B, T, HW, C, heads, F, tgroups;
  float scale;
};

// grid (ceil(HW/32), heads, B * tgroups); block = TA_NW warps.  Warp w owns query frame t = tg*TA_NW + w for 32 pixels
// (lane <-> pixel).  Per head-dim chunk of 8: K (then V) of ALL T key frames is staged once per block as [s][f][px]
// (conflict-free per-lane reads); every warp stages ITS rows Rk[t,:,chunk], Rq[:,t,chunk] (then Rv[t,:,chunk]) into a
// private shared-memory slice with lane-parallel loads (T independent loads in flight instead of T dependent broadcast
// loads from L2 in the inner loop), and reads them back as float4 broadcasts.
template <int TP, typename QT, typename OT>
__global__ void __launch_bounds__(TA_NW * 32) attn_temporal_kernel(TAParams p) {
  extern __shared__ __align__(16) float ta_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int T = p.T, C = p.C, F = p.F, HW = p.HW;
  float* kv = ta_smem;                                  // [T][TA_FC][32]
  float* ra = ta_smem + (size_t)T * TA_FC * 32 + (size_t)w * 2 * T * TA_FC;  // this warp's [T][TA_FC]: Rk (scores) / Rv (output)
  float* rb = ra + (size_t)T * TA_FC;                   // this warp's [T][TA_FC]: Rq
  const int px0 = blockIdx.x * 32, h = blockIdx.y;
  const int b = blockIdx.z / p.tgroups, tg = blockIdx.z - b * p.tgroups;
  const int px = min(px0 + lane, HW - 1);
  const bool px_ok = px0 + lane < HW;
  const QT* qkv = reinterpret_cast<const QT*>(p.qkv);
  const size_t tok_stride = (size_t)3 * C;  // per (frame, pixel)
  const float* maskb = p.mask ? p.mask + (size_t)b * T : nullptr;
  const int t = tg * TA_NW + w;
  const bool act = t < T;
  const int tt = act ? t : 0;

  float S[TP];
#pragma unroll
  for (int s = 0; s < TP; ++s) S[s] = 0.f;
  const QT* qrow = qkv + ((size_t)(b * T + tt) * HW + px) * tok_stride + h * F;
  // ---------------- scores
  for (int f0 = 0; f0 < F; f0 += TA_FC) {
    __syncthreads();
    for (int i = threadIdx.x; i < T * 32 * (TA_FC / 4); i += TA_NW * 32) {
      int fq = i % (TA_FC / 4);
      int pl = (i / (TA_FC / 4)) % 32;
      int s = i / (32 * (TA_FC / 4));
      int pp = min(px0 + pl, HW - 1);
      float4 v = OpType<QT>::load4(qkv + ((size_t)(b * T + s) * HW + pp) * tok_stride + C + h * F + f0 + fq * 4);
      float* d = kv + ((size_t)s * TA_FC + fq * 4) * 32 + pl;
      d[0] = v.x; d[32] = v.y; d[64] = v.z; d[96] = v.w;
    }
    for (int i = lane; i < T * TA_FC; i += 32) {
      int s = i / TA_FC, f = i - s * TA_FC;
      ra[i] = __ldg(p.Rk + (((size_t)(b * T + tt) * T + s) * C + h * F + f0 + f));
      rb[i] = __ldg(p.Rq + (((size_t)(b * T + s) * T + tt) * C + h * F + f0 + f));
    }
    float q[TA_FC];
    {
      float4 a = OpType<QT>::load4(qrow + f0), c = OpType<QT>::load4(qrow + f0 + 4);
      q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = c.x; q[5] = c.y; q[6] = c.z; q[7] = c.w;
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int s = 0; s < TP; ++s) {
        if (s < T) {
          const float4 rk0 = *reinterpret_cast<const float4*>(ra + s * TA_FC), rk1 = *reinterpret_cast<const float4*>(ra + s * TA_FC + 4);
          const float4 rq0 = *reinterpret_cast<const float4*>(rb + s * TA_FC), rq1 = *reinterpret_cast<const float4*>(rb + s * TA_FC + 4);
          const float rkv[8] = {rk0.x, rk0.y, rk0.z, rk0.w, rk1.x, rk1.y, rk1.z, rk1.w};
          const float rqv[8] = {rq0.x, rq0.y, rq0.z, rq0.w, rq1.x, rq1.y, rq1.z, rq1.w};
          float acc = S[s];
#pragma unroll
          for (int f = 0; f < TA_FC; ++f) {
            const float kk = kv[((size_t)s * TA_FC + f) * 32 + lane];
            acc = fmaf(q[f], kk + rkv[f], acc);
            acc = fmaf(kk, rqv[f], acc);
          }
          S[s] = acc;
        }
      }
    }
  }
  // ---------------- masked softmax (fp32)
  if (act) {
    const bool gt = maskb ? maskb[t] > 0.5f : true;
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < TP; ++s) {
      if (s < T) {
        bool ok = maskb ? ((maskb[s] > 0.5f) == gt) : true;
        S[s] = ok ? S[s] * p.scale : -INFINITY;
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
  // ---------------- output
  OT* orow = reinterpret_cast<OT*>(p.out) + ((size_t)(b * T + tt) * HW + px) * C + h * F;
  for (int f0 = 0; f0 < F; f0 += TA_FC) {
    __syncthreads();
    for (int i = threadIdx.x; i < T * 32 * (TA_FC / 4); i += TA_NW * 32) {
      int fq = i % (TA_FC / 4);
      int pl = (i / (TA_FC / 4)) % 32;
      int s = i / (32 * (TA_FC / 4));
      int pp = min(px0 + pl, HW - 1);
      float4 v = OpType<QT>::load4(qkv + ((size_t)(b * T + s) * HW + pp) * tok_stride + 2 * C + h * F + f0 + fq * 4);
      float* d = kv + ((size_t)s * TA_FC + fq * 4) * 32 + pl;
      d[0] = v.x; d[32] = v.y; d[64] = v.z; d[96] = v.w;
    }
    for (int i = lane; i < T * TA_FC; i += 32) {
      int s = i / TA_FC, f = i - s * TA_FC;
      ra[i] = __ldg(p.Rv + (((size_t)(b * T + tt) * T + s) * C + h * F + f0 + f));
    }
    __syncthreads();
    if (act) {
      float o[TA_FC];
#pragma unroll
      for (int f = 0; f < TA_FC; ++f) o[f] = 0.f;
#pragma unroll
      for (int s = 0; s < TP; ++s) {
        if (s < T) {
          const float4 r0 = *reinterpret_cast<const float4*>(ra + s * TA_FC), r1 = *reinterpret_cast<const float4*>(ra + s * TA_FC + 4);
          const float rvv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          const float pr = S[s];
#pragma unroll
          for (int f = 0; f < TA_FC; ++f) o[f] = fmaf(pr, kv[((size_t)s * TA_FC + f) * 32 + lane] + rvv[f], o[f]);
        }
      }
      if (px_ok) {
        OpType<OT>::store4(orow + f0, make_float4(o[0], o[1], o[2], o[3]));
        OpType<OT>::store4(orow + f0 + 4, make_float4(o[4], o[5], o[6], o[7]));
      }
    }
  }
}

template <typename QT, typename OT>
static int launch_temporal(TAParams& p, cudaStream_t st) {
  p.tgroups = (p.T + TA_NW - 1) / TA_NW;
  dim3 grid((p.HW + 31) / 32, p.heads, p.B * p.tgroups);
  size_t smem = ((size_t)p.T * TA_FC * 32 + (size_t)TA_NW * 2 * p.T * TA_FC) * sizeof(float);
#define FDM_TA_LAUNCH(TPV)                                                                                              \
  do {                                                                                                                  \
    if (smem > 48 * 1024)                                                                                               \
      cudaFuncSetAttribute(attn_temporal_kernel<TPV, QT, OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    attn_temporal_kernel<TPV, QT, OT><<<grid, TA_NW * 32, smem, st>>>(p);                                               \
  } while (0)
  if (p.T <= 8) FDM_TA_LAUNCH(8);
  else if (p.T <= 16) FDM_TA_LAUNCH(16);
  else if (p.T <= 24) FDM_TA_LAUNCH(24);
  else if (p.T <= 32) FDM_TA_LAUNCH(32);
  else if (p.T <= 40) FDM_TA_LAUNCH(40);
  else return FDM_ERR_UNSUPPORTED;
  return check_launch();
}

// =====================================================================================================
// spatial attention: grid (ceil(L/64), heads, N); block 256 = 64 queries x 4 lanes (each lane owns F/4 dims)
// =====================================================================================================
constexpr int SA_KC = 32;  // keys per shared-memory chunk
constexpr int SA_QB = 64;  // queries per block

struct SAParams {
  const void* qkv; void* out;
  int N, L, C, heads, F;
  float scale;
};

template <int FQ, typename QT, typename OT>
__global__ void __launch_bounds__(256) attn_spatial_kernel(SAParams p) {
  constexpr int V = (FQ % 4 == 0) ? 4 : 2;   // vector width of shared-memory accesses (F = 24 -> FQ = 6 -> float2)
  constexpr int SUB = FQ + V;                // padded sub-vector stride: lane `sub` starts at sub*SUB
  constexpr int ROW = 4 * SUB;
  __shared__ __align__(16) float ks[SA_KC * ROW];
  __shared__ __align__(16) float vs[SA_KC * ROW];
  const int tid = threadIdx.x, sub = tid & 3, ql = tid >> 2;
  const int n = blockIdx.z, h = blockIdx.y, C = p.C, L = p.L;
  constexpr int F = FQ * 4;
  const int qi = blockIdx.x * SA_QB + ql;
  const bool q_ok = qi < L;
  const QT* base = reinterpret_cast<const QT*>(p.qkv) + (size_t)n * L * 3 * C + h * F;
  float q[FQ], o[FQ];
  {
    const QT* qp = base + (size_t)(q_ok ? qi : 0) * 3 * C + sub * FQ;
#pragma unroll
    for (int f = 0; f < FQ; ++f) {
      q[f] = OpType<QT>::load(qp + f) * p.scale;
      o[f] = 0.f;
    }
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int j0 = 0; j0 < L; j0 += SA_KC) {
    __syncthreads();
    for (int i = tid; i < SA_KC * (F / V); i += 256) {
      int j = i / (F / V), f = (i - j * (F / V)) * V;
      int jj = min(j0 + j, L - 1);
      const QT* kp = base + (size_t)jj * 3 * C + C + f;
      int off = j * ROW + (f / FQ) * SUB + (f % FQ);
      if constexpr (V == 4) {
        *reinterpret_cast<float4*>(&ks[off]) = OpType<QT>::load4(kp);
        *reinterpret_cast<float4*>(&vs[off]) = OpType<QT>::load4(kp + C);
      } else {
        *reinterpret_cast<float2*>(&ks[off]) = make_float2(OpType<QT>::load(kp), OpType<QT>::load(kp + 1));
        *reinterpret_cast<float2*>(&vs[off]) = make_float2(OpType<QT>::load(kp + C), OpType<QT>::load(kp + C + 1));
      }
    }
    __syncthreads();
    float s[SA_KC];
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < SA_KC; ++j) {
      const float* kr = &ks[j * ROW + sub * SUB];
      float acc = 0.f;
      if constexpr (V == 4) {
#pragma unroll
        for (int f = 0; f < FQ; f += 4) {
          float4 kk = *reinterpret_cast<const float4*>(kr + f);
          acc = fmaf(q[f], kk.x, acc); acc = fmaf(q[f + 1], kk.y, acc);
          acc = fmaf(q[f + 2], kk.z, acc); acc = fmaf(q[f + 3], kk.w, acc);
        }
      } else {
#pragma unroll
        for (int f = 0; f < FQ; f += 2) {
          float2 kk = *reinterpret_cast<const float2*>(kr + f);
          acc = fmaf(q[f], kk.x, acc); acc = fmaf(q[f + 1], kk.y, acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      s[j] = (j0 + j < L) ? acc : -INFINITY;
      cmax = fmaxf(cmax, s[j]);
    }
    const float mnew = fmaxf(mrun, cmax);
    const float corr = expf(mrun - mnew);  // exp(-inf) = 0 on the first chunk
    lrun *= corr;
#pragma unroll
    for (int f = 0; f < FQ; ++f) o[f] *= corr;
#pragma unroll
    for (int j = 0; j < SA_KC; ++j) {
      const float pj = expf(s[j] - mnew);
      lrun += pj;
      const float* vr = &vs[j * ROW + sub * SUB];
      if constexpr (V == 4) {
#pragma unroll
        for (int f = 0; f < FQ; f += 4) {
          float4 vv = *reinterpret_cast<const float4*>(vr + f);
          o[f] = fmaf(pj, vv.x, o[f]); o[f + 1] = fmaf(pj, vv.y, o[f + 1]);
          o[f + 2] = fmaf(pj, vv.z, o[f + 2]); o[f + 3] = fmaf(pj, vv.w, o[f + 3]);
        }
      } else {
#pragma unroll
        for (int f = 0; f < FQ; f += 2) {
          float2 vv = *reinterpret_cast<const float2*>(vr + f);
          o[f] = fmaf(pj, vv.x, o[f]); o[f + 1] = fmaf(pj, vv.y, o[f + 1]);
        }
      }
    }
    mrun = mnew;
  }
  if (q_ok) {
    const float inv = 1.f / lrun;
    OT* op = reinterpret_cast<OT*>(p.out) + ((size_t)n * L + qi) * C + h * F + sub * FQ;
#pragma unroll
    for (int f = 0; f < FQ; ++f) OpType<OT>::store(op + f, o[f] * inv);
  }
}

template <typename QT, typename OT>
static int launch_spatial(const SAParams& p, cudaStream_t st) {
  dim3 grid((p.L + SA_QB - 1) / SA_QB, p.heads, p.N);
  switch (p.F) {
    case 8: attn_spatial_kernel<2, QT, OT><<<grid, 256, 0, st>>>(p); break;
    case 16: attn_spatial_kernel<4, QT, OT><<<grid, 256, 0, st>>>(p); break;
    case 24: attn_spatial_kernel<6, QT, OT><<<grid, 256, 0, st>>>(p); break;
    case 32: attn_spatial_kernel<8, QT, OT><<<grid, 256, 0, st>>>(p); break;
    case 48: attn_spatial_kernel<12, QT, OT><<<grid, 256, 0, st>>>(p); break;
    case 64: attn_spatial_kernel<16, QT, OT><<<grid, 256, 0, st>>>(p); break;
    case 96: attn_spatial_kernel<24, QT, OT><<<grid, 256, 0, st>>>(p); break;
    case 128: attn_spatial_kernel<32, QT, OT><<<grid, 256, 0, st>>>(p); break;
    default: return FDM_ERR_UNSUPPORTED;
  }
  return check_launch();
}

}  // namespace fdm

using namespace fdm;

extern "C" int fdm_attn_temporal(const fdm_attn_temporal_args* a, void* stream) {
  FDM_REQUIRE(a && a->qkv && a->Rq && a->Rk && a->Rv && a->out, FDM_ERR_BAD_ARG);
  FDM_REQUIRE(a->B > 0 && a->T > 0 && a->HW > 0 && a->heads > 0 && a->C % a->heads == 0, FDM_ERR_BAD_ARG);
  const int F = a->C / a->heads;
  FDM_REQUIRE(F % TA_FC == 0 && a->T <= 40, FDM_ERR_UNSUPPORTED);
  TAParams p{a->qkv, a->Rq, a->Rk, a->Rv, a->mask, a->out, a->B, a->T, a->HW, a->C, a->heads, F, 1, 1.0f / sqrtf((float)F)};
  cudaStream_t st = (cudaStream_t)stream;
  const bool qb = a->qkv_dtype == FDM_BF16, ob = a->out_dtype == FDM_BF16;
  if (!qb && !ob) return launch_temporal<float, float>(p, st);
  if (!qb && ob) return launch_temporal<float, __nv_bfloat16>(p, st);
  if (qb && !ob) return launch_temporal<__nv_bfloat16, float>(p, st);
  return launch_temporal<__nv_bfloat16, __nv_bfloat16>(p, st);
}

namespace fdm {
int attn_spatial_tc_launch(const fdm_attn_spatial_args* a, cudaStream_t st);  // attn_tc.cu
}

extern "C" int fdm_attn_spatial(const fdm_attn_spatial_args* a, void* stream) {
  FDM_REQUIRE(a && a->qkv && a->out, FDM_ERR_BAD_ARG);
  FDM_REQUIRE(a->N > 0 && a->L > 0 && a->heads > 0 && a->C % a->heads == 0, FDM_ERR_BAD_ARG);
  if (a->engine == 0 && a->qkv_dtype == FDM_BF16 && a->out_dtype == FDM_BF16) {
    int rc = attn_spatial_tc_launch(a, (cudaStream_t)stream);
    if (rc != FDM_ERR_UNSUPPORTED) return rc;
  }
  const int F = a->C / a->heads;
  SAParams p{a->qkv, a->out, a->N, a->L, a->C, a->heads, F, 1.0f / sqrtf((float)F)};
  cudaStream_t st = (cudaStream_t)stream;
  const bool qb = a->qkv_dtype == FDM_BF16, ob = a->out_dtype == FDM_BF16;
  if (!qb && !ob) return launch_spatial<float, float>(p, st);
  if (!qb && ob) return launch_spatial<float, __nv_bfloat16>(p, st);
  if (qb && !ob) return launch_spatial<__nv_bfloat16, float>(p, st);
  return launch_spatial<__nv_bfloat16, __nv_bfloat16>(p, st);
}
