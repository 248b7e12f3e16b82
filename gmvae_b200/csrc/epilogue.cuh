// epilogue.cuh -- GEMM epilogue functors shared by the tcgen05 (bf16) and SIMT (fp32) kernels.
//
// A GEMM kernel hands an epilogue one row fragment at a time:
//     epi.row<NV>(m, n0, acc, nvalid)   acc[0..NV) = C[m, n0..n0+NV), the first `nvalid` are in range
// and calls epi.finish_warp() once per warp (all 32 lanes converged) when the tile is done.
// The ELBO body lives here: bias+ReLU (base.py MLP layers), the Bernoulli log-likelihood and its
// gradient (gmvae.py:254 / vae.py:177), ReLU masks of the backward pass, atomic accumulation of
// split-K weight gradients.
#pragma once
#include "common.cuh"

namespace gmvae {

// ---- row-fragment stores -------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void store_frag(float* dst, const float* v, int nvalid) {
  if (nvalid == NV && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}
template <int NV>
__device__ __forceinline__ void store_frag(bf16* dst, const float* v, int nvalid) {
  if (NV % 8 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      uint4 p;
      p.x = pack_bf16x2(v[i], v[i + 1]); p.y = pack_bf16x2(v[i + 2], v[i + 3]);
      p.z = pack_bf16x2(v[i + 4], v[i + 5]); p.w = pack_bf16x2(v[i + 6], v[i + 7]);
      *reinterpret_cast<uint4*>(dst + i) = p;
    }
  } else if (nvalid == NV && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 2) *reinterpret_cast<uint32_t*>(dst + i) = pack_bf16x2(v[i], v[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) dst[i] = __float2bfloat16_rn(v[i]);
  }
}

// ---- row-fragment loads (converted to fp32) ------------------------------------------------
template <int NV>
__device__ __forceinline__ void load_frag(const float* src, float* v, int nvalid) {
  if (nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(src + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = i < nvalid ? src[i] : 0.f;
  }
}
template <int NV>
__device__ __forceinline__ void load_frag(const bf16* src, float* v, int nvalid) {
  if (NV % 8 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      uint4 t = *reinterpret_cast<const uint4*>(src + i);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(p[j]);
        v[i + 2 * j] = f.x; v[i + 2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = i < nvalid ? __bfloat162float(src[i]) : 0.f;
  }
}
template <int NV>
__device__ __forceinline__ void load_frag(const uint8_t* src, float* v, int nvalid) {
  if (NV % 16 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 16) {
      uint4 t = *reinterpret_cast<const uint4*>(src + i);
      const uint8_t* p = reinterpret_cast<const uint8_t*>(&t);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[i + j] = (float)p[j];
    }
  } else if (NV % 4 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      uint32_t t = *reinterpret_cast<const uint32_t*>(src + i);
      v[i] = (float)(t & 0xff); v[i + 1] = (float)((t >> 8) & 0xff);
      v[i + 2] = (float)((t >> 16) & 0xff); v[i + 3] = (float)(t >> 24);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = i < nvalid ? (float)src[i] : 0.f;
  }
}

// ---- out = act(acc + bias[n] + addend[m,n])  (+= when accumulate) -----------------------------
// Forward MLP layers (base.py:46-60: relu(h W + b), last layer linear) and plain stores.
template <typename OutT>
struct EpiStore {
  OutT* out; int64_t ld;
  const float* bias;        // [N] or null
  const float* addend;      // [M, ld_add] fp32 or null (the y-part of encoder_gmm layer 0 in fp32 mode)
  int64_t ld_add;
  int relu;
  int accumulate;           // out += value (fp32 outputs only; rows are owned by one thread)
  float scale;

  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid) {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = acc[i] * scale;
    if (bias) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i < nvalid) v[i] += __ldg(bias + n0 + i);
    }
    if (addend) {
      float a[NV];
      load_frag<NV>(addend + (int64_t)m * ld_add + n0, a, nvalid);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += a[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    OutT* dst = out + (int64_t)m * ld + n0;
    if (accumulate) {
      float o[NV];
      load_frag<NV>(dst, o, nvalid);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += o[i];
    }
    store_frag<NV>(dst, v, nvalid);
  }
  __device__ __forceinline__ void finish_warp() {}
};

// ---- decoder output layer fused with the Bernoulli log-likelihood ---------------------------
// logits = acc + b + gen_bias_init (base.py:135); log p(x|z) = sum_d [x l - softplus(l)]
// (gmvae.py:254, vae.py:177; TFP Bernoulli = -sigmoid_cross_entropy); d nll / d logits =
// (sigmoid(l) - x) / B.  Logits never reach HBM; only the gradient does.
template <typename OutT>
struct EpiBCE {
  OutT* dlogits; int64_t ld;
  const float* bias; float gen_bias;
  const uint8_t* x; int64_t ldx;
  int x_row_div;            // x row = m / x_row_div (objective M: K consecutive rows share one image)
  const float* row_weight;  // [M] q(y=k|x) weights (objective M) or null -> 1
  float* row_sum;           // [M] per-row log-likelihood (objective M) or null
  float* nll_acc;           // scalar accumulator: += -inv_bg * sum(w * loglik)
  float inv_bg;             // 1 / global batch
  float partial;

  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid) {
    float xv[NV], d[NV];
    load_frag<NV>(x + (int64_t)(m / x_row_div) * ldx + n0, xv, nvalid);
    float w = row_weight ? __ldg(row_weight + m) : 1.f;
    float ll = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float l = acc[i] + gen_bias;
      if (bias && i < nvalid) l += __ldg(bias + n0 + i);
      float e = __expf(-fabsf(l));
      float sp = fmaxf(l, 0.f) + log1pf(e);
      float inv1pe = 1.f / (1.f + e);
      float sg = l >= 0.f ? inv1pe : e * inv1pe;
      if (i < nvalid) ll += xv[i] * l - sp;
      d[i] = (sg - xv[i]) * (w * inv_bg);
    }
    store_frag<NV>(dlogits + (int64_t)m * ld + n0, d, nvalid);
    if (row_sum) atomicAdd(row_sum + m, ll);
    partial += w * ll;
  }
  __device__ __forceinline__ void finish_warp() {
    float s = warp_sum(partial);
    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(nll_acc, -inv_bg * s);
    partial = 0.f;
  }
};

// ---- backward through a ReLU layer: out = acc * [h > 0] -------------------------------------
template <typename OutT, typename HT>
struct EpiReluMask {
  OutT* out; int64_t ld;
  const HT* h; int64_t ldh;
  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid) {
    float hv[NV], v[NV];
    load_frag<NV>(h + (int64_t)m * ldh + n0, hv, nvalid);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = hv[i] > 0.f ? acc[i] : 0.f;
    store_frag<NV>(out + (int64_t)m * ld + n0, v, nvalid);
  }
  __device__ __forceinline__ void finish_warp() {}
};

// ---- split-K weight gradient: out[m,n] += acc (fp32 atomics into the flat gradient buffer) ----
struct EpiAtomicAdd {
  float* out; int64_t ld;
  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid) {
    float* dst = out + (int64_t)m * ld + n0;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) atomicAdd(dst + i, acc[i]);
  }
  __device__ __forceinline__ void finish_warp() {}
};

}  // namespace gmvae
