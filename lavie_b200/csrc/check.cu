// fp32-accumulate CHECK MODE (BASELINE north star: noise-prediction rel-L2 <= 1e-3 against the reference's fp32 forward).
//
// Representation: every activation is a "split-bf16 triple" [rows, 3C] = [hi | lo | hi] with hi = bf16(x),
// lo = bf16(x - hi), i.e. ~16 mantissa bits.  With weights packed as W' = [Wh | Wh | Wl] along K (Wh = bf16(W),
// Wl = bf16(W - Wh)) the UNCHANGED tcgen05 GEMM / implicit-GEMM conv mainloop computes
//      A' W'^T = Ah Wh + Al Wh + Ah Wl        (error ~2^-17, fp32 accumulation in TMEM)
// so the tensor-core kernels under test are the product's own.  Their accumulators leave through the split-K partial
// path as fp32 and the epilogue (bias, time bias, residual, GEGLU) runs here in fp32 (gemm.cu: check reduce).
// Everything that is not a GEMM in check mode lives in this file as plain fp32 SIMT kernels reading hi + lo:
// LayerNorm, the attention cores (one generic kernel: spatial, cross, SparseCausal, temporal with RoPE + bias, strided
// frame attention), the fp32-weight small-M linear, and the fp32 -> triple splitter.  (GroupNorm statistics / apply,
// conv_in and conv_out take the triple through flags of their regular kernels in norm.cu / misc.cu.)
#include "common.cuh"

namespace {

__device__ __forceinline__ float ld_hl(const __nv_bfloat16* p, int lo_off) {
  return __bfloat162float(p[0]) + __bfloat162float(p[lo_off]);
}
__device__ __forceinline__ void st_triple(__nv_bfloat16* p, int c, float v) {   // p -> hi element, c = C (block width)
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  p[0] = hi;
  p[c] = lo;
  p[2 * c] = hi;
}

__global__ void split3_kernel(const float* __restrict__ x, long long rows, int C, __nv_bfloat16* __restrict__ y) {
  pdl_prologue();
  const long long total = rows * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C;
    const int c = static_cast<int>(i - r * C);
    st_triple(y + r * 3 * C + c, C, x[i]);
  }
}

// LayerNorm over C channels of a triple row -> triple row; one warp per row, two-pass fp32 statistics.
__global__ void __launch_bounds__(256)
layernorm_check_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int ldy, int rows,
                       int C) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const __nv_bfloat16* src = x + static_cast<size_t>(warp) * ldx;
  constexpr int MAXP = 64;                       // C <= 2048
  float v[MAXP];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int c = lane + i * 32;
    v[i] = c < C ? ld_hl(src + c, C) : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / static_cast<float>(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int c = lane + i * 32;
    if (c < C) {
      const float d = v[i] - mean;
      sq += d * d;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / static_cast<float>(C) + eps);
  __nv_bfloat16* dst = y + static_cast<size_t>(warp) * ldy;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int c = lane + i * 32;
    if (c < C) st_triple(dst + c, C, (v[i] - mean) * rstd * gamma[c] + beta[c]);
  }
}

// out[m, n] = act_out( sum_k act_in(x[m,k]) * w[n,k] + bias[n] ) with fp32 weights; one warp per output feature
__global__ void __launch_bounds__(256)
linear_smallm_f32_kernel(const float* __restrict__ x, int M, int K, const float* __restrict__ w,
                         const float* __restrict__ bias, float* __restrict__ out, int N, int silu_in, int silu_out) {
  pdl_prologue();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  for (int m = 0; m < M; ++m) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
      float xv = x[static_cast<size_t>(m) * K + k];
      if (silu_in) xv = xv / (1.0f + expf(-xv));
      acc = fmaf(xv, w[static_cast<size_t>(n) * K + k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float r = acc + (bias ? bias[n] : 0.f);
      if (silu_out) r = r / (1.0f + expf(-r));
      out[static_cast<size_t>(m) * N + n] = r;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic fp32 attention core for check mode.  Block = 128 threads = 32 queries x 4 key lanes of one (batch, head);
// K / V stream through shared memory in 32-key tiles (fp32, hi + lo summed on load), online softmax, exact expf.
// ---------------------------------------------------------------------------------------------------------------
struct AttnCheckParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* o;
  long long q_seq, q_batch, kv_seq, kv_batch, o_seq, o_batch;   // element strides
  int lo_q, lo_kv, lo_o;       // column offset of the lo block (and width of the output hi block)
  int Sq, Sk, d, head_pitch, kv_batch_div, sc_frames;
  float scale;                 // q is multiplied by it BEFORE the rotary embedding (attention.py:640-646)
  const float* rope;           // [max(Sq,Sk), rot_pairs, 2] (cos, sin) or nullptr
  int rot_pairs;
  const float* bias;           // [heads, Sq, Sk] or nullptr
};

constexpr int CK_Q = 32, CK_K = 32, CK_MAXD = 160;

__global__ void __launch_bounds__(128)
attention_check_kernel(const AttnCheckParams p) {
  pdl_prologue();
  extern __shared__ float csm[];
  const int d = p.d, dp = d + 1;
  float* sQ = csm;                       // [CK_Q][dp]
  float* sK = sQ + CK_Q * dp;            // [CK_K][dp]
  float* sV = sK + CK_K * dp;            // [CK_K][dp]
  float* sP = sV + CK_K * dp;            // [CK_Q][CK_K + 1]
  const int tid = threadIdx.x;
  const int qi = tid >> 2, kl = tid & 3;
  const int q0 = blockIdx.x * CK_Q, head = blockIdx.y, batch = blockIdx.z;
  const int col0 = head * p.head_pitch;
  const int n_seg = p.sc_frames > 0 ? 2 : 1;
  const int sc_f = p.sc_frames > 0 ? batch % p.sc_frames : 0;
  // ---- Q tile: scale, rotate ----
  for (int i = tid; i < CK_Q * d; i += 128) {
    const int r = i / d, c = i - r * d;
    float val = 0.f;
    if (q0 + r < p.Sq)
      val = ld_hl(p.q + static_cast<size_t>(batch) * p.q_batch + static_cast<size_t>(q0 + r) * p.q_seq + col0 + c, p.lo_q) *
            p.scale;
    sQ[r * dp + c] = val;
  }
  __syncthreads();
  if (p.rope) {
    for (int i = tid; i < CK_Q * p.rot_pairs; i += 128) {
      const int r = i / p.rot_pairs, pr = i - r * p.rot_pairs;
      if (q0 + r < p.Sq) {
        const float cs = p.rope[(static_cast<size_t>(q0 + r) * p.rot_pairs + pr) * 2];
        const float sn = p.rope[(static_cast<size_t>(q0 + r) * p.rot_pairs + pr) * 2 + 1];
        const float x = sQ[r * dp + 2 * pr], y = sQ[r * dp + 2 * pr + 1];
        sQ[r * dp + 2 * pr] = x * cs - y * sn;
        sQ[r * dp + 2 * pr + 1] = y * cs + x * sn;
      }
    }
    __syncthreads();
  }
  float m_run = -INFINITY, l_run = 0.f;
  float acc[CK_MAXD / 4];
#pragma unroll
  for (int i = 0; i < CK_MAXD / 4; ++i) acc[i] = 0.f;

  for (int seg = 0; seg < n_seg; ++seg) {
    int kvb = batch / p.kv_batch_div;
    if (p.sc_frames > 0) kvb = seg == 0 ? batch - sc_f : (sc_f > 0 ? batch - 1 : batch);
    const __nv_bfloat16* kbase = p.k + static_cast<size_t>(kvb) * p.kv_batch + col0;
    const __nv_bfloat16* vbase = p.v + static_cast<size_t>(kvb) * p.kv_batch + col0;
    for (int k0 = 0; k0 < p.Sk; k0 += CK_K) {
      __syncthreads();                                     // previous tile fully consumed
      for (int i = tid; i < CK_K * d; i += 128) {
        const int r = i / d, c = i - r * d;
        float kv = 0.f, vv = 0.f;
        if (k0 + r < p.Sk) {
          kv = ld_hl(kbase + static_cast<size_t>(k0 + r) * p.kv_seq + c, p.lo_kv);
          vv = ld_hl(vbase + static_cast<size_t>(k0 + r) * p.kv_seq + c, p.lo_kv);
        }
        sK[r * dp + c] = kv;
        sV[r * dp + c] = vv;
      }
      __syncthreads();
      if (p.rope) {
        for (int i = tid; i < CK_K * p.rot_pairs; i += 128) {
          const int r = i / p.rot_pairs, pr = i - r * p.rot_pairs;
          if (k0 + r < p.Sk) {
            const float cs = p.rope[(static_cast<size_t>(k0 + r) * p.rot_pairs + pr) * 2];
            const float sn = p.rope[(static_cast<size_t>(k0 + r) * p.rot_pairs + pr) * 2 + 1];
            const float x = sK[r * dp + 2 * pr], y = sK[r * dp + 2 * pr + 1];
            sK[r * dp + 2 * pr] = x * cs - y * sn;
            sK[r * dp + 2 * pr + 1] = y * cs + x * sn;
          }
        }
        __syncthreads();
      }
      // ---- scores of this thread's 8 keys ----
      float s[CK_K / 4];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < CK_K / 4; ++j) {
        const int kj = kl + 4 * j;
        float dot = 0.f;
        for (int c = 0; c < d; ++c) dot = fmaf(sQ[qi * dp + c], sK[kj * dp + c], dot);
        if (p.bias && q0 + qi < p.Sq && k0 + kj < p.Sk)
          dot += p.bias[(static_cast<size_t>(head) * p.Sq + q0 + qi) * p.Sk + k0 + kj];
        s[j] = (k0 + kj < p.Sk) ? dot : -INFINITY;
        mx = fmaxf(mx, s[j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);                // finite: every tile has at least one valid key
      const float alpha = expf(m_run - m_new);             // exp(-inf) = 0 on the first tile
      float ls = 0.f;
#pragma unroll
      for (int j = 0; j < CK_K / 4; ++j) {
        const float e = expf(s[j] - m_new);
        sP[qi * (CK_K + 1) + kl + 4 * j] = e;
        ls += e;
      }
      ls += __shfl_xor_sync(0xffffffffu, ls, 1);
      ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      l_run = l_run * alpha + ls;
      m_run = m_new;
      __syncwarp();                                        // the 4 key lanes of a query sit in one warp
      // ---- O += P V for this thread's channels c = kl, kl + 4, ... ----
#pragma unroll
      for (int i = 0; i < CK_MAXD / 4; ++i) {
        const int c = kl + 4 * i;
        if (c < d) {
          float a = acc[i] * alpha;
          for (int kj = 0; kj < CK_K; ++kj) a = fmaf(sP[qi * (CK_K + 1) + kj], sV[kj * dp + c], a);
          acc[i] = a;
        }
      }
    }
  }
  if (q0 + qi < p.Sq) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = p.o + static_cast<size_t>(batch) * p.o_batch + static_cast<size_t>(q0 + qi) * p.o_seq + head * d;
#pragma unroll
    for (int i = 0; i < CK_MAXD / 4; ++i) {
      const int c = kl + 4 * i;
      if (c < d) st_triple(dst + c, p.lo_o, acc[i] * inv);
    }
  }
}

int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  if (b > 148LL * 32) b = 148LL * 32;
  return b < 1 ? 1 : static_cast<int>(b);
}

}  // namespace

extern "C" int lavie_check_split3(const float* x, long long rows, int C, void* y, cudaStream_t stream) {
  LAVIE_REQUIRE(rows > 0 && C > 0, LAVIE_ERR_SHAPE, "check_split3: empty input");
  launch_pdl(split3_kernel, grid_for(rows * C, 256), 256, 0, stream, x, rows, C, static_cast<__nv_bfloat16*>(y));
  return lavie_check_launch("split3_kernel");
}

extern "C" int lavie_check_layernorm(const void* x, int ldx, const float* gamma, const float* beta, float eps, void* y,
                                     int ldy, int rows, int C, cudaStream_t stream) {
  LAVIE_REQUIRE(C > 0 && C <= 2048 && ldx >= 2 * C && ldy >= 3 * C, LAVIE_ERR_SHAPE,
                "check_layernorm: C=%d (<= 2048), triples need ldx >= 2C, ldy >= 3C", C);
  if (rows <= 0) return LAVIE_OK;
  launch_pdl(layernorm_check_kernel, (rows + 7) / 8, 256, 0, stream, static_cast<const __nv_bfloat16*>(x), ldx, gamma,
             beta, eps, static_cast<__nv_bfloat16*>(y), ldy, rows, C);
  return lavie_check_launch("layernorm_check_kernel");
}

extern "C" int lavie_check_linear_smallm(const float* x, int M, int K, const float* w, const float* bias, float* out,
                                         int N, int silu_in, int silu_out, cudaStream_t stream) {
  LAVIE_REQUIRE(M >= 1 && K > 0 && N > 0, LAVIE_ERR_SHAPE, "check_linear_smallm: M=%d K=%d N=%d", M, K, N);
  launch_pdl(linear_smallm_f32_kernel, (N + 7) / 8, 256, 0, stream, x, M, K, w, bias, out, N, silu_in, silu_out);
  return lavie_check_launch("linear_smallm_f32_kernel");
}

extern "C" int lavie_check_attention(const void* q, long long q_seq_stride, long long q_batch_stride, int q_lo_offset,
                                     const void* k, const void* v, long long kv_seq_stride, long long kv_batch_stride,
                                     int kv_lo_offset, void* o, long long o_seq_stride, long long o_batch_stride,
                                     int o_lo_offset, int batch, int heads, int Sq, int Sk, int d, int head_pitch,
                                     int kv_batch_div, int sparse_causal_frames, float scale, const float* rope,
                                     int rot_pairs, const float* bias, cudaStream_t stream) {
  LAVIE_REQUIRE(batch > 0 && heads > 0 && Sq > 0 && Sk > 0 && kv_batch_div > 0 && batch % kv_batch_div == 0 &&
                    d > 0 && d <= CK_MAXD && batch <= 65535 && heads <= 65535,
                LAVIE_ERR_SHAPE, "check_attention: bad sizes batch=%d heads=%d Sq=%d Sk=%d d=%d", batch, heads, Sq, Sk, d);
  LAVIE_REQUIRE(sparse_causal_frames == 0 || (kv_batch_div == 1 && batch % sparse_causal_frames == 0), LAVIE_ERR_SHAPE,
                "check_attention: sparse-causal mode needs batch %% frames == 0");
  LAVIE_REQUIRE(rope == nullptr || 2 * rot_pairs <= d, LAVIE_ERR_SHAPE, "check_attention: rot_pairs");
  AttnCheckParams p;
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v); p.o = static_cast<__nv_bfloat16*>(o);
  p.q_seq = q_seq_stride; p.q_batch = q_batch_stride; p.kv_seq = kv_seq_stride; p.kv_batch = kv_batch_stride;
  p.o_seq = o_seq_stride; p.o_batch = o_batch_stride;
  p.lo_q = q_lo_offset; p.lo_kv = kv_lo_offset; p.lo_o = o_lo_offset;
  p.Sq = Sq; p.Sk = Sk; p.d = d; p.head_pitch = head_pitch; p.kv_batch_div = kv_batch_div;
  p.sc_frames = sparse_causal_frames; p.scale = scale; p.rope = rope; p.rot_pairs = rope ? rot_pairs : 0; p.bias = bias;
  const int smem = ((CK_Q + 2 * CK_K) * (d + 1) + CK_Q * (CK_K + 1)) * static_cast<int>(sizeof(float));
  static LavieSmemConfig configured;
  const int rc = lavie_config_smem(attention_check_kernel, smem, &configured, "attention_check_kernel");
  if (rc) return rc;
  dim3 grid((Sq + CK_Q - 1) / CK_Q, heads, batch);
  launch_pdl(attention_check_kernel, grid, 128, smem, stream, p);
  return lavie_check_launch("attention_check_kernel");
}
