// Small / bandwidth-bound helpers of the denoiser step: time-embedding path, first and last convolution,
// patch matrix for strided convs, nearest upsampling, and the caller-side CFG + DDIM update.
#include "common.cuh"

namespace {

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------------------
// Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): [cos | sin]   (unet.py:153,428; diffusers 0.16)
// ---------------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B, int dim, float* __restrict__ out) {
  pdl_prologue();
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float freq = expf(-logf(10000.0f) * static_cast<float>(k) / static_cast<float>(half));
  const float arg = t[b] * freq;
  out[b * dim + k] = cosf(arg);
  out[b * dim + half + k] = sinf(arg);
}

// ---------------------------------------------------------------------------------------------------------
// out[m, n] = act_out( sum_k act_in(x[m,k]) * w[n,k] + bias[n] ),  m <= 8.  One warp per output feature.
// ---------------------------------------------------------------------------------------------------------
template <int MAXM>
__global__ void __launch_bounds__(256)
linear_smallm_kernel(const float* __restrict__ x, int M, int K, const __nv_bfloat16* __restrict__ w,
                     const float* __restrict__ bias, float* __restrict__ out, int N, int silu_in, int silu_out) {
  pdl_prologue();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[MAXM];
#pragma unroll
  for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
  const __nv_bfloat16* wr = w + static_cast<size_t>(n) * K;
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + k));
    const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
    float wf[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16(ww[e]);
      wf[2 * e] = f.x;
      wf[2 * e + 1] = f.y;
    }
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < M) {
        const float4 a = *reinterpret_cast<const float4*>(x + static_cast<size_t>(m) * K + k);
        const float4 b = *reinterpret_cast<const float4*>(x + static_cast<size_t>(m) * K + k + 4);
        float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = silu_in ? silu_f(xv[e]) : xv[e];
          acc[m] += v * wf[e];
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MAXM; ++m) {
    const float s = warp_sum(acc[m]);
    if (lane == 0 && m < M) {
      float r = s + (bias ? bias[n] : 0.f);
      if (silu_out) r = silu_f(r);
      out[static_cast<size_t>(m) * N + n] = r;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// conv_in: fp32 [B,Cin,F,H,W] -> bf16 channels-last [B*F*H*W, Cout]; thread = (pixel, 8 output channels)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ x, int B, int Cin, int F, int H, int W, const float* __restrict__ w,
               const float* __restrict__ bias, int Cout, __nv_bfloat16* __restrict__ out, int ldo) {
  pdl_prologue();
  extern __shared__ float s_w[];     // transposed to [Cin*9][Cout]: a warp reads 1 KiB contiguous per tap
  const int kk = Cin * 9;
  for (int i = threadIdx.x; i < Cout * kk; i += blockDim.x) {
    const int o = i / kk, r = i - o * kk;
    s_w[r * Cout + o] = w[i];
  }
  __syncthreads();
  const int groups = Cout >> 3;
  const long long total = static_cast<long long>(B) * F * H * W * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long pix = i / groups;
    const int xw = static_cast<int>(pix % W);
    const int yh = static_cast<int>((pix / W) % H);
    const int f = static_cast<int>((pix / (static_cast<long long>(W) * H)) % F);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H * F));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias[g * 8 + e];
    for (int c = 0; c < Cin; ++c) {
      const float* plane = x + ((static_cast<size_t>(b) * Cin + c) * F + f) * H * W;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = yh + t / 3 - 1, xx = xw + t % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float v = __ldg(plane + yy * W + xx);
          const float4* wp = reinterpret_cast<const float4*>(s_w + (c * 9 + t) * Cout + g * 8);
          const float4 w0 = wp[0], w1 = wp[1];
          acc[0] += v * w0.x; acc[1] += v * w0.y; acc[2] += v * w0.z; acc[3] += v * w0.w;
          acc[4] += v * w1.x; acc[5] += v * w1.y; acc[6] += v * w1.z; acc[7] += v * w1.w;
        }
      }
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]);
    o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]);
    o.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(pix) * ldo + g * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------------------
// conv_norm_out -> SiLU -> conv_out, fused: one warp per output pixel, lanes split the channel vectors.
// Zero padding applies to the ACTIVATED map, so out-of-image taps are simply skipped.
// ---------------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256)
conv_out_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ scale_shift, int B, int F,
                int H, int W, int C, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ out) {
  pdl_prologue();
  extern __shared__ float s_w[];     // [COUT][9][C]
  for (int i = threadIdx.x; i < COUT * 9 * C; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const long long total = static_cast<long long>(B) * F * H * W;
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long pix = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < total;
       pix += warps_total) {
    const int xw = static_cast<int>(pix % W);
    const int yh = static_cast<int>((pix / W) % H);
    const long long bf = pix / (static_cast<long long>(W) * H);
    const int b = static_cast<int>(bf / F);
    const int f = static_cast<int>(bf % F);
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float sc[8], sh[8];
      const float4* ss = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(b) * C + v * 8) * 2);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 p = __ldg(ss + e);
        sc[2 * e] = p.x; sh[2 * e] = p.y; sc[2 * e + 1] = p.z; sh[2 * e + 1] = p.w;
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = yh + t / 3 - 1, xx = xw + t % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const size_t row = (static_cast<size_t>(bf) * H + yy) * W + xx;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * ldx + v * 8));
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
        float a[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t2 = unpack_bf16(uw[e]);
          a[2 * e] = silu_f(t2.x * sc[2 * e] + sh[2 * e]);
          a[2 * e + 1] = silu_f(t2.y * sc[2 * e + 1] + sh[2 * e + 1]);
        }
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float4* wp = reinterpret_cast<const float4*>(s_w + (o * 9 + t) * C + v * 8);
          const float4 w0 = wp[0], w1 = wp[1];
          acc[o] += a[0] * w0.x + a[1] * w0.y + a[2] * w0.z + a[3] * w0.w + a[4] * w1.x + a[5] * w1.y + a[6] * w1.z +
                    a[7] * w1.w;
        }
      }
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      const float s = warp_sum(acc[o]);
      if (lane == 0)
        out[(((static_cast<size_t>(b) * COUT + o) * F + f) * H + yh) * W + xw] = s + bias[o];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// im2col for 3x3 pad-1 convs with stride (Downsample3D) or geometries the TMA path does not take.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const __nv_bfloat16* __restrict__ x, int NF, int H, int W, int C, int stride, int Ho, int Wo,
                 __nv_bfloat16* __restrict__ col) {
  pdl_prologue();
  const int nvec = C >> 3;
  const long long total = static_cast<long long>(NF) * Ho * Wo * 9 * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    const int t = static_cast<int>((i / nvec) % 9);
    const long long opix = i / (9LL * nvec);
    const int xo = static_cast<int>(opix % Wo);
    const int yo = static_cast<int>((opix / Wo) % Ho);
    const int n = static_cast<int>(opix / (static_cast<long long>(Wo) * Ho));
    const int yy = yo * stride + t / 3 - 1, xx = xo * stride + t % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      val = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + yy) * W + xx) * C + v * 8));
    *reinterpret_cast<uint4*>(col + (static_cast<size_t>(opix) * 9 + t) * C + v * 8) = val;
  }
}

__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int NF, int H, int W, int C, __nv_bfloat16* __restrict__ y) {
  pdl_prologue();
  const int nvec = C >> 3;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = static_cast<long long>(NF) * Ho * Wo * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    const long long opix = i / nvec;
    const int xo = static_cast<int>(opix % Wo);
    const int yo = static_cast<int>((opix / Wo) % Ho);
    const int n = static_cast<int>(opix / (static_cast<long long>(Wo) * Ho));
    const uint4 val =
        __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + (yo >> 1)) * W + (xo >> 1)) * C + v * 8));
    *reinterpret_cast<uint4*>(y + static_cast<size_t>(opix) * C + v * 8) = val;
  }
}

__global__ void cfg_ddim_kernel(const float* __restrict__ nu, const float* __restrict__ nt, float g, float sa_t,
                                float s1a_t, float sa_p, float s1a_p, const float* __restrict__ lat,
                                float* __restrict__ out, long long n) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float eps = nu[i] + g * (nt[i] - nu[i]);
    const float x0 = (lat[i] - s1a_t * eps) / sa_t;
    out[i] = sa_p * x0 + s1a_p * eps;
  }
}

int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  if (b > 148LL * 32) b = 148LL * 32;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

extern "C" int lavie_timestep_embedding(const float* t, int B, int dim, float* out, cudaStream_t stream) {
  LAVIE_REQUIRE(B > 0 && dim > 0 && dim % 2 == 0, LAVIE_ERR_SHAPE, "timestep_embedding: dim must be even");
  const int total = B * (dim / 2);
  launch_pdl(timestep_embedding_kernel, (total + 127) / 128, 128, 0, stream, t, B, dim, out);
  return lavie_check_launch("timestep_embedding_kernel");
}

extern "C" int lavie_linear_smallm(const float* x, int M, int K, const void* w, const float* bias, float* out, int N,
                                   int silu_in, int silu_out, cudaStream_t stream) {
  LAVIE_REQUIRE(M >= 1 && M <= 8 && K % 8 == 0 && N > 0, LAVIE_ERR_SHAPE, "linear_smallm: M=%d (<=8) K=%d N=%d", M, K, N);
  LAVIE_REQUIRE(al16(x) && al16(w), LAVIE_ERR_ALIGN, "linear_smallm: alignment");
  const int blocks = (N + 7) / 8;
  const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w);
  if (M <= 2) launch_pdl(linear_smallm_kernel<2>, blocks, 256, 0, stream, x, M, K, wp, bias, out, N, silu_in, silu_out);
  else launch_pdl(linear_smallm_kernel<8>, blocks, 256, 0, stream, x, M, K, wp, bias, out, N, silu_in, silu_out);
  return lavie_check_launch("linear_smallm_kernel");
}

extern "C" int lavie_conv_in(const float* x, int B, int Cin, int F, int H, int W, const float* w, const float* bias,
                             int Cout, void* out, int ldo, cudaStream_t stream) {
  LAVIE_REQUIRE(Cout % 8 == 0 && ldo % 8 == 0 && al16(out), LAVIE_ERR_SHAPE, "conv_in: Cout/ldo must be multiples of 8");
  const int smem = Cout * Cin * 9 * static_cast<int>(sizeof(float));
  LAVIE_REQUIRE(smem <= 200 * 1024, LAVIE_ERR_SHAPE, "conv_in: weights do not fit shared memory");
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    LAVIE_REQUIRE(e == cudaSuccess, LAVIE_ERR_CUDA, "cudaFuncSetAttribute(conv_in): %s", cudaGetErrorString(e));
    configured = smem;
  }
  const long long total = static_cast<long long>(B) * F * H * W * (Cout / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  launch_pdl(conv_in_kernel, static_cast<int>(blocks), 256, smem, stream, x, B, Cin, F, H, W, w, bias, Cout,
                                                                  static_cast<__nv_bfloat16*>(out), ldo);
  return lavie_check_launch("conv_in_kernel");
}

extern "C" int lavie_conv_out(const void* x, int ldx, const float* scale_shift, int B, int F, int H, int W, int C,
                              const float* w, const float* bias, int Cout, float* out, cudaStream_t stream) {
  LAVIE_REQUIRE(Cout == 4 && C % 8 == 0 && ldx % 8 == 0, LAVIE_ERR_SHAPE, "conv_out: Cout must be 4, C %% 8 == 0");
  LAVIE_REQUIRE(al16(x) && al16(scale_shift), LAVIE_ERR_ALIGN, "conv_out: alignment");
  const int smem = Cout * 9 * C * static_cast<int>(sizeof(float));
  LAVIE_REQUIRE(smem <= 200 * 1024, LAVIE_ERR_SHAPE, "conv_out: weights do not fit shared memory");
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_out_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    LAVIE_REQUIRE(e == cudaSuccess, LAVIE_ERR_CUDA, "cudaFuncSetAttribute(conv_out): %s", cudaGetErrorString(e));
    configured = smem;
  }
  const long long total = static_cast<long long>(B) * F * H * W;
  long long blocks = (total + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  launch_pdl(conv_out_kernel<4>, static_cast<int>(blocks), 256, smem, stream, static_cast<const __nv_bfloat16*>(x), ldx,
                                                                      scale_shift, B, F, H, W, C, w, bias, out);
  return lavie_check_launch("conv_out_kernel");
}

extern "C" int lavie_im2col3x3_bf16(const void* x, int NF, int H, int W, int C, int stride, void* col,
                                    cudaStream_t stream) {
  LAVIE_REQUIRE(C % 8 == 0 && (stride == 1 || stride == 2), LAVIE_ERR_SHAPE, "im2col: C %% 8 == 0, stride 1 or 2");
  LAVIE_REQUIRE(al16(x) && al16(col), LAVIE_ERR_ALIGN, "im2col: alignment");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = static_cast<long long>(NF) * Ho * Wo * 9 * (C / 8);
  launch_pdl(im2col3x3_kernel, grid_for(total, 256), 256, 0, stream, static_cast<const __nv_bfloat16*>(x), NF, H, W, C, stride,
                                                             Ho, Wo, static_cast<__nv_bfloat16*>(col));
  return lavie_check_launch("im2col3x3_kernel");
}

extern "C" int lavie_upsample_nearest2x(const void* x, int NF, int H, int W, int C, void* y, cudaStream_t stream) {
  LAVIE_REQUIRE(C % 8 == 0, LAVIE_ERR_SHAPE, "upsample: C %% 8 == 0");
  LAVIE_REQUIRE(al16(x) && al16(y), LAVIE_ERR_ALIGN, "upsample: alignment");
  const long long total = static_cast<long long>(NF) * 4 * H * W * (C / 8);
  launch_pdl(upsample2x_kernel, grid_for(total, 256), 256, 0, stream, static_cast<const __nv_bfloat16*>(x), NF, H, W, C,
                                                              static_cast<__nv_bfloat16*>(y));
  return lavie_check_launch("upsample2x_kernel");
}

extern "C" int lavie_cfg_ddim_step(const float* noise_uncond, const float* noise_text, float guidance, float alpha_t,
                                   float alpha_prev, const float* latents, float* latents_out, long long n,
                                   cudaStream_t stream) {
  LAVIE_REQUIRE(n > 0 && alpha_t > 0.f && alpha_t <= 1.f && alpha_prev > 0.f && alpha_prev <= 1.f, LAVIE_ERR_SHAPE,
                "cfg_ddim_step: bad arguments");
  launch_pdl(cfg_ddim_kernel, grid_for(n, 256), 256, 0, stream, noise_uncond, noise_text, guidance, sqrtf(alpha_t),
                                                        sqrtf(1.f - alpha_t), sqrtf(alpha_prev),
                                                        sqrtf(1.f - alpha_prev), latents, latents_out, n);
  return lavie_check_launch("cfg_ddim_kernel");
}
