// Bandwidth-bound normalisation kernels (channels-last bf16, 16-byte vector accesses, fp32 statistics).
//   GroupNorm  : nn.GroupNorm of resnet.py:144,160,180,191 / unet.py:288,504 (5-D: statistics span frames) and
//                attention.py:324,369 (4-D per frame), optionally followed by SiLU (resnet.py:181,193).
//   LayerNorm  : attention.py:444-477 (norm1 / norm2 / norm_temp / norm3).
#include "common.cuh"

// CTAs per launch (tools/bench_norm.py): the statistics kernel ends in a serial tail (block reduce, ticket, the last
// block folds every chunk partial), so it wants few, long chunks; the apply kernel is a pure stream and wants many.
int g_lavie_gn_target_ctas = 148 * 2;      // statistics (lavie_debug_set(6, n) for tuning)
int g_lavie_gn_apply_ctas = 148 * 8;       // apply

namespace {

// Streaming (multi-wave, non-persistent) kernels: A/B switch for WHEN the dependent grid may start.  Default: at our start
// (its prologue overlaps our whole run); -DLAVIE_PDL_LATE_NORMS: only when we exit.
#ifdef LAVIE_PDL_LATE_NORMS
#define PDL_STREAM_PROLOGUE() pdl_wait()
#else
#define PDL_STREAM_PROLOGUE() pdl_prologue()
#endif

constexpr int GN_THREADS = 256;
constexpr int GN_MIN_ROWS_PER_CHUNK = 16;

__device__ __forceinline__ uint4 ld_vec8(const __nv_bfloat16* x0, int ld0, int c0, const __nv_bfloat16* x1, int ld1,
                                         size_t row, int col) {
  const __nv_bfloat16* src = (col < c0) ? x0 + row * ld0 + col : x1 + row * ld1 + (col - c0);
  return __ldg(reinterpret_cast<const uint4*>(src));
}

// check mode (split-bf16 triples, check.cu): the lo half of the same 8 channels sits c0 (c1) columns further
__device__ __forceinline__ uint4 ld_vec8_lo(const __nv_bfloat16* x0, int ld0, int c0, const __nv_bfloat16* x1, int ld1,
                                            int c1, size_t row, int col) {
  const __nv_bfloat16* src = (col < c0) ? x0 + row * ld0 + c0 + col : x1 + row * ld1 + c1 + (col - c0);
  return __ldg(reinterpret_cast<const uint4*>(src));
}

// combine the chunk partials of one sample in fp64 (fixed order) and emit per-channel (scale, shift); whole block
__device__ __forceinline__ void gn_finalize_sample(const float* __restrict__ partial, int sample, int chunks, int groups,
                                                   int C, double inv_count, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float eps,
                                                   float* __restrict__ scale_shift) {
  __shared__ float s_mean[64], s_rstd[64];
  // 8 consecutive lanes share one group: each sums every 8th chunk in fp64, then a fixed-order shuffle tree
  const int sub = threadIdx.x & 7;
  for (int g = threadIdx.x >> 3; g < groups; g += blockDim.x >> 3) {
    double a = 0.0, b = 0.0;
    for (int k0 = sub; k0 < chunks; k0 += 64) {          // 8 independent loads in flight, summed in a fixed order
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + 8 * u;
        v[u] = k < chunks ? __ldcg(reinterpret_cast<const float2*>(
                                partial + ((static_cast<size_t>(sample) * chunks + k) * groups + g) * 2))
                          : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a += v[u].x;
        b += v[u].y;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (sub == 0) {
      const double mean = a * inv_count;
      double var = b * inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = static_cast<float>(mean);
      s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * gamma[c];
    float* dst = scale_shift + (static_cast<size_t>(sample) * C + c) * 2;
    dst[0] = sc;
    dst[1] = beta[c] - s_mean[g] * sc;
  }
}

// one block per sample
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int chunks, int groups, int C,
                                   double inv_count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float* __restrict__ scale_shift) {
  pdl_prologue();
  gn_finalize_sample(partial, blockIdx.x, chunks, groups, C, inv_count, gamma, beta, eps, scale_shift);
}

// Optional fused finalize: the LAST block of a sample to publish its partials (atomic ticket) folds all of them into
// (scale, shift) -- same arithmetic and order as gn_finalize_kernel, one launch less per GroupNorm.
struct GnFused {
  int* tickets;            // [samples], zero on entry, reset to zero by the finalizing block; nullptr = stats only
  double inv_count;
  const float* gamma;
  const float* beta;
  float eps;
  float* scale_shift;
};

// partial[sample][chunk][group][2]
template <bool TRIPLE>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x0, int ld0, int c0, const __nv_bfloat16* __restrict__ x1, int ld1,
                int c1, int rows_per_sample, int rows_per_chunk, int groups, int chunks, float* __restrict__ partial,
                const GnFused fused) {
  PDL_STREAM_PROLOGUE();
  extern __shared__ float sm[];          // [row_lanes][2][C]  (one private slot per row lane: deterministic)
  const int C = c0 + c1;
  const int sample = blockIdx.y;
  const int chunk = blockIdx.x;
  const int vec_per_row = C >> 3;
  const int r_begin = chunk * rows_per_chunk;
  const int r_end = min(r_begin + rows_per_chunk, rows_per_sample);
  // column slots: a thread owns one 8-channel vector per slot; narrow rows are covered by several row lanes
  const int lanes_per_row = min(vec_per_row, static_cast<int>(blockDim.x));
  const int row_lanes = blockDim.x / lanes_per_row;            // rows processed concurrently
  const int my_rl = threadIdx.x / lanes_per_row;
  if (my_rl < row_lanes) {
    float* s_sum = sm + static_cast<size_t>(my_rl) * 2 * C;
    float* s_sq = s_sum + C;
    for (int my_vec = threadIdx.x % lanes_per_row; my_vec < vec_per_row; my_vec += lanes_per_row) {
      float s[8], q[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
      constexpr int U = 8;                     // independent 16-byte loads in flight per thread
      for (int r = r_begin + my_rl; r < r_end; r += row_lanes * U) {
        uint4 v[U], vl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int rr = r + u * row_lanes;
          const size_t row = static_cast<size_t>(sample) * rows_per_sample + (rr < r_end ? rr : r);
          v[u] = ld_vec8(x0, ld0, c0, x1, ld1, row, my_vec * 8);
          if (TRIPLE) vl[u] = ld_vec8_lo(x0, ld0, c0, x1, ld1, c1, row, my_vec * 8);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (r + u * row_lanes < r_end) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            const uint32_t wl[4] = {vl[u].x, vl[u].y, vl[u].z, vl[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = unpack_bf16(w[e]);
              if (TRIPLE) {
                const float2 fl = unpack_bf16(wl[e]);
                f.x += fl.x;
                f.y += fl.y;
              }
              s[2 * e] += f.x; q[2 * e] += f.x * f.x;
              s[2 * e + 1] += f.y; q[2 * e + 1] += f.y * f.y;
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s_sum[my_vec * 8 + e] = s[e];
        s_sq[my_vec * 8 + e] = q[e];
      }
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rl = 0; rl < row_lanes; ++rl) {
      const float* ps = sm + static_cast<size_t>(rl) * 2 * C;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += ps[c]; b += ps[C + c]; }
    }
    float* dst = partial + ((static_cast<size_t>(sample) * chunks + chunk) * groups + g) * 2;
    dst[0] = a;
    dst[1] = b;
  }
  if (fused.tickets == nullptr) return;
  __shared__ int s_last;
  __threadfence();                                 // partials visible device-wide before the ticket
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(fused.tickets + sample, 1) == chunks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  gn_finalize_sample(partial, sample, chunks, groups, C, fused.inv_count, fused.gamma, fused.beta, fused.eps,
                     fused.scale_shift);
  if (threadIdx.x == 0) fused.tickets[sample] = 0;
}

// y = act(x * scale + shift).  A thread owns one 8-channel column vector and walks down the rows of ONE sample, so the
// 16 (scale, shift) floats stay in registers (they are 4x the bytes of the data they apply to); 4 independent 16-byte
// loads are in flight per thread.  grid = (row chunks, samples).
template <bool TRIPLE>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, int ld0, int c0, const __nv_bfloat16* __restrict__ x1, int ld1,
                int c1, int rows_per_sample, int rows_per_chunk, const float* __restrict__ scale_shift, int silu,
                __nv_bfloat16* __restrict__ y, int ldy) {
  PDL_STREAM_PROLOGUE();
  const int C = c0 + c1;
  const int vec_per_row = C >> 3;
  const int sample = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(r_begin + rows_per_chunk, rows_per_sample);
  const int lanes_per_row = min(vec_per_row, static_cast<int>(blockDim.x));
  const int row_lanes = blockDim.x / lanes_per_row;
  const int my_rl = threadIdx.x / lanes_per_row;
  if (my_rl >= row_lanes) return;
  constexpr int U = 4;
  for (int my_vec = threadIdx.x % lanes_per_row; my_vec < vec_per_row; my_vec += lanes_per_row) {
    const int col = my_vec * 8;
    float sc[8], sh[8];
    const float4* ss = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(sample) * C + col) * 2);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 p = __ldg(ss + e);
      sc[2 * e] = p.x; sh[2 * e] = p.y; sc[2 * e + 1] = p.z; sh[2 * e + 1] = p.w;
    }
    for (int r = r_begin + my_rl; r < r_end; r += row_lanes * U) {
      uint4 v[U], vl[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = r + u * row_lanes;
        const size_t row = static_cast<size_t>(sample) * rows_per_sample + (rr < r_end ? rr : r);
        v[u] = ld_vec8(x0, ld0, c0, x1, ld1, row, col);
        if (TRIPLE) vl[u] = ld_vec8_lo(x0, ld0, c0, x1, ld1, c1, row, col);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = r + u * row_lanes;
        if (rr >= r_end) break;
        const size_t row = static_cast<size_t>(sample) * rows_per_sample + rr;
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const uint32_t wl[4] = {vl[u].x, vl[u].y, vl[u].z, vl[u].w};
        uint32_t o[4], ol[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = unpack_bf16(w[e]);
          if (TRIPLE) {
            const float2 fl = unpack_bf16(wl[e]);
            f.x += fl.x;
            f.y += fl.y;
          }
          float a = f.x * sc[2 * e] + sh[2 * e];
          float b = f.y * sc[2 * e + 1] + sh[2 * e + 1];
          if (silu) {
            if (TRIPLE) { a = a / (1.0f + expf(-a)); b = b / (1.0f + expf(-b)); }
            else { a = silu_f(a); b = silu_f(b); }
          }
          o[e] = pack_bf16(a, b);
          if (TRIPLE) {
            const float2 hi = unpack_bf16(o[e]);
            ol[e] = pack_bf16(a - hi.x, b - hi.y);
          }
        }
        *reinterpret_cast<uint4*>(y + row * ldy + col) = make_uint4(o[0], o[1], o[2], o[3]);
        if (TRIPLE) {      // [hi | lo | hi]
          *reinterpret_cast<uint4*>(y + row * ldy + C + col) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
          *reinterpret_cast<uint4*>(y + row * ldy + 2 * C + col) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
}

// LayerNorm for C = 40*L (320 / 640 / 1280): L lanes share a row (5 independent 16-byte loads per lane), so a warp
// keeps 32/L rows = 2.5 KB in flight instead of one row.  Same arithmetic (two-pass, fp32) as layernorm_kernel.
template <int L>
__global__ void __launch_bounds__(256)
layernorm_grouped_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int ldy, int rows,
                         int perm_hw, int perm_hwp) {
  PDL_STREAM_PROLOGUE();
  constexpr int VPL = 5;
  constexpr int C = 40 * L;
  constexpr int RPW = 32 / L;
  const int lane = threadIdx.x & 31;
  const int sub = lane % L;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long row = warp * RPW + lane / L;
  const bool ok = row < rows;
  const __nv_bfloat16* src = x + static_cast<size_t>(ok ? row : 0) * ldx;
  uint4 u[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(src + (sub + i * L) * 8));
  float f[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = unpack_bf16(w[e]);
      f[i][2 * e] = t.x;
      f[i][2 * e + 1] = t.y;
      sum += t.x + t.y;
    }
  }
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = f[i][e] - mean;
      sq += d * d;
    }
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(C) + eps);
  if (!ok) return;
  long long drow = row;
  if (perm_hw > 0) {
    const long long fr = row / perm_hw, pix = row - fr * perm_hw;
    const long long blk = pix / perm_hwp;
    drow = (blk * (rows / perm_hw) + fr) * perm_hwp + (pix - blk * perm_hwp);
  }
  __nv_bfloat16* dst = y + static_cast<size_t>(drow) * ldy;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = sub + i * L;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
    uint4 o;
    o.x = pack_bf16((f[i][0] - mean) * rstd * g0.x + b0.x, (f[i][1] - mean) * rstd * g0.y + b0.y);
    o.y = pack_bf16((f[i][2] - mean) * rstd * g0.z + b0.z, (f[i][3] - mean) * rstd * g0.w + b0.w);
    o.z = pack_bf16((f[i][4] - mean) * rstd * g1.x + b1.x, (f[i][5] - mean) * rstd * g1.y + b1.y);
    o.w = pack_bf16((f[i][6] - mean) * rstd * g1.z + b1.z, (f[i][7] - mean) * rstd * g1.w + b1.w);
    *reinterpret_cast<uint4*>(dst + v * 8) = o;
  }
}

// One warp per row; the row (C <= 2048) lives in registers between the two reduction passes.
template <int MAX_VEC>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int ldy, int rows, int C,
                 int perm_hw, int perm_hwp) {
  PDL_STREAM_PROLOGUE();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = C >> 3;
  const __nv_bfloat16* src = x + static_cast<size_t>(warp) * ldx;
  float f[MAX_VEC][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + v * 8));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_bf16(w[e]);
        f[i][2 * e] = t.x;
        f[i][2 * e + 1] = t.y;
        sum += t.x + t.y;
      }
    }
  }
  const float mean = warp_sum(sum) / static_cast<float>(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = f[i][e] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(C) + eps);
  // Optional scatter for frame sharding: source row (f, pixel) of a [F_loc, HW] shard goes to row
  // (pixel_block * F_loc + f) * HWp + pixel_in_block, i.e. the send buffer of the all-to-all that turns frame
  // sharding into pixel sharding around the temporal attention.
  int drow = warp;
  if (perm_hw > 0) {
    const int f = warp / perm_hw, pix = warp - f * perm_hw;
    const int blk = pix / perm_hwp;
    drow = (blk * (rows / perm_hw) + f) * perm_hwp + (pix - blk * perm_hwp);
  }
  __nv_bfloat16* dst = y + static_cast<size_t>(drow) * ldy;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      uint4 o;
      o.x = pack_bf16((f[i][0] - mean) * rstd * g0.x + b0.x, (f[i][1] - mean) * rstd * g0.y + b0.y);
      o.y = pack_bf16((f[i][2] - mean) * rstd * g0.z + b0.z, (f[i][3] - mean) * rstd * g0.w + b0.w);
      o.z = pack_bf16((f[i][4] - mean) * rstd * g1.x + b1.x, (f[i][5] - mean) * rstd * g1.y + b1.y);
      o.w = pack_bf16((f[i][6] - mean) * rstd * g1.z + b1.z, (f[i][7] - mean) * rstd * g1.w + b1.w);
      *reinterpret_cast<uint4*>(dst + v * 8) = o;
    }
  }
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int check_sources(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1) {
  LAVIE_REQUIRE(c0 > 0 && c0 % 8 == 0 && c1 % 8 == 0 && ld0 % 8 == 0 && (c1 == 0 || ld1 % 8 == 0), LAVIE_ERR_SHAPE,
                "groupnorm: channel counts and row strides must be multiples of 8");
  LAVIE_REQUIRE(al16(x0) && (c1 == 0 || al16(x1)), LAVIE_ERR_ALIGN, "groupnorm: inputs must be 16-byte aligned");
  return LAVIE_OK;
}

}  // namespace

namespace {
int gn_rows_per_chunk(int samples, int rows_per_sample, int target_ctas) {
  int chunks = (target_ctas + samples - 1) / samples;
  const int max_chunks = (rows_per_sample + GN_MIN_ROWS_PER_CHUNK - 1) / GN_MIN_ROWS_PER_CHUNK;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  return (rows_per_sample + chunks - 1) / chunks;
}
}  // namespace

extern "C" int lavie_groupnorm_chunks(int samples, int rows_per_sample) {
  const int rpc = gn_rows_per_chunk(samples, rows_per_sample, g_lavie_gn_target_ctas);
  return (rows_per_sample + rpc - 1) / rpc;
}

namespace {
int gn_stats_impl(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1, int samples, int rows_per_sample,
                  int groups, float* partial, bool triple, cudaStream_t stream) {
  int rc = check_sources(x0, ld0, c0, x1, ld1, c1);
  if (rc) return rc;
  const int C = c0 + c1;
  LAVIE_REQUIRE(C % groups == 0 && groups <= 64 && C <= 8192, LAVIE_ERR_SHAPE,
                "groupnorm: C=%d groups=%d unsupported", C, groups);
  const int rows_per_chunk = gn_rows_per_chunk(samples, rows_per_sample, g_lavie_gn_target_ctas);
  const int chunks = (rows_per_sample + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid(chunks, samples);
  const int vec_per_row = C >> 3;
  const int row_lanes = GN_THREADS / (vec_per_row < GN_THREADS ? vec_per_row : GN_THREADS);
  if (triple)
    launch_pdl(gn_stats_kernel<true>, grid, GN_THREADS, static_cast<size_t>(row_lanes) * 2 * C * sizeof(float), stream,
               static_cast<const __nv_bfloat16*>(x0), ld0, c0, static_cast<const __nv_bfloat16*>(x1), ld1, c1,
               rows_per_sample, rows_per_chunk, groups, chunks, partial, GnFused{});
  else
    launch_pdl(gn_stats_kernel<false>, grid, GN_THREADS, static_cast<size_t>(row_lanes) * 2 * C * sizeof(float), stream,
               static_cast<const __nv_bfloat16*>(x0), ld0, c0, static_cast<const __nv_bfloat16*>(x1), ld1, c1,
               rows_per_sample, rows_per_chunk, groups, chunks, partial, GnFused{});
  return lavie_check_launch("gn_stats_kernel");
}
}  // namespace

extern "C" int lavie_groupnorm_stats(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1, int samples,
                                     int rows_per_sample, int groups, float* partial, cudaStream_t stream) {
  return gn_stats_impl(x0, ld0, c0, x1, ld1, c1, samples, rows_per_sample, groups, partial, false, stream);
}

/* check mode: the sources are split-bf16 triples (row = [hi | lo | hi], c0 / c1 = width of ONE block) */
extern "C" int lavie_check_groupnorm_stats(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1,
                                           int samples, int rows_per_sample, int groups, float* partial,
                                           cudaStream_t stream) {
  LAVIE_REQUIRE(ld0 >= 2 * c0 && (c1 == 0 || ld1 >= 2 * c1), LAVIE_ERR_SHAPE, "check_groupnorm_stats: triple strides");
  return gn_stats_impl(x0, ld0, c0, x1, ld1, c1, samples, rows_per_sample, groups, partial, true, stream);
}

extern "C" int lavie_groupnorm_finalize(const float* partial, int samples, int chunks, int groups, int C,
                                        long long count_per_group, const float* gamma, const float* beta, float eps,
                                        float* scale_shift, cudaStream_t stream) {
  LAVIE_REQUIRE(groups <= 64 && groups % 4 == 0 && C % groups == 0 && count_per_group > 0, LAVIE_ERR_SHAPE,
                "groupnorm_finalize: groups must be a multiple of 4 (<= 64) dividing C");
  launch_pdl(gn_finalize_kernel, samples, 256, 0, stream, partial, chunks, groups, C, 1.0 / static_cast<double>(count_per_group),
                                                  gamma, beta, eps, scale_shift);
  return lavie_check_launch("gn_finalize_kernel");
}

extern "C" int lavie_groupnorm_scale_shift(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1,
                                           int samples, int rows_per_sample, int groups, const float* gamma,
                                           const float* beta, float eps, float* partial, int* tickets,
                                           float* scale_shift, cudaStream_t stream) {
  int rc = check_sources(x0, ld0, c0, x1, ld1, c1);
  if (rc) return rc;
  const int C = c0 + c1;
  LAVIE_REQUIRE(C % groups == 0 && groups <= 64 && groups % 4 == 0 && C <= 8192, LAVIE_ERR_SHAPE,
                "groupnorm: C=%d groups=%d unsupported", C, groups);
  LAVIE_REQUIRE(tickets != nullptr && partial != nullptr && scale_shift != nullptr, LAVIE_ERR_SHAPE,
                "groupnorm_scale_shift: null buffer");
  const int rows_per_chunk = gn_rows_per_chunk(samples, rows_per_sample, g_lavie_gn_target_ctas);
  const int chunks = (rows_per_sample + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid(chunks, samples);
  const int vec_per_row = C >> 3;
  const int row_lanes = GN_THREADS / (vec_per_row < GN_THREADS ? vec_per_row : GN_THREADS);
  GnFused fused;
  fused.tickets = tickets;
  fused.inv_count = 1.0 / (static_cast<double>(rows_per_sample) * (C / groups));
  fused.gamma = gamma;
  fused.beta = beta;
  fused.eps = eps;
  fused.scale_shift = scale_shift;
  launch_pdl(gn_stats_kernel<false>, grid, GN_THREADS, static_cast<size_t>(row_lanes) * 2 * C * sizeof(float), stream,
             static_cast<const __nv_bfloat16*>(x0), ld0, c0, static_cast<const __nv_bfloat16*>(x1), ld1, c1,
             rows_per_sample, rows_per_chunk, groups, chunks, partial, fused);
  return lavie_check_launch("gn_stats_kernel(fused)");
}

namespace {
int gn_apply_impl(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1, int samples, int rows_per_sample,
                  const float* scale_shift, int silu, void* y, int ldy, bool triple, cudaStream_t stream) {
  int rc = check_sources(x0, ld0, c0, x1, ld1, c1);
  if (rc) return rc;
  LAVIE_REQUIRE(al16(y) && ldy % 8 == 0 && al16(scale_shift), LAVIE_ERR_ALIGN, "groupnorm_apply: output alignment");
  const int rows_per_chunk = gn_rows_per_chunk(samples, rows_per_sample, g_lavie_gn_apply_ctas);
  const int chunks = (rows_per_sample + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid(chunks, samples);
  if (triple)
    launch_pdl(gn_apply_kernel<true>, grid, 256, 0, stream, static_cast<const __nv_bfloat16*>(x0), ld0, c0,
               static_cast<const __nv_bfloat16*>(x1), ld1, c1, rows_per_sample, rows_per_chunk, scale_shift, silu,
               static_cast<__nv_bfloat16*>(y), ldy);
  else
    launch_pdl(gn_apply_kernel<false>, grid, 256, 0, stream, static_cast<const __nv_bfloat16*>(x0), ld0, c0,
               static_cast<const __nv_bfloat16*>(x1), ld1, c1, rows_per_sample, rows_per_chunk, scale_shift, silu,
               static_cast<__nv_bfloat16*>(y), ldy);
  return lavie_check_launch("gn_apply_kernel");
}
}  // namespace

extern "C" int lavie_groupnorm_apply(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1, int samples,
                                     int rows_per_sample, const float* scale_shift, int silu, void* y, int ldy,
                                     cudaStream_t stream) {
  return gn_apply_impl(x0, ld0, c0, x1, ld1, c1, samples, rows_per_sample, scale_shift, silu, y, ldy, false, stream);
}

/* check mode: triple sources in, triple [rows, 3C] out (ldy >= 3C) */
extern "C" int lavie_check_groupnorm_apply(const void* x0, int ld0, int c0, const void* x1, int ld1, int c1,
                                           int samples, int rows_per_sample, const float* scale_shift, int silu,
                                           void* y, int ldy, cudaStream_t stream) {
  LAVIE_REQUIRE(ld0 >= 2 * c0 && (c1 == 0 || ld1 >= 2 * c1) && ldy >= 3 * (c0 + c1), LAVIE_ERR_SHAPE,
                "check_groupnorm_apply: triple strides");
  return gn_apply_impl(x0, ld0, c0, x1, ld1, c1, samples, rows_per_sample, scale_shift, silu, y, ldy, true, stream);
}

namespace {
int layernorm_impl(const void* x, int ldx, const float* gamma, const float* beta, float eps, void* y, int ldy, int rows,
                   int C, int perm_hw, int perm_hwp, cudaStream_t stream) {
  LAVIE_REQUIRE(C % 8 == 0 && C <= 2048 && ldx % 8 == 0 && ldy % 8 == 0, LAVIE_ERR_SHAPE,
                "layernorm: C=%d must be a multiple of 8 and <= 2048", C);
  LAVIE_REQUIRE(al16(x) && al16(y) && al16(gamma) && al16(beta), LAVIE_ERR_ALIGN, "layernorm: alignment");
  if (rows <= 0) return LAVIE_OK;
  const int blocks = (rows + 7) / 8;     // 8 warps (rows) per 256-thread block
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  const int nvec = C >> 3;
  if (C == 320 || C == 640 || C == 1280) {
    const int lanes = C / 40;
    const int rpw = 32 / lanes;
    const long long warps = (static_cast<long long>(rows) + rpw - 1) / rpw;
    const int gblocks = static_cast<int>((warps + 7) / 8);
    if (lanes == 8)
      launch_pdl(layernorm_grouped_kernel<8>, gblocks, 256, 0, stream, xp, ldx, gamma, beta, eps, yp, ldy, rows, perm_hw, perm_hwp);
    else if (lanes == 16)
      launch_pdl(layernorm_grouped_kernel<16>, gblocks, 256, 0, stream, xp, ldx, gamma, beta, eps, yp, ldy, rows, perm_hw, perm_hwp);
    else
      launch_pdl(layernorm_grouped_kernel<32>, gblocks, 256, 0, stream, xp, ldx, gamma, beta, eps, yp, ldy, rows, perm_hw, perm_hwp);
    return lavie_check_launch("layernorm_grouped_kernel");
  }
  if (nvec <= 64)
    launch_pdl(layernorm_kernel<2>, blocks, 256, 0, stream, xp, ldx, gamma, beta, eps, yp, ldy, rows, C, perm_hw, perm_hwp);
  else if (nvec <= 160)
    launch_pdl(layernorm_kernel<5>, blocks, 256, 0, stream, xp, ldx, gamma, beta, eps, yp, ldy, rows, C, perm_hw, perm_hwp);
  else
    launch_pdl(layernorm_kernel<8>, blocks, 256, 0, stream, xp, ldx, gamma, beta, eps, yp, ldy, rows, C, perm_hw, perm_hwp);
  return lavie_check_launch("layernorm_kernel");
}

// out[(f, blk*HWp + j)] = res[(f, blk*HWp + j)] + z[(blk, f, j)]: the receive side of the all-to-all back to frame sharding
__global__ void __launch_bounds__(256)
add_gathered_kernel(const __nv_bfloat16* __restrict__ res, int ldr, const __nv_bfloat16* __restrict__ z, int ldz,
                    __nv_bfloat16* __restrict__ out, int ldo, int rows, int C, int hw, int hwp) {
  pdl_prologue();
  const int nvec = C >> 3;
  const int f_loc = rows / hw;
  const long long total = static_cast<long long>(rows) * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / nvec), v = static_cast<int>(i % nvec);
    const int f = row / hw, pix = row - f * hw;
    const int blk = pix / hwp;
    const size_t zrow = (static_cast<size_t>(blk) * f_loc + f) * hwp + (pix - blk * hwp);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(res + static_cast<size_t>(row) * ldr + v * 8));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(z + zrow * ldz + v * 8));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(bw[e]);
      o[e] = pack_bf16(x.x + y.x, x.y + y.y);
    }
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * ldo + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// sums[sample][group][2] (fp64) = ordered sum of the chunk partials: what gets all-reduced across frame shards
__global__ void gn_reduce_kernel(const float* __restrict__ partial, int chunks, int groups, double* __restrict__ sums) {
  pdl_prologue();
  const int sample = blockIdx.x;
  const int sub = threadIdx.x & 7;
  for (int g = threadIdx.x >> 3; g < groups; g += blockDim.x >> 3) {
    double a = 0.0, b = 0.0;
    for (int k0 = sub; k0 < chunks; k0 += 64) {          // 8 independent loads in flight, summed in a fixed order
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + 8 * u;
        v[u] = k < chunks ? *reinterpret_cast<const float2*>(
                                partial + ((static_cast<size_t>(sample) * chunks + k) * groups + g) * 2)
                          : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a += v[u].x;
        b += v[u].y;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (sub == 0) {
      sums[(static_cast<size_t>(sample) * groups + g) * 2] = a;
      sums[(static_cast<size_t>(sample) * groups + g) * 2 + 1] = b;
    }
  }
}

__global__ void gn_finalize_sums_kernel(const double* __restrict__ sums, int groups, int C, double inv_count,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                        float* __restrict__ scale_shift) {
  pdl_prologue();
  const int sample = blockIdx.x;
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double mean = sums[(static_cast<size_t>(sample) * groups + g) * 2] * inv_count;
    double var = sums[(static_cast<size_t>(sample) * groups + g) * 2 + 1] * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = rstd * gamma[c];
    float* dst = scale_shift + (static_cast<size_t>(sample) * C + c) * 2;
    dst[0] = sc;
    dst[1] = beta[c] - static_cast<float>(mean) * sc;
  }
}
}  // namespace

namespace {
// GroupNorm statistics from the producers' per-slab micro-group sums (gemm.cu epilogue): one block per (group, sample).
// A group is a run of whole decades (10 channels) of the concatenated sources; a decade lies in one or two 32-column
// chunks of its source, so every (slab, decade) costs one or two 8-byte loads.  Fixed summation order: thread-strided
// fp32 partials over a few terms each, then an fp64 tree -> deterministic.  sums != nullptr: stop at the fp64
// (sum, sumsq) per (sample, group) (frame-sharded exchange).
__global__ void __launch_bounds__(256)
gn_colsums_kernel(const float* __restrict__ cs0, int c0, const float* __restrict__ cs1, int c1, int slabs_per_sample,
                  int groups, int mg, int segs0, int segs1, double inv_count, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float eps, float* __restrict__ scale_shift, double* __restrict__ sums) {
  pdl_prologue();
  const int g = blockIdx.x, sample = blockIdx.y;
  const int C = c0 + c1, cpg = C / groups;
  const int ch0 = g * cpg;
  const int decs = cpg / mg;                                   // micro-groups (mg = 10 or 8 channels) per group
  const long long total = static_cast<long long>(slabs_per_sample) * decs;
  double a = 0.0, b = 0.0;
  auto load = [&](long long i) -> float2 {
    if (i >= total) return make_float2(0.f, 0.f);
    const long long sl = i / decs;                             // slab of this sample, in row order
    int ch = ch0 + static_cast<int>(i % decs) * mg;            // first channel of the micro-group in the concat
    const float* cs = cs0;
    int cn = c0, segs = segs0;
    if (ch >= c0) {
      ch -= c0;
      cs = cs1;
      cn = c1;
      segs = segs1;
    }
    // a source written by the fused upsample conv (gemm.cu conv == 3) keeps its slabs in `segs` = 4 phase segments, each
    // holding 1 / segs of every sample's slabs; any order works for a sum as long as it is fixed
    long long slab;
    if (segs > 1) {
      const long long sps = slabs_per_sample / segs;
      const long long seg = sl / sps;
      slab = seg * (static_cast<long long>(gridDim.y) * sps) + static_cast<long long>(sample) * sps + (sl - seg * sps);
    } else {
      slab = static_cast<long long>(sample) * slabs_per_sample + sl;
    }
    const int dec = ch / mg;
    const int k_lo = ch >> 5, k_hi = (ch + mg - 1) >> 5;       // the chunk(s) the micro-group touches
    const float* row = cs + slab * (cn >> 5) * 8;
    float2 v = __ldg(reinterpret_cast<const float2*>(row + (k_lo * 4 + (dec - (k_lo * 32) / mg)) * 2));
    if (k_hi != k_lo) {
      const float2 w = __ldg(reinterpret_cast<const float2*>(row + (k_hi * 4 + (dec - (k_hi * 32) / mg)) * 2));
      v.x += w.x;
      v.y += w.y;
    }
    return v;
  };
  for (long long i = threadIdx.x; i < total; i += 256 * 8) {   // 8 independent loads in flight, summed in a fixed order
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = load(i + 256LL * u);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a += v[u].x;
      b += v[u].y;
    }
  }
  __shared__ double sa[256], sb[256];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (sums != nullptr) {
    if (threadIdx.x == 0) {
      sums[(static_cast<size_t>(sample) * groups + g) * 2] = sa[0];
      sums[(static_cast<size_t>(sample) * groups + g) * 2 + 1] = sb[0];
    }
    return;
  }
  const double mean = sa[0] * inv_count;
  double var = sb[0] * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
    const float sc = rstd * gamma[ch0 + c];
    float* dst = scale_shift + (static_cast<size_t>(sample) * C + ch0 + c) * 2;
    dst[0] = sc;
    dst[1] = beta[ch0 + c] - static_cast<float>(mean) * sc;
  }
}

int colsums_impl(const float* cs0, int c0, const float* cs1, int c1, int samples, int rows_per_sample, int groups,
                 const float* gamma, const float* beta, float eps, float* scale_shift, double* sums,
                 cudaStream_t stream, int segs0 = 1, int segs1 = 1) {
  const int C = c0 + c1;
  LAVIE_REQUIRE(segs0 >= 1 && segs1 >= 1 && (rows_per_sample / 32) % segs0 == 0 && (rows_per_sample / 32) % segs1 == 0,
                LAVIE_ERR_SHAPE, "groupnorm colsums: %d slabs per sample do not split into %d / %d phase segments",
                rows_per_sample / 32, segs0, segs1);
  LAVIE_REQUIRE(samples > 0 && rows_per_sample > 0 && rows_per_sample % 32 == 0 && groups > 0 && C % groups == 0 &&
                    c0 > 0 && c0 % 32 == 0 && c1 % 32 == 0, LAVIE_ERR_SHAPE,
                "groupnorm colsums: rows_per_sample=%d must be a multiple of 32, C=%d %% groups=%d == 0, sources multiples "
                "of 32 channels", rows_per_sample, C, groups);
  // micro-group width the PRODUCERS used (gemm.cu fill_epilogue): 10 channels when their N is a multiple of 10, else 8
  const int mg = (c0 % 10 == 0) ? 10 : 8;
  LAVIE_REQUIRE((c1 == 0 || ((c1 % 10 == 0) == (mg == 10))) && (C / groups) % mg == 0 && c0 % mg == 0, LAVIE_ERR_SHAPE,
                "groupnorm colsums: sources c0=%d c1=%d with groups of %d channels do not share a micro-group width", c0,
                c1, C / groups);
  LAVIE_REQUIRE(cs0 != nullptr && (c1 == 0 || cs1 != nullptr), LAVIE_ERR_SHAPE, "groupnorm colsums: null statistics");
  dim3 grid(groups, samples);
  launch_pdl(gn_colsums_kernel, grid, 256, 0, stream, cs0, c0, cs1, c1, rows_per_sample / 32, groups, mg, segs0, segs1,
             1.0 / (static_cast<double>(rows_per_sample) * (C / groups)), gamma, beta, eps, scale_shift, sums);
  return lavie_check_launch("gn_colsums_kernel");
}
}  // namespace

extern "C" int lavie_groupnorm_finalize_colsums(const float* cs0, int c0, const float* cs1, int c1, int samples,
                                                int rows_per_sample, int groups, const float* gamma, const float* beta,
                                                float eps, float* scale_shift, cudaStream_t stream) {
  LAVIE_REQUIRE(scale_shift != nullptr && gamma != nullptr && beta != nullptr, LAVIE_ERR_SHAPE, "groupnorm colsums: null");
  return colsums_impl(cs0, c0, cs1, c1, samples, rows_per_sample, groups, gamma, beta, eps, scale_shift, nullptr, stream);
}

extern "C" int lavie_groupnorm_finalize_colsums_seg(const float* cs0, int c0, int segs0, const float* cs1, int c1,
                                                    int segs1, int samples, int rows_per_sample, int groups,
                                                    const float* gamma, const float* beta, float eps, float* scale_shift,
                                                    cudaStream_t stream) {
  LAVIE_REQUIRE(scale_shift != nullptr && gamma != nullptr && beta != nullptr, LAVIE_ERR_SHAPE, "groupnorm colsums: null");
  return colsums_impl(cs0, c0, cs1, c1, samples, rows_per_sample, groups, gamma, beta, eps, scale_shift, nullptr, stream,
                      segs0, segs1 > 0 ? segs1 : 1);
}

extern "C" int lavie_groupnorm_reduce_colsums(const float* cs0, int c0, const float* cs1, int c1, int samples,
                                              int rows_per_sample, int groups, double* sums, cudaStream_t stream) {
  LAVIE_REQUIRE(sums != nullptr, LAVIE_ERR_SHAPE, "groupnorm colsums: null sums");
  return colsums_impl(cs0, c0, cs1, c1, samples, rows_per_sample, groups, nullptr, nullptr, 0.f, nullptr, sums, stream);
}

extern "C" int lavie_layernorm_bf16(const void* x, int ldx, const float* gamma, const float* beta, float eps, void* y,
                                    int ldy, int rows, int C, cudaStream_t stream) {
  return layernorm_impl(x, ldx, gamma, beta, eps, y, ldy, rows, C, 0, 0, stream);
}

extern "C" int lavie_layernorm_scatter_bf16(const void* x, int ldx, const float* gamma, const float* beta, float eps,
                                            void* y, int ldy, int rows, int C, int hw, int hwp, cudaStream_t stream) {
  LAVIE_REQUIRE(hw > 0 && hwp > 0 && hw % hwp == 0 && rows % hw == 0, LAVIE_ERR_SHAPE,
                "layernorm_scatter: rows=%d must be whole frames of hw=%d, hw %% hwp=%d == 0", rows, hw, hwp);
  return layernorm_impl(x, ldx, gamma, beta, eps, y, ldy, rows, C, hw, hwp, stream);
}

extern "C" int lavie_add_gathered_bf16(const void* res, int ldr, const void* z, int ldz, void* out, int ldo, int rows,
                                       int C, int hw, int hwp, cudaStream_t stream) {
  LAVIE_REQUIRE(C % 8 == 0 && ldr % 8 == 0 && ldz % 8 == 0 && ldo % 8 == 0 && hw > 0 && hwp > 0 && hw % hwp == 0 &&
                    rows % hw == 0,
                LAVIE_ERR_SHAPE, "add_gathered: bad shape rows=%d C=%d hw=%d hwp=%d", rows, C, hw, hwp);
  LAVIE_REQUIRE(al16(res) && al16(z) && al16(out), LAVIE_ERR_ALIGN, "add_gathered: alignment");
  const long long total = static_cast<long long>(rows) * (C >> 3);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  launch_pdl(add_gathered_kernel, static_cast<int>(blocks), 256, 0, stream, static_cast<const __nv_bfloat16*>(res), ldr, static_cast<const __nv_bfloat16*>(z), ldz,
      static_cast<__nv_bfloat16*>(out), ldo, rows, C, hw, hwp);
  return lavie_check_launch("add_gathered_kernel");
}

extern "C" int lavie_groupnorm_reduce(const float* partial, int samples, int chunks, int groups, double* sums,
                                      cudaStream_t stream) {
  LAVIE_REQUIRE(groups <= 64 && groups % 4 == 0 && samples > 0 && chunks > 0, LAVIE_ERR_SHAPE, "groupnorm_reduce: shape");
  launch_pdl(gn_reduce_kernel, samples, 256, 0, stream, partial, chunks, groups, sums);
  return lavie_check_launch("gn_reduce_kernel");
}

extern "C" int lavie_groupnorm_finalize_sums(const double* sums, int samples, int groups, int C,
                                             long long count_per_group, const float* gamma, const float* beta,
                                             float eps, float* scale_shift, cudaStream_t stream) {
  LAVIE_REQUIRE(groups <= 64 && C % groups == 0 && count_per_group > 0, LAVIE_ERR_SHAPE, "groupnorm_finalize_sums: shape");
  launch_pdl(gn_finalize_sums_kernel, samples, 256, 0, stream, sums, groups, C, 1.0 / static_cast<double>(count_per_group), gamma,
                                                       beta, eps, scale_shift);
  return lavie_check_launch("gn_finalize_sums_kernel");
}

