// tcgen05 GEMM / implicit-GEMM 3x3 convolution for the LaVie denoiser (sm_100a).
//
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )        bf16 operands, fp32 accumulation in TMEM, bf16 out
//
// Replaces every nn.Linear / 1x1 conv / 3x3 InflatedConv3d call of the reference hot path
// (base/models/attention.py:95-104,328,356 ; resnet.py:13-21,146,162,171-175 ; diffusers FeedForward).
//
// Structure: persistent, warp-specialised CTA of 192 threads, one CTA per SM.
//   warp 0      : TMA producer   (A tile 128x64, W tile BLOCK_N x 64, 128-byte swizzle, ring of STAGES)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma M=128,N=BLOCK_N,K=16) + TMEM owner
//   warps 2..5  : epilogue       (tcgen05.ld 32 lanes x 32 columns -> bias / time-bias / GEGLU / residual -> bf16)
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// A-operand modes
//   plain : A is [M, K] row-major (row stride lda); optionally split along K over two sources
//           (the folded torch.cat([h, skip], dim=1) in front of a 1x1 shortcut, unet_blocks.py:538,630).
//   conv3 : A is the channels-last feature map [NF, H, W, C]; K = 9*C ordered (kh, kw, c); every K block is one
//           (tap, 64-channel) slab fetched with 4-D TMA boxes whose out-of-bounds rows/columns are zero-filled
//           by the hardware = the conv's zero padding.  An M tile is 128 consecutive output pixels =
//           128/W whole image rows.
#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARP0 = 2;

struct GemmParams {
  int M, N, K;
  int num_k_blocks;
  int k_split_blocks;        // plain mode: K blocks [0, k_split) come from A0, the rest from A1
  int k_rot;                 // K-loop rotation stride per M tile (0 = off)
  int m_tiles, n_tiles;
  // conv3 mode
  int conv;                  // 0 plain, 1 conv3x3 stride 1 pad 1
  int c_blocks;              // Cin / 64
  int img_h, img_w;
  int box_h;                 // image rows per TMA box
  int boxes_per_tile;        // 128 / (box_h * W)
  int row_groups_per_img;    // H / box_h
  // epilogue
  const float* bias;         // [N] or null
  const float* row_bias;     // [M / rows_per_batch, N] or null  (time embedding add, resnet.py:187-190)
  int rows_per_batch;
  int ld_row_bias;
  const __nv_bfloat16* residual;   // [M, ldr] or null
  int ldr;
  int geglu;                 // 1: tile columns [0,BN/2) = value, [BN/2,BN) = gate -> out = value * gelu(gate)
  __nv_bfloat16* out;
  int ldo;
};

template <int BLOCK_N>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int MAX_SMEM = 227 * 1024 - 2048;   // leave room for barriers + alignment slack
  static constexpr int STAGES_RAW = MAX_SMEM / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                  const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using L = SmemLayout<BLOCK_N>;
  constexpr int STAGES = L::STAGES;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    tma_prefetch_desc(&tmap_a1);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);   // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles;
        const int n_blk = tile % p.n_tiles;
        // K-loop rotation: CTAs start their K sweep at different blocks so that the ~74 CTAs sharing one weight
        // tile do not hammer the same L2 lines in lockstep (the sum is order-independent up to fp32 rounding and
        // the mapping is fixed, so results stay deterministic).
        const int kb0 = (m_blk * p.k_rot) % p.num_k_blocks;
        for (int it = 0; it < p.num_k_blocks; ++it) {
          int kb = kb0 + it;
          if (kb >= p.num_k_blocks) kb -= p.num_k_blocks;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if (p.conv) {
            const int tap = kb / p.c_blocks;
            const int c0 = (kb - tap * p.c_blocks) * BLOCK_K;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int box_bytes = p.box_h * p.img_w * BLOCK_K * 2;
            for (int i = 0; i < p.boxes_per_tile; ++i) {
              const int g = m_blk * p.boxes_per_tile + i;
              const int img = g / p.row_groups_per_img;
              const int y0 = (g - img * p.row_groups_per_img) * p.box_h;
              tma_load_4d(a_dst + i * box_bytes, &tmap_a0, &full_bar[stage], c0, dx, y0 + dy, img);
            }
          } else if (kb < p.k_split_blocks) {
            tma_load_2d(a_dst, &tmap_a0, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          } else {
            tma_load_2d(a_dst, &tmap_a1, &full_bar[stage], (kb - p.k_split_blocks) * BLOCK_K, m_blk * BLOCK_M);
          }
          tma_load_2d(b_dst, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          const uint64_t a_desc = umma_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = umma_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advancing 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the >>4 address field
            umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);          // frees the smem stage once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);              // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile % p.n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      const float* rb = (p.row_bias != nullptr && row_ok)
                            ? p.row_bias + static_cast<size_t>(row / p.rows_per_batch) * p.ld_row_bias
                            : nullptr;
      if (!p.geglu) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c * 32, v);
          tmem_wait_ld();
          const int col0 = n_blk * BLOCK_N + c * 32;
          if (row_ok && col0 < p.N) {
            const __nv_bfloat16* res = p.residual ? p.residual + static_cast<size_t>(row) * p.ldr + col0 : nullptr;
            __nv_bfloat16* dst = p.out + static_cast<size_t>(row) * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < p.N) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
                if (p.bias) {
                  const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col0 + j);
                  const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col0 + j + 4);
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                if (rb) {
                  const float4 b0 = *reinterpret_cast<const float4*>(rb + col0 + j);
                  const float4 b1 = *reinterpret_cast<const float4*>(rb + col0 + j + 4);
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                if (res) {
                  const uint4 r = *reinterpret_cast<const uint4*>(res + j);
                  const float2 r0 = unpack_bf16(r.x), r1 = unpack_bf16(r.y), r2 = unpack_bf16(r.z),
                               r3 = unpack_bf16(r.w);
                  f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
                  f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
                }
                uint4 o;
                o.x = pack_bf16(f[0], f[1]);
                o.y = pack_bf16(f[2], f[3]);
                o.z = pack_bf16(f[4], f[5]);
                o.w = pack_bf16(f[6], f[7]);
                *reinterpret_cast<uint4*>(dst + j) = o;
              }
            }
          }
        }
      } else {
        // GEGLU: value columns [0, BN/2), gate columns [BN/2, BN) of the same tile (weights interleaved on the
        // host); out[:, n_blk*BN/2 + j] = (value + b) * gelu_erf(gate + b')   (diffusers GEGLU, mirror at
        // vsr/models/diffusers_attention.py:811-822)
        constexpr int HALF = BLOCK_N / 2;
#pragma unroll 1
        for (int c = 0; c < HALF / 32; ++c) {
          uint32_t v[32], g[32];
          tmem_ld_32x32(t_row + c * 32, v);
          tmem_ld_32x32(t_row + HALF + c * 32, g);
          tmem_wait_ld();
          const int colv = n_blk * BLOCK_N + c * 32;          // column in the interleaved weight space
          const int colo = n_blk * HALF + c * 32;             // output column
          if (row_ok && colv < p.N) {
            __nv_bfloat16* dst = p.out + static_cast<size_t>(row) * p.ldo + colo;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float val = __uint_as_float(v[j + e]);
                float gate = __uint_as_float(g[j + e]);
                if (p.bias) {
                  val += p.bias[colv + j + e];
                  gate += p.bias[colv + HALF + j + e];
                }
                f[e] = val * gelu_erf_f(gate);
              }
              uint4 o;
              o.x = pack_bf16(f[0], f[1]);
              o.y = pack_bf16(f[2], f[3]);
              o.z = pack_bf16(f[4], f[5]);
              o.w = pack_bf16(f[6], f[7]);
              *reinterpret_cast<uint4*>(dst + j) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N>
int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmParams& p,
                int num_sms, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL);
    LAVIE_REQUIRE(e == cudaSuccess, LAVIE_ERR_CUDA, "cudaFuncSetAttribute(gemm<%d>): %s", BLOCK_N,
                  cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_tcgen05<BLOCK_N><<<grid, NUM_THREADS, L::TOTAL, stream>>>(a0, a1, b, p);
  return lavie_check_launch("gemm_bf16_tcgen05");
}

int pick_block_n(int M, int N, int forced) {
  if (forced) return forced;
  // minimise (waves x per-tile cost); per-tile cost ~ BLOCK_N (MMA time) with a small fixed overhead
  const int cands[5] = {256, 192, 160, 128, 64};
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  double best = 1e30;
  int best_bn = 128;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    const int n_tiles = (N + bn - 1) / bn;
    const long tiles = static_cast<long>(m_tiles) * n_tiles;
    const long waves = (tiles + 147) / 148;
    const double cost = static_cast<double>(waves) * (bn + 24);
    if (cost < best - 1e-9) {
      best = cost;
      best_bn = bn;
    }
  }
  return best_bn;
}

int g_k_rot = 7;
int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int dispatch(int bn, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmParams& p,
             cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_gemm<64>(a0, a1, b, p, num_sms(), stream);
    case 128: return launch_gemm<128>(a0, a1, b, p, num_sms(), stream);
    case 160: return launch_gemm<160>(a0, a1, b, p, num_sms(), stream);
    case 192: return launch_gemm<192>(a0, a1, b, p, num_sms(), stream);
    case 256: return launch_gemm<256>(a0, a1, b, p, num_sms(), stream);
    default: lavie_set_error("unsupported BLOCK_N %d", bn); return LAVIE_ERR_SHAPE;
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int make_weight_map(CUtensorMap* map, const void* w, int N, int K, int bn) {
  const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
  const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
  const uint32_t box[2] = {BLOCK_K, static_cast<uint32_t>(bn)};
  return lavie_make_tmap(map, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int fill_epilogue(GemmParams& p, const lavie_epilogue* ep, int M, int N, void* out, int ldo) {
  p.bias = nullptr; p.row_bias = nullptr; p.rows_per_batch = 1; p.ld_row_bias = N; p.residual = nullptr; p.ldr = 0; p.geglu = 0;
  if (ep) {
    p.bias = ep->bias;
    p.row_bias = ep->row_bias;
    p.rows_per_batch = ep->rows_per_batch > 0 ? ep->rows_per_batch : 1;
    p.ld_row_bias = ep->ld_row_bias > 0 ? ep->ld_row_bias : N;
    p.residual = static_cast<const __nv_bfloat16*>(ep->residual);
    p.ldr = ep->ld_residual;
    p.geglu = ep->geglu;
    LAVIE_REQUIRE(!p.residual || (aligned16(p.residual) && p.ldr % 8 == 0), LAVIE_ERR_ALIGN,
                  "gemm: residual must be 16-byte aligned with ld %% 8 == 0");
    LAVIE_REQUIRE(!p.bias || aligned16(p.bias), LAVIE_ERR_ALIGN, "gemm: bias must be 16-byte aligned");
    LAVIE_REQUIRE(!p.row_bias || (aligned16(p.row_bias) && p.ld_row_bias % 4 == 0), LAVIE_ERR_ALIGN,
                  "gemm: row_bias must be 16-byte aligned with ld %% 4 == 0");
    LAVIE_REQUIRE(!(p.geglu && (p.residual || p.row_bias)), LAVIE_ERR_SHAPE,
                  "gemm: GEGLU epilogue cannot be combined with residual/row_bias");
  }
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  LAVIE_REQUIRE(aligned16(out) && ldo % 8 == 0, LAVIE_ERR_ALIGN, "gemm: out must be 16-byte aligned, ldo %% 8 == 0");
  (void)M; (void)N;
  return LAVIE_OK;
}

}  // namespace

extern "C" int lavie_gemm_bf16(const void* a0, int lda0, int k0, const void* a1, int lda1, int k1, const void* w,
                               void* out, int ldo, int M, int N, const lavie_epilogue* ep, int block_n,
                               cudaStream_t stream) {
  const int K = k0 + k1;
  LAVIE_REQUIRE(M > 0 && N > 0 && K > 0, LAVIE_ERR_SHAPE, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  LAVIE_REQUIRE(N % 8 == 0 && k0 % 8 == 0 && k1 % 8 == 0, LAVIE_ERR_SHAPE, "gemm: N, K must be multiples of 8");
  LAVIE_REQUIRE(k1 == 0 || k0 % BLOCK_K == 0, LAVIE_ERR_SHAPE, "gemm: split-K source boundary must be a multiple of 64");
  LAVIE_REQUIRE(aligned16(a0) && aligned16(w) && lda0 % 8 == 0 && (k1 == 0 || (aligned16(a1) && lda1 % 8 == 0)),
                LAVIE_ERR_ALIGN, "gemm: operands must be 16-byte aligned with ld %% 8 == 0");
  GemmParams p{};
  int bn = ep && ep->geglu ? 256 : pick_block_n(M, N, block_n);
  LAVIE_REQUIRE(!(ep && ep->geglu) || N % 256 == 0, LAVIE_ERR_SHAPE, "gemm: GEGLU needs N %% 256 == 0");
  p.M = M; p.N = N; p.K = K;
  p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.k_split_blocks = k1 ? k0 / BLOCK_K : p.num_k_blocks;
  p.k_rot = g_k_rot;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = (N + bn - 1) / bn;
  p.conv = 0;
  int rc = fill_epilogue(p, ep, M, N, out, ldo);
  if (rc) return rc;
  CUtensorMap ma0, ma1, mb;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(k0), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda0) * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_M};
    rc = lavie_make_tmap(&ma0, a0, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (k1) {
    const uint64_t dims[2] = {static_cast<uint64_t>(k1), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda1) * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_M};
    rc = lavie_make_tmap(&ma1, a1, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    ma1 = ma0;
  }
  rc = make_weight_map(&mb, w, N, K, bn);
  if (rc) return rc;
  return dispatch(bn, ma0, ma1, mb, p, stream);
}

extern "C" int lavie_debug_set(int what, int value) {
  if (what == 0) g_k_rot = value;
  return 0;
}

extern "C" int lavie_conv3x3_supported(int H, int W, int C) {
  if (C % BLOCK_K != 0) return 0;
  if (W > BLOCK_M || BLOCK_M % W != 0 || W % 8 != 0) return 0;
  return 1;
}

extern "C" int lavie_conv3x3_bf16(const void* x, int NF, int H, int W, int C, const void* w, void* out, int ldo,
                                  int N, const lavie_epilogue* ep, int block_n, cudaStream_t stream) {
  LAVIE_REQUIRE(lavie_conv3x3_supported(H, W, C), LAVIE_ERR_SHAPE,
                "conv3x3: TMA path needs C %% 64 == 0 and W in {8,16,32,64,128} (got H=%d W=%d C=%d)", H, W, C);
  LAVIE_REQUIRE(N % 8 == 0 && aligned16(x) && aligned16(w), LAVIE_ERR_ALIGN, "conv3x3: alignment");
  const int M = NF * H * W;
  GemmParams p{};
  const int bn = pick_block_n(M, N, block_n);
  p.M = M; p.N = N; p.K = 9 * C;
  p.num_k_blocks = 9 * (C / BLOCK_K);
  p.k_split_blocks = p.num_k_blocks;
  p.k_rot = g_k_rot;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = (N + bn - 1) / bn;
  p.conv = 1;
  p.c_blocks = C / BLOCK_K;
  p.img_h = H; p.img_w = W;
  // largest number of whole image rows per TMA box that divides both H and the 128/W rows of an M tile
  const int rows_per_tile = BLOCK_M / W;
  int box_h = 1;
  for (int h = rows_per_tile; h >= 1; --h) {
    if (H % h == 0 && rows_per_tile % h == 0) { box_h = h; break; }
  }
  p.box_h = box_h;
  p.boxes_per_tile = rows_per_tile / box_h;
  p.row_groups_per_img = H / box_h;
  int rc = fill_epilogue(p, ep, M, N, out, ldo);
  if (rc) return rc;
  LAVIE_REQUIRE(!(ep && ep->geglu), LAVIE_ERR_SHAPE, "conv3x3: GEGLU epilogue not supported");
  CUtensorMap ma, mb;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(NF)};
    const uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2,
                                 static_cast<uint64_t>(H) * W * C * 2};
    const uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(W), static_cast<uint32_t>(box_h), 1};
    rc = lavie_make_tmap(&ma, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  rc = make_weight_map(&mb, w, N, 9 * C, bn);
  if (rc) return rc;
  return dispatch(bn, ma, ma, mb, p, stream);
}
