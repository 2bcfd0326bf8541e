// Attention cores of the LaVie transformer block (sm_100a).
//
//  lavie_attention_bf16          flash-style softmax(q k^T * scale) v on tcgen05 / TMEM with TMA-staged Q, K, V tiles:
//                                spatial self-attention (S = H*W tokens per frame) and CLIP cross-attention
//                                (77 keys, shared across the frames of a batch item).
//                                Replaces CrossAttention._attention (base/models/attention.py:209-239).
//  lavie_temporal_attention_bf16 per-pixel attention over the F frames with fused q-scale, RoPE and rel-pos bias.
//                                Replaces TemporalAttention._attention (attention.py:634-667) and the two
//                                (b f) d c <-> (b d) f c transposes around it (:549-555): frames are simply read with
//                                a row stride of H*W.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// Flash attention, one CTA = one 128-query tile of one (batch, head).  128 threads, thread t owns query row t
// (TMEM lane t).  Per 128-key tile:  S = Q K^T (tcgen05, fp32 in TMEM) -> online softmax in registers ->
// P (bf16, 128B-swizzled K-major in smem) -> O_tile = P V (tcgen05, V is the MN-major B operand) -> rescaled
// accumulation in registers.  Two co-resident CTAs per SM (d = 40) overlap one CTA's softmax with the other's MMAs.
// ------------------------------------------------------------------------------------------------------------
constexpr int ATT_M = 128;       // queries per CTA
constexpr int ATT_N = 64;        // keys per tile: small tiles -> 4 co-resident CTAs per SM hide the MMA round trips
constexpr int CHUNK_BYTES = 128 * 128;      // one 64-column (128-byte) slab of a 128-row tile (Q, P)
constexpr int KV_CHUNK_BYTES = ATT_N * 128;  // the same slab of a K / V tile

struct AttnParams {
  int Sq, Sk, d, head_pitch, kv_batch_div;
  long long o_seq_stride, o_batch_stride;   // elements: output row of (batch, query) = batch*o_batch_stride + query*o_seq_stride
  // SparseCausalAttention (interpolation/models/attention.py:611-664): batch = (video, frame); the keys of frame f are
  // [all Sk keys of frame 0 | all Sk keys of frame max(f - 1, 0)] of the same video.  0 = ordinary attention.
  int sc_frames;
  int sc_halo;         // frame-sharded SparseCausal: k / v = [frame 0 | previous rank's last frame | local frames]; 2 = first rank
  int swap_dims;       // tensor maps are (col, batch, seq) instead of (col, seq, batch): batch stride < sequence stride
  float scale_log2;
  __nv_bfloat16* o;
  long long* timeline;   // debug (tools/attn_timeline.py): clock64 stamps of one mid-grid CTA, else nullptr
};

template <int DK>
struct AttnCfg {
  static constexpr int NC = (DK + 63) / 64;
  static constexpr int P_CHUNKS = (ATT_N + 63) / 64;
  static constexpr int SMEM = NC * CHUNK_BYTES + 2 * NC * KV_CHUNK_BYTES + 1024 + 128;   // Q, K, V (+ align, barriers)
  static constexpr uint32_t TMEM_COLS = (ATT_N + DK) <= 128 ? 128 : (ATT_N + DK) <= 256 ? 256 : 512;
  static constexpr int MIN_BLOCKS = DK <= 64 ? 4 : (DK <= 96 ? 2 : 1);
};

constexpr int ATT_THREADS = 160;     // warps 0-3: softmax (thread = query row), warp 4: TMA + MMA issue

template <int DK, int POLY, bool ONES>
__global__ void __launch_bounds__(ATT_THREADS, AttnCfg<DK>::MIN_BLOCKS)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  pdl_launch_dependents();
  using Cfg = AttnCfg<DK>;
  constexpr int NC = Cfg::NC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NC * CHUNK_BYTES;
  uint8_t* sV = sK + NC * KV_CHUNK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NC * KV_CHUNK_BYTES);
  uint64_t* bar_q = bars + 0;        // Q landed (TMA)
  uint64_t* bar_k = bars + 1;        // K(j) landed, phase j
  uint64_t* bar_v = bars + 2;        // V(j) landed, phase j
  uint64_t* bar_s_full = bars + 3;   // S(j) = Q K(j)^T complete in TMEM (tcgen05.commit), phase j
  uint64_t* bar_p_full = bars + 5;   // all 4 softmax warps stored P(j) to TMEM (and rescaled O if needed), phase j
  uint64_t* bar_pv = bars + 6;       // O += P(j) V(j) retired (tcgen05.commit), phase j: V buffer free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
  const int kv_batch = batch / p.kv_batch_div;
  // key tiles: one segment of ceil(Sk / 64) tiles, or two of them for SparseCausal attention (first frame | former frame);
  // a segment's last tile may be partial (rows past Sk are zero-filled by the TMA and masked in the softmax)
  const int seg_tiles = (p.Sk + ATT_N - 1) / ATT_N;
  const int n_kv = p.sc_frames > 0 ? 2 * seg_tiles : seg_tiles;
  const int sc_f = p.sc_frames > 0 ? batch % p.sc_frames : 0;
  // sc_halo: one video per launch, local frame b lives at k/v index b + 2, halo frame 0 at index 0, the frame in front
  // of local frame 0 at index 1 (on the first rank frame 0 is its own former frame: index 2)
  const int kv_batch_seg0 = p.sc_halo ? 0 : (p.sc_frames > 0 ? batch - sc_f : kv_batch);
  const int kv_batch_seg1 = p.sc_halo ? (batch > 0 ? batch + 1 : (p.sc_halo == 2 ? 2 : 1))
                                      : (p.sc_frames > 0 ? (sc_f > 0 ? batch - 1 : batch) : kv_batch);
  const int col0 = head * p.head_pitch;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s_full, 1);
    mbar_init(bar_p_full, 4);
    mbar_init(bar_pv, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                    // prologue above overlaps the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;             // columns [0, ATT_N)
  const uint32_t tmem_o = tmem_base + ATT_N;     // columns [ATT_N, ATT_N + DK)
  const uint32_t tmem_p = tmem_base;             // bf16 P(j) overwrites the first ATT_N / 2 columns of S(j) in place
  const bool tl = p.timeline != nullptr && lane == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2 &&
                  blockIdx.z == gridDim.z / 2;

  if (warp == 4) {
    // =============================== control warp: TMA loads + MMA issue ===============================
    // The whole warp runs the loop (descriptors stay in uniform registers), one elected lane issues.  The
    // probabilities never touch shared memory: the softmax warps store P(j) as bf16 over S(j) in TENSOR memory and the
    // P.V MMA takes its A operand from there; S(j+1) = Q K(j+1)^T is issued right behind it (tcgen05.mma executes in
    // issue order, so it cannot overwrite P(j) before P(j) V(j) has consumed it).
    constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_M, ATT_N, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_M, DK, 0, 1);
    const uint64_t desc_q = umma_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t desc_k = umma_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t desc_v = umma_desc_sw128(smem_u32(sV), KV_CHUNK_BYTES, 1024);
    auto issue_qk = [&]() {
#pragma unroll
      for (int kk = 0; kk < DK / 16; ++kk) {
        const uint32_t qoff = (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32;
        const uint32_t koff = (kk >> 2) * KV_CHUNK_BYTES + (kk & 3) * 32;
        umma_bf16(tmem_s, desc_q + (qoff >> 4), desc_k + (koff >> 4), idesc_qk, kk != 0);
      }
    };
    auto load_rows = [&](uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int col, int row, int b) {
      if (p.swap_dims) tma_load_3d(dst, m, bar, col, b, row);
      else tma_load_3d(dst, m, bar, col, row, b);
    };
    auto load_k = [&](int j) {
      const int seg = j >= seg_tiles, jr = j - seg * seg_tiles;
      mbar_expect_tx(bar_k, NC * KV_CHUNK_BYTES);
      for (int c = 0; c < NC; ++c)
        load_rows(sK + c * KV_CHUNK_BYTES, &tmap_k, bar_k, col0 + c * 64, jr * ATT_N, seg ? kv_batch_seg1 : kv_batch_seg0);
    };
    auto load_v = [&](int j) {
      const int seg = j >= seg_tiles, jr = j - seg * seg_tiles;
      mbar_expect_tx(bar_v, NC * KV_CHUNK_BYTES);
      for (int c = 0; c < NC; ++c)
        load_rows(sV + c * KV_CHUNK_BYTES, &tmap_v, bar_v, col0 + c * 64, jr * ATT_N, seg ? kv_batch_seg1 : kv_batch_seg0);
    };
    if (elect_one()) {
      mbar_expect_tx(bar_q, NC * CHUNK_BYTES);
      for (int c = 0; c < NC; ++c) load_rows(sQ + c * CHUNK_BYTES, &tmap_q, bar_q, col0 + c * 64, q_tile * ATT_M, batch);
      load_k(0);
      load_v(0);
    }
    __syncwarp();
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    if (elect_one()) {
      issue_qk();
      umma_commit(bar_s_full);
    }
    __syncwarp();
    if (n_kv > 1) {
      mbar_wait(bar_s_full, 0);                  // Q K(0)^T retired: the K buffer is free
      if (elect_one()) load_k(1);
      __syncwarp();
    }
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int kv_len = min(ATT_N, p.Sk - (j >= seg_tiles ? j - seg_tiles : j) * ATT_N);
      long long* tl_row = p.timeline + j * 8;
      mbar_wait(bar_v, ph);                      // V(j) landed
      if (ONES) {                                // V(j)[key][column d] = 1.0 (bf16), 128-byte-swizzled address
#pragma unroll
        for (int key = lane; key < ATT_N; key += 32) {
          const int unit = ((p.d >> 3) & 7) ^ (key & 7);
          *reinterpret_cast<uint16_t*>(sV + (p.d >> 6) * KV_CHUNK_BYTES + key * 128 + unit * 16 + (p.d & 7) * 2) = 0x3F80;
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (j + 1 < n_kv) mbar_wait(bar_k, ph ^ 1);  // K(j+1) landed (requested a whole tile ago)
      if (tl) tl_row[5] = clock64();
      mbar_wait(bar_p_full, ph);                 // P(j) in tensor memory, O rescaled where needed
      tc_fence_after();
      if (elect_one()) {
        const int nk = (kv_len + 15) >> 4;       // 16 keys = 8 packed columns of P per MMA
        if (kv_len == ATT_N) {
#pragma unroll
          for (int kk = 0; kk < ATT_N / 16; ++kk)
            umma_bf16_ts(tmem_o, tmem_p + kk * 8, desc_v + ((kk * 2048) >> 4), idesc_pv, (j > 0 || kk != 0));
        } else {
          for (int kk = 0; kk < nk; ++kk)
            umma_bf16_ts(tmem_o, tmem_p + kk * 8, desc_v + ((kk * 2048) >> 4), idesc_pv, (j > 0 || kk != 0));
        }
        umma_commit(bar_pv);
        if (j + 1 < n_kv) {                      // in issue order behind P(j) V(j): S(j+1) overwrites P(j) only afterwards
          issue_qk();
          umma_commit(bar_s_full);
        }
      }
      __syncwarp();
      if (tl) tl_row[6] = clock64();
      if (j + 1 < n_kv) {
        mbar_wait(bar_pv, ph);                   // P(j) V(j) retired: V buffer free
        if (elect_one()) load_v(j + 1);
        __syncwarp();
      }
      if (j + 2 < n_kv) {
        mbar_wait(bar_s_full, ph ^ 1);           // Q K(j+1)^T retired: K buffer free -> K(j+2) has a whole tile to land
        if (elect_one()) load_k(j + 2);
        __syncwarp();
      }
      if (tl) tl_row[7] = clock64();
    }
  } else {
    // =============================== softmax warps: one query row per thread ===========================
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    // The output accumulator stays in TMEM across key tiles (P.V accumulates with use_acc); m_used is the exponent
    // offset currently baked into it.  It is only moved -- and O rescaled in TMEM -- when the running maximum has
    // grown by more than 2^8 since (lazy rescale): probabilities are then at most 256, harmless in bf16 / fp32, and
    // the result O / l is unchanged because numerator and denominator use the same offset.
    float m_used = -INFINITY, l_run = 0.f;
    // d < DK (d = 40 in a 48-wide head slot): a padded V column is set to 1.0 so that the P.V MMA itself produces the
    // softmax denominator (sum of the bf16-rounded probabilities) in accumulator column d -- no per-element FADD.
    constexpr bool ones_col = ONES;              // host: p.d < DK
    const float sc = p.scale_log2;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int kv_len = min(ATT_N, p.Sk - (j >= seg_tiles ? j - seg_tiles : j) * ATT_N);
      long long* tl_row = p.timeline + j * 8;
      if (tl && warp == 0) tl_row[0] = clock64();
      mbar_wait(bar_s_full, ph);
      tc_fence_after();
      if (tl && warp == 0) tl_row[1] = clock64();
      // ---- S(j) -> registers, read once ----
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(tmem_s + lane_base, s0);
      tmem_ld_32x32(tmem_s + lane_base + 32, s1);
      tmem_wait_ld();
      if (tl && warp == 0) tl_row[2] = clock64();
      const bool full = (kv_len == ATT_N);         // CTA-uniform
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (full) {
        float mx2 = -INFINITY, mx3 = -INFINITY;    // four short FMNMX3 chains instead of two long ones
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          mx0 = fmax3(mx0, __uint_as_float(s0[e]), __uint_as_float(s0[e + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s1[e]), __uint_as_float(s1[e + 1]));
          mx2 = fmax3(mx2, __uint_as_float(s0[e + 2]), __uint_as_float(s0[e + 3]));
          mx3 = fmax3(mx3, __uint_as_float(s1[e + 2]), __uint_as_float(s1[e + 3]));
        }
        mx0 = fmaxf(mx0, mx2);
        mx1 = fmaxf(mx1, mx3);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e < kv_len) mx0 = fmaxf(mx0, __uint_as_float(s0[e]));
          if (32 + e < kv_len) mx1 = fmaxf(mx1, __uint_as_float(s1[e]));
        }
      }
      const float m_tile = fmaxf(mx0, mx1) * sc;   // sc > 0
      const bool grow = m_tile > m_used + 8.0f;    // always true for j == 0 (m_used = -inf)
      if (j > 0 && __any_sync(0xffffffffu, grow)) {  // S(j) was issued behind P(j-1) V(j-1): O is stable here
        const float alpha = grow ? fast_exp2(m_used - m_tile) : 1.0f;
#pragma unroll
        for (int c = 0; c < DK / 16; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(tmem_o + lane_base + c * 16, v);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st_32x16(tmem_o + lane_base + c * 16, v);
        }
        tmem_wait_st();
        l_run *= alpha;
      }
      if (grow) m_used = m_tile;
      const float neg_m = -m_used;

      // ---- p = exp2(s*scale - m_used), packed to bf16 in place (s0/s1[0..15] hold the packed pairs) ----
      // Full tiles: every POLY-th exponential runs on the FMA pipe (exp2_poly), the rest on the MUFU.
      float rs0 = 0.f, rs1 = 0.f;
      auto soft_half = [&](uint32_t (&sv)[32], int c) {
        if (full) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float x0 = fmaf(__uint_as_float(sv[e]), sc, neg_m);
            const float x1 = fmaf(__uint_as_float(sv[e + 1]), sc, neg_m);
            const float p0 = (POLY > 0 && (e % POLY) == POLY - 1) ? exp2_poly(x0) : fast_exp2(x0);
            const float p1 = (POLY > 0 && ((e + 1) % POLY) == POLY - 1) ? exp2_poly(x1) : fast_exp2(x1);
            if (!ones_col) {
              rs0 += p0;
              rs1 += p1;
            }
            sv[e >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = (c * 32 + e < kv_len) ? fast_exp2(fmaf(__uint_as_float(sv[e]), sc, neg_m)) : 0.f;
            const float p1 = (c * 32 + e + 1 < kv_len) ? fast_exp2(fmaf(__uint_as_float(sv[e + 1]), sc, neg_m)) : 0.f;
            rs0 += p0;
            rs1 += p1;
            sv[e >> 1] = pack_bf16(p0, p1);
          }
        }
      };
      soft_half(s0, 0);
      soft_half(s1, 1);
      l_run += rs0 + rs1;
      if (tl && warp == 0) tl_row[3] = clock64();
      // P(j) -> tensor memory, bf16 pairs over the first 32 columns of S(j) (this warp's own 32 lanes only)
      {
        uint32_t pk[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          pk[e] = s0[e];
          pk[16 + e] = s1[e];
        }
        tmem_st_32x32(tmem_p + lane_base, pk);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full);
      if (tl && warp == 0) tl_row[4] = clock64();
    }
    // ---- O complete ----
    mbar_wait(bar_pv, (n_kv - 1) & 1);
    tc_fence_after();
    float denom = l_run;
    if (ones_col) {                            // the ones column of V accumulated the denominator in O[:, d]
      uint32_t v[16];
      tmem_ld_32x16(tmem_o + lane_base + (p.d >> 4) * 16, v);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (e == (p.d & 15)) denom = __uint_as_float(v[e]);
    }
    const float inv = 1.f / denom;
    const int qrow = q_tile * ATT_M + tid;
    __nv_bfloat16* dst = p.o + static_cast<size_t>(batch) * p.o_batch_stride + static_cast<size_t>(qrow) * p.o_seq_stride +
                         head * p.d;
#pragma unroll
    for (int c = 0; c < DK / 16; ++c) {
      uint32_t v[16];
      tmem_ld_32x16(tmem_o + lane_base + c * 16, v);
      tmem_wait_ld();
      if (qrow < p.Sq) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c * 16 + h * 8 < p.d) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[8 * h]) * inv, __uint_as_float(v[8 * h + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(v[8 * h + 2]) * inv, __uint_as_float(v[8 * h + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(v[8 * h + 4]) * inv, __uint_as_float(v[8 * h + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(v[8 * h + 6]) * inv, __uint_as_float(v[8 * h + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c * 16 + h * 8) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int DK, int POLY, bool ONES>
int launch_attn_inst(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& p, int batch,
                     int heads, cudaStream_t stream) {
  using Cfg = AttnCfg<DK>;
  static LavieSmemConfig configured;
  const int rc_cfg = lavie_config_smem(attn_fwd_kernel<DK, POLY, ONES>, Cfg::SMEM, &configured, "attn_fwd_kernel");
  if (rc_cfg) return rc_cfg;
  dim3 grid((p.Sq + ATT_M - 1) / ATT_M, heads, batch);
  launch_pdl(attn_fwd_kernel<DK, POLY, ONES>, grid, ATT_THREADS, Cfg::SMEM, stream, mq, mk, mv, p);
  return lavie_check_launch("attn_fwd_kernel");
}

// every g_lavie_attn_poly-th exponential of a full key tile is evaluated on the FMA pipe (0: all on the MUFU)
template <int DK>
int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& p, int batch,
                int heads, cudaStream_t stream) {
  if (p.d < DK) {                 // a spare padded column carries the softmax denominator through the P.V MMA
    // default (-1): every 4th exponential on the FMA pipe -- 2 % faster at 2560 keys now that P stays in tensor memory
    if (g_lavie_attn_poly == 4 || g_lavie_attn_poly < 0) return launch_attn_inst<DK, 4, true>(mq, mk, mv, p, batch, heads, stream);
    return launch_attn_inst<DK, 0, true>(mq, mk, mv, p, batch, heads, stream);
  }
  if (g_lavie_attn_poly == 4) return launch_attn_inst<DK, 4, false>(mq, mk, mv, p, batch, heads, stream);
  return launch_attn_inst<DK, 0, false>(mq, mk, mv, p, batch, heads, stream);
}

int make_qkv_map(CUtensorMap* map, const void* base, long long seq_stride, long long batch_stride, int cols, int S,
                 int nbatch, int box_rows, bool swap) {
  if (swap) {      // strides must ascend with the dimension index: (col, batch, seq); the box is still box_rows tokens
    const uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(nbatch), static_cast<uint64_t>(S)};
    const uint64_t strides[2] = {static_cast<uint64_t>(batch_stride) * 2, static_cast<uint64_t>(seq_stride) * 2};
    const uint32_t box[3] = {64, 1, static_cast<uint32_t>(box_rows)};
    return lavie_make_tmap(map, base, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  const uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(S), static_cast<uint64_t>(nbatch)};
  const uint64_t strides[2] = {static_cast<uint64_t>(seq_stride) * 2, static_cast<uint64_t>(batch_stride) * 2};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
  return lavie_make_tmap(map, base, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// ------------------------------------------------------------------------------------------------------------
// Temporal attention: one warp per (batch item, pixel, head); F <= 64 frames.
// ------------------------------------------------------------------------------------------------------------
struct TempParams {
  const __nv_bfloat16* qkv;
  int ld, k_off, v_off;
  __nv_bfloat16* o;
  int ldo, B, F, HW, heads, d, head_pitch;
  float scale;
  const float* rope;   // [F, rot_pairs, 2]
  int rot_pairs;
  const float* bias;   // [heads, F, F]
  long long items;
};

__global__ void __launch_bounds__(128)
temporal_attn_kernel(const TempParams p) {
  pdl_prologue();
  extern __shared__ float tsm[];
  const int warp_in_blk = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int F = p.F, d = p.d;
  const int dp = d + 1;                       // padded pitch: conflict-free column walks
  float* base = tsm + warp_in_blk * (3 * F * dp + F * (F + 1));
  float* sq = base;
  float* sk = sq + F * dp;
  float* sv = sk + F * dp;
  float* ss = sv + F * dp;                    // [F][F+1]
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long item = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp_in_blk; item < p.items;
       item += warps_total) {
    const int h = static_cast<int>(item % p.heads);
    const long long bp = item / p.heads;
    const int pix = static_cast<int>(bp % p.HW);
    const int b = static_cast<int>(bp / p.HW);
    // ---- load q (scaled, rotated), k (rotated), v ----
    const int half = d >> 1;                  // bf16 pairs per head row
    for (int idx = lane; idx < F * half; idx += 32) {
      const int f = idx / half, pr = idx - f * half;
      const size_t row = (static_cast<size_t>(b) * F + f) * p.HW + pix;
      const __nv_bfloat16* src = p.qkv + row * p.ld + h * p.head_pitch + 2 * pr;
      float2 q2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(src));
      float2 k2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(src + p.k_off));
      const float2 v2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(src + p.v_off));
      q2.x *= p.scale;                        // q scaled BEFORE the rotation (attention.py:640-646)
      q2.y *= p.scale;
      if (pr < p.rot_pairs) {
        const float cs = p.rope[(f * p.rot_pairs + pr) * 2], sn = p.rope[(f * p.rot_pairs + pr) * 2 + 1];
        const float qx = q2.x * cs - q2.y * sn, qy = q2.y * cs + q2.x * sn;
        const float kx = k2.x * cs - k2.y * sn, ky = k2.y * cs + k2.x * sn;
        q2 = make_float2(qx, qy);
        k2 = make_float2(kx, ky);
      }
      sq[f * dp + 2 * pr] = q2.x; sq[f * dp + 2 * pr + 1] = q2.y;
      sk[f * dp + 2 * pr] = k2.x; sk[f * dp + 2 * pr + 1] = k2.y;
      sv[f * dp + 2 * pr] = v2.x; sv[f * dp + 2 * pr + 1] = v2.y;
    }
    __syncwarp();
    // ---- scores + bias ----
    for (int idx = lane; idx < F * F; idx += 32) {
      const int i = idx / F, jn = idx - i * F;
      float s = 0.f;
      for (int c = 0; c < d; ++c) s += sq[i * dp + c] * sk[jn * dp + c];
      ss[i * (F + 1) + jn] = s + p.bias[(h * F + i) * F + jn];
    }
    __syncwarp();
    // ---- softmax per query frame ----
    for (int i = lane; i < F; i += 32) {
      float mx = -INFINITY;
      for (int jn = 0; jn < F; ++jn) mx = fmaxf(mx, ss[i * (F + 1) + jn]);
      float sum = 0.f;
      for (int jn = 0; jn < F; ++jn) {
        const float e = __expf(ss[i * (F + 1) + jn] - mx);
        ss[i * (F + 1) + jn] = e;
        sum += e;
      }
      const float inv = 1.f / sum;
      for (int jn = 0; jn < F; ++jn) ss[i * (F + 1) + jn] *= inv;
    }
    __syncwarp();
    // ---- out = P V ----
    for (int idx = lane; idx < F * half; idx += 32) {
      const int i = idx / half, pr = idx - i * half;
      float ox = 0.f, oy = 0.f;
      for (int jn = 0; jn < F; ++jn) {
        const float w = ss[i * (F + 1) + jn];
        ox += w * sv[jn * dp + 2 * pr];
        oy += w * sv[jn * dp + 2 * pr + 1];
      }
      const size_t row = (static_cast<size_t>(b) * F + i) * p.HW + pix;
      *reinterpret_cast<uint32_t*>(p.o + row * p.ldo + h * d + 2 * pr) = pack_bf16(ox, oy);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------
// Temporal attention, F <= 16 frames: one warp per (batch item, pixel, head) on mma.sync m16n8k16 (the 16x16xd problem
// is far below the 64-row minimum of tcgen05; the kernel is bound by streaming q/k/v once from HBM).
// Q and K are loaded straight from global memory into the A / B fragment layouts; the contraction index is permuted
// (identically for Q and K, so the dot products are unchanged) such that every lane reads 16 contiguous bytes.
// The exp(S) registers are re-used as the A operand of P.V (accumulator layout == A layout); V goes through shared
// memory for the transposing ldmatrix.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_addr));
}
// scale (q only) and rotate one bf16 pair; pair index pr < rot_pairs is rotated by the frame's angle
__device__ __forceinline__ uint32_t scale_rope_pair(uint32_t w, float scale, bool rotate, const float* cs) {
  float2 f = unpack_bf16(w);
  f.x *= scale;
  f.y *= scale;
  if (rotate) {
    const float c = cs[0], sn = cs[1];
    const float x = f.x * c - f.y * sn, y = f.y * c + f.x * sn;
    f.x = x;
    f.y = y;
  }
  return pack_bf16(f.x, f.y);
}

// same with the angle's (cos, sin) already in registers ((1, 0) = no rotation)
__device__ __forceinline__ uint32_t scale_rope_pair_cs(uint32_t w, float scale, float c, float sn) {
  float2 f = unpack_bf16(w);
  f.x *= scale;
  f.y *= scale;
  return pack_bf16(f.x * c - f.y * sn, f.y * c + f.x * sn);
}

template <int DP>   // head pitch (d rounded up to 16): 48, 64, 80, 96, 128, 160
__global__ void __launch_bounds__(128, 4)
temporal_attn_mma_kernel(const TempParams p) {
  pdl_prologue();
  constexpr int VP = DP + 8;                      // smem V row pitch (elements): conflict-free ldmatrix rows
  constexpr int NB32 = DP / 32;                   // 32-column blocks (two k-steps each)
  constexpr bool TAIL16 = (DP % 32) != 0;         // one trailing 16-column block
  constexpr int KSTEPS = DP / 16;
  __shared__ __align__(16) __nv_bfloat16 s_v[4][16 * VP];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int F = p.F;
  __nv_bfloat16* sv = s_v[wib];
  const long long warps_total = static_cast<long long>(gridDim.x) * 4;
  // Per-lane constants hoisted out of the item loop (they were ~50 of the ~65 load instructions per item):
  //  * the rotary (cos, sin) of the lane's two frames for the 4 channel pairs of the first 32-wide block -- with the
  //    model's 32 rotated dims that is every rotated pair; pairs >= rot_pairs get (1, 0);
  //  * the relative-position bias of the lane's 8 score elements (-inf on padded frames, so adding it also masks),
  //    when the warp keeps the same head for all its items (warps_total % heads == 0; the host sizes the grid so).
  float cs_lo0[4][2], cs_hi0[4][2];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int pr = t * 4 + w;
    const bool rot = pr < p.rot_pairs;
    cs_lo0[w][0] = (rot && g < F) ? __ldg(p.rope + (static_cast<size_t>(g) * p.rot_pairs + pr) * 2) : 1.f;
    cs_lo0[w][1] = (rot && g < F) ? __ldg(p.rope + (static_cast<size_t>(g) * p.rot_pairs + pr) * 2 + 1) : 0.f;
    cs_hi0[w][0] = (rot && g + 8 < F) ? __ldg(p.rope + (static_cast<size_t>(g + 8) * p.rot_pairs + pr) * 2) : 1.f;
    cs_hi0[w][1] = (rot && g + 8 < F) ? __ldg(p.rope + (static_cast<size_t>(g + 8) * p.rot_pairs + pr) * 2 + 1) : 0.f;
  }
  float bias_lo[4], bias_hi[4];                   // index nt * 2 + e  <->  key frame nt * 8 + 2 t + e
  auto load_bias = [&](int h) {
    const float* bias_h = p.bias + static_cast<size_t>(h) * F * F;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jn = nt * 8 + 2 * t + e;
        bias_lo[nt * 2 + e] = (jn < F && g < F) ? __ldg(bias_h + g * F + jn) : -INFINITY;
        bias_hi[nt * 2 + e] = (jn < F && g + 8 < F) ? __ldg(bias_h + (g + 8) * F + jn) : -INFINITY;
      }
  };
  const bool fixed_head = (warps_total % p.heads) == 0;
  if (fixed_head) load_bias(static_cast<int>((static_cast<long long>(blockIdx.x) * 4 + wib) % p.heads));
  for (long long item = static_cast<long long>(blockIdx.x) * 4 + wib; item < p.items; item += warps_total) {
    const int h = static_cast<int>(item % p.heads);
    if (!fixed_head) load_bias(h);
    const long long bp = item / p.heads;
    const int pix = static_cast<int>(bp % p.HW);
    const int b = static_cast<int>(bp / p.HW);
    const size_t row0 = static_cast<size_t>(b) * F * p.HW + pix;             // frame f lives at row0 + f*HW
    const __nv_bfloat16* base = p.qkv + row0 * p.ld + h * p.head_pitch;
    const size_t fstride = static_cast<size_t>(p.HW) * p.ld;
    const bool lo_ok = g < F, hi_ok = g + 8 < F;
    const __nv_bfloat16* r_lo = base + static_cast<size_t>(g) * fstride;
    const __nv_bfloat16* r_hi = base + static_cast<size_t>(g + 8) * fstride;

    // ---- every global load of the item is issued before the first use: V (to registers), then Q / K ----
    // (a rolled load -> store loop for V followed by per-block Q/K loads cost ~5 dependent DRAM round trips per item)
    constexpr int VTOT = 16 * (DP / 8);
    constexpr int VIT = (VTOT + 31) / 32;
    uint4 vreg[VIT];
#pragma unroll
    for (int it = 0; it < VIT; ++it) {
      const int i = lane + it * 32;
      const int f = i / (DP / 8), vc = i - f * (DP / 8);
      vreg[it] = make_uint4(0, 0, 0, 0);
      if (i < VTOT && f < F)
        vreg[it] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(f) * fstride + p.v_off + vc * 8));
    }
    constexpr bool PRELOAD = DP <= 96;              // registers: 16 per 32-wide block (+ 8 for the 16-wide tail)
    constexpr int NPRE = PRELOAD ? NB32 : 1;
    uint4 pq_lo[NPRE], pq_hi[NPRE], pk_lo[NPRE], pk_hi[NPRE];
    uint2 tq_lo = make_uint2(0, 0), tq_hi = tq_lo, tk_lo = tq_lo, tk_hi = tq_lo;
    if (PRELOAD) {
#pragma unroll
      for (int blk = 0; blk < NB32; ++blk) {
        const int col = blk * 32 + t * 8;
        pq_lo[blk] = pq_hi[blk] = pk_lo[blk] = pk_hi[blk] = make_uint4(0, 0, 0, 0);
        if (lo_ok) {
          pq_lo[blk] = __ldg(reinterpret_cast<const uint4*>(r_lo + col));
          pk_lo[blk] = __ldg(reinterpret_cast<const uint4*>(r_lo + p.k_off + col));
        }
        if (hi_ok) {
          pq_hi[blk] = __ldg(reinterpret_cast<const uint4*>(r_hi + col));
          pk_hi[blk] = __ldg(reinterpret_cast<const uint4*>(r_hi + p.k_off + col));
        }
      }
      if (TAIL16) {
        const int col = NB32 * 32 + t * 4;
        if (lo_ok) {
          tq_lo = __ldg(reinterpret_cast<const uint2*>(r_lo + col));
          tk_lo = __ldg(reinterpret_cast<const uint2*>(r_lo + p.k_off + col));
        }
        if (hi_ok) {
          tq_hi = __ldg(reinterpret_cast<const uint2*>(r_hi + col));
          tk_hi = __ldg(reinterpret_cast<const uint2*>(r_hi + p.k_off + col));
        }
      }
    }
    // ---- V tile -> shared memory (row-major [frame][DP], 16-byte vectors) ----
    __syncwarp();                                   // the previous item's ldmatrix reads of this buffer are done
#pragma unroll
    for (int it = 0; it < VIT; ++it) {
      const int i = lane + it * 32;
      const int f = i / (DP / 8), vc = i - f * (DP / 8);
      if (i < VTOT) *reinterpret_cast<uint4*>(sv + f * VP + vc * 8) = vreg[it];
    }

    // ---- S = (scale * rope(Q)) . rope(K)^T ----
    float s[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
    const float* cs_lo = p.rope + static_cast<size_t>(g) * p.rot_pairs * 2;
    const float* cs_hi = p.rope + static_cast<size_t>(g + 8) * p.rot_pairs * 2;
#pragma unroll
    for (int blk = 0; blk < NB32; ++blk) {
      const int col = blk * 32 + t * 8;           // this lane's 8 contiguous channels of the 32-wide block
      uint4 q_lo = make_uint4(0, 0, 0, 0), q_hi = q_lo, k_lo = q_lo, k_hi = q_lo;
      if (PRELOAD) {
        q_lo = pq_lo[PRELOAD ? blk : 0]; q_hi = pq_hi[PRELOAD ? blk : 0];
        k_lo = pk_lo[PRELOAD ? blk : 0]; k_hi = pk_hi[PRELOAD ? blk : 0];
      } else {
        if (lo_ok) {
          q_lo = __ldg(reinterpret_cast<const uint4*>(r_lo + col));
          k_lo = __ldg(reinterpret_cast<const uint4*>(r_lo + p.k_off + col));
        }
        if (hi_ok) {
          q_hi = __ldg(reinterpret_cast<const uint4*>(r_hi + col));
          k_hi = __ldg(reinterpret_cast<const uint4*>(r_hi + p.k_off + col));
        }
      }
      uint32_t ql[4] = {q_lo.x, q_lo.y, q_lo.z, q_lo.w}, qh[4] = {q_hi.x, q_hi.y, q_hi.z, q_hi.w};
      uint32_t kl[4] = {k_lo.x, k_lo.y, k_lo.z, k_lo.w}, kh[4] = {k_hi.x, k_hi.y, k_hi.z, k_hi.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        if (blk == 0) {                           // hoisted angles (compile-time branch: the loop is unrolled)
          ql[w] = scale_rope_pair_cs(ql[w], p.scale, cs_lo0[w][0], cs_lo0[w][1]);
          qh[w] = scale_rope_pair_cs(qh[w], p.scale, cs_hi0[w][0], cs_hi0[w][1]);
          kl[w] = scale_rope_pair_cs(kl[w], 1.f, cs_lo0[w][0], cs_lo0[w][1]);
          kh[w] = scale_rope_pair_cs(kh[w], 1.f, cs_hi0[w][0], cs_hi0[w][1]);
        } else {
          const int pr = (col >> 1) + w;          // bf16-pair index inside the head
          const bool rot = pr < p.rot_pairs;
          ql[w] = scale_rope_pair(ql[w], p.scale, rot && lo_ok, cs_lo + 2 * (rot ? pr : 0));
          qh[w] = scale_rope_pair(qh[w], p.scale, rot && hi_ok, cs_hi + 2 * (rot ? pr : 0));
          kl[w] = scale_rope_pair(kl[w], 1.f, rot && lo_ok, cs_lo + 2 * (rot ? pr : 0));
          kh[w] = scale_rope_pair(kh[w], 1.f, rot && hi_ok, cs_hi + 2 * (rot ? pr : 0));
        }
      }
      // words (0,1) form one k-step, words (2,3) the next: a0/a2 = row g, a1/a3 = row g+8; b0/b1 = key frame g (+8)
      const uint32_t a_first[4] = {ql[0], qh[0], ql[1], qh[1]};
      const uint32_t a_second[4] = {ql[2], qh[2], ql[3], qh[3]};
      mma_bf16_16816(s[0], a_first, kl[0], kl[1]);
      mma_bf16_16816(s[1], a_first, kh[0], kh[1]);
      mma_bf16_16816(s[0], a_second, kl[2], kl[3]);
      mma_bf16_16816(s[1], a_second, kh[2], kh[3]);
    }
    if (TAIL16) {
      const int col = NB32 * 32 + t * 4;          // 4 contiguous channels of the trailing 16-wide block
      uint2 q_lo = tq_lo, q_hi = tq_hi, k_lo = tk_lo, k_hi = tk_hi;
      if (!PRELOAD) {
        if (lo_ok) {
          q_lo = __ldg(reinterpret_cast<const uint2*>(r_lo + col));
          k_lo = __ldg(reinterpret_cast<const uint2*>(r_lo + p.k_off + col));
        }
        if (hi_ok) {
          q_hi = __ldg(reinterpret_cast<const uint2*>(r_hi + col));
          k_hi = __ldg(reinterpret_cast<const uint2*>(r_hi + p.k_off + col));
        }
      }
      uint32_t ql[2] = {q_lo.x, q_lo.y}, qh[2] = {q_hi.x, q_hi.y}, kl[2] = {k_lo.x, k_lo.y}, kh[2] = {k_hi.x, k_hi.y};
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int pr = (col >> 1) + w;
        const bool rot = pr < p.rot_pairs;
        ql[w] = scale_rope_pair(ql[w], p.scale, rot && lo_ok, cs_lo + 2 * (rot ? pr : 0));
        qh[w] = scale_rope_pair(qh[w], p.scale, rot && hi_ok, cs_hi + 2 * (rot ? pr : 0));
        kl[w] = scale_rope_pair(kl[w], 1.f, rot && lo_ok, cs_lo + 2 * (rot ? pr : 0));
        kh[w] = scale_rope_pair(kh[w], 1.f, rot && hi_ok, cs_hi + 2 * (rot ? pr : 0));
      }
      const uint32_t a_tail[4] = {ql[0], qh[0], ql[1], qh[1]};
      mma_bf16_16816(s[0], a_tail, kl[0], kl[1]);
      mma_bf16_16816(s[1], a_tail, kh[0], kh[1]);
    }
    (void)KSTEPS;

    // ---- + rel-pos bias, mask padded frames, softmax over the key frames (a row lives in one lane quad) ----
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] += bias_lo[nt * 2 + e];          // -inf on padded frames
        s[nt][2 + e] += bias_hi[nt * 2 + e];
        mx_lo = fmaxf(mx_lo, s[nt][e]);
        mx_hi = fmaxf(mx_hi, s[nt][2 + e]);
      }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    if (!lo_ok) mx_lo = 0.f;
    if (!hi_ok) mx_hi = 0.f;
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = __expf(s[nt][e] - mx_lo);
        s[nt][2 + e] = __expf(s[nt][2 + e] - mx_hi);
        sum_lo += s[nt][e];
        sum_hi += s[nt][2 + e];
      }
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
    const float inv_lo = lo_ok ? 1.f / sum_lo : 0.f, inv_hi = hi_ok ? 1.f / sum_hi : 0.f;
    // probabilities -> A fragment of P.V (row g: a0 / a2, row g+8: a1 / a3)
    const uint32_t pa[4] = {pack_bf16(s[0][0] * inv_lo, s[0][1] * inv_lo), pack_bf16(s[0][2] * inv_hi, s[0][3] * inv_hi),
                            pack_bf16(s[1][0] * inv_lo, s[1][1] * inv_lo), pack_bf16(s[1][2] * inv_hi, s[1][3] * inv_hi)};

    // ---- O = P . V, 8 output channels per mma ----
    __syncwarp();                                   // V tile visible to the whole warp
    const uint32_t sv_addr = smem_u32(sv) + ((lane & 15) * VP) * 2;   // ldmatrix row address: key frame (lane & 15)
    __nv_bfloat16* o_lo = p.o + (row0 + static_cast<size_t>(g) * p.HW) * p.ldo + h * p.d;
    __nv_bfloat16* o_hi = p.o + (row0 + static_cast<size_t>(g + 8) * p.HW) * p.ldo + h * p.d;
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
      if (nt * 8 < p.d) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, sv_addr + nt * 16);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16_16816(o, pa, b0, b1);
        const int c = nt * 8 + 2 * t;
        if (lo_ok) *reinterpret_cast<uint32_t*>(o_lo + c) = pack_bf16(o[0], o[1]);
        if (hi_ok) *reinterpret_cast<uint32_t*>(o_hi + c) = pack_bf16(o[2], o[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Cross-attention against a SHORT key sequence (the text: Sk <= 80 keys, attention.py:364-366 / vsr attn1 + attn2) in a
// single pass: no online softmax, no TMEM, no tile loop.  K and V of one (text, head) live in shared memory; a warp owns
// 16 query rows at a time: S[16 x 80] = Q K^T on mma.sync m16n8k16 (Q straight from global memory into the A fragments
// with the same contraction-index permutation as the temporal kernel, K fragments by 16-byte shared loads), softmax in
// registers (a row lives in one lane quad), the exp(S) registers are the A operand of P.V (V through ldmatrix.trans).
// A block serves one (text, head) and strides over that group's query tiles, so K / V are staged once per block.  The op
// is bound by streaming Q in and O out once; the tcgen05 kernel it replaces for these shapes spent its time in per-CTA
// set-up for two mostly padded 64-key tiles (64 us at 81920 queries x 8 heads, 115 TFLOP/s).
// ------------------------------------------------------------------------------------------------------------
struct XAttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  long long q_seq, q_batch, kv_seq, kv_batch, o_seq, o_batch;   // strides in elements
  int Sq, Sk, heads, d, head_pitch, kv_div, tiles_per_batch, blocks_per_group;
  float scale_log2;
};

constexpr int XATT_KEYS = 80;          // 10 n-tiles of 8 keys
constexpr int XATT_NT = XATT_KEYS / 8;

template <int DP>
__global__ void __launch_bounds__(128)
cross_attn_mma_kernel(const XAttnParams p) {
  pdl_prologue();
  constexpr int KP = DP + 8;                      // smem row pitch (elements) of K and V
  constexpr int NB32 = DP / 32;
  constexpr bool TAIL16 = (DP % 32) != 0;
  extern __shared__ __align__(16) unsigned char xsm[];
  __nv_bfloat16* sk = reinterpret_cast<__nv_bfloat16*>(xsm);
  __nv_bfloat16* sv = sk + XATT_KEYS * KP;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int group = blockIdx.x / p.blocks_per_group;          // (kv batch, head)
  const int part = blockIdx.x - group * p.blocks_per_group;
  const int kvb = group / p.heads, h = group - kvb * p.heads;
  // ---- stage K and V of this (text, head): rows >= Sk are zero ----
  {
    const __nv_bfloat16* kb = p.k + static_cast<size_t>(kvb) * p.kv_batch + h * p.head_pitch;
    const __nv_bfloat16* vb = p.v + static_cast<size_t>(kvb) * p.kv_batch + h * p.head_pitch;
    constexpr int VEC = DP / 8;
    for (int i = threadIdx.x; i < XATT_KEYS * VEC; i += blockDim.x) {
      const int r = i / VEC, c = i - r * VEC;
      uint4 kk = make_uint4(0, 0, 0, 0), vv = kk;
      if (r < p.Sk) {
        kk = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<size_t>(r) * p.kv_seq + c * 8));
        vv = __ldg(reinterpret_cast<const uint4*>(vb + static_cast<size_t>(r) * p.kv_seq + c * 8));
      }
      *reinterpret_cast<uint4*>(sk + r * KP + c * 8) = kk;
      *reinterpret_cast<uint4*>(sv + r * KP + c * 8) = vv;
    }
  }
  __syncthreads();
  const long long tiles = static_cast<long long>(p.kv_div) * p.tiles_per_batch;     // query tiles of this group
  const uint32_t sv_addr = smem_u32(sv) + ((lane & 15) * KP) * 2;
  for (long long tile = static_cast<long long>(part) * 4 + wib; tile < tiles; tile += static_cast<long long>(p.blocks_per_group) * 4) {
    const int bl = static_cast<int>(tile / p.tiles_per_batch);
    const int r0 = static_cast<int>(tile - static_cast<long long>(bl) * p.tiles_per_batch) * 16;
    const long long b = static_cast<long long>(kvb) * p.kv_div + bl;
    const bool lo_ok = r0 + g < p.Sq, hi_ok = r0 + g + 8 < p.Sq;
    const __nv_bfloat16* q_lo_p = p.q + b * p.q_batch + static_cast<long long>(r0 + g) * p.q_seq + h * p.head_pitch;
    const __nv_bfloat16* q_hi_p = q_lo_p + 8 * p.q_seq;
    // ---- all Q loads of the tile first ----
    uint4 pq_lo[NB32 > 0 ? NB32 : 1], pq_hi[NB32 > 0 ? NB32 : 1];
    uint2 tq_lo = make_uint2(0, 0), tq_hi = tq_lo;
#pragma unroll
    for (int blk = 0; blk < NB32; ++blk) {
      pq_lo[blk] = lo_ok ? __ldg(reinterpret_cast<const uint4*>(q_lo_p + blk * 32 + t * 8)) : make_uint4(0, 0, 0, 0);
      pq_hi[blk] = hi_ok ? __ldg(reinterpret_cast<const uint4*>(q_hi_p + blk * 32 + t * 8)) : make_uint4(0, 0, 0, 0);
    }
    if (TAIL16) {
      if (lo_ok) tq_lo = __ldg(reinterpret_cast<const uint2*>(q_lo_p + NB32 * 32 + t * 4));
      if (hi_ok) tq_hi = __ldg(reinterpret_cast<const uint2*>(q_hi_p + NB32 * 32 + t * 4));
    }
    // ---- S = Q K^T ----
    float s[XATT_NT][4];
#pragma unroll
    for (int nt = 0; nt < XATT_NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
    for (int blk = 0; blk < NB32; ++blk) {
      const uint32_t a_first[4] = {pq_lo[blk].x, pq_hi[blk].x, pq_lo[blk].y, pq_hi[blk].y};
      const uint32_t a_second[4] = {pq_lo[blk].z, pq_hi[blk].z, pq_lo[blk].w, pq_hi[blk].w};
#pragma unroll
      for (int nt = 0; nt < XATT_NT; ++nt) {
        const uint4 kk = *reinterpret_cast<const uint4*>(sk + (nt * 8 + g) * KP + blk * 32 + t * 8);
        mma_bf16_16816(s[nt], a_first, kk.x, kk.y);
        mma_bf16_16816(s[nt], a_second, kk.z, kk.w);
      }
    }
    if (TAIL16) {
      const uint32_t a_tail[4] = {tq_lo.x, tq_hi.x, tq_lo.y, tq_hi.y};
#pragma unroll
      for (int nt = 0; nt < XATT_NT; ++nt) {
        const uint2 kk = *reinterpret_cast<const uint2*>(sk + (nt * 8 + g) * KP + NB32 * 32 + t * 4);
        mma_bf16_16816(s[nt], a_tail, kk.x, kk.y);
      }
    }
    // ---- softmax over the keys (row g: elements 0,1; row g+8: elements 2,3 of every n-tile) ----
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < XATT_NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool key_ok = nt * 8 + 2 * t + e < p.Sk;
        s[nt][e] = key_ok ? s[nt][e] * p.scale_log2 : -INFINITY;
        s[nt][2 + e] = key_ok ? s[nt][2 + e] * p.scale_log2 : -INFINITY;
        mx_lo = fmaxf(mx_lo, s[nt][e]);
        mx_hi = fmaxf(mx_hi, s[nt][2 + e]);
      }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < XATT_NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = fast_exp2(s[nt][e] - mx_lo);
        s[nt][2 + e] = fast_exp2(s[nt][2 + e] - mx_hi);
        sum_lo += s[nt][e];
        sum_hi += s[nt][2 + e];
      }
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
    const float inv_lo = 1.f / sum_lo, inv_hi = 1.f / sum_hi;
    // ---- O = P V: the probabilities (unnormalised, bf16) are the A fragments; 1 / sum is applied to the fp32 result ----
    uint32_t pa[XATT_NT / 2][4];
#pragma unroll
    for (int j = 0; j < XATT_NT / 2; ++j) {
      pa[j][0] = pack_bf16(s[2 * j][0], s[2 * j][1]);
      pa[j][1] = pack_bf16(s[2 * j][2], s[2 * j][3]);
      pa[j][2] = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[j][3] = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
    }
    __nv_bfloat16* o_lo = p.o + b * p.o_batch + static_cast<long long>(r0 + g) * p.o_seq + h * p.d;
    __nv_bfloat16* o_hi = o_lo + 8 * p.o_seq;
#pragma unroll
    for (int c = 0; c < DP / 8; ++c) {
      if (c * 8 < p.d) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < XATT_NT / 2; ++j) {
          uint32_t b0, b1;
          ldmatrix_x2_trans(b0, b1, sv_addr + (j * 16 * KP + c * 8) * 2);
          mma_bf16_16816(o, pa[j], b0, b1);
        }
        const int col = c * 8 + 2 * t;
        if (lo_ok) *reinterpret_cast<uint32_t*>(o_lo + col) = pack_bf16(o[0] * inv_lo, o[1] * inv_lo);
        if (hi_ok) *reinterpret_cast<uint32_t*>(o_hi + col) = pack_bf16(o[2] * inv_hi, o[3] * inv_hi);
      }
    }
  }
}

template <int DP>
int launch_cross_attn(const XAttnParams& p0, int kv_batches, cudaStream_t stream) {
  XAttnParams p = p0;
  constexpr int KP = DP + 8;
  const int smem = 2 * XATT_KEYS * KP * 2;
  static LavieSmemConfig configured;
  int rc = lavie_config_smem(cross_attn_mma_kernel<DP>, smem, &configured, "cross_attn_mma_kernel");
  if (rc) return rc;
  const int groups = kv_batches * p.heads;
  const long long tiles = static_cast<long long>(p.kv_div) * p.tiles_per_batch;
  // ~4 blocks of 4 warps per SM, at least 2 tiles per warp, K / V staged once per block
  long long per_group = (static_cast<long long>(lavie_num_sms()) * 4 + groups - 1) / groups;
  const long long max_per_group = (tiles + 7) / 8;
  if (per_group > max_per_group) per_group = max_per_group;
  if (per_group < 1) per_group = 1;
  p.blocks_per_group = static_cast<int>(per_group);
  launch_pdl(cross_attn_mma_kernel<DP>, static_cast<int>(groups * per_group), 128, smem, stream, p);
  return lavie_check_launch("cross_attn_mma_kernel");
}

template <int DP>
int launch_temporal_mma(const TempParams& p, cudaStream_t stream) {
  long long blocks = (p.items + 3) / 4;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  // keep (4 * blocks) % heads == 0 when the grid is capped, so every warp serves a single head (hoisted bias)
  if (blocks * 4 < p.items) {
    while (blocks > 1 && (blocks * 4) % p.heads != 0) --blocks;
  }
  launch_pdl(temporal_attn_mma_kernel<DP>, static_cast<int>(blocks), 128, 0, stream, p);
  return lavie_check_launch("temporal_attn_mma_kernel");
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int lavie_attention_strided_bf16(const void* q, long long q_seq_stride, long long q_batch_stride,
                                            const void* k, const void* v, long long kv_seq_stride,
                                            long long kv_batch_stride, void* o, long long o_seq_stride,
                                            long long o_batch_stride, int batch, int heads, int Sq, int Sk, int d,
                                            int head_pitch, int kv_batch_div, int sparse_causal_frames, int sc_halo,
                                            float scale, cudaStream_t stream) {
  LAVIE_REQUIRE(batch > 0 && heads > 0 && Sq > 0 && Sk > 0 && kv_batch_div > 0 && batch % kv_batch_div == 0,
                LAVIE_ERR_SHAPE, "attention: bad sizes batch=%d heads=%d Sq=%d Sk=%d div=%d", batch, heads, Sq, Sk,
                kv_batch_div);
  LAVIE_REQUIRE(sparse_causal_frames >= 0 && (sparse_causal_frames == 0 ||
                                              (kv_batch_div == 1 && batch % sparse_causal_frames == 0)),
                LAVIE_ERR_SHAPE, "attention: sparse-causal mode needs batch %% frames == 0 and kv_batch_div == 1");
  const int dk = (d + 15) & ~15;
  LAVIE_REQUIRE(d % 8 == 0 && head_pitch >= dk && head_pitch % 8 == 0, LAVIE_ERR_SHAPE,
                "attention: d=%d head_pitch=%d (pitch must be >= d rounded up to 16)", d, head_pitch);
  LAVIE_REQUIRE(al16(q) && al16(k) && al16(v) && al16(o) && q_seq_stride % 8 == 0 && q_batch_stride % 8 == 0 &&
                    kv_seq_stride % 8 == 0 && kv_batch_stride % 8 == 0 && o_seq_stride % 8 == 0 &&
                    o_batch_stride % 8 == 0,
                LAVIE_ERR_ALIGN, "attention: 16-byte alignment required");
  LAVIE_REQUIRE(heads <= 65535 && batch <= 65535, LAVIE_ERR_SHAPE, "attention: grid too large (batch=%d heads=%d)",
                batch, heads);
  AttnParams p;
  p.Sq = Sq; p.Sk = Sk; p.d = d; p.head_pitch = head_pitch; p.kv_batch_div = kv_batch_div;
  p.o_seq_stride = o_seq_stride; p.o_batch_stride = o_batch_stride;
  p.sc_frames = sparse_causal_frames;
  p.sc_halo = sc_halo;
  LAVIE_REQUIRE(sc_halo == 0 || (sc_halo >= 1 && sc_halo <= 2 && sparse_causal_frames == batch), LAVIE_ERR_SHAPE,
                "attention: sc_halo needs sparse-causal mode with one video per launch (frames == batch)");
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.timeline = static_cast<long long*>(g_lavie_debug_buf);
  if (Sk <= XATT_KEYS && sparse_causal_frames == 0 && g_lavie_xattn && q_batch_stride >= q_seq_stride &&
      kv_batch_stride >= kv_seq_stride) {
    // short key sequence (text cross-attention): the single-pass mma.sync kernel
    XAttnParams x;
    x.q = static_cast<const __nv_bfloat16*>(q); x.k = static_cast<const __nv_bfloat16*>(k);
    x.v = static_cast<const __nv_bfloat16*>(v); x.o = static_cast<__nv_bfloat16*>(o);
    x.q_seq = q_seq_stride; x.q_batch = q_batch_stride; x.kv_seq = kv_seq_stride; x.kv_batch = kv_batch_stride;
    x.o_seq = o_seq_stride; x.o_batch = o_batch_stride;
    x.Sq = Sq; x.Sk = Sk; x.heads = heads; x.d = d; x.head_pitch = head_pitch; x.kv_div = kv_batch_div;
    x.tiles_per_batch = (Sq + 15) / 16; x.blocks_per_group = 1;
    x.scale_log2 = scale * 1.4426950408889634f;
    const int kvb = batch / kv_batch_div;
    switch (dk) {
      case 48: return launch_cross_attn<48>(x, kvb, stream);
      case 64: return launch_cross_attn<64>(x, kvb, stream);
      case 80: return launch_cross_attn<80>(x, kvb, stream);
      case 96: return launch_cross_attn<96>(x, kvb, stream);
      case 128: return launch_cross_attn<128>(x, kvb, stream);
      case 160: return launch_cross_attn<160>(x, kvb, stream);
      default: break;
    }
  }
  CUtensorMap mq, mk, mv;
  const int cols = heads * head_pitch;
  const bool swap = q_batch_stride < q_seq_stride;
  LAVIE_REQUIRE(swap == (kv_batch_stride < kv_seq_stride), LAVIE_ERR_SHAPE,
                "attention: q and k/v must use the same dimension order (batch stride vs sequence stride)");
  p.swap_dims = swap ? 1 : 0;
  int rc = make_qkv_map(&mq, q, q_seq_stride, q_batch_stride, cols, Sq, batch, ATT_M, swap);
  if (rc) return rc;
  const int kv_batches = batch / kv_batch_div + (sc_halo ? 2 : 0);
  rc = make_qkv_map(&mk, k, kv_seq_stride, kv_batch_stride, cols, Sk, kv_batches, ATT_N, swap);
  if (rc) return rc;
  rc = make_qkv_map(&mv, v, kv_seq_stride, kv_batch_stride, cols, Sk, kv_batches, ATT_N, swap);
  if (rc) return rc;
  switch (dk) {
    case 48: return launch_attn<48>(mq, mk, mv, p, batch, heads, stream);
    case 64: return launch_attn<64>(mq, mk, mv, p, batch, heads, stream);
    case 80: return launch_attn<80>(mq, mk, mv, p, batch, heads, stream);
    case 96: return launch_attn<96>(mq, mk, mv, p, batch, heads, stream);
    case 128: return launch_attn<128>(mq, mk, mv, p, batch, heads, stream);
    case 160: return launch_attn<160>(mq, mk, mv, p, batch, heads, stream);
    default:
      lavie_set_error("attention: head dim %d (padded %d) not instantiated (48/64/80/96/128/160)", d, dk);
      return LAVIE_ERR_SHAPE;
  }
}

extern "C" int lavie_attention_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o,
                                    int ldo, int batch, int heads, int Sq, int Sk, int d, int head_pitch,
                                    int kv_batch_div, float scale, cudaStream_t stream) {
  LAVIE_REQUIRE(ldk == ldv, LAVIE_ERR_SHAPE, "attention: k and v must share their row stride (ldk=%d ldv=%d)", ldk, ldv);
  return lavie_attention_strided_bf16(q, ldq, static_cast<long long>(Sq) * ldq, k, v, ldk,
                                      static_cast<long long>(Sk) * ldk, o, ldo, static_cast<long long>(Sq) * ldo, batch,
                                      heads, Sq, Sk, d, head_pitch, kv_batch_div, 0, 0, scale, stream);
}

extern "C" int lavie_temporal_attention_bf16(const void* qkv, int ld, int k_off, int v_off, void* o, int ldo, int B,
                                             int F, int HW, int heads, int d, int head_pitch, float scale,
                                             const float* rope, int rot_pairs, const float* bias,
                                             cudaStream_t stream) {
  LAVIE_REQUIRE(F >= 1 && F <= 64 && d % 2 == 0 && 2 * rot_pairs <= d, LAVIE_ERR_SHAPE,
                "temporal attention: F=%d (<=64) d=%d rot_pairs=%d", F, d, rot_pairs);
  LAVIE_REQUIRE(ld % 2 == 0 && ldo % 2 == 0 && k_off % 2 == 0 && v_off % 2 == 0 && head_pitch % 2 == 0,
                LAVIE_ERR_ALIGN, "temporal attention: even strides required");
  TempParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.ld = ld; p.k_off = k_off; p.v_off = v_off;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.ldo = ldo; p.B = B; p.F = F; p.HW = HW; p.heads = heads; p.d = d; p.head_pitch = head_pitch;
  p.scale = scale; p.rope = rope; p.rot_pairs = rot_pairs; p.bias = bias;
  p.items = static_cast<long long>(B) * HW * heads;
  if (F <= 16 && d % 8 == 0 && head_pitch == ((d + 15) & ~15) && ld % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0 &&
      al16(qkv)) {
    switch (head_pitch) {
      case 48: return launch_temporal_mma<48>(p, stream);
      case 64: return launch_temporal_mma<64>(p, stream);
      case 80: return launch_temporal_mma<80>(p, stream);
      case 96: return launch_temporal_mma<96>(p, stream);
      case 128: return launch_temporal_mma<128>(p, stream);
      case 160: return launch_temporal_mma<160>(p, stream);
      default: break;     // fall through to the general kernel
    }
  }
  const int per_warp = (3 * F * (d + 1) + F * (F + 1)) * static_cast<int>(sizeof(float));
  int warps = 4;
  while (warps > 1 && warps * per_warp > 200 * 1024) warps >>= 1;
  const int smem = warps * per_warp;
  LAVIE_REQUIRE(smem <= 227 * 1024, LAVIE_ERR_SHAPE, "temporal attention: F*d too large for shared memory");
  static LavieSmemConfig configured_smem;
  const int rc_cfg = lavie_config_smem(temporal_attn_kernel, smem, &configured_smem, "temporal_attn_kernel");
  if (rc_cfg) return rc_cfg;
  long long blocks = (p.items + warps - 1) / warps;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  launch_pdl(temporal_attn_kernel, static_cast<int>(blocks), warps * 32, smem, stream, p);
  return lavie_check_launch("temporal_attn_kernel");
}
