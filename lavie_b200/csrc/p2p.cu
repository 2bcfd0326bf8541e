// Frame-sharding collectives fused into compute kernels over NVLink peer memory (one process per GPU, peers mapped
// by the caller, e.g. through CUDA IPC / torch symmetric memory; the library only sees raw peer pointers).
//
//   lavie_gn_exchange_finalize   GroupNorm statistics: reduce the local partials, publish the 1 KB of (sum, sumsq)
//                                to every peer with plain stores, wait for theirs, add in rank order, emit
//                                (scale, shift).  One kernel instead of reduce + ncclAllReduce(512 B) + finalize.
//   lavie_layernorm_scatter_p2p  LayerNorm whose normalised rows are stored straight into the peers' receive buffers:
//                                the frame->pixel all-to-all in front of the temporal attention (attention.py:549-550)
//                                becomes the store side of the LayerNorm kernel.
//   lavie_add_gathered_p2p       residual add that loads the attention output rows from the peers' buffers: the
//                                all-to-all back (attention.py:554-555) becomes the load side of the residual kernel.
//   lavie_rank_barrier           flag barrier between the ranks of a frame group (publishes "my previous kernels are
//                                done", waits for everybody).
//
// Synchronisation: every peer owns a flag array flags[P] (uint32) in its symmetric buffer; rank r writes epoch e into
// flags_of_peer[r] with a system-scope release store after its data, and waits on its own array with acquire loads.
// Epochs come from a per-rank device counter, so the kernels can be replayed from a CUDA graph.  Waits are bounded:
// a lost peer traps instead of hanging the GPU.
#include "common.cuh"

namespace {

constexpr int MAX_PEERS = 8;

struct PeerPtrs {
  void* p[MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
// Fault report: a peer that never arrives must not hang the GPU box, so the wait is bounded by WALL time
// (%globaltimer, default 30 s: a rank may legitimately be late by seconds -- lazy module load, a host stall, a
// debugger).  Before trapping, the waiting rank writes {magic, my rank, missing peer, epoch wanted, epoch seen} into a
// caller-provided HOST-mapped word array (lavie_p2p_fault_buffer), which survives the sticky context error the trap
// raises on this rank -- and, through the missing flag, on every other rank of the frame group.
struct FaultCtx {
  uint32_t* report;            // host-mapped uint32[8] or nullptr
  unsigned long long timeout_ns;
};
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_flag(const uint32_t* addr, uint32_t epoch, const FaultCtx& fc, int my_rank,
                                          int peer) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  uint32_t seen;
  while (static_cast<int32_t>((seen = ld_acquire_sys(addr)) - epoch) < 0) {
    __nanosleep(64);
    if ((++spins & 1023u) == 0) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > fc.timeout_ns) {
        if (fc.report) {
          fc.report[1] = static_cast<uint32_t>(my_rank);
          fc.report[2] = static_cast<uint32_t>(peer);
          fc.report[3] = epoch;
          fc.report[4] = seen;
          __threadfence_system();
          fc.report[0] = 0x4C564945u;     // "LVIE": report valid
          __threadfence_system();
        }
        __trap();
      }
    }
  }
}

// all threads of ONE block call this; thread r < P signals peer r and waits for peer r
__device__ __forceinline__ void block_rank_barrier(const PeerPtrs& flags, int P, int my_rank, uint32_t epoch,
                                                   const FaultCtx& fc) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < P) {
    st_release_sys(static_cast<uint32_t*>(flags.p[threadIdx.x]) + my_rank, epoch);
    wait_flag(static_cast<const uint32_t*>(flags.p[my_rank]) + threadIdx.x, epoch, fc, my_rank, threadIdx.x);
  }
  __syncthreads();
}

__global__ void rank_barrier_kernel(PeerPtrs flags, uint32_t* epoch_counter, int P, int my_rank, FaultCtx fc) {
  pdl_prologue();
  const uint32_t epoch = *epoch_counter + 1;
  block_rank_barrier(flags, P, my_rank, epoch, fc);
  if (threadIdx.x == 0) *epoch_counter = epoch;
}

// Barrier folded into the END of a multi-block kernel: every block publishes its peer stores (system fence) and takes a
// ticket; the LAST block to finish signals the peers, waits for theirs and bumps the epoch, so kernel completion implies
// "all ranks passed the barrier" and no separate barrier launch is needed.  Every thread of every block must call it.
struct SyncCtx {
  PeerPtrs flags;
  uint32_t* epoch_counter;
  int* ticket;               // device int, zero on entry, left zero
  int P, my_rank;
  FaultCtx fc;
};
__device__ __forceinline__ void grid_end_barrier(const SyncCtx& sc) {
  __shared__ int s_last_blk;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last_blk = (atomicAdd(sc.ticket, 1) == static_cast<int>(gridDim.x) - 1);
  __syncthreads();
  if (!s_last_blk) return;
  __threadfence();
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(sc.epoch_counter) + 1;
  block_rank_barrier(sc.flags, sc.P, sc.my_rank, epoch, sc.fc);
  if (threadIdx.x == 0) {
    *sc.epoch_counter = epoch;
    *sc.ticket = 0;
  }
}
// Barrier folded into the START of a multi-block kernel whose blocks all read peer memory: block 0 signals, every block
// waits for all peers; the last block to FINISH bumps the epoch (every block has read it by then).
__device__ __forceinline__ uint32_t grid_start_barrier(const SyncCtx& sc) {
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(sc.epoch_counter) + 1;
  if (blockIdx.x == 0 && threadIdx.x < sc.P) {
    __threadfence_system();
    st_release_sys(static_cast<uint32_t*>(sc.flags.p[threadIdx.x]) + sc.my_rank, epoch);
  }
  if (threadIdx.x < sc.P)
    wait_flag(static_cast<const uint32_t*>(sc.flags.p[sc.my_rank]) + threadIdx.x, epoch, sc.fc, sc.my_rank, threadIdx.x);
  __syncthreads();
  return epoch;
}
__device__ __forceinline__ void grid_start_barrier_finish(const SyncCtx& sc, uint32_t epoch) {
  __shared__ int s_last_blk2;
  __syncthreads();
  if (threadIdx.x == 0) s_last_blk2 = (atomicAdd(sc.ticket, 1) == static_cast<int>(gridDim.x) - 1);
  __syncthreads();
  if (s_last_blk2 && threadIdx.x == 0) {
    *sc.epoch_counter = epoch;
    *sc.ticket = 0;
  }
}

// this rank's GroupNorm sums straight from the producers' micro-group statistics (gemm.cu epilogue; layout and address
// arithmetic of norm.cu's gn_colsums_kernel): warp per (sample, group), fp64 lanes, fixed shuffle tree
struct ColSrc {
  const float* cs0;
  const float* cs1;
  int c0, c1, slabs_per_sample;
};
__device__ __forceinline__ void block_colsums(const ColSrc& src, int samples, int groups, double* s_sums) {
  const int C = src.c0 + src.c1, cpg = C / groups, decs = cpg / 10;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int sg = warp; sg < samples * groups; sg += nwarps) {
    const int sample = sg / groups, g = sg - sample * groups;
    const long long total = static_cast<long long>(src.slabs_per_sample) * decs;
    double a = 0.0, b = 0.0;
    for (long long i = lane; i < total; i += 32) {
      const long long slab = static_cast<long long>(sample) * src.slabs_per_sample + i / decs;
      int ch = g * cpg + static_cast<int>(i % decs) * 10;
      const float* cs = src.cs0;
      int cn = src.c0;
      if (ch >= src.c0) {
        ch -= src.c0;
        cs = src.cs1;
        cn = src.c1;
      }
      const int dec = ch / 10, k_lo = ch >> 5, k_hi = (ch + 9) >> 5;
      const float* row = cs + slab * (cn >> 5) * 8;
      float2 v = *reinterpret_cast<const float2*>(row + (k_lo * 4 + (dec - (k_lo * 32) / 10)) * 2);
      if (k_hi != k_lo) {
        const float2 w = *reinterpret_cast<const float2*>(row + (k_hi * 4 + (dec - (k_hi * 32) / 10)) * 2);
        v.x += w.x;
        v.y += w.y;
      }
      a += v.x;
      b += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      s_sums[sg * 2] = a;
      s_sums[sg * 2 + 1] = b;
    }
  }
}

// partial[samples][chunks][groups][2] -> exchange -> scale_shift[samples][C][2]
// slots: per peer a buffer double[2 (epoch parity)][P][samples*groups*2]
__global__ void __launch_bounds__(256)
gn_exchange_finalize_kernel(const float* __restrict__ partial, const double* __restrict__ local_sums, ColSrc colsrc,
                            int samples, int chunks, int groups, int C,
                            double inv_count, const float* __restrict__ gamma, const float* __restrict__ beta,
                            float eps, float* __restrict__ scale_shift, PeerPtrs slots, PeerPtrs flags,
                            uint32_t* epoch_counter, int P, int my_rank, FaultCtx fc) {
  pdl_prologue();
  __shared__ double s_sums[2 * 64 * 2];          // [samples <= 2][groups <= 64][2]
  __shared__ float s_mean[2 * 64], s_rstd[2 * 64];
  const uint32_t epoch = *epoch_counter + 1;
  const int n = samples * groups * 2;
  // 1. ordered local reduction of the chunk partials (8 lanes per (sample, group)) -- or the local sums as they come
  //    from the producers' column statistics (lavie_groupnorm_reduce_colsums)
  const int sub = threadIdx.x & 7;
  if (colsrc.cs0 != nullptr) {
    block_colsums(colsrc, samples, groups, s_sums);
  } else if (local_sums != nullptr) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_sums[i] = local_sums[i];
  } else
  for (int sg = threadIdx.x >> 3; sg < samples * groups; sg += blockDim.x >> 3) {
    const int sample = sg / groups, g = sg - sample * groups;
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
      s_sums[sg * 2] = a;
      s_sums[sg * 2 + 1] = b;
    }
  }
  __syncthreads();
  // 2. publish to every peer (slot [parity][my_rank]), then flag barrier
  const size_t slot_off = (static_cast<size_t>(epoch & 1) * P + my_rank) * n;
  for (int i = threadIdx.x; i < n * P; i += blockDim.x) {
    const int peer = i / n, e = i - peer * n;
    static_cast<double*>(slots.p[peer])[slot_off + e] = s_sums[e];
  }
  block_rank_barrier(flags, P, my_rank, epoch, fc);
  // 3. add the P contributions in rank order (bit-identical on every rank) and finalize
  const double* mine = static_cast<const double*>(slots.p[my_rank]) + static_cast<size_t>(epoch & 1) * P * n;
  for (int sg = threadIdx.x; sg < samples * groups; sg += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int r = 0; r < P; ++r) {
      a += mine[static_cast<size_t>(r) * n + sg * 2];
      b += mine[static_cast<size_t>(r) * n + sg * 2 + 1];
    }
    const double mean = a * inv_count;
    double var = b * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[sg] = static_cast<float>(mean);
    s_rstd[sg] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int i = threadIdx.x; i < samples * C; i += blockDim.x) {
    const int sample = i / C, c = i - sample * C;
    const int sg = sample * groups + c / cpg;
    const float sc = s_rstd[sg] * gamma[c];
    scale_shift[static_cast<size_t>(i) * 2] = sc;
    scale_shift[static_cast<size_t>(i) * 2 + 1] = beta[c] - s_mean[sg] * sc;
  }
  if (threadIdx.x == 0) *epoch_counter = epoch;
}

// LayerNorm (C = 40*L) with peer-scattered output: row (f, pixel) -> peer (pixel / hwp), row (frame_off + f, pixel % hwp)
// (frame_off = global index of this rank's first frame; shards may hold different numbers of frames)
template <int L>
__global__ void __launch_bounds__(256)
layernorm_scatter_p2p_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float eps, PeerPtrs recv, int rows, int hw, int hwp,
                             int frame_off, SyncCtx sync) {
  pdl_prologue();
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
  if (ok) {
  const int fr = static_cast<int>(row / hw), pix = static_cast<int>(row - static_cast<long long>(fr) * hw);
  const int blk = pix / hwp;
  const size_t drow = (static_cast<size_t>(frame_off) + fr) * hwp + (pix - blk * hwp);     // global frame index
  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(recv.p[blk]) + drow * C;
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
  if (sync.ticket != nullptr) grid_end_barrier(sync);    // kernel completion == every rank's rows have landed
}

// out[(f, blk*hwp + j)] = res[(f, blk*hwp + j)] + y_of_peer_blk[(frame_off + f)*hwp + j]   (peer loads over NVLink)
__global__ void __launch_bounds__(256)
add_gathered_p2p_kernel(const __nv_bfloat16* __restrict__ res, int ldr, PeerPtrs ybuf, __nv_bfloat16* __restrict__ out,
                        int ldo, int rows, int C, int hw, int hwp, int frame_off, SyncCtx sync) {
  pdl_prologue();
  uint32_t epoch = 0;
  if (sync.ticket != nullptr) epoch = grid_start_barrier(sync);      // every rank's y buffer is complete
  const int nvec = C >> 3;
  const long long total = static_cast<long long>(rows) * nvec;
  constexpr int U = 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long base = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; base < total; base += stride * U) {
    uint4 a[U], b[U];
    long long idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + u * stride;
      idx[u] = i < total ? i : base;
      const int row = static_cast<int>(idx[u] / nvec), v = static_cast<int>(idx[u] % nvec);
      const int f = row / hw, pix = row - f * hw;
      const int blk = pix / hwp;
      const size_t yrow = (static_cast<size_t>(frame_off) + f) * hwp + (pix - blk * hwp);
      a[u] = __ldg(reinterpret_cast<const uint4*>(res + static_cast<size_t>(row) * ldr + v * 8));
      b[u] = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(ybuf.p[blk]) + yrow * C + v * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * stride >= total) break;
      const int row = static_cast<int>(idx[u] / nvec), v = static_cast<int>(idx[u] % nvec);
      const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, bw[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(bw[e]);
        o[e] = pack_bf16(x.x + y.x, x.y + y.y);
      }
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * ldo + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (sync.ticket != nullptr) grid_start_barrier_finish(sync, epoch);
}

// SparseCausalAttention under frame sharding: every rank needs the projected q|k|v rows of
// frame 0 of the video (from the first rank) and of the frame in front of its first frame (from its left neighbour).
// The receiving buffer on every rank is [halo block 0 = frame 0 | halo block 1 = previous frame | local frames].
__global__ void __launch_bounds__(256)
halo_push_kernel(const uint4* __restrict__ first_frame, const uint4* __restrict__ last_frame, long long n16,
                 PeerPtrs dst, int P, int my_rank, SyncCtx sync) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (my_rank == 0) {
      const uint4 v = __ldg(first_frame + i);
      for (int r = 0; r < P; ++r) static_cast<uint4*>(dst.p[r])[i] = v;
    }
    if (my_rank + 1 < P) static_cast<uint4*>(dst.p[my_rank + 1])[n16 + i] = __ldg(last_frame + i);
  }
  if (sync.ticket != nullptr) grid_end_barrier(sync);
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

uint32_t* g_fault_report = nullptr;              // host-mapped, set by lavie_p2p_fault_buffer
unsigned long long g_wait_timeout_ns = 30ull * 1000000000ull;
FaultCtx fault_ctx() { return FaultCtx{g_fault_report, g_wait_timeout_ns}; }

int fill_peers(PeerPtrs& pp, void* const* ptrs, int P);
// optional fused barrier: flag_ptrs == nullptr -> no barrier inside the kernel (caller issues lavie_rank_barrier)
int fill_sync(SyncCtx& sc, void* const* flag_ptrs, unsigned int* epoch_counter, int* ticket, int P, int my_rank) {
  sc.epoch_counter = epoch_counter;
  sc.ticket = nullptr;
  sc.P = P;
  sc.my_rank = my_rank;
  sc.fc = fault_ctx();
  for (int i = 0; i < MAX_PEERS; ++i) sc.flags.p[i] = nullptr;
  if (flag_ptrs == nullptr) return LAVIE_OK;
  LAVIE_REQUIRE(epoch_counter != nullptr && ticket != nullptr, LAVIE_ERR_SHAPE,
                "p2p: a fused barrier needs the epoch counter and a ticket word");
  int rc = fill_peers(sc.flags, flag_ptrs, P);
  if (rc) return rc;
  sc.ticket = ticket;
  return LAVIE_OK;
}

int fill_peers(PeerPtrs& pp, void* const* ptrs, int P) {
  LAVIE_REQUIRE(P >= 1 && P <= MAX_PEERS, LAVIE_ERR_SHAPE, "p2p: 1 <= peers <= %d", MAX_PEERS);
  for (int i = 0; i < MAX_PEERS; ++i) pp.p[i] = i < P ? ptrs[i] : nullptr;
  for (int i = 0; i < P; ++i) LAVIE_REQUIRE(ptrs[i] != nullptr && al16(ptrs[i]), LAVIE_ERR_ALIGN, "p2p: peer pointer %d", i);
  return LAVIE_OK;
}

}  // namespace

extern "C" int lavie_rank_barrier(void* const* flag_ptrs, unsigned int* epoch_counter, int P, int my_rank,
                                  cudaStream_t stream) {
  PeerPtrs f;
  int rc = fill_peers(f, flag_ptrs, P);
  if (rc) return rc;
  launch_pdl(rank_barrier_kernel, 1, 32, 0, stream, f, epoch_counter, P, my_rank, fault_ctx());
  return lavie_check_launch("rank_barrier_kernel");
}

namespace {
int gn_exchange_impl(const float* partial, const double* local_sums, ColSrc colsrc, int samples, int chunks, int groups,
                     int C,
                     long long count_per_group_global, const float* gamma, const float* beta, float eps,
                     float* scale_shift, void* const* slot_ptrs, void* const* flag_ptrs, unsigned int* epoch_counter,
                     int P, int my_rank, cudaStream_t stream) {
  LAVIE_REQUIRE(samples >= 1 && samples <= 2 && groups <= 64 && groups % 4 == 0 && C % groups == 0 &&
                    count_per_group_global > 0,
                LAVIE_ERR_SHAPE, "gn_exchange_finalize: samples <= 2, groups <= 64 (multiple of 4)");
  PeerPtrs s, f;
  int rc = fill_peers(s, slot_ptrs, P);
  if (rc) return rc;
  rc = fill_peers(f, flag_ptrs, P);
  if (rc) return rc;
  launch_pdl(gn_exchange_finalize_kernel, 1, 256, 0, stream, partial, local_sums, colsrc, samples, chunks, groups, C,
                                                     1.0 / static_cast<double>(count_per_group_global), gamma, beta, eps,
                                                     scale_shift, s, f, epoch_counter, P, my_rank, fault_ctx());
  return lavie_check_launch("gn_exchange_finalize_kernel");
}
}  // namespace

extern "C" int lavie_gn_exchange_finalize(const float* partial, int samples, int chunks, int groups, int C,
                                          long long count_per_group_global, const float* gamma, const float* beta,
                                          float eps, float* scale_shift, void* const* slot_ptrs, void* const* flag_ptrs,
                                          unsigned int* epoch_counter, int P, int my_rank, cudaStream_t stream) {
  return gn_exchange_impl(partial, nullptr, ColSrc{nullptr, nullptr, 0, 0, 0}, samples, chunks, groups, C,
                          count_per_group_global, gamma, beta, eps, scale_shift, slot_ptrs, flag_ptrs, epoch_counter, P,
                          my_rank, stream);
}

extern "C" int lavie_gn_exchange_finalize_sums(const double* local_sums, int samples, int groups, int C,
                                               long long count_per_group_global, const float* gamma, const float* beta,
                                               float eps, float* scale_shift, void* const* slot_ptrs,
                                               void* const* flag_ptrs, unsigned int* epoch_counter, int P, int my_rank,
                                               cudaStream_t stream) {
  LAVIE_REQUIRE(local_sums != nullptr, LAVIE_ERR_SHAPE, "gn_exchange_finalize_sums: null sums");
  return gn_exchange_impl(nullptr, local_sums, ColSrc{nullptr, nullptr, 0, 0, 0}, samples, 1, groups, C,
                          count_per_group_global, gamma, beta, eps, scale_shift, slot_ptrs, flag_ptrs, epoch_counter, P,
                          my_rank, stream);
}

extern "C" int lavie_gn_exchange_finalize_colsums(const float* cs0, int c0, const float* cs1, int c1, int samples,
                                                  int rows_local, int groups, long long count_per_group_global,
                                                  const float* gamma, const float* beta, float eps, float* scale_shift,
                                                  void* const* slot_ptrs, void* const* flag_ptrs,
                                                  unsigned int* epoch_counter, int P, int my_rank, cudaStream_t stream) {
  const int C = c0 + c1;
  LAVIE_REQUIRE(cs0 != nullptr && (c1 == 0 || cs1 != nullptr) && rows_local > 0 && rows_local % 32 == 0 && groups > 0 &&
                    C % groups == 0 && (C / groups) % 10 == 0 && c0 % 32 == 0 && c1 % 32 == 0 && c0 % 10 == 0,
                LAVIE_ERR_SHAPE, "gn_exchange_finalize_colsums: rows_local=%d C=%d groups=%d", rows_local, C, groups);
  return gn_exchange_impl(nullptr, nullptr, ColSrc{cs0, cs1, c0, c1, rows_local / 32}, samples, 1, groups, C,
                          count_per_group_global, gamma, beta, eps, scale_shift, slot_ptrs, flag_ptrs, epoch_counter, P,
                          my_rank, stream);
}

extern "C" int lavie_layernorm_scatter_p2p(const void* x, int ldx, const float* gamma, const float* beta, float eps,
                                           void* const* recv_ptrs, int rows, int C, int hw, int hwp, int P, int my_rank,
                                           int frame_off, void* const* flag_ptrs, unsigned int* epoch_counter,
                                           int* ticket, cudaStream_t stream) {
  if (frame_off < 0) frame_off = my_rank * (hw > 0 ? rows / hw : 0);        // equal shards
  SyncCtx sync;
  int rcs = fill_sync(sync, flag_ptrs, epoch_counter, ticket, P, my_rank);
  if (rcs) return rcs;
  LAVIE_REQUIRE(C == 320 || C == 640 || C == 1280, LAVIE_ERR_SHAPE, "layernorm_scatter_p2p: C must be 320/640/1280");
  LAVIE_REQUIRE(hw > 0 && hwp > 0 && hw == hwp * P && rows % hw == 0 && ldx % 8 == 0, LAVIE_ERR_SHAPE,
                "layernorm_scatter_p2p: rows=%d hw=%d hwp=%d P=%d", rows, hw, hwp, P);
  LAVIE_REQUIRE(al16(x) && al16(gamma) && al16(beta), LAVIE_ERR_ALIGN, "layernorm_scatter_p2p: alignment");
  PeerPtrs r;
  int rc = fill_peers(r, recv_ptrs, P);
  if (rc) return rc;
  const int lanes = C / 40;
  const int rpw = 32 / lanes;
  const long long warps = (static_cast<long long>(rows) + rpw - 1) / rpw;
  const int blocks = static_cast<int>((warps + 7) / 8);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  if (lanes == 8)
    launch_pdl(layernorm_scatter_p2p_kernel<8>, blocks, 256, 0, stream, xp, ldx, gamma, beta, eps, r, rows, hw, hwp, frame_off, sync);
  else if (lanes == 16)
    launch_pdl(layernorm_scatter_p2p_kernel<16>, blocks, 256, 0, stream, xp, ldx, gamma, beta, eps, r, rows, hw, hwp, frame_off, sync);
  else
    launch_pdl(layernorm_scatter_p2p_kernel<32>, blocks, 256, 0, stream, xp, ldx, gamma, beta, eps, r, rows, hw, hwp, frame_off, sync);
  return lavie_check_launch("layernorm_scatter_p2p_kernel");
}

extern "C" int lavie_add_gathered_p2p(const void* res, int ldr, void* const* y_ptrs, void* out, int ldo, int rows, int C,
                                      int hw, int hwp, int P, int my_rank, int frame_off, void* const* flag_ptrs,
                                      unsigned int* epoch_counter, int* ticket, cudaStream_t stream) {
  if (frame_off < 0) frame_off = my_rank * (hw > 0 ? rows / hw : 0);        // equal shards
  SyncCtx sync;
  int rcs = fill_sync(sync, flag_ptrs, epoch_counter, ticket, P, my_rank);
  if (rcs) return rcs;
  LAVIE_REQUIRE(C % 8 == 0 && ldr % 8 == 0 && ldo % 8 == 0 && hw > 0 && hwp > 0 && hw == hwp * P && rows % hw == 0,
                LAVIE_ERR_SHAPE, "add_gathered_p2p: rows=%d C=%d hw=%d hwp=%d P=%d", rows, C, hw, hwp, P);
  LAVIE_REQUIRE(al16(res) && al16(out), LAVIE_ERR_ALIGN, "add_gathered_p2p: alignment");
  PeerPtrs y;
  int rc = fill_peers(y, y_ptrs, P);
  if (rc) return rc;
  const long long total = static_cast<long long>(rows) * (C >> 3);
  long long blocks = (total + 1023) / 1024;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(add_gathered_p2p_kernel, static_cast<int>(blocks), 256, 0, stream, static_cast<const __nv_bfloat16*>(res), ldr, y, static_cast<__nv_bfloat16*>(out), ldo, rows, C, hw, hwp, frame_off, sync);
  return lavie_check_launch("add_gathered_p2p_kernel");
}

extern "C" int lavie_p2p_fault_buffer(void* host_mapped_words, int timeout_seconds) {
  g_fault_report = static_cast<uint32_t*>(host_mapped_words);
  if (timeout_seconds > 0) g_wait_timeout_ns = static_cast<unsigned long long>(timeout_seconds) * 1000000000ull;
  return LAVIE_OK;
}

extern "C" int lavie_halo_push_p2p(const void* first_frame, const void* last_frame, long long frame_bytes,
                                   void* const* ext_ptrs, int P, int my_rank, void* const* flag_ptrs,
                                   unsigned int* epoch_counter, int* ticket, cudaStream_t stream) {
  SyncCtx sync;
  int rcs = fill_sync(sync, flag_ptrs, epoch_counter, ticket, P, my_rank);
  if (rcs) return rcs;
  LAVIE_REQUIRE(frame_bytes > 0 && frame_bytes % 16 == 0 && al16(first_frame) && al16(last_frame), LAVIE_ERR_ALIGN,
                "halo_push: frame size and pointers must be 16-byte multiples");
  PeerPtrs d;
  int rc = fill_peers(d, ext_ptrs, P);
  if (rc) return rc;
  const long long n16 = frame_bytes / 16;
  long long blocks = (n16 + 255) / 256;
  if (blocks > 148LL * 4) blocks = 148LL * 4;
  launch_pdl(halo_push_kernel, static_cast<int>(blocks), 256, 0, stream, static_cast<const uint4*>(first_frame),
             static_cast<const uint4*>(last_frame), n16, d, P, my_rank, sync);
  return lavie_check_launch("halo_push_kernel");
}
