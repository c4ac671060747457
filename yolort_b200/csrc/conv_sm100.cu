// Conv2d(+folded BN)+SiLU(+residual) as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[m, co] = sum_{r,s,ci} X[n, ho*stride - pad + r, wo*stride - pad + s, ci] * W[co, r, s, ci]
//   m = (n*Ho + ho)*Wo + wo      (NHWC activations, K-major weights)
//
// Replaces the arithmetic of yolort/v5/models/common.py:42-73 (Conv = conv2d -> BN(eps 1e-3) -> SiLU),
// :94-116 (Bottleneck residual) and the 1x1 head convs of yolort/models/box_head.py:35-37,68-82.
//
// Persistent kernel, one CTA per SM, static round-robin over 128 x block_n output tiles:
//   warp 0     : TMA producer.  A tiles come from a 4-D im2col tensor map (the TMA engine walks 128
//                output pixels, applies padding/stride and zero-fills the halo) or, for 1x1/s1 convs,
//                from a 2-D tiled map; B tiles (weights) from a 2-D tiled map.  Both land in shared
//                memory in the 32/64/128-byte swizzled K-major layout UMMA expects.  The producer
//                runs ahead across tile boundaries, so the next tile's operands stream in while the
//                current tile is still being multiplied / drained.
//   warp 1     : allocates TMEM (two accumulator stages) and issues tcgen05.mma (M=128, N=block_n,
//                K=16) from one thread; tcgen05.commit releases smem stages and publishes accumulators.
//   warp 2     : idle (a second MMA-issuer thread was tried here and removed: DESIGN.md section 3)
//   warps 3-10 : two epilogue groups of 4 warps; group g drains accumulator stage g (tiles alternate),
//                so one tile's epilogue overlaps the next tile's MMAs and the other group's epilogue.
//                tcgen05.ld (one output pixel per thread) -> +bias -> activation -> (+residual) -> fp16/bf16
//                -> swizzled shared-memory staging -> TMA store into the NHWC destination view (a
//                channel window of a concat buffer is just a strided tensor map; ragged M is clipped
//                by the TMA unit).
// The kernel is a template over (dtype, store-box width, activation family, fused decode, shortcut, chained tail):
// every variant the hot path launches carries exactly one inlined epilogue per output (conv_epilogue.cuh,
// select_conv_kernel below).  A chained tail (conv_chain.cuh) is a second, pointwise GEMM over the tile the epilogue has
// just staged: the MMA warp issues it into a tail accumulator of the same epilogue group and a second epilogue pass stores
// it.  Opt-in build / descriptor variants kept for A/B timing and measured equal or slower (profiles/r02_ab_variants.txt):
// four accumulator stages, four epilogue groups (kGroups), a store warp (-DYB_STORE_WARP), clock64 instrumentation
// of the epilogue (-DYB_EPI_TIMING).
#include <cstdlib>

#include "common.cuh"
#include "conv_sm100.h"
#include "conv_epilogue.cuh"
#include "conv_chain.cuh"
#include "decode_common.cuh"

namespace yb {

#ifdef YB_EPI_TIMING
// Instrumented build (make variant EXTRA=-DYB_EPI_TIMING, scripts/epi_timing.py): clock64 totals of the epilogue phases of
// CTA 0 / group 0 / thread 0, accumulated over the tiles of one launch.  Never part of the shipped library.
__device__ unsigned long long g_epi_ticks[16];
#define YB_EPI_TICK(slot)                                                     \
  do {                                                                        \
    if (blockIdx.x == 0 && gtid == 0 && g == 0) {                             \
      const long long t_now_ = clock64();                                     \
      g_epi_ticks[slot] += static_cast<unsigned long long>(t_now_ - t_prev_); \
      t_prev_ = t_now_;                                                       \
    }                                                                         \
  } while (0)
#else
#define YB_EPI_TICK(slot) do { } while (0)
#endif

namespace {

constexpr int kBlockM = 128;
constexpr int kMaxStages = 12;
constexpr int kEpiGroups = 2;                         // epilogue groups of the standard variants
constexpr int kMaxGroups = 4;                         // ... of the wide variant (kGroups template parameter)
constexpr int kMaxAccStages = 4;
constexpr int kFirstEpiWarp = 3;                     // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warp 2: idle
constexpr int kThreads = 32 * kFirstEpiWarp + kEpiGroups * 128;
constexpr int kThreadsWide = 32 * kFirstEpiWarp + kMaxGroups * 128;   // 608: 104 registers per thread
constexpr int kStageBufBytes = 128 * 128;  // 128 rows x (up to) 64 columns x 2 B
constexpr int kMaxBlockN = 256;
constexpr size_t kSmemBudget = 222 * 1024;  // dynamic shared memory per CTA (227 KB limit minus static)

struct ConvKernelParams {
  int M, block_n, block_k;
  int ksize, chunks, num_k_iters;
  int mode;  // 0: 2-D tiled rows (1x1 stride 1), 1: 4-D im2col
  int HoWo, Wo, stride, pad;
  int stages;
  int kpg;         // k-iterations (A/B sub-tiles) carried by one pipeline stage
  int b_resident;  // weights of the (single) N tile stay in shared memory for the CTA's lifetime
  uint32_t b_res_bytes;
  int n_tiles, num_tiles;
  int store_cols;  // columns per TMA store box: 64 / 32 / 16
  int bias_len;    // length of the (padded) bias vector
  int kk_last;     // K=16 steps of the LAST channel chunk (Cin need not fill it: TMA zero-fills, the MMA skips)
  int dbg;         // ablation knobs, -DYB_ABLATION builds only (common.cuh)
  uint32_t a_stage_bytes, b_stage_bytes, tmem_cols, idesc;
  int acc_stride;  // TMEM columns between accumulator stages (= block_n)
  int epi_groups;  // 2, or 4 for the wide variant (host-side: selects the kernel and the block size)
  int acc_stages;  // 2 or 4 accumulator stages: with 4, each epilogue group owns two and the MMAs of its next tile have
                   // completed by the time it has stored the current one (with 2 the group waited out MMA + commit latency
                   // at the start of every tile)
  int acc2_base;   // chain: first TMEM column of the tail's two accumulators (n2 columns each), after the first conv's
  const float* bias;
  EpilogueParams ep;
  ChainParams ch;  // chained pointwise tail (kStore2 != 0 kernels)
  int decode_on;        // detection head with the fused decode epilogue (no logits are stored)
  int dec_H, dec_W;     // level extent (output pixels)
  yb_head_decode dec;   // copied from the op descriptor
};

// The k-iterations of one pipeline stage as (near) straight-line code: KK K=16 steps each, constant stride between the
// operand sub-tiles of consecutive iterations.
template <int KK>
__device__ __forceinline__ void issue_group(int cnt, bool first_group, uint32_t tmem_d, uint32_t a_lo0, uint32_t a_step16,
                                            uint32_t b_lo0, uint32_t b_step16, uint32_t desc_hi, uint32_t idesc) {
  umma_ksteps<KK>(tmem_d, a_lo0, desc_hi, b_lo0, desc_hi, idesc, first_group);
  for (int j = 1; j < cnt; ++j) umma_ksteps<KK>(tmem_d, a_lo0 + j * a_step16, desc_hi, b_lo0 + j * b_step16, desc_hi, idesc, false);
}

// kRes: the layer adds a shortcut (fp32 epilogue tail, conv_epilogue.cuh).  kStore2 != 0: a pointwise tail is chained
// onto every tile (conv_chain.cuh), its output stored in boxes of kStore2 columns.
// kGroups = 4 ("wide" variant, opt-in): four epilogue groups of four warps, each owning one accumulator stage and two
// staging buffers, 608 threads at <= 104 registers (TMEM / shortcut loads batched 16 columns at a time).  Built to test
// whether the shallow layers are bound by the latency chain of their epilogue; they are not (see conv_configure).
template <bool kBf16, int kStoreCols, bool kRareAct, bool kDecode, bool kRes = true, int kStore2 = 0, int kGroups = kEpiGroups>
__global__ void __launch_bounds__(32 * kFirstEpiWarp + kGroups * 128, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_w2,
                 const __grid_constant__ CUtensorMap tmap_out2, const ConvKernelParams p) {
  constexpr bool kChain = kStore2 != 0;
  // -DYB_STORE_WARP (A/B build only): stores issued by the otherwise idle warp 2 instead of a thread of the epilogue
  // group, so that the group's threads neither meet at a named barrier per box nor wait for the issuing thread (clock64
  // profile of the 1x1 layers, CTA 0 / group 0: store issue 11-17 %, barrier 7-10 %, store-read wait 5 % of the group's
  // time).  Measured on B200: 1.317 ms per yolov5s plan against 1.311 ms with the in-group issue -- those slices were not
  // on the kernels' critical path -- so the shipped kernels keep the in-group issue and carry none of this.
#ifdef YB_STORE_WARP
  constexpr bool kStoreWarp = !kChain && !kDecode && kGroups == kEpiGroups;
#else
  constexpr bool kStoreWarp = false;
#endif
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t acc_full[kMaxAccStages];
  __shared__ __align__(8) uint64_t acc_empty[kMaxAccStages];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[kGroups][kGroups == kEpiGroups ? kMaxBlockN : 128];   // wide variant: N tile <= 128
  __shared__ __align__(8) uint64_t a2_full[kEpiGroups];     // chain: the tile's output boxes are in shared memory
  __shared__ __align__(8) uint64_t acc2_full[kEpiGroups];   // chain: the tail's accumulator is complete
  __shared__ __align__(8) uint64_t w2_full;
  // store warp (warp 2): box_ready[g][b] = the group's box in staging buffer b is written and fenced; buf_free[g][b] = the
  // TMA store that read it has finished with the buffer
  __shared__ __align__(8) uint64_t box_ready[kEpiGroups][2];
  __shared__ __align__(8) uint64_t buf_free[kEpiGroups][2];
  __shared__ __align__(16) float s_bias2[kChain ? kEpiGroups : 1][kChain ? kMaxBlockN : 4];

  // Swizzled tiles need 1024-byte alignment.
  uint8_t* tiles = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // stage = kpg A sub-tiles followed (unless the weights are resident) by kpg B sub-tiles
  const uint32_t stage_bytes = p.kpg * (p.a_stage_bytes + (p.b_resident ? 0u : p.b_stage_bytes));
  uint8_t* b_res = tiles + static_cast<size_t>(p.stages) * stage_bytes;   // resident weights (optional)
  uint8_t* staging = b_res + p.b_res_bytes;                                // [kEpiGroups][2][kStageBufBytes]
  uint8_t* w2_res = staging + kGroups * 2 * kStageBufBytes;                // chain: resident tail weights

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&b_full, 1);
    for (int g = 0; g < kMaxAccStages; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&acc_empty[g], 4);  // one arrival per epilogue warp of the group that drains the stage
    }
    for (int g = 0; g < kEpiGroups; ++g) {
      mbar_init(&a2_full[g], 1);
      mbar_init(&acc2_full[g], 1);
    }
    mbar_init(&w2_full, 1);
    for (int g = 0; g < kEpiGroups; ++g)
      for (int b = 0; b < 2; ++b) {
        mbar_init(&box_ready[g][b], 4);   // one arrival per epilogue warp
        mbar_init(&buf_free[g][b], 1);
      }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Programmatic dependent launch: everything above overlapped the tail of the previous kernel in the
  // stream; from here on we touch memory it produced.  Let our own dependent start its prologue too.
  // Resident weights do not depend on the previous kernel: their loads are issued BEFORE the grid-dependency wait
  // and stream in while the previous kernel drains.
#ifndef YB_NO_WEIGHT_PREFETCH
  if (warp == 0 && lane == 0 && p.b_resident) {
    const uint32_t b_bytes = p.block_n * p.block_k * 2;
    mbar_expect_tx(&b_full, p.num_k_iters * b_bytes);
    for (int it = 0; it < p.num_k_iters; ++it)
      tma_load_2d(&tmap_b, &b_full, b_res + it * p.b_stage_bytes, it * p.block_k, 0);
  }
#endif
  if constexpr (kChain) {
    if (warp == 0 && lane == 0) {   // tail weights: [n2][kc] chunks, resident for the CTA's lifetime
      tma_prefetch_desc(&tmap_w2);
      tma_prefetch_desc(&tmap_out2);
      mbar_expect_tx(&w2_full, p.ch.w2_chunks * p.ch.n2 * p.ch.w2_row_bytes);
      for (int j = 0; j < p.ch.w2_chunks; ++j)
        tma_load_2d(&tmap_w2, &w2_full, w2_res + j * p.ch.w2_sub_bytes, j * (p.ch.w2_row_bytes >> 1), 0);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef YB_NO_WEIGHT_PREFETCH      // A/B build: weights fetched after the wait (scripts/ab_step.sh)
  if (warp == 0 && lane == 0 && p.b_resident) {
    const uint32_t b_bytes = p.block_n * p.block_k * 2;
    mbar_expect_tx(&b_full, p.num_k_iters * b_bytes);
    for (int it = 0; it < p.num_k_iters; ++it)
      tma_load_2d(&tmap_b, &b_full, b_res + it * p.b_stage_bytes, it * p.block_k, 0);
  }
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Warp-uniform role loops (all 32 lanes walk them, one elected lane issues): TMA / tcgen05 instructions take
    // uniform-register operands, and inside a one-lane branch ptxas wraps each of them in an elect/branch convergence
    // loop with R2UR moves (~10 instructions per MMA instead of ~3; the issuing thread is the critical path of the
    // shallow layers).
    if (YB_ROLE_LANES(lane)) {
      const uint32_t a_bytes = kBlockM * p.block_k * 2, b_bytes = p.block_n * p.block_k * 2;
      int kit = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int n0 = (tile - m_tile * p.n_tiles) * p.block_n;
        const int m0 = m_tile * kBlockM;
        int cw = 0, ch = 0, cn = 0;
        if (p.mode == 1) {
          cn = m0 / p.HoWo;
          const int rem = m0 - cn * p.HoWo;
          const int ho = rem / p.Wo;
          const int wo = rem - ho * p.Wo;
          ch = ho * p.stride - p.pad;
          cw = wo * p.stride - p.pad;
        }
        for (int it0 = 0; it0 < p.num_k_iters; it0 += p.kpg, ++kit) {
          const int cnt = min(p.kpg, p.num_k_iters - it0);
          const int s = kit % p.stages;
          const uint32_t ph = (kit / p.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = tiles + s * stage_bytes;
          uint8_t* b_dst = a_dst + p.kpg * p.a_stage_bytes;
          if (YB_DBG(p, 8)) {   // ablation: no loads at all, only the pipeline handshake
            if (YB_ELECT()) mbar_arrive(&full_bar[s]);
            continue;
          }
          if (YB_ELECT()) {
            mbar_expect_tx(&full_bar[s], cnt * (a_bytes + (p.b_resident ? 0u : b_bytes)));
            for (int j = 0; j < cnt; ++j) {
              const int it = it0 + j;
              const int tap = it / p.chunks;
              const int chunk = it - tap * p.chunks;
              if (p.mode == 0) {
                tma_load_2d(&tmap_a, &full_bar[s], a_dst + j * p.a_stage_bytes, chunk * p.block_k, m0);
              } else {
                const int r = tap / p.ksize;
                const int sx = tap - r * p.ksize;
                tma_load_im2col_4d(&tmap_a, &full_bar[s], a_dst + j * p.a_stage_bytes, chunk * p.block_k, cw, ch, cn,
                                   static_cast<uint16_t>(sx), static_cast<uint16_t>(r));
              }
              if (!p.b_resident) tma_load_2d(&tmap_b, &full_bar[s], b_dst + j * p.b_stage_bytes, it * p.block_k, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // (A second issuing thread for alternate tiles was tried and removed: two consumers that are several phases
    // apart on the same full/empty mbarriers alias under parity waits.)
    if (YB_ROLE_LANES(lane)) {
      const uint32_t row_bytes = p.block_k * 2;
      const int kk = p.block_k >> 4;
      if (p.b_resident) {
        mbar_wait(&b_full, 0);
        tc_fence_after();
      }
      const uint32_t b_res_addr = smem_u32(b_res);
      const uint32_t desc_hi = static_cast<uint32_t>(make_kmajor_desc(0, row_bytes) >> 32);
      const uint32_t a_step16 = p.a_stage_bytes >> 4, b_step16 = p.b_stage_bytes >> 4;
      int kit = 0, lt = 0;
      // chain: the tail GEMM of tile t is issued after the MMAs of tile t + 1 (its operand is what the epilogue of
      // tile t writes, which takes about as long as the next tile's MMAs).  The tail has its own accumulator columns
      // (one per epilogue group): with D2 aliased onto the drained first accumulator, the next-but-one tile's MMAs had
      // to wait for the tail's epilogue and -- this thread issuing in order -- the two epilogue groups ran in lock step
      // (measured: 64->64 + 32->32 at 160x160 74 us fused vs 82 us as two launches).  No "tail accumulator empty"
      // barrier is needed: group g drains tail t - 2 before it writes the boxes of tile t, whose a2_full arrival
      // this thread waits for.
      int pend = -1;
      uint32_t ph2 = 0;
      const uint32_t a2_hi = static_cast<uint32_t>(make_kmajor_desc(0, p.ch.own_row_bytes) >> 32);
      const uint32_t w2_hi = static_cast<uint32_t>(make_kmajor_desc(0, p.ch.w2_row_bytes) >> 32);
      const uint32_t w2_lo0 = (smem_u32(w2_res) & 0x3FFFFu) >> 4;
      auto issue_tail = [&](int gsel) {
        mbar_wait(&a2_full[gsel], (ph2 >> gsel) & 1u);
        ph2 ^= 1u << gsel;
        tc_fence_after();
        const uint32_t d2 = tmem_base + p.acc2_base + gsel * p.ch.n2;
        const uint32_t stag_lo = (smem_u32(staging + static_cast<size_t>(gsel) * 2 * kStageBufBytes) & 0x3FFFFu) >> 4;
        if (YB_ELECT()) {
          for (int j = 0; j < p.ch.own_chunks; ++j)
            umma_ksteps_rt(p.ch.ksteps, d2, stag_lo + j * (kStageBufBytes >> 4), a2_hi, w2_lo0 + j * (p.ch.w2_sub_bytes >> 4), w2_hi,
                           p.ch.idesc2, j == 0);
          umma_commit(&acc2_full[gsel]);
        }
      };
      if constexpr (kChain) {
        mbar_wait(&w2_full, 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
        const int as = lt % p.acc_stages;   // accumulator stage; the epilogue group is lt & 1 (stages are even / odd alike)

        const uint32_t aph = (lt / p.acc_stages) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * p.acc_stride;
        int chunk = 0;   // channel chunk of the running k-iteration (it = tap * chunks + chunk)
        for (int it0 = 0; it0 < p.num_k_iters; it0 += p.kpg, ++kit) {
          const int cnt = min(p.kpg, p.num_k_iters - it0);
          const int s = kit % p.stages;
          const uint32_t ph = (kit / p.stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(tiles + s * stage_bytes);
          const uint32_t b_base = a_base + p.kpg * p.a_stage_bytes;
          // descriptors differ only in their 14-bit start-address field (16-byte units)
          const uint32_t a_lo0 = (a_base & 0x3FFFFu) >> 4;
          const uint32_t b_lo0 = ((p.b_resident ? b_res_addr + it0 * p.b_stage_bytes : b_base) & 0x3FFFFu) >> 4;
          if (YB_ELECT()) {
            if (!YB_DBG(p, 2)) {
              if (p.kk_last == kk) {   // every chunk is full (the common case): no per-iteration dispatch
                if (kk == 4)
                  issue_group<4>(cnt, it0 == 0, tmem_d, a_lo0, a_step16, b_lo0, b_step16, desc_hi, p.idesc);
                else if (kk == 2)
                  issue_group<2>(cnt, it0 == 0, tmem_d, a_lo0, a_step16, b_lo0, b_step16, desc_hi, p.idesc);
                else
                  issue_group<1>(cnt, it0 == 0, tmem_d, a_lo0, a_step16, b_lo0, b_step16, desc_hi, p.idesc);
              } else {
                int ch = chunk;
                for (int j = 0; j < cnt; ++j) {
                  umma_ksteps_rt(ch == p.chunks - 1 ? p.kk_last : kk, tmem_d, a_lo0 + j * a_step16, desc_hi,
                                 b_lo0 + j * b_step16, desc_hi, p.idesc, (it0 | j) == 0);
                  if (++ch == p.chunks) ch = 0;
                }
              }
            }
            umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
          }
          chunk = (chunk + cnt) % p.chunks;
        }
        if (YB_ELECT()) umma_commit(&acc_full[as]);  // accumulator of this tile complete
        if constexpr (kChain) {
          if (pend >= 0) issue_tail(pend);
          pend = lt & 1;   // the group that drains this tile (chained kernels run two groups)
        }
      }
      if constexpr (kChain) {
        if (pend >= 0) issue_tail(pend);
      }
    }
  } else if (warp == 2) {
    // ===================== store warp =====================
    if constexpr (kStoreWarp) {
      if (YB_ROLE_LANES(lane)) {
        const int store_cols = kStoreCols != 0 ? kStoreCols : p.store_cols;
        int lt = 0;
        int cnt[kEpiGroups] = {0, 0};          // boxes stored so far per group
        int pend_g = -1, pend_b = 0;           // the store issued last: its buffer is released one store later
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
          const int g = lt & 1;
          const int m_tile = tile / p.n_tiles;
          const int n0 = (tile - m_tile * p.n_tiles) * p.block_n;
          const int m0 = m_tile * kBlockM;
          for (int c0 = 0; c0 < p.block_n; c0 += store_cols) {
            const int b = cnt[g] & 1;
            mbar_wait(&box_ready[g][b], (cnt[g] >> 1) & 1);
            if (lane == 0) {   // bulk async-groups are per thread: the same lane issues and waits
              if (n0 + c0 < p.ep.Cout) tma_store_2d(&tmap_out, staging + (static_cast<size_t>(g) * 2 + b) * kStageBufBytes, n0 + c0, m0);
              tma_store_commit();
              if (pend_g >= 0) {               // every store but the one just issued has finished reading shared memory
                tma_store_wait_read<1>();
                mbar_arrive(&buf_free[pend_g][pend_b]);
              }
            }
            pend_g = g;
            pend_b = b;
            ++cnt[g];
          }
        }
        if (lane == 0) {
          tma_store_wait_all<0>();             // the kernel must not end with stores in flight
          if (pend_g >= 0) mbar_arrive(&buf_free[pend_g][pend_b]);
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue groups =====================
    const int g = (warp - kFirstEpiWarp) >> 2;   // epilogue group (two groups: drains stages g, g + 2; four: stage g)
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int gtid = threadIdx.x - 32 * kFirstEpiWarp - g * 128;
    const int row_in_tile = q * 32 + lane;
    const bool issuer = (gtid == 0);
    const uint32_t bar_id = 1 + g;
    const int store_cols = kStoreCols != 0 ? kStoreCols : p.store_cols;
    const int row_bytes = store_cols * 2;
    uint8_t* my_staging = staging + static_cast<size_t>(g) * 2 * kStageBufBytes;
    float* bias_s = s_bias[g];
    int lt = 0, store_idx = 0;
    uint32_t ph2 = 0;
    if constexpr (kChain) {   // the tail has a single N tile: its bias is the same for every tile (visible after the first barrier)
      for (int i = gtid; i < p.ch.n2; i += 128) s_bias2[g][i] = (i < p.ch.bias2_len) ? __ldg(p.ch.bias2 + i) : 0.f;
    }
    // Every tile of this CTA has the same N tile when the grid is a multiple of the N-tile count (always with one N
    // tile): the bias is then loaded ONCE instead of per tile (a global-load latency plus a barrier per tile).
#ifdef YB_NO_BIAS_HOIST       // A/B build (scripts/ab_step.sh)
    const bool fixed_n = false;
#else
    const bool fixed_n = (gridDim.x % p.n_tiles) == 0;
#endif
    if (fixed_n) {
      const int n0f = (blockIdx.x % p.n_tiles) * p.block_n;
      for (int i = gtid; i < p.block_n; i += 128) bias_s[i] = (n0f + i < p.bias_len) ? __ldg(p.bias + n0f + i) : 0.f;
      named_bar_sync(bar_id, 128);
    }
#ifdef YB_EPI_TIMING
    long long t_prev_ = clock64();
#endif
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      if ((lt % kGroups) != g) continue;
      YB_EPI_TICK(0);   // loop overhead / previous tile's tail
      const int as = lt % p.acc_stages;
      const uint32_t aph = (lt / p.acc_stages) & 1;
      const int m_tile = tile / p.n_tiles;
      const int n0 = (tile - m_tile * p.n_tiles) * p.block_n;
      const int m0 = m_tile * kBlockM;
      const long long row = static_cast<long long>(m0) + row_in_tile;
      const bool row_ok = row < p.M;
      if (YB_DBG(p, 16)) {   // ablation: accumulator handshake only
        mbar_wait(&acc_full[as], aph);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        continue;
      }
      if (!fixed_n) {
        for (int i = gtid; i < p.block_n; i += 128) bias_s[i] = (n0 + i < p.bias_len) ? __ldg(p.bias + n0 + i) : 0.f;
      }
      if constexpr (kChain) {
        // chain tiles index the staging buffers by box (box b of the tile stays in buffer b until the tail GEMM has
        // read it), so the previous tile's stores must have drained both buffers before this tile writes them
        if (issuer) tma_store_wait_read<0>();
      }
      if (kChain || !fixed_n) named_bar_sync(bar_id, 128);
      YB_EPI_TICK(1);   // tile set-up
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      YB_EPI_TICK(2);   // waiting for the accumulator
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.acc_stride;
      if constexpr (kDecode) {
        // ---- fused post-processing front end (yolort/models/box_head.py:328-360,418) ----
        // This thread owns one output pixel: all A*(nc+5) logits of its anchors sit in its TMEM lane.  Per anchor:
        // objectness first (score = cls*obj <= obj, so most anchors stop there), then the classes in 16-column
        // TMEM reads; candidates go straight into the NMS arena.  TMEM reads are warp-collective, so the class
        // scan of an anchor runs whenever ANY pixel of the warp passed the objectness test.
        const yb_head_decode& D = p.dec;
        const int K = D.n_classes + 5;
        const int hw = p.dec_H * p.dec_W;
        const int n_img = static_cast<int>(row / hw);
        const int rem = static_cast<int>(row - static_cast<long long>(n_img) * hw);
        const int py = rem / p.dec_W, px = rem - py * p.dec_W;
        // pass 1: objectness of every anchor (16-column ALIGNED TMEM reads: unaligned start columns are an order
        // of magnitude slower)
        constexpr int kMaxA = 4;
        float obj[kMaxA], lt[kMaxA], bx[kMaxA][4];
        bool pass[kMaxA], emitted[kMaxA];
        uint32_t anchors_alive = 0;   // warp-uniform: bit a set iff some pixel of the warp passed anchor a
#pragma unroll
        for (int a = 0; a < kMaxA; ++a) {
          obj[a] = 0.f; lt[a] = INFINITY; pass[a] = false; emitted[a] = false;
          bx[a][0] = bx[a][1] = bx[a][2] = bx[a][3] = 0.f;
          if (a < D.n_anchors) {
            const int oc = a * K + 4;
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + (oc & ~15), v);
            tmem_ld_wait();
            float xo = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) xo = (j == (oc & 15)) ? __uint_as_float(v[j]) : xo;
            obj[a] = sigmoidf_ref(xo + bias_s[oc]);
            pass[a] = row_ok && obj[a] > D.score_thresh;
            if (pass[a]) {
              // cheap conservative pre-test on raw class logits: sigmoid(x)*obj > thr  <=>  x > logit(thr/obj)
              const float r = D.score_thresh / obj[a];
              lt[a] = r < 1.0f ? __logf(r / (1.0f - r)) - 1e-2f : INFINITY;
              if (!(D.score_thresh > 0.f)) lt[a] = -INFINITY;
            }
            if (__any_sync(0xffffffffu, pass[a])) anchors_alive |= 1u << a;
          }
        }
        // pass 2: aligned chunks that overlap an alive anchor
        if (anchors_alive) {
          int ca = 0, cr = 0;   // anchor / offset-in-anchor of the chunk's first column
          for (int c = 0; c < D.n_anchors * K; c += 16) {
            // anchors overlapping [c, c+16)
            const int a_first = ca;
            const int a_last = (cr + 15 >= K) ? ca + 1 : ca;
            const bool need = ((anchors_alive >> a_first) & 1u) || (a_last < D.n_anchors && ((anchors_alive >> a_last) & 1u));
            if (need) {
              uint32_t u[16];
              tmem_ld_32x32b_x16(taddr + c, u);
              tmem_ld_wait();
              int a = ca, r = cr;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (a < D.n_anchors) {
                  const float x = __uint_as_float(u[j]) + bias_s[c + j];
#pragma unroll
                  for (int aa = 0; aa < kMaxA; ++aa) {
                    if (aa == a && pass[aa]) {
                      if (r < 4) {
                        bx[aa][r & 3] = x;
                      } else if (r >= 5 && x > lt[aa]) {
                        const float score = __fmul_rn(sigmoidf_ref(x), obj[aa]);
                        if (score > D.score_thresh) {
                          const int anchor_flat = D.level_start + (aa * p.dec_H + py) * p.dec_W + px;
                          emit_candidate(D.keys, D.img_count, D.cap_per_image, n_img, anchor_flat, D.n_classes, r - 5, score);
                          emitted[aa] = true;
                        }
                      }
                    }
                  }
                }
                if (++r == K) {
                  r = 0;
                  ++a;
                }
              }
            }
            cr += 16;
            if (cr >= K) {
              cr -= K;
              ++ca;
            }
          }
#pragma unroll
          for (int aa = 0; aa < kMaxA; ++aa) {
            if (emitted[aa]) {
              const int anchor_flat = D.level_start + (aa * p.dec_H + py) * p.dec_W + px;
              const float4 b = decode_box(sigmoidf_ref(bx[aa][0]), sigmoidf_ref(bx[aa][1]), sigmoidf_ref(bx[aa][2]),
                                          sigmoidf_ref(bx[aa][3]), px, py, D.stride_px, D.anchors_px[2 * aa],
                                          D.anchors_px[2 * aa + 1]);
              reinterpret_cast<float4*>(D.boxes)[static_cast<long long>(n_img) * D.anchors_per_image + anchor_flat] = b;
              atomicMax(&D.img_maxc[n_img], float_to_ordered_int(fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        continue;
      }
      for (int c0 = 0; c0 < p.block_n; c0 += store_cols, ++store_idx) {
        // Two staging buffers, one barrier per box: before the barrier below the issuer waits until the PREVIOUS
        // store has finished reading its buffer, which is the one the next box will overwrite.
        uint8_t* buf = my_staging + (kChain ? ((c0 / store_cols) & 1) : (store_idx & 1)) * kStageBufBytes;
        uint8_t* my_row = buf + row_in_tile * row_bytes;
        if constexpr (kStoreWarp) {   // the store that read this buffer two boxes ago has released it
          mbar_wait(&buf_free[g][store_idx & 1], ((store_idx >> 1) & 1) ^ 1);
        }
        if (!YB_DBG(p, 1)) {
          epilogue_box_select<kBf16, kStoreCols, kRareAct, kRes, kGroups == kEpiGroups ? 32 : 16>(p.ep, store_cols, taddr + c0, bias_s + c0, row, row_ok, n0 + c0, my_row, row_in_tile);
        }
        YB_EPI_TICK(3);   // TMEM load + bias/activation(/shortcut) + swizzled shared-memory writes of one box
        if (c0 + store_cols >= p.block_n) {
          // all TMEM reads of this tile are done: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        fence_proxy_async_smem();
        YB_EPI_TICK(4);   // fences
        if constexpr (kStoreWarp) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&box_ready[g][store_idx & 1]);   // warp 2 issues the store
          YB_EPI_TICK(7);
          continue;
        }
        if constexpr (!kChain) {
          if (issuer) tma_store_wait_read<0>();
        }
        YB_EPI_TICK(5);   // previous store's shared-memory read
        named_bar_sync(bar_id, 128);
        YB_EPI_TICK(6);   // group barrier
        if (issuer) {
          if ((!kChain || p.ch.store_first) && n0 + c0 < p.ep.Cout && !YB_DBG(p, 5)) tma_store_2d(&tmap_out, buf, n0 + c0, m0);
          tma_store_commit();
        }
        YB_EPI_TICK(7);   // store issue
      }
      if constexpr (kChain) {
        // every box of the tile is in shared memory, visible to the async proxy (fence + barrier above), and every
        // TMEM read of the group has retired: the MMA warp may run the tail GEMM
        if (issuer) mbar_arrive(&a2_full[g]);
        mbar_wait(&acc2_full[g], ph2);
        ph2 ^= 1u;
        tc_fence_after();
        const uint32_t taddr2 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + p.acc2_base + g * p.ch.n2;
        constexpr int kRow2 = kStore2 * 2;
        // the tail's boxes reuse the staging buffers: the operand boxes are dead (the tail GEMM has completed), but the
        // stores of the first output may still be reading them
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
        for (int c0 = 0; c0 < p.ch.n2; c0 += kStore2) {
          uint8_t* buf = my_staging + ((c0 / kStore2) & 1) * kStageBufBytes;
          epilogue_box<kBf16, kStore2, false, kBf16>(p.ch.ep2, taddr2 + c0, s_bias2[g] + c0, row, row_ok, c0, buf + row_in_tile * kRow2, row_in_tile);
          if (c0 + kStore2 >= p.ch.n2) tc_fence_before();   // ordered before this group's next a2_full arrival
          fence_proxy_async_smem();
          if (issuer) tma_store_wait_read<0>();   // box k + 1 overwrites the buffer of box k - 1
          named_bar_sync(bar_id, 128);
          if (issuer) {
            if (c0 < p.ch.ep2.Cout) tma_store_2d(&tmap_out2, buf, c0, m0);
            tma_store_commit();
          }
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

int load_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col) return YB_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  YB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  YB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
             "cuTensorMapEncodeTiled not available from the driver");
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  YB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
  YB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
             "cuTensorMapEncodeIm2col not available from the driver");
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  return YB_OK;
}

CUtensorMapSwizzle swizzle_for_row_bytes(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while (static_cast<int>(c) < n) c <<= 1;
  return c;
}

}  // namespace

using ConvKernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                              const ConvKernelParams);

// One kernel per (dtype, store-box width, shortcut) for the SiLU / linear epilogue; the r3.1 activations and the fused
// decode epilogue are separate kernels that pick the store width at run time (see conv_epilogue.cuh).  fp16 layers
// without a shortcut take the packed half2 epilogue tail (kRes = false); bf16 always runs the fp32 tail, so it has one
// variant.  Chained tails (conv_chain.cuh): 64-column boxes for the first output, 64 / 32 for the tail.
template <bool kBf16>
ConvKernelFn select_conv_kernel_t(const ConvKernelParams& kp) {
  if (kp.decode_on) return conv_umma_kernel<kBf16, 0, false, true>;
  if (kp.ep.act >= YB_ACT_HARDSWISH) return conv_umma_kernel<kBf16, 0, true, false>;
  constexpr bool kResAlways = kBf16;
  const bool res = kResAlways || kp.ep.residual != nullptr;
  if (kp.ch.on) {   // conv_configure admits exactly these shapes
    const int s2 = chain_store2_cols(kp.ch.n2);
    if (res) return s2 == 64 ? conv_umma_kernel<kBf16, 64, false, false, true, 64> : conv_umma_kernel<kBf16, 64, false, false, true, 32>;
    return s2 == 64 ? conv_umma_kernel<kBf16, 64, false, false, kResAlways, 64> : conv_umma_kernel<kBf16, 64, false, false, kResAlways, 32>;
  }
  if (kp.epi_groups == kMaxGroups)   // conv_configure admits only 64-column boxes here
    return res ? conv_umma_kernel<kBf16, 64, false, false, true, 0, kMaxGroups>
               : conv_umma_kernel<kBf16, 64, false, false, kResAlways, 0, kMaxGroups>;
  if (res) {
    switch (kp.store_cols) {
      case 64: return conv_umma_kernel<kBf16, 64, false, false, true>;
      case 32: return conv_umma_kernel<kBf16, 32, false, false, true>;
      default: return conv_umma_kernel<kBf16, 16, false, false, true>;
    }
  }
  switch (kp.store_cols) {
    case 64: return conv_umma_kernel<kBf16, 64, false, false, kResAlways>;
    case 32: return conv_umma_kernel<kBf16, 32, false, false, kResAlways>;
    default: return conv_umma_kernel<kBf16, 16, false, false, kResAlways>;
  }
}
ConvKernelFn select_conv_kernel(const ConvKernelParams& kp) {
  return kp.ep.is_bf16 ? select_conv_kernel_t<true>(kp) : select_conv_kernel_t<false>(kp);
}

struct ConvOp {
  PatchConvOp* patch = nullptr;  // non-null: this conv runs on the halo-patch kernel
  CUtensorMap tmap_a, tmap_b, tmap_out, tmap_w2, tmap_out2;
  ConvKernelParams kp;
  ConvKernelFn fn = nullptr;
  dim3 grid;
  size_t smem_bytes;
};

// Pure host logic: validates the op and derives tiling, pipeline depth, shared-memory layout and launch shape
// (no driver calls: yb_conv_chain_supported runs this without a GPU).
static int conv_configure(const yb_op_desc& d, ConvKernelParams& kp, dim3& grid, size_t& smem_bytes) {
  YB_REQUIRE(d.dtype == YB_F16 || d.dtype == YB_BF16, "conv: dtype must be f16 or bf16");
  YB_REQUIRE(d.ksize >= 1 && d.ksize <= 7 && d.stride >= 1 && d.stride <= 2, "conv: ksize/stride");
  YB_REQUIRE(d.act >= YB_ACT_NONE && d.act <= YB_ACT_LEAKY01, "conv: unknown activation %d", d.act);
  YB_REQUIRE(d.Cin % 8 == 0 && d.in_cstride % 8 == 0 && d.in_cstride >= d.Cin,
             "conv: Cin/in_cstride must be multiples of 8 (16-byte TMA granularity), got %d/%d", d.Cin,
             d.in_cstride);
  YB_REQUIRE(d.Cout % 8 == 0 && d.out_cstride % 8 == 0 && d.out_cstride >= d.Cout,
             "conv: Cout/out_cstride must be multiples of 8, got %d/%d", d.Cout, d.out_cstride);
  YB_REQUIRE(d.Cin_pad % 16 == 0 && d.Cin_pad >= d.Cin, "conv: Cin_pad must be a multiple of 16");
  YB_REQUIRE(d.Cout_pad % 16 == 0 && d.Cout_pad >= d.Cout, "conv: Cout_pad must be a multiple of 16");
  YB_REQUIRE((reinterpret_cast<uintptr_t>(d.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d.weight) & 15) == 0,
             "conv: tensors must be 16-byte aligned");
  YB_REQUIRE(d.residual == nullptr ||
                 ((reinterpret_cast<uintptr_t>(d.residual) & 15) == 0 && d.res_cstride % 8 == 0),
             "conv: residual alignment");
  const int Ho = (d.H + 2 * d.pad - d.ksize) / d.stride + 1;
  const int Wo = (d.W + 2 * d.pad - d.ksize) / d.stride + 1;
  YB_REQUIRE(Ho == d.Ho && Wo == d.Wo, "conv: output extent mismatch (%d,%d) vs (%d,%d)", Ho, Wo, d.Ho,
             d.Wo);
  const long long M_ll = static_cast<long long>(d.N) * Ho * Wo;
  YB_REQUIRE(M_ll > 0 && M_ll < (1ll << 31), "conv: M out of range");

  YB_REQUIRE(!(d.reserved & 2) || patch_conv_eligible(d),
             "conv: banded stem weights (reserved bit 1) need the halo-patch kernel, which this %dx%d map does not qualify for",
             d.H, d.W);
  if (patch_conv_eligible(d)) return YB_OK;   // configured by patch_conv_configure
  kp = ConvKernelParams();
  kp.M = static_cast<int>(M_ll);
  kp.ep.Cout = d.Cout;
  const int m_tiles = (kp.M + kBlockM - 1) / kBlockM;
  const int sms = num_sms();
  // N tile: the whole Cout up to 256 columns (fewest A re-reads); halve it when that leaves fewer
  // than two tiles per SM so the persistent grid balances better.
  int n_tiles = (d.Cout + kMaxBlockN - 1) / kMaxBlockN;
  int block_n = (((d.Cout + n_tiles - 1) / n_tiles) + 15) / 16 * 16;
  if (m_tiles * n_tiles < 2 * sms && block_n > 128 && block_n % 32 == 0 && d.chain == nullptr) {
    block_n /= 2;
    n_tiles = (d.Cout + block_n - 1) / block_n;
  }
  kp.block_n = block_n;
  kp.n_tiles = n_tiles;
  kp.num_tiles = m_tiles * n_tiles;
  kp.block_k = (d.Cin_pad % 64 == 0) ? 64 : ((d.Cin_pad % 32 == 0) ? 32 : 16);
  kp.ksize = d.ksize;
  kp.chunks = d.Cin_pad / kp.block_k;
  kp.num_k_iters = d.ksize * d.ksize * kp.chunks;
  kp.mode = (d.ksize == 1 && d.stride == 1 && d.pad == 0) ? 0 : 1;
  kp.HoWo = Ho * Wo;
  kp.Wo = Wo;
  kp.stride = d.stride;
  kp.pad = d.pad;
  kp.decode_on = 0;
  if (d.decode != nullptr) {
    const yb_head_decode& dd = *d.decode;
    const int width = dd.n_anchors * (dd.n_classes + 5);
    if (!(dd.n_anchors > 0 && dd.n_anchors <= 4 && width <= kMaxBlockN && width <= d.Cout_pad && d.ksize == 1 &&
          dd.keys && dd.boxes && dd.img_count && dd.img_maxc)) {
      set_error("conv: fused decode needs a 1x1 head with n_anchors*(n_classes+5) <= %d and a candidate arena", kMaxBlockN);
      return YB_ERR_INVALID;
    }
    // all anchors of a pixel must sit in one accumulator row: one N tile covering the whole head
    n_tiles = 1;
    block_n = (d.Cout + 15) / 16 * 16;
    kp.block_n = block_n;
    kp.n_tiles = 1;
    kp.num_tiles = m_tiles;
    kp.decode_on = 1;
    kp.dec = dd;
    kp.dec_H = Ho;
    kp.dec_W = Wo;
  }
  kp.store_cols = (block_n % 64 == 0) ? 64 : ((block_n % 32 == 0) ? 32 : 16);
  kp.bias_len = d.Cout_pad;
  kp.dbg = 0;
#ifdef YB_ABLATION
  if (const char* e = getenv("YB_CONV_DBG")) kp.dbg = atoi(e);
#endif
  kp.kk_last = (d.Cin - (kp.chunks - 1) * kp.block_k + 15) / 16;
  if (kp.kk_last < 1) kp.kk_last = 1;
  if (kp.kk_last > (kp.block_k >> 4)) kp.kk_last = kp.block_k >> 4;
  kp.a_stage_bytes = kBlockM * kp.block_k * 2;
  kp.b_stage_bytes = (static_cast<uint32_t>(kp.block_n * kp.block_k * 2) + 1023u) & ~1023u;
  kp.acc_stride = kp.block_n;
  kp.acc_stages = 2;
  kp.ch.on = 0;
  size_t chain_bytes = 0;
  if (d.chain != nullptr) {
    YB_REQUIRE(kp.store_cols == 64, "conv: a chained tail needs 64-column output boxes (Cout %% 64 == 0), got Cout=%d", d.Cout);
    const char* why = chain_setup(d, kp.block_n, kp.n_tiles, kp.store_cols, /*allow_extra=*/false, &kp.ch);
    YB_REQUIRE(why == nullptr, "conv: chained tail not supported here: %s", why);
    const int s2 = chain_store2_cols(kp.ch.n2);
    YB_REQUIRE(s2 == 64 || s2 == 32, "conv: the tail's Cout_pad must be a multiple of 32, got %d", kp.ch.n2);
    kp.acc_stages = ((d.reserved & 16) && 4 * kp.acc_stride + 2 * kp.ch.n2 <= 512) ? 4 : 2;
    kp.acc2_base = kp.acc_stages * kp.acc_stride;
    YB_REQUIRE(kp.acc2_base + 2 * kp.ch.n2 <= 512, "conv: accumulators of the convolution and its tail exceed TMEM (%d + %d columns)",
               kp.acc2_base, 2 * kp.ch.n2);
    chain_bytes = static_cast<size_t>(kp.ch.w2_chunks) * kp.ch.w2_sub_bytes;
  }
  // Wide variant (four epilogue groups, see the kernel's header), opt-in through reserved bit 5: plain SiLU / linear
  // layers with 64-column store boxes, an N tile of at most 128 columns and at least four tiles per CTA; it needs 128 KB
  // of staging, so it is dropped again below if fewer than three pipeline stages would be left.  Measured on B200
  // (yolov5s batch 32, every convolution its own launch): 1.333 ms per plan with it, 1.327 ms without -- the 1x1 layers at
  // 160 x 160 / 80 x 80 already move 4.9 TB/s of mixed read + write traffic, which is what this part sustains.
  const int grid_x = kp.num_tiles < sms ? kp.num_tiles : sms;
  bool wide = d.chain == nullptr && !kp.decode_on && d.act < YB_ACT_HARDSWISH && kp.store_cols == 64 && kp.block_n <= 128 &&
              kp.num_tiles >= 4 * grid_x && (d.reserved & 32);
  // shared-memory pipeline for a given number of epilogue groups (each owns two staging buffers)
  struct Pipe {
    size_t fixed;
    uint32_t stage_bytes;
    int stages;
  };
  auto size_pipeline = [&](int epi_groups) {
    Pipe pp;
    pp.fixed = static_cast<size_t>(epi_groups) * 2 * kStageBufBytes + 1024 + chain_bytes;
    // Weights stay resident in shared memory when the layer has a single N tile and they are small:
    // the persistent CTA then streams only activations (halves the L2->SM traffic of the shallow layers).
    const size_t b_total = static_cast<size_t>(kp.num_k_iters) * kp.b_stage_bytes;
    kp.b_resident = (n_tiles == 1 && b_total <= 80 * 1024) ? 1 : 0;
    kp.b_res_bytes = kp.b_resident ? static_cast<uint32_t>(b_total) : 0u;
    // k-iterations per pipeline stage: aim at ~32 KB per stage so that one mbarrier round trip moves
    // enough bytes (a 16-channel tap is only 4 KB), in near-equal groups.
    const uint32_t per_iter = kp.a_stage_bytes + (kp.b_resident ? 0u : kp.b_stage_bytes);
    const size_t avail = kSmemBudget - pp.fixed - kp.b_res_bytes;
    size_t target = avail / 3 < 32 * 1024 ? avail / 3 : 32 * 1024;   // keep at least three stages in flight
    int kpg_max = static_cast<int>(target / per_iter);
    if (kpg_max < 1) kpg_max = 1;
    if (kpg_max > kp.num_k_iters) kpg_max = kp.num_k_iters;
    const int groups = (kp.num_k_iters + kpg_max - 1) / kpg_max;
    kp.kpg = (kp.num_k_iters + groups - 1) / groups;
    pp.stage_bytes = kp.kpg * per_iter;
    pp.stages = static_cast<int>(avail / pp.stage_bytes);
    if (pp.stages > kMaxStages) pp.stages = kMaxStages;
    return pp;
  };
  Pipe pp = size_pipeline(wide ? kMaxGroups : kEpiGroups);
  if (wide && pp.stages < 3) {   // 128 KB of staging would starve the operand pipeline: back to two groups
    wide = false;
    pp = size_pipeline(kEpiGroups);
  }
  kp.epi_groups = wide ? kMaxGroups : kEpiGroups;
  const size_t fixed = pp.fixed;
  const uint32_t stage_bytes = pp.stage_bytes;
  int stages = pp.stages;
  if (stages < 2) stages = 2;
  kp.stages = stages;
  if (wide) kp.acc_stages = 4;
  if (!kp.ch.on && 4 * kp.acc_stride <= 512 && (d.reserved & 16)) kp.acc_stages = 4;   // reserved bit 4: four stages (measured equal or slower: opt-in)
  kp.tmem_cols = pow2_cols(kp.acc_stages * kp.acc_stride + (kp.ch.on ? 2 * kp.ch.n2 : 0));
  kp.ep.is_bf16 = d.dtype == YB_BF16;
  const uint32_t fmt = kp.ep.is_bf16 ? 1u : 0u;
  kp.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(kp.block_n >> 3) << 17) |
             (static_cast<uint32_t>(kBlockM >> 4) << 24);
  kp.ep.act = d.act;
  kp.bias = d.bias;
  kp.ep.residual = d.residual;
  kp.ep.res_cstride = d.res_cstride;
  grid = dim3(kp.num_tiles < sms ? kp.num_tiles : sms, 1, 1);
  // >= 120 KB so that two CTAs can never share an SM (each owns up to all 512 TMEM columns)
  size_t smem = static_cast<size_t>(stages) * stage_bytes + kp.b_res_bytes + fixed;
  YB_REQUIRE(smem <= kSmemBudget, "conv: %zu bytes of shared memory needed, %zu available", smem, kSmemBudget);
  if (smem < 120 * 1024) smem = 120 * 1024;
  smem_bytes = smem;
  return YB_OK;
}

int conv_configure_check(const yb_op_desc& d, int* info) {
  ConvKernelParams kp;
  dim3 grid;
  size_t smem = 0;
  const int rc = conv_configure(d, kp, grid, smem);
  if (rc == YB_OK && info && !patch_conv_eligible(d)) {   // yb_conv_config: see include/yolort_b200.h
    info[0] = 0;
    info[1] = kp.block_n;
    info[2] = kp.n_tiles;
    info[3] = kp.b_resident;
    info[4] = 1;
    info[5] = kp.stages;
    info[6] = kp.kpg;
    info[7] = kp.store_cols;
    info[8] = kp.epi_groups;      // (im2col / 1x1 kernel: epilogue groups; each has two staging buffers)
    info[9] = static_cast<int>(smem);
    info[10] = static_cast<int>(grid.x);
    info[11] = kp.ch.on;
  }
  return rc;
}

int conv_op_create(const yb_op_desc& d, ConvOp** out) {
  int rc = load_driver_entry_points();
  if (rc != YB_OK) return rc;
  ConvOp* op = new ConvOp();
  rc = conv_configure(d, op->kp, op->grid, op->smem_bytes);
  if (rc != YB_OK) {
    delete op;
    return rc;
  }
  if (patch_conv_eligible(d)) {
    rc = patch_conv_create(d, g_encode_tiled, &op->patch);
    if (rc != YB_OK) {
      delete op;
      return rc;
    }
    *out = op;
    return YB_OK;
  }
  ConvKernelParams& kp = op->kp;
  const int Ho = d.Ho, Wo = d.Wo;
  (void)Ho; (void)Wo;

  const CUtensorMapDataType dt =
      kp.ep.is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = swizzle_for_row_bytes(kp.block_k * 2);
  CUresult cr;
  if (kp.mode == 0) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.Cin), static_cast<cuuint64_t>(kp.M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.in_cstride) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.block_k), kBlockM};
    cuuint32_t estr[2] = {1, 1};
    cr = g_encode_tiled(&op->tmap_a, dt, 2, const_cast<void*>(d.in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.Cin), static_cast<cuuint64_t>(d.W),
                          static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.N)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.in_cstride) * 2,
                             static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W,
                             static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W * d.H};
    int lower[2] = {-d.pad, -d.pad};
    int upper[2] = {d.pad - (d.ksize - 1), d.pad - (d.ksize - 1)};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(d.stride), static_cast<cuuint32_t>(d.stride), 1};
    cr = g_encode_im2col(&op->tmap_a, dt, 4, const_cast<void*>(d.in), dims, strides, lower, upper,
                         static_cast<cuuint32_t>(kp.block_k), kBlockM, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // Driver workaround also applied by CUTLASS (copy_traits_sm90_im2col.hpp): for tensors smaller
    // than 128 KiB, drivers <= 13.1 set a descriptor bit that makes im2col loads fault.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t span = static_cast<size_t>(d.in_cstride) * 2 * d.W * d.H * d.N;
    if (cr == CUDA_SUCCESS && drv <= 13010 && span < 131072) {
      reinterpret_cast<uint64_t*>(&op->tmap_a)[1] &= ~(1ull << 21);
    }
  }
  if (cr != CUDA_SUCCESS) {
    set_error("conv: cuTensorMapEncode (A, mode %d) failed with CUresult %d (Cin=%d cs=%d H=%d W=%d N=%d k=%d s=%d bk=%d)",
              kp.mode, static_cast<int>(cr), d.Cin, d.in_cstride, d.H, d.W, d.N, d.ksize, d.stride,
              kp.block_k);
    delete op;
    return YB_ERR_CUDA;
  }
  {
    const int ktot = d.ksize * d.ksize * d.Cin_pad;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(d.Cout_pad)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.block_k), static_cast<cuuint32_t>(kp.block_n)};
    cuuint32_t estr[2] = {1, 1};
    cr = g_encode_tiled(&op->tmap_b, dt, 2, const_cast<void*>(d.weight), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeTiled (weights) failed with CUresult %d", static_cast<int>(cr));
      delete op;
      return YB_ERR_CUDA;
    }
  }
  {
    // destination view [M rows, Cout channels], row pitch = out_cstride; boxes of 128 rows x store_cols
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.Cout), static_cast<cuuint64_t>(kp.M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.out_cstride) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.store_cols), kBlockM};
    cuuint32_t estr[2] = {1, 1};
    cr = g_encode_tiled(&op->tmap_out, dt, 2, d.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_for_row_bytes(kp.store_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeTiled (output) failed with CUresult %d", static_cast<int>(cr));
      delete op;
      return YB_ERR_CUDA;
    }
  }
  op->tmap_w2 = op->tmap_b;      // placeholders when nothing is chained (never dereferenced)
  op->tmap_out2 = op->tmap_out;
  if (kp.ch.on) {
    const yb_conv_chain& c = *d.chain;
    const int kc = kp.ch.w2_row_bytes / 2;
    cuuint64_t wdims[2] = {static_cast<cuuint64_t>(c.K_pad), static_cast<cuuint64_t>(c.Cout_pad)};
    cuuint64_t wstrides[1] = {static_cast<cuuint64_t>(c.K_pad) * 2};
    cuuint32_t wbox[2] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(kp.ch.n2)};
    cuuint32_t estr[2] = {1, 1};
    cr = g_encode_tiled(&op->tmap_w2, dt, 2, const_cast<void*>(c.weight), wdims, wstrides, wbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_for_row_bytes(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr == CUDA_SUCCESS) {
      const int s2 = chain_store2_cols(kp.ch.n2);
      cuuint64_t odims[2] = {static_cast<cuuint64_t>(c.Cout), static_cast<cuuint64_t>(kp.M)};
      cuuint64_t ostrides[1] = {static_cast<cuuint64_t>(c.out_cstride) * 2};
      cuuint32_t obox[2] = {static_cast<cuuint32_t>(s2), kBlockM};
      cr = g_encode_tiled(&op->tmap_out2, dt, 2, c.out, odims, ostrides, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle_for_row_bytes(s2 * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cr != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeTiled (chained tail) failed with CUresult %d", static_cast<int>(cr));
      delete op;
      return YB_ERR_CUDA;
    }
  }
  op->fn = select_conv_kernel(kp);
  cudaError_t e = cudaFuncSetAttribute(op->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget));
  if (e != cudaSuccess) {
    set_error("conv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    delete op;
    return YB_ERR_CUDA;
  }
  *out = op;
  return YB_OK;
}

int conv_op_launch(const ConvOp* op, cudaStream_t stream) {
  if (op->patch) return patch_conv_launch(op->patch, stream);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = op->grid;
  cfg.blockDim = dim3(op->kp.epi_groups == kMaxGroups ? kThreadsWide : kThreads, 1, 1);
  cfg.dynamicSmemBytes = op->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  YB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, op->fn, op->tmap_a, op->tmap_b, op->tmap_out, op->tmap_w2, op->tmap_out2, op->kp));
  return YB_OK;
}

void conv_op_destroy(ConvOp* op) {
  if (op && op->patch) patch_conv_destroy(op->patch);
  delete op;
}

}  // namespace yb

#ifdef YB_EPI_TIMING
extern "C" int yb_debug_epi_ticks(unsigned long long* out16, int reset) {
  if (out16) {
    if (cudaMemcpyFromSymbol(out16, yb::g_epi_ticks, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  }
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(yb::g_epi_ticks, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif
