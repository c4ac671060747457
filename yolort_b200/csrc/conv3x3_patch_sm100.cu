// 3x3 / stride 1 / pad 1 convolution with the input halo patch staged ONCE per output tile.
//
// Why a second kernel: with im2col tensor maps every filter tap re-fetches its 128 pixels, i.e. the SM
// ingests 9x the tile's input (plus the weights) through the TMA path, and that path saturates at
// ~18 B/clk/SM (profiles/: 3x3 layers sat at Cin FLOP/B x ~5 TB/s).  Here an output tile is a 16 x 8 pixel
// block of one image; its (16+2) x (8+2) input halo is fetched by ONE tiled 4-D TMA load per 64-channel
// chunk (box 18 x 10 pixels, halo zero-filled by the TMA unit), and the nine filter taps are nine
// *views* of that patch: tap (dy,dx) starts (dy*10 + dx) pixel-rows into the patch, rows of one 8-pixel
// group are contiguous and consecutive groups are exactly one patch row (10 pixel-rows) apart, which is
// what a K-major UMMA shared-memory descriptor expresses (SBO = 10 * row_bytes; the swizzle phase follows the
// absolute shared-memory address, so a view may start inside a swizzle atom as long as base_offset stays 0).
// Input traffic per tile drops from 9 x 128 to 180 pixel-rows (6.4x less).
//
// Same arithmetic as conv_sm100.cu (yolort/v5/models/common.py:42-73,94-116): BN folded, bias + SiLU
// (+ residual) epilogue, fp32 accumulation in TMEM.
//
// Roles (one persistent CTA per SM): warp 0 = patch (A) producer, warp 1 = MMA issuer + TMEM owner,
// warp 2 = weight (B) producer (unless the weights are resident in shared memory), warps 3-10 = two epilogue groups
// (352 threads).  The kCpAsync variants insert four cooperative cp.async patch-loader warps before the epilogue
// groups (YB_PATCH_LOADER=1, an experiment measured equal to the TMA loads).  Like conv_sm100.cu the kernel is
// specialised per (dtype, store-box width, activation family).
#include <cstdlib>

#include "common.cuh"
#include "conv_epilogue.cuh"
#include "conv_sm100.h"

namespace yb {
namespace {

constexpr int kTileH = 16, kTileW = 8;           // output tile (pixels)
constexpr int kPatchH = 18, kPatchW = 10;        // TMA box (pixels): the tile plus its 1-pixel halo
constexpr int kMaxA = 4, kMaxB = 12;
constexpr int kEpiGroups = 2;
constexpr int kFirstLoadWarp = 3, kLoadWarps = 4;   // cooperative cp.async patch loaders
// The cp.async loader warps exist only in the kCpAsync kernel variants (YB_PATCH_LOADER=1, measured equal to TMA);
// the default variants start the epilogue warps right after the three producer / MMA warps: 352 threads instead
// of 480, which also lifts the per-thread register cap of __launch_bounds__ from 128 to 186.
constexpr int first_epi_warp(bool cp_async) { return cp_async ? kFirstLoadWarp + kLoadWarps : kFirstLoadWarp; }
constexpr int block_threads(bool cp_async) { return 32 * first_epi_warp(cp_async) + kEpiGroups * 128; }
constexpr int kStageBufBytes = 128 * 128;
constexpr int kMaxBlockN = 256;
constexpr size_t kSmemBudget = 222 * 1024;

struct PatchParams {
  int N, H, W;
  int tiles_x, tiles_y, m_tiles, n_tiles, num_tiles;
  int block_n, block_k, chunks;
  int a_stages, b_stages, b_resident;
  int band;               // 1: banded super-pixel weights (stem), see the kBand MMA loop
  int store_cols, store_bufs, bias_len;
  int dbg;                // ablation knobs (YB_CONV_DBG), see conv_sm100.cu
  int a_loader;           // 0: TMA box loads, 1: cooperative cp.async loads (4 warps)
  const void* in;         // NHWC input view (cp.async loader)
  int in_cstride, Cin;
  int view_mode;          // 0/1: one 18x16 patch, taps are shifted views (1 = also set the descriptor's base-offset
                          // field); 2: three 18x8 patches, one per dx (every view starts on a swizzle-atom boundary)
  uint32_t a_bytes, a_stride, b_sub_bytes, b_res_bytes, tmem_cols, idesc;   // a_stride: a_bytes rounded up to 1 KB
  const float* bias;
  EpilogueParams ep;
};

__device__ __forceinline__ void tma_load_tiled_4d(const void* desc, uint64_t* bar, void* smem_dst, int c, int w,
                                                  int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* desc, const void* smem_src, int c, int w, int h, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c), "r"(w), "r"(h), "r"(n)
               : "memory");
}

// Descriptor of a tap view: K-major, rows `row_bytes` apart inside an 8-row group, groups `sbo` bytes apart.
// The view starts inside a swizzle atom (dx pixel-rows in); the swizzle phase is a function of the absolute
// shared-memory address of each row, which is how the TMA unit laid the patch out.
__device__ __forceinline__ uint64_t make_view_desc(uint32_t addr, uint32_t row_bytes, uint32_t sbo, int mode) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  if (mode == 1) d |= static_cast<uint64_t>((addr >> 7) & 7u) << 49;  // matrix base offset field
  d |= layout << 61;
  return d;
}

template <bool kBf16, int kStoreCols, bool kRareAct, bool kCpAsync, bool kBand = false>
__global__ void __launch_bounds__(block_threads(kCpAsync), 1)
conv3x3_patch_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_out, const PatchParams p) {
  constexpr int kFirstEpiWarp = first_epi_warp(kCpAsync);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxA], a_empty[kMaxA];
  __shared__ __align__(8) uint64_t b_full[kMaxB], b_empty[kMaxB];
  __shared__ __align__(8) uint64_t acc_full[kEpiGroups], acc_empty[kEpiGroups];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[kEpiGroups][kMaxBlockN];

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_buf = base;                                                      // [a_stages][a_bytes]
  uint8_t* b_buf = a_buf + static_cast<size_t>(p.a_stages) * p.a_stride;       // resident [9*chunks] or ring [b_stages]
  const size_t b_region = p.b_resident ? p.b_res_bytes : static_cast<size_t>(p.b_stages) * p.b_sub_bytes;
  uint8_t* staging = b_buf + b_region;                                        // [kEpiGroups][store_bufs][16 KB]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps_total = 9 * p.chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], p.a_loader ? kLoadWarps * 32 : 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kMaxB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int g = 0; g < kEpiGroups; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&acc_empty[g], 4);
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
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== patch (A) producer, TMA variant =====================
    if (lane == 0 && p.a_loader == 0) {
      int ka = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int n_img = m_tile / tiles_per_img;
        const int t = m_tile - n_img * tiles_per_img;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        for (int c = 0; c < p.chunks; ++c, ++ka) {
          const int s = ka % p.a_stages;
          const uint32_t ph = (ka / p.a_stages) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          if (p.dbg & 8) {   // ablation: no loads at all, only the pipeline handshake
            mbar_arrive(&a_full[s]);
            continue;
          }
          mbar_expect_tx(&a_full[s], p.a_bytes);
          uint8_t* dst = a_buf + static_cast<size_t>(s) * p.a_stride;
          if (p.view_mode == 2) {
            for (int dx = 0; dx < 3; ++dx)
              tma_load_tiled_4d(&tmap_a, &a_full[s], dst + dx * (p.a_bytes / 3), c * p.block_k, tx * kTileW - 1 + dx,
                                ty * kTileH - 1, n_img);
          } else {
            tma_load_tiled_4d(&tmap_a, &a_full[s], dst, c * p.block_k, tx * kTileW - 1, ty * kTileH - 1, n_img);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weight (B) producer =====================
    if (lane == 0) {
      const uint32_t b_bytes = p.block_n * p.block_k * 2;
      if constexpr (kBand) {
        // banded stem weights: 3 filter rows x 2 blocks of 64 K-columns, consecutive in the weight matrix
        mbar_expect_tx(&b_full[0], 6 * b_bytes);
        for (int i = 0; i < 6; ++i)
          tma_load_2d(&tmap_b, &b_full[0], b_buf + static_cast<size_t>(i) * p.b_sub_bytes, i * p.block_k, 0);
      } else if (p.b_resident) {
        mbar_expect_tx(&b_full[0], taps_total * b_bytes);
        for (int i = 0; i < taps_total; ++i)   // i = chunk*9 + tap ; weight column block = tap*chunks + chunk
          tma_load_2d(&tmap_b, &b_full[0], b_buf + static_cast<size_t>(i) * p.b_sub_bytes,
                      ((i % 9) * p.chunks + i / 9) * p.block_k, 0);
      } else {
        int kb = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          const int n0 = (tile % p.n_tiles) * p.block_n;
          for (int i = 0; i < taps_total; ++i, ++kb) {
            const int s = kb % p.b_stages;
            const uint32_t ph = (kb / p.b_stages) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_expect_tx(&b_full[s], b_bytes);
            tma_load_2d(&tmap_b, &b_full[s], b_buf + static_cast<size_t>(s) * p.b_sub_bytes,
                        ((i % 9) * p.chunks + i / 9) * p.block_k, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t row_bytes = p.block_k * 2;
      const int pitch = p.view_mode == 2 ? kTileW : kPatchW;   // pixel-rows per patch row
      const uint32_t sbo = pitch * row_bytes;
      const int kk = p.block_k >> 4;
      if (p.b_resident) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      // per-tap start-address offsets of the A views (16-byte units) and the constant descriptor halves
      uint32_t tap_off16[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap - dy * 3;
        tap_off16[tap] = (p.view_mode == 2 ? dx * (p.a_bytes / 3) + dy * pitch * row_bytes : (dy * pitch + dx) * row_bytes) >> 4;
      }
      const uint32_t a_hi = static_cast<uint32_t>(make_view_desc(0, row_bytes, sbo, p.view_mode == 1 ? 0 : p.view_mode) >> 32);
      const uint32_t b_hi = static_cast<uint32_t>(make_kmajor_desc(0, row_bytes) >> 32);
      const uint32_t b_res_lo0 = (smem_u32(b_buf) & 0x3FFFFu) >> 4;
      const uint32_t b_step16 = p.b_sub_bytes >> 4;
      int ka = 0, kb = 0, lt = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;

        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * p.block_n;
        for (int c = 0; c < p.chunks; ++c, ++ka) {
          const int sa = ka % p.a_stages;
          mbar_wait(&a_full[sa], (ka / p.a_stages) & 1);
          tc_fence_after();
          const uint32_t patch = smem_u32(a_buf + static_cast<size_t>(sa) * p.a_stride);
          const uint32_t a_lo0 = (patch & 0x3FFFFu) >> 4;
          if constexpr (kBand) {
            // Super-pixel stem (engine.stem_superpixel, pack 4, 16 channels per pixel, one 64-channel chunk): the
            // expanded weight matrix is block-banded -- a group of 4 output pixels reads, per filter row, exactly
            // the 6 input pixels x 16 channels that sit in 192 CONTIGUOUS bytes of the patch, starting 96 bytes into
            // the left neighbour super-pixel.  Six K=16 steps per filter row walk that span (every start is a
            // (dx, k-step) address the plain 9-tap loop also uses: (0,3), (1,0..3), (2,0)); the weights hold only
            // those K-slices: [ky][2 blocks of 64], the last 32 columns of the second block are zero padding that
            // is never multiplied.  18 MMAs per tile instead of 36, 96 KB of resident weights instead of a
            // 147 KB ring that is re-streamed from L2 for every tile.
            if (!(p.dbg & 2)) {
              const uint32_t row16 = (static_cast<uint32_t>(pitch) * row_bytes) >> 4;
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                  const uint32_t a_lo = a_lo0 + ky * row16 + 6 + 2 * j;
                  const uint32_t b_lo = b_res_lo0 + static_cast<uint32_t>(ky * 2 + (j >> 2)) * b_step16 + 2 * (j & 3);
                  if (ky == 0 && j == 0)
                    umma_f16_lohi<false>(tmem_d, a_lo, a_hi, b_lo, b_hi, p.idesc);
                  else
                    umma_f16_lohi<true>(tmem_d, a_lo, a_hi, b_lo, b_hi, p.idesc);
                }
              }
            }
          } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t b_lo;
            int sb = 0;
            if (p.b_resident) {
              b_lo = b_res_lo0 + static_cast<uint32_t>(c * 9 + tap) * b_step16;
            } else {
              sb = kb % p.b_stages;
              mbar_wait(&b_full[sb], (kb / p.b_stages) & 1);
              tc_fence_after();
              b_lo = b_res_lo0 + static_cast<uint32_t>(sb) * b_step16;
            }
            if (!(p.dbg & 2))
              umma_ksteps_rt(kk, tmem_d, a_lo0 + tap_off16[tap], a_hi, b_lo, b_hi, p.idesc, (c | tap) == 0);
            if (!p.b_resident) {
              umma_commit(&b_empty[sb]);
              ++kb;
            }
          }
          }
          umma_commit(&a_empty[sa]);
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else if (kCpAsync && warp >= kFirstLoadWarp && warp < kFirstEpiWarp) {
    // ===================== patch (A) producer, cooperative cp.async variant =====================
    // The TMA unit handles a box row by row (measured ~19 clk per 128-byte row for these 4-D boxes, ~15 clk
    // even for 32-byte rows); 128 threads issuing 16-byte cp.async copies move the same patch several times
    // faster.  Each thread writes its chunks to the swizzled position the UMMA descriptor expects and
    // zero-fills the halo (src-size 0).  Completion: wait_group -> proxy fence -> mbarrier arrive.
    if constexpr (kCpAsync) {
      const int ltid = threadIdx.x - 32 * kFirstLoadWarp;
      const int row_bytes = p.block_k * 2;
      const int cpr = row_bytes >> 4;                    // 16-byte chunks per pixel-row
      const int total = kPatchH * kPatchW * cpr;
      // The copy pattern of a patch is the same for every tile: each thread precomputes, once, the shared-memory
      // offset (swizzled), the global element offset relative to the patch origin and the (row, column) of its
      // <= kMaxOps chunks; per tile only the origin pointer and two validity bit-masks change.
      constexpr int kMaxOps = (kPatchH * kPatchW * 8 + kLoadWarps * 32 - 1) / (kLoadWarps * 32);
      uint32_t soff[kMaxOps];
      int goff[kMaxOps];
      uint32_t hw[kMaxOps];
      int nops = 0;
#pragma unroll
      for (int k = 0; k < kMaxOps; ++k) {
        const int i = ltid + k * kLoadWarps * 32;
        soff[k] = 0; goff[k] = 0; hw[k] = 0;
        if (i < total) {
          const int row = i / cpr, ch = i - row * cpr;
          const int hh = row / kPatchW, ww = row - hh * kPatchW;
          soff[k] = row * row_bytes + swizzle_chunk(row, ch, row_bytes) * 16;
          goff[k] = (hh * p.W + ww) * p.in_cstride + ch * 8;
          hw[k] = (static_cast<uint32_t>(hh) << 8) | static_cast<uint32_t>(ww) | (static_cast<uint32_t>(ch) << 16);
          nops = k + 1;
        }
      }
      const uint16_t* in = reinterpret_cast<const uint16_t*>(p.in);
      int ka = 0;
      int pending_stage = -1;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int n_img = m_tile / tiles_per_img;
        const int t = m_tile - n_img * tiles_per_img;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        const int y_base = ty * kTileH - 1, x_base = tx * kTileW - 1;
        // bit hh of rowmask: input row y_base+hh exists; bit ww of colmask: input column x_base+ww exists
        uint32_t rowmask = 0, colmask = 0;
        for (int hh = 0; hh < kPatchH; ++hh) rowmask |= (static_cast<unsigned>(y_base + hh) < static_cast<unsigned>(p.H) ? 1u : 0u) << hh;
        for (int ww = 0; ww < kPatchW; ++ww) colmask |= (static_cast<unsigned>(x_base + ww) < static_cast<unsigned>(p.W) ? 1u : 0u) << ww;
        const long long origin = ((static_cast<long long>(n_img) * p.H + y_base) * p.W + x_base) * p.in_cstride;
        for (int c = 0; c < p.chunks; ++c, ++ka) {
          const int s = ka % p.a_stages;
          const uint32_t ph = (ka / p.a_stages) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          const uint32_t dst_base = smem_u32(a_buf + static_cast<size_t>(s) * p.a_stride);
          const uint16_t* src0 = in + origin + c * p.block_k;
          const int cmax = p.Cin - c * p.block_k;          // channels of this chunk that exist (zero-fill the rest)
#pragma unroll
          for (int k = 0; k < kMaxOps; ++k) {
            if (k < nops) {
              const uint32_t hh = (hw[k] >> 8) & 0xff, ww = hw[k] & 0xff, ch = hw[k] >> 16;
              const bool ok = ((rowmask >> hh) & (colmask >> ww) & 1u) && static_cast<int>(ch * 8) < cmax;
              const uint16_t* src = ok ? src0 + goff[k] : in;
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_base + soff[k]), "l"(src), "r"(ok ? 16 : 0) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (pending_stage >= 0) {   // publish the previous patch while this one is in flight
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            fence_proxy_async_smem();
            mbar_arrive(&a_full[pending_stage]);
          }
          pending_stage = s;
        }
      }
      if (pending_stage >= 0) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async_smem();
        mbar_arrive(&a_full[pending_stage]);
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue groups =====================
    const int g = (warp - kFirstEpiWarp) >> 2;
    const int q = warp & 3;
    const int gtid = threadIdx.x - 32 * kFirstEpiWarp - g * 128;
    const int row_in_tile = q * 32 + lane;     // j = r*8 + x  (r: output row in tile, x: column in tile)
    const bool issuer = (gtid == 0);
    const uint32_t bar_id = 1 + g;
    const int store_cols = kStoreCols != 0 ? kStoreCols : p.store_cols;
    const int row_bytes = store_cols * 2;
    uint8_t* my_staging = staging + static_cast<size_t>(g) * p.store_bufs * kStageBufBytes;
    float* bias_s = s_bias[g];
    int lt = 0, store_idx = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      if ((lt & 1) != g) continue;
      const uint32_t aph = (lt >> 1) & 1;
      const int m_tile = tile / p.n_tiles;
      const int n0 = (tile - m_tile * p.n_tiles) * p.block_n;
      const int n_img = m_tile / tiles_per_img;
      const int t = m_tile - n_img * tiles_per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      const int y = ty * kTileH + (row_in_tile >> 3), x = tx * kTileW + (row_in_tile & 7);
      const bool row_ok = y < p.H && x < p.W;
      const long long row = (static_cast<long long>(n_img) * p.H + y) * p.W + x;
      if (p.dbg & 16) {   // ablation: accumulator handshake only
        mbar_wait(&acc_full[g], aph);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[g]);
        continue;
      }
      for (int i = gtid; i < p.block_n; i += 128) bias_s[i] = (n0 + i < p.bias_len) ? __ldg(p.bias + n0 + i) : 0.f;
      named_bar_sync(bar_id, 128);
      mbar_wait(&acc_full[g], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * p.block_n;
      for (int c0 = 0; c0 < p.block_n; c0 += store_cols, ++store_idx) {
        // Two staging buffers, one barrier per box: before the barrier below the issuer waits until the PREVIOUS
        // store has finished reading its buffer, which is the one the next box will overwrite.
        uint8_t* buf = my_staging + (store_idx & 1) * kStageBufBytes;
        uint8_t* my_row = buf + row_in_tile * row_bytes;
        if (!(p.dbg & 1)) {
          epilogue_box_select<kBf16, kStoreCols, kRareAct>(p.ep, store_cols, taddr + c0, bias_s + c0, row, row_ok, n0 + c0, my_row, row_in_tile);
        }
        if (c0 + store_cols >= p.block_n) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[g]);
        }
        fence_proxy_async_smem();
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
        if (issuer) {
          if (n0 + c0 < p.ep.Cout && !(p.dbg & 5)) tma_store_4d(&tmap_out, buf, n0 + c0, tx * kTileW, ty * kTileH, n_img);
          tma_store_commit();
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

}  // namespace

using PatchKernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const PatchParams);

template <bool kBf16>
PatchKernelFn select_patch_kernel_t(const PatchParams& kp) {
  if (kp.band) return conv3x3_patch_kernel<kBf16, 64, false, false, true>;    // banded stem (opt-in, YB_STEM_BAND=1)
  if (kp.a_loader == 1) return conv3x3_patch_kernel<kBf16, 0, false, true>;   // experiment knob: one generic variant
  if (kp.ep.act >= YB_ACT_HARDSWISH) return conv3x3_patch_kernel<kBf16, 0, true, false>;
  switch (kp.store_cols) {
    case 64: return conv3x3_patch_kernel<kBf16, 64, false, false>;
    case 32: return conv3x3_patch_kernel<kBf16, 32, false, false>;
    default: return conv3x3_patch_kernel<kBf16, 16, false, false>;
  }
}
PatchKernelFn select_patch_kernel(const PatchParams& kp) {
  return kp.ep.is_bf16 ? select_patch_kernel_t<true>(kp) : select_patch_kernel_t<false>(kp);
}

struct PatchConvOp {
  CUtensorMap tmap_a, tmap_b, tmap_out;
  PatchParams kp;
  PatchKernelFn fn = nullptr;
  dim3 grid;
  size_t smem_bytes;
};

// Eligibility: 3x3 / stride 1 / pad 1 and a feature map that 16x8 tiles cover with little waste.
bool patch_conv_eligible(const yb_op_desc& d) {
  if (d.kind != YB_OP_CONV || d.ksize != 3 || d.stride != 1 || d.pad != 1) return false;
  if (d.reserved & 1) return false;   // caller asked for the generic im2col kernel
  const char* env = getenv("YB_DISABLE_PATCH_CONV");
  if (env && env[0] == '1') return false;
  const int ty = (d.H + kTileH - 1) / kTileH, tx = (d.W + kTileW - 1) / kTileW;
  const double eff = static_cast<double>(d.H) * d.W / (static_cast<double>(ty) * kTileH * tx * kTileW);
  return eff >= 0.7;
}

int patch_conv_create(const yb_op_desc& d, EncodeTiledFn encode_tiled, PatchConvOp** out) {
  PatchConvOp* op = new PatchConvOp();
  PatchParams& kp = op->kp;
  kp.N = d.N;
  kp.H = d.H;
  kp.W = d.W;
  kp.tiles_x = (d.W + kTileW - 1) / kTileW;
  kp.tiles_y = (d.H + kTileH - 1) / kTileH;
  kp.m_tiles = d.N * kp.tiles_x * kp.tiles_y;
  const int sms = num_sms();
  int n_tiles = (d.Cout + kMaxBlockN - 1) / kMaxBlockN;
  int block_n = (((d.Cout + n_tiles - 1) / n_tiles) + 15) / 16 * 16;
  if (kp.m_tiles * n_tiles < 2 * sms && block_n > 128 && block_n % 32 == 0) {
    block_n /= 2;
    n_tiles = (d.Cout + block_n - 1) / block_n;
  }
  kp.block_n = block_n;
  kp.n_tiles = n_tiles;
  kp.num_tiles = kp.m_tiles * n_tiles;
  kp.block_k = (d.Cin_pad % 64 == 0) ? 64 : ((d.Cin_pad % 32 == 0) ? 32 : 16);
  kp.chunks = d.Cin_pad / kp.block_k;
  const char* env_mode = getenv("YB_PATCH_MODE");
  kp.view_mode = env_mode ? atoi(env_mode) : 0;
  if (kp.view_mode < 0 || kp.view_mode > 2) kp.view_mode = 0;
  kp.a_bytes = (kp.view_mode == 2 ? 3 * kPatchH * kTileW : kPatchH * kPatchW) * kp.block_k * 2;
  kp.a_stride = (kp.a_bytes + 1023u) & ~1023u;
  kp.b_sub_bytes = (static_cast<uint32_t>(block_n * kp.block_k * 2) + 1023u) & ~1023u;
  kp.store_cols = (block_n % 64 == 0) ? 64 : ((block_n % 32 == 0) ? 32 : 16);
  kp.store_bufs = 2;
  kp.bias_len = d.Cout_pad;
  {
    const char* e = getenv("YB_CONV_DBG");
    kp.dbg = e ? atoi(e) : 0;
  }
  const size_t staging = static_cast<size_t>(kEpiGroups) * kp.store_bufs * kStageBufBytes;
  // reserved bit 1: the weights are the banded super-pixel stem matrix [Cout_pad][3 rows][2 x 64] (engine.stem_band)
  kp.band = (d.reserved & 2) ? 1 : 0;
  if (kp.band && !(d.Cin_pad == 64 && n_tiles == 1 && kp.store_cols == 64 && kp.view_mode == 0 && d.act < YB_ACT_HARDSWISH &&
                   d.residual == nullptr)) {
    set_error("patch conv: banded stem weights need Cin_pad 64, one N tile with 64-column store boxes, SiLU/linear epilogue");
    delete op;
    return YB_ERR_INVALID;
  }
  const size_t b_total = static_cast<size_t>(kp.band ? 6 : 9 * kp.chunks) * kp.b_sub_bytes;
  const size_t avail = kSmemBudget - staging - 1024;
  kp.b_resident = (n_tiles == 1 && b_total + 2 * kp.a_stride <= avail) ? 1 : 0;
  if (kp.band && !kp.b_resident) {
    set_error("patch conv: banded stem weights do not fit in shared memory (block_n=%d)", block_n);
    delete op;
    return YB_ERR_INVALID;
  }
  kp.b_res_bytes = kp.b_resident ? static_cast<uint32_t>(b_total) : 0u;
  if (kp.b_resident) {
    int a_st = static_cast<int>((avail - b_total) / kp.a_stride);
    kp.a_stages = a_st > kMaxA ? kMaxA : a_st;
    kp.b_stages = 1;
  } else {
    kp.a_stages = 2;
    size_t rem = avail - 2 * kp.a_stride;
    if (rem >= static_cast<size_t>(kp.a_stride) + 6 * kp.b_sub_bytes) {
      kp.a_stages = 3;
      rem -= kp.a_stride;
    }
    int b_st = static_cast<int>(rem / kp.b_sub_bytes);
    kp.b_stages = b_st > kMaxB ? kMaxB : b_st;
    if (kp.b_stages < 2) {
      set_error("patch conv: not enough shared memory for the weight ring (block_n=%d)", block_n);
      delete op;
      return YB_ERR_INVALID;
    }
  }
  uint32_t cols = 32;
  while (static_cast<int>(cols) < 2 * block_n) cols <<= 1;
  kp.tmem_cols = cols;
  kp.ep.Cout = d.Cout;
  kp.ep.act = d.act;
  kp.ep.is_bf16 = d.dtype == YB_BF16;
  kp.ep.residual = d.residual;
  kp.ep.res_cstride = d.res_cstride;
  kp.bias = d.bias;
  kp.in = d.in;
  kp.in_cstride = d.in_cstride;
  kp.Cin = d.Cin;
  {
    const char* e = getenv("YB_PATCH_LOADER");
    kp.a_loader = e ? atoi(e) : 0;
    if (kp.view_mode == 2 || kp.band) kp.a_loader = 0;   // the dx-split layout / banded weights exist only for the TMA variant
    if (d.act >= YB_ACT_HARDSWISH) kp.a_loader = 0;   // ... and so do the r3.1 activation variants
  }
  const uint32_t fmt = kp.ep.is_bf16 ? 1u : 0u;
  kp.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(block_n >> 3) << 17) | (8u << 24);
  op->grid = dim3(kp.num_tiles < sms ? kp.num_tiles : sms, 1, 1);
  const size_t b_region = kp.b_resident ? kp.b_res_bytes : static_cast<size_t>(kp.b_stages) * kp.b_sub_bytes;
  size_t smem = static_cast<size_t>(kp.a_stages) * kp.a_stride + b_region + staging + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;
  op->smem_bytes = smem;

  const CUtensorMapDataType dt = kp.ep.is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const int rb = kp.block_k * 2;
  const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult cr;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.Cin), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H),
                          static_cast<cuuint64_t>(d.N)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.in_cstride) * 2, static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W,
                             static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W * d.H};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(kp.block_k), static_cast<cuuint32_t>(kp.view_mode == 2 ? kTileW : kPatchW),
                         kPatchH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    cr = encode_tiled(&op->tmap_a, dt, 4, const_cast<void*>(d.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("patch conv: cuTensorMapEncodeTiled (input patch) failed with CUresult %d", static_cast<int>(cr));
      delete op;
      return YB_ERR_CUDA;
    }
  }
  {
    const int ktot = kp.band ? 6 * 64 : 9 * d.Cin_pad;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(d.Cout_pad)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.block_k), static_cast<cuuint32_t>(block_n)};
    cuuint32_t estr[2] = {1, 1};
    cr = encode_tiled(&op->tmap_b, dt, 2, const_cast<void*>(d.weight), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("patch conv: cuTensorMapEncodeTiled (weights) failed with CUresult %d", static_cast<int>(cr));
      delete op;
      return YB_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.Cout), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H),
                          static_cast<cuuint64_t>(d.N)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.out_cstride) * 2, static_cast<cuuint64_t>(d.out_cstride) * 2 * d.W,
                             static_cast<cuuint64_t>(d.out_cstride) * 2 * d.W * d.H};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(kp.store_cols), kTileW, kTileH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const int srb = kp.store_cols * 2;
    cr = encode_tiled(&op->tmap_out, dt, 4, d.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      srb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (srb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("patch conv: cuTensorMapEncodeTiled (output) failed with CUresult %d", static_cast<int>(cr));
      delete op;
      return YB_ERR_CUDA;
    }
  }
  op->fn = select_patch_kernel(kp);
  cudaError_t e = cudaFuncSetAttribute(op->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget));
  if (e != cudaSuccess) {
    set_error("patch conv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    delete op;
    return YB_ERR_CUDA;
  }
  *out = op;
  return YB_OK;
}

int patch_conv_launch(const PatchConvOp* op, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = op->grid;
  cfg.blockDim = dim3(block_threads(op->kp.a_loader == 1), 1, 1);
  cfg.dynamicSmemBytes = op->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  YB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, op->fn, op->tmap_a, op->tmap_b, op->tmap_out, op->kp));
  return YB_OK;
}

void patch_conv_destroy(PatchConvOp* op) { delete op; }

}  // namespace yb
