// Dense multi-term contraction on TMA + tcgen05 (sm_100a):  out[R, c] = epi( sum_t A_t[R, 0:F_t] . B_t[0:F_t, c] )
//
// This is cape_cheb_fwd for calls whose terms are all PLAIN tensors (identity operators): 1x1 convs, the two halves
// of the split convolution forms the host uses -- "contract first" (Z = X . [W_0 | W_1 | ...], the operators are then
// applied to the narrower Z by cape_apply) and "basis first" (cape_apply writes the Chebyshev basis / the narrow-side
// tensors op_k^T G, which are then contracted here) -- and the GroupNorm blocks' linear layers.  Nothing is gathered
// in this kernel, so the operand pipeline is pure TMA:
//
//   * warp 13 issues the A boxes (128 rows x 32 k of term t, SWIZZLE_128B: a box IS a K-major UMMA operand tile; the
//     raw fp32 words are the "hi" operand because kind::tf32 reads their top 19 bits), ring of `sr` stages;
//   * warp 14 issues the weight boxes (raw K-major copy, BN columns x 32 k), ring of `sbr` stages -- only the raw words
//     travel: with the pre-split low parts a CTA would stream 8 bytes per weight element and 128-row tile, 43 B/clk per
//     SM at the tensor-core rate, which is the whole L2 -> SM bandwidth of the chip (~42 B/clk/SM);
//   * warps 0-7 derive the lo tiles of A and of the weights (x - trunc_tf32(x)) from shared memory, rings of `sl` /
//     `sbl` stages;
//   * warp 8 issues tcgen05.mma kind::tf32 (3xTF32: hi*hi + lo*hi + hi*lo) into TMEM;
//   * warps 9-12 drain TMEM: condition broadcast / bias / activation / backward masks, float4 stores.
//
// Output columns are processed in GROUPS of `gw` columns (<= 512 TMEM columns per group, `nsub` MMA sub-tiles of BN
// columns); a persistent CTA walks (row tile, column group) pairs, so any ncols % 16 == 0 works (544-wide GroupNorm
// blocks included) and a group's accumulators can be double-buffered against the previous group's epilogue.
//
// PRECISE mode (cape_conv_args.precise): the tensor core's fp32 accumulator TRUNCATES on every accumulate, which
// shrinks a dot product by ~2^-25 per MMA on its chain -- 8.6e-6 relative for a 1024-long reduction at three MMAs per
// k-step (measured, tests/gpu_tf32_accuracy.py), 2e-5 after the eight encoder layers, and the VAE's exp(logvar)
// amplifies exactly that.  Precise groups are 128 columns wide and keep FOUR accumulators: the hi*hi products go
// round-robin to three of them (chains a third as long), the small lo*hi + hi*lo corrections to the fourth (their
// truncation error is 2^-11 smaller); the epilogue adds the four in fp32 with round-to-nearest.  Same MMA count,
// more A traffic (one pass over the A tiles per 128 columns); used for the encoder's forward convs.
#include "common.cuh"
#include "ellconv_params.cuh"
#include "tc_common.cuh"

namespace cape {

namespace {

using namespace tc;

constexpr int G_CONV_WARPS = 8;
constexpr int G_CONV_THREADS = G_CONV_WARPS * 32;
constexpr int G_MMA_WARP = 8, G_EPI_WARP0 = 9, G_TMA_A_WARP = 13, G_TMA_B_WARP = 14;
constexpr int G_EPI_WARPS = 4;
constexpr int G_THREADS = 15 * 32;
constexpr int G_A_TILE = BM * 128;          // 128 rows x 32 fp32
constexpr int G_MAX_STAGES = 8;
constexpr int G_TERMS = 4;
constexpr int G_QS_FLOATS = 2048;           // condition vectors of the group's columns, double-buffered halves
constexpr int G_SMEM_LIMIT = 227 * 1024;

struct GMaps {
  CUtensorMap a[G_TERMS];      // source rows of term t: boxes of 32 f x 128 rows
  CUtensorMap bh[G_TERMS];     // K-major weight copy (raw fp32 = hi operand): boxes of 32 f x BN columns
  CUtensorMap bl[G_TERMS];     // its tf32 low part (cape_term.wT_lo), when every term has one
};

struct GPlan {
  int gw, nsub, ngroups, ntiles;       // group width (columns), BN-wide sub-tiles per group, groups per row tile, row tiles
  int nchain, corr, nbuf, tmem_cols;   // main accumulators per group, 1 = separate correction accumulator, TMEM buffers
  int sr, sl, sbr, sbl;                // ring depths: A raw (TMA), A lo (converters), B raw (TMA), B lo (converters)
  int rotate;                          // 1: every row tile starts its reduction at a different chunk (see chunk_of)
  int blo_tma;                         // 1: the weight lo tiles come by TMA from the pre-split copy (one B ring pair, depth
                                       // sbr, one full/empty barrier pair); 0: the converters derive them from the raw tiles
};

__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {     // K-major SWIZZLE_128B operand tile
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

template <int BN, bool PRECISE>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_tc_kernel(const __grid_constant__ ConvParams p,
                                                               const __grid_constant__ GMaps maps,
                                                               const __grid_constant__ GPlan g) {
  constexpr int B_TILE = BN * 128;             // raw (= hi) or lo tile of BN weight columns x 32 k
  extern __shared__ uint8_t smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  char* raw_ring = smem;
  char* lo_ring = raw_ring + g.sr * G_A_TILE;
  char* braw_ring = lo_ring + g.sl * G_A_TILE;
  char* blo_ring = braw_ring + g.sbr * B_TILE;
  float* qs_all = reinterpret_cast<float*>(blo_ring + g.sbl * B_TILE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(qs_all + G_QS_FLOATS);
  // raw_full[8] raw_empty[8] lo_full[8] lo_empty[8] braw_full[8] braw_empty[8] blo_full[8] blo_empty[8] t_full[2] t_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 * G_MAX_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_rf = smem_u32(bars), bar_re = smem_u32(bars + G_MAX_STAGES);
  const uint32_t bar_lf = smem_u32(bars + 2 * G_MAX_STAGES), bar_le = smem_u32(bars + 3 * G_MAX_STAGES);
  const uint32_t bar_bf = smem_u32(bars + 4 * G_MAX_STAGES), bar_be = smem_u32(bars + 5 * G_MAX_STAGES);
  const uint32_t bar_blf = smem_u32(bars + 6 * G_MAX_STAGES), bar_ble = smem_u32(bars + 7 * G_MAX_STAGES);
  const uint32_t bar_tf = smem_u32(bars + 8 * G_MAX_STAGES), bar_te = smem_u32(bars + 8 * G_MAX_STAGES + 2);

  if (warp == G_MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < G_MAX_STAGES; ++s) {
        mbar_init(bar_rf + 8 * s, 1); mbar_init(bar_re + 8 * s, 1);
        mbar_init(bar_lf + 8 * s, G_CONV_WARPS); mbar_init(bar_le + 8 * s, 1);
        mbar_init(bar_bf + 8 * s, 1); mbar_init(bar_be + 8 * s, 1);
        mbar_init(bar_blf + 8 * s, G_CONV_WARPS); mbar_init(bar_ble + 8 * s, 1);
      }
      for (int s = 0; s < 2; ++s) { mbar_init(bar_tf + 8 * s, 1); mbar_init(bar_te + 8 * s, G_EPI_WARPS); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)g.tmem_cols);
  }
  if (warp == G_TMA_A_WARP && lane == 0)
    for (int t = 0; t < p.nterms; ++t) tma_prefetch_desc(&maps.a[t]);
  if (warp == G_TMA_B_WARP && lane == 0)
    for (int t = 0; t < p.nterms; ++t) tma_prefetch_desc(&maps.bh[t]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nacc = g.nchain + g.corr;
  const int nwork = g.ntiles * g.ngroups;
  // sub-tiles of work item w's column group (the last group of a layer may be narrower than gw)
  auto nsub_of = [&](int w) { return (min(g.gw, p.ncols - (w % g.ngroups) * g.gw) + BN - 1) / BN; };
  // The reduction of a work item is a list of 32-deep chunks over all terms.  All CTAs walk the same weight tiles, so
  // with a common order the whole chip asks the same few L2 lines at the same moment; row tile i therefore starts at
  // chunk (5 i mod nchunks) and wraps around.  chunk_of maps the j-th chunk of work item w to (term, f0).
  int nchunks = 0;
  for (int t = 0; t < p.nterms; ++t) nchunks += (p.terms[t].F + BK - 1) / BK;
  auto chunk_of = [&](int w, int j, int& t, int& f0) {
    int c = j + (g.rotate ? ((w / g.ngroups) * 5) % nchunks : 0);
    if (c >= nchunks) c -= nchunks;
    for (t = 0;; ++t) {
      const int nt = (p.terms[t].F + BK - 1) / BK;
      if (c < nt) break;
      c -= nt;
    }
    f0 = c * BK;
  };

  if (warp < G_CONV_WARPS) {
    // =========================== converters: lo tiles of every A chunk and every weight sub-tile ===========================
    const int l8 = tid & 7, rs = tid >> 3;
    int sr = 0, sl = 0, sbr = 0, sbl = 0;
    uint32_t phr = 0, phl = 0, phb = 0, phbl = 0;
    auto lo4 = [](float4 v) {
      float4 l;
      l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
      l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
      l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
      l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
      return l;
    };
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      const int nsub = nsub_of(w);
      for (int j = 0; j < nchunks; ++j) {
        {
          mbar_wait(bar_rf + 8 * sr, (phr >> sr) & 1u);
          mbar_wait(bar_le + 8 * sl, ((phl >> sl) & 1u) ^ 1u);
          {
            const char* hi = raw_ring + (size_t)sr * G_A_TILE;
            char* lo = lo_ring + (size_t)sl * G_A_TILE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = rs + 32 * i;
              const uint32_t off = (uint32_t)(row * 128 + ((l8 ^ (row & 7)) << 4));
              *reinterpret_cast<float4*>(lo + off) = lo4(*reinterpret_cast<const float4*>(hi + off));
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_lf + 8 * sl);
          phr ^= 1u << sr; phl ^= 1u << sl;
          if (++sr == g.sr) sr = 0;
          if (++sl == g.sl) sl = 0;
          for (int s = 0; s < (g.blo_tma ? 0 : nsub); ++s) {
            mbar_wait(bar_bf + 8 * sbr, (phb >> sbr) & 1u);
            mbar_wait(bar_ble + 8 * sbl, ((phbl >> sbl) & 1u) ^ 1u);
            const char* hi = braw_ring + (size_t)sbr * B_TILE;
            char* lo = blo_ring + (size_t)sbl * B_TILE;
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) {
              const uint32_t off = (uint32_t)((i * G_CONV_THREADS + tid) * 16);   // the lo tile mirrors the hi tile byte for byte
              *reinterpret_cast<float4*>(lo + off) = lo4(*reinterpret_cast<const float4*>(hi + off));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_blf + 8 * sbl);
            phb ^= 1u << sbr; phbl ^= 1u << sbl;
            if (++sbr == g.sbr) sbr = 0;
            if (++sbl == g.sbl) sbl = 0;
          }
        }
      }
    }
  } else if (warp == G_TMA_A_WARP) {
    // =========================== TMA: A tiles ===========================
    {
      int sr = 0;
      uint32_t phr = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int row0 = (w / g.ngroups) * BM;
        for (int j = 0; j < nchunks; ++j) {
          {
            int t, f0;
            chunk_of(w, j, t, f0);
            mbar_wait(bar_re + 8 * sr, ((phr >> sr) & 1u) ^ 1u);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar_rf + 8 * sr, (uint32_t)G_A_TILE);
              tma_load_2d(smem_u32(raw_ring + (size_t)sr * G_A_TILE), &maps.a[t], f0, row0, bar_rf + 8 * sr);
            }
            __syncwarp();
            phr ^= 1u << sr;
            if (++sr == g.sr) sr = 0;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == G_TMA_B_WARP) {
    // =========================== TMA: raw weight tiles ===========================
    {
      int sb = 0;
      uint32_t phb = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int col0 = (w % g.ngroups) * g.gw;
        const int nsub = nsub_of(w);
        for (int j = 0; j < nchunks; ++j) {
          {
            int t, f0;
            chunk_of(w, j, t, f0);
            for (int s = 0; s < nsub; ++s) {
              mbar_wait(bar_be + 8 * sb, ((phb >> sb) & 1u) ^ 1u);
              if (elect_one()) {
                mbar_arrive_expect_tx(bar_bf + 8 * sb, (uint32_t)(g.blo_tma ? 2 * B_TILE : B_TILE));
                tma_load_2d(smem_u32(braw_ring + (size_t)sb * B_TILE), &maps.bh[t], f0, col0 + s * BN, bar_bf + 8 * sb);
                if (g.blo_tma)
                  tma_load_2d(smem_u32(blo_ring + (size_t)sb * B_TILE), &maps.bl[t], f0, col0 + s * BN, bar_bf + 8 * sb);
              }
              __syncwarp();
              phb ^= 1u << sb;
              if (++sb == g.sbr) sb = 0;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == G_MMA_WARP) {
    // =========================== MMA issuer (whole warp, one elected lane issues) ===========================
    {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128; N (a multiple of 16) is set per sub-tile
      constexpr uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BM >> 4) << 24);
      int sr = 0, sl = 0, sbr = 0, sbl = 0, it = 0;
      uint32_t phr = 0, phl = 0, phb = 0, phbl = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
        const int buf = it % g.nbuf;
        const uint32_t use = (uint32_t)(it / g.nbuf);
        mbar_wait(bar_te + 8 * buf, (use & 1u) ^ 1u);            // the epilogue has drained this buffer
        tc_fence_after();
        const uint32_t tb = tmem_base + (uint32_t)(buf * g.gw * nacc);
        const int gcols = min(g.gw, p.ncols - (w % g.ngroups) * g.gw);
        const int nsub = (gcols + BN - 1) / BN;
        // PRECISE: the hi*hi products of k-step q go to chain q mod 3, the corrections to the fourth accumulator.  This one
        // thread feeds the tensor core, so the loop below is kept free of divisions and data-dependent branches: the chain
        // of a chunk's first k-step is carried along (4 k-steps per chunk: it advances by one), and "first write
        // overwrites" only concerns chunk 0 (k-steps 0..2 start the three chains, k-step 0 the correction accumulator).
        uint32_t c0 = 0;
        for (int j = 0; j < nchunks; ++j) {
          {
            const uint32_t later = j > 0 ? 1u : 0u;
            mbar_wait(bar_rf + 8 * sr, (phr >> sr) & 1u);         // TMA bytes of the raw tile
            mbar_wait(bar_lf + 8 * sl, (phl >> sl) & 1u);         // its lo tile
            tc_fence_after();
            const uint64_t a_hi = make_desc_k(smem_u32(raw_ring + (size_t)sr * G_A_TILE));
            const uint64_t a_lo = make_desc_k(smem_u32(lo_ring + (size_t)sl * G_A_TILE));
            const uint32_t c1 = c0 == 2 ? 0u : c0 + 1, c2 = c1 == 2 ? 0u : c1 + 1;
            const uint32_t chain_off[4] = {c0 * (uint32_t)g.gw, c1 * (uint32_t)g.gw, c2 * (uint32_t)g.gw, c0 * (uint32_t)g.gw};
            for (int s = 0; s < nsub; ++s) {
              // the last sub-tile of a group may be narrower: N = its real columns (ncols % 16 == 0)
              const uint32_t idesc = idesc0 | ((uint32_t)(min(BN, gcols - s * BN) >> 3) << 17);
              mbar_wait(bar_bf + 8 * sbr, (phb >> sbr) & 1u);      // TMA bytes of the raw weight tile (and of its lo tile)
              if (!g.blo_tma) mbar_wait(bar_blf + 8 * sbl, (phbl >> sbl) & 1u);    // lo tile from the converters
              tc_fence_after();
              const uint64_t b_hi = make_desc_k(smem_u32(braw_ring + (size_t)sbr * B_TILE));
              const uint64_t b_lo = make_desc_k(smem_u32(blo_ring + (size_t)(g.blo_tma ? sbr : sbl) * B_TILE));
              const uint32_t d0 = tb + (uint32_t)(s * BN);
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                  const uint64_t adv = (uint64_t)(ks * 2);        // +32 bytes along K inside the swizzle row
                  if (PRECISE) {
                    const uint32_t d_corr = d0 + 3u * (uint32_t)g.gw;
                    umma_tf32(d0 + chain_off[ks], a_hi + adv, b_hi + adv, idesc, ks < 3 ? later : 1u);
                    umma_tf32(d_corr, a_lo + adv, b_hi + adv, idesc, ks == 0 ? later : 1u);
                    umma_tf32(d_corr, a_hi + adv, b_lo + adv, idesc, 1u);
                  } else {
                    umma_tf32(d0, a_hi + adv, b_hi + adv, idesc, ks == 0 ? later : 1u);
                    umma_tf32(d0, a_lo + adv, b_hi + adv, idesc, 1u);
                    umma_tf32(d0, a_hi + adv, b_lo + adv, idesc, 1u);
                  }
                }
                umma_commit(bar_be + 8 * sbr);
                if (!g.blo_tma) umma_commit(bar_ble + 8 * sbl);
                if (s == nsub - 1) {                              // last sub-tile: the A tiles of this chunk are free too
                  umma_commit(bar_re + 8 * sr);
                  umma_commit(bar_le + 8 * sl);
                  if (j == nchunks - 1) umma_commit(bar_tf + 8 * buf);
                }
              }
              __syncwarp();
              phb ^= 1u << sbr; phbl ^= 1u << sbl;
              if (++sbr == g.sbr) sbr = 0;
              if (++sbl == g.sbl) sbl = 0;
            }
            c0 = c1;
            phr ^= 1u << sr; phl ^= 1u << sl;
            if (++sr == g.sr) sr = 0;
            if (++sl == g.sl) sl = 0;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue warps ===========================
    const int et = tid - G_EPI_WARP0 * 32;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int total_ksteps = 0;
    for (int t = 0; t < p.nterms; ++t) total_ksteps += (p.terms[t].F + BK - 1) / BK * (BK / 8);
    const int nused = PRECISE ? min(3, total_ksteps) : 1;         // main chains that received data
    int it = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
      const int buf = it % g.nbuf;
      const uint32_t use = (uint32_t)(it / g.nbuf);
      const int tile = w / g.ngroups, col0 = (w % g.ngroups) * g.gw;
      const int gcols = min(g.gw, p.ncols - col0);                // real columns of this group (multiple of 16)
      const long long row0 = (long long)tile * BM;
      const long long R = row0 + row;
      const bool valid = R < p.total_rows;
      const int n = valid ? (int)(R / p.rows_out) : -1, r = valid ? (int)(R % p.rows_out) : 0;
      const int n_first = (int)(row0 / p.rows_out);
      float* qs = qs_all + (size_t)(it & 1) * (G_QS_FLOATS / 2);
      if (p.nslots > 0) {
        // condition broadcast vectors of this tile, group columns only: q[s][slot][c]
        const long long rlast = min(p.total_rows, row0 + BM) - 1;
        const int S = (int)(rlast / p.rows_out) - n_first + 1;
        const int total = S * p.nslots * gcols;
        for (int o = et; o < total; o += G_EPI_WARPS * 32) {
          const int c = o % gcols;
          const int slot = (o / gcols) % p.nslots;
          const int s = o / (gcols * p.nslots);
          const float* y = p.cond + (size_t)(n_first + s) * p.C;
          const float* wc = p.slot_w[slot] + col0 + c;
          const int ws = p.terms[p.slot_term[slot]].w_stride;
          float q = 0.f;
          for (int j = 0; j < p.C; ++j) q = fmaf(__ldg(y + j), __ldg(wc + (size_t)j * ws), q);
          qs[o] = q;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
      }
      const uint32_t taddr_row = tmem_base + (uint32_t)(buf * g.gw * nacc) + ((uint32_t)(quad * 32) << 16);
      const size_t orow = (size_t)R * p.ncols + col0;
      const bool use_aux = valid && (p.epilogue == CAPE_EPI_SLOPE || p.epilogue == CAPE_EPI_DUALMASK);
      const bool linear = p.epilogue == CAPE_EPI_LINEAR;
      const int act = p.act;
      const float alpha = p.alpha;
      const float* bias_row = (linear && p.bias != nullptr) ? p.bias + (p.bias_per_row ? (size_t)r * p.ncols : 0) + col0 : nullptr;
      const bool bias_vec = bias_row != nullptr && (reinterpret_cast<uintptr_t>(bias_row) & 15u) == 0;
      float4 axn[4];
      if (use_aux) {           // saved activation of this lane's row: into L2 and the first group on its way while the MMAs run
        for (int c = 0; c < gcols; c += 32) prefetch_l2(p.aux + orow + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) axn[j] = ldg4(p.aux + orow + 4 * j);
      }
      mbar_wait(bar_tf + 8 * buf, use & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < gcols; c0 += 16) {
        float v0[16];
        float4 axc[4];
        if (use_aux) {
#pragma unroll
          for (int j = 0; j < 4; ++j) axc[j] = axn[j];
          if (c0 + 16 < gcols) {
#pragma unroll
            for (int j = 0; j < 4; ++j) axn[j] = ldg4(p.aux + orow + c0 + 16 + 4 * j);
          }
        }
        if (PRECISE) {
          // precise mode: the main chains and the correction accumulator are added here, in fp32 with round-to-nearest;
          // all four TMEM loads are in flight before the one wait
          uint32_t r0[16], r1[16], r2[16], r3[16];
          tmem_ld16_async(taddr_row + (uint32_t)c0, r0);
          if (nused > 1) tmem_ld16_async(taddr_row + (uint32_t)(g.gw + c0), r1);
          if (nused > 2) tmem_ld16_async(taddr_row + (uint32_t)(2 * g.gw + c0), r2);
          tmem_ld16_async(taddr_row + (uint32_t)(3 * g.gw + c0), r3);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = __uint_as_float(r0[j]);
            if (nused > 1) v += __uint_as_float(r1[j]);
            if (nused > 2) v += __uint_as_float(r2[j]);
            v0[j] = v + __uint_as_float(r3[j]);
          }
        } else {
          tmem_ld16(taddr_row + (uint32_t)c0, v0);
        }
        if (!valid) continue;
        for (int slot = 0; slot < p.nslots; ++slot) {
          const TermDev& tm = p.terms[p.slot_term[slot]];
          const float coef = tm.op.rowsum ? __ldg(tm.op.rowsum + r) : 1.f;
          const float* q = qs + ((size_t)(n - n_first) * p.nslots + slot) * gcols + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) v0[j] = fmaf(coef, q[j], v0[j]);
        }
        float o1[16], o2[16];
        bool write2 = false;
        if (linear) {
          // 16 columns of one row: bias as four 16-byte loads, the activation chosen once per tile (not per element)
          if (bias_vec) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = ldg4(bias_row + c0 + j);
              v0[j] += b4.x; v0[j + 1] += b4.y; v0[j + 2] += b4.z; v0[j + 3] += b4.w;
            }
          } else if (bias_row != nullptr) {              // a bias that is a 4-byte-aligned view into a flat parameter buffer
#pragma unroll
            for (int j = 0; j < 16; ++j) v0[j] += __ldg(bias_row + c0 + j);
          }
          if (act == CAPE_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o1[j] = v0[j] > 0.f ? v0[j] : alpha * v0[j];
          } else if (act == CAPE_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o1[j] = fmaxf(v0[j], 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) o1[j] = v0[j];
          }
        } else {
          float ax[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 a4 = axc[j >> 2];
            ax[j] = a4.x; ax[j + 1] = a4.y; ax[j + 2] = a4.z; ax[j + 3] = a4.w;
          }
          if (p.epilogue == CAPE_EPI_SLOPE) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o1[j] = v0[j] * (ax[j] > 0.f ? 1.f : alpha);
          } else {
            write2 = p.out2 != nullptr;
#pragma unroll
            for (int j = 0; j < 16; ++j) { o1[j] = v0[j]; o2[j] = ax[j] > 0.f ? v0[j] : 0.f; }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4*>(p.out + orow + c0 + j) = make_float4(o1[j], o1[j + 1], o1[j + 2], o1[j + 3]);
          if (write2)
            *reinterpret_cast<float4*>(p.out2 + orow + c0 + j) = make_float4(o2[j], o2[j + 1], o2[j + 2], o2[j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_te + 8 * buf);
    }
  }

  __syncthreads();
  if (warp == G_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
  }
}

bool make_map(CUtensorMap* m, const float* base, unsigned long long inner, unsigned long long outer, int stride_floats,
              int box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || base == nullptr || !aligned16(base) || stride_floats % 4 != 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)stride_floats * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool PRECISE>
int launch_gemm(const cape_topology* t, const ConvParams& p, const GMaps& maps, GPlan g, cudaStream_t st) {
  constexpr int B_TILE = BN * 128;
  // Ring depths from the shared-memory budget.  An A chunk lives for nsub sub-tiles (>= 768 clocks each), so three raw
  // stages cover the TMA latency; a weight sub-tile lives for one, so the weight rings get what is left (two lo stages
  // each: the converters run one tile ahead of the MMAs).
  const int fixed = 1024 + G_QS_FLOATS * 4 + 1024;
  g.sl = 2; g.sbl = 2; g.sbr = 2;
  int left = G_SMEM_LIMIT - fixed - g.sl * G_A_TILE - (g.sbl + g.sbr) * B_TILE;
  g.sr = left / G_A_TILE;
  const int sr_want = g.nsub >= 2 ? 3 : 4;
  if (g.sr < 2) return 0;
  if (g.sr > sr_want) g.sr = sr_want;
  left -= g.sr * G_A_TILE;
  if (g.blo_tma) {                     // raw and lo weight tiles travel together: equal depths
    while (g.sbr < 4 && left >= 2 * B_TILE) { ++g.sbr; left -= 2 * B_TILE; }
    g.sbl = g.sbr;
  } else {
    while (g.sbr < G_MAX_STAGES && left >= B_TILE) { ++g.sbr; left -= B_TILE; }
  }
  while (g.sr < G_MAX_STAGES && left >= G_A_TILE) { ++g.sr; left -= G_A_TILE; }
  if (g_tuning[11] > 0 || g_tuning[12] > 0 || g_tuning[13] > 0 || g_tuning[14] > 0) {    // experiment knobs: ring depths
    const int sl = g_tuning[11] > 0 ? g_tuning[11] : g.sl, sbl = g_tuning[12] > 0 ? g_tuning[12] : g.sbl;
    const int sr = g_tuning[13] > 0 ? g_tuning[13] : g.sr, sbr = g_tuning[14] > 0 ? g_tuning[14] : g.sbr;
    if (!g.blo_tma && sl >= 2 && sbl >= 2 && sr >= 2 && sbr >= 2 && sl <= G_MAX_STAGES && sbl <= G_MAX_STAGES && sr <= G_MAX_STAGES &&
        sbr <= G_MAX_STAGES && fixed + (sr + sl) * G_A_TILE + (sbr + sbl) * B_TILE <= G_SMEM_LIMIT) {
      g.sl = sl; g.sbl = sbl; g.sr = sr; g.sbr = sbr;
    }
  }
  const int smem = fixed + (g.sr + g.sl) * G_A_TILE + (g.sbr + g.sbl) * B_TILE;
  static bool configured = false;
  if (!configured) {
    CAPE_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_LIMIT));
    configured = true;
  }
  const int nwork = g.ntiles * g.ngroups;
  const int grid = nwork < t->sm_count ? nwork : t->sm_count;
  gemm_tc_kernel<BN, PRECISE><<<grid, G_THREADS, smem, st>>>(p, maps, g);
  CAPE_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 1;
}

}  // namespace

// 1 = launched, 0 = not eligible (the caller falls through to the gather kernels), <0 = error
int launch_gemm_tc(const cape_topology* t, const ConvParams& p, bool dual, cudaStream_t st) {
  if (!tensor_cores_enabled() || g_tuning[8] == 1) return 0;
  if (dual || p.epilogue == CAPE_EPI_AFFINE) return 0;
  if (p.nterms > G_TERMS || p.ncols % 16 != 0 || p.ncols < 32 || !p.ovec) return 0;
  if (p.total_rows >= (1LL << 31)) return 0;
  long long kred = 0;
  for (int i = 0; i < p.nterms; ++i) {
    const TermDev& tm = p.terms[i];
    if (tm.op.idx != nullptr || tm.src_rows != p.rows_out || !tm.vec || tm.stash != nullptr) return 0;
    if (tm.wT == nullptr || (tm.wT_stride % 4) != 0 || !aligned16(tm.wT)) return 0;
    kred += tm.F;
  }
  if (kred < 32) return 0;
  GPlan g{};
  // MMA N: 256 halves the A-operand reads per flop (the tensor core fetches both operands from shared memory for
  // every instruction); the precise mode keeps four 128-wide accumulators
  const int BN = p.precise ? (p.ncols > 64 ? 128 : (p.ncols > 32 ? 64 : 32))
                           : (p.ncols > 128 ? 256 : (p.ncols > 64 ? 128 : (p.ncols > 32 ? 64 : 32)));
  g.ntiles = (int)((p.total_rows + BM - 1) / BM);
  const int ncols_r = (p.ncols + BN - 1) / BN * BN;      // TMEM columns are reserved in whole sub-tiles
  if (p.precise) {
    if (p.epilogue != CAPE_EPI_LINEAR) return 0;
    g.gw = BN; g.nsub = 1; g.nchain = 3; g.corr = 1;
    g.ngroups = ncols_r / BN;
  } else {
    g.nchain = 1; g.corr = 0;
    g.ngroups = (ncols_r + 511) / 512;
    const int per = (ncols_r / BN + g.ngroups - 1) / g.ngroups;      // sub-tiles per group
    g.gw = per * BN; g.nsub = per;
    g.ngroups = (ncols_r + g.gw - 1) / g.gw;
  }
  const int nacc = g.nchain + g.corr;
  g.rotate = g_tuning[9] != 1;          // experiment knob 9 = 1: every row tile walks the reduction in the same order
  g.nbuf = (2 * g.gw * nacc <= 512) ? 2 : 1;
  g.tmem_cols = 32;
  while (g.tmem_cols < g.gw * nacc * g.nbuf) g.tmem_cols *= 2;
  if (g.tmem_cols > 512) return 0;
  if (p.nslots > 0) {
    const long long max_samples = (BM - 1) / p.rows_out + 2;
    if (max_samples * p.nslots * g.gw > G_QS_FLOATS / 2) return 0;
    for (int s = 0; s < p.nslots; ++s)
      if (p.slot_acc[s] != 0) return 0;
  }
  static GMaps maps;
  g.blo_tma = g_tuning[15] != 1;        // experiment knob 15 = 1: converters derive the weight lo tiles
  for (int i = 0; i < p.nterms; ++i) {
    const TermDev& tm = p.terms[i];
    if (!make_map(&maps.a[i], tm.src, (unsigned long long)tm.F, (unsigned long long)p.total_rows, tm.src_stride, BM) ||
        !make_map(&maps.bh[i], tm.wT, (unsigned long long)tm.F, (unsigned long long)p.ncols, tm.wT_stride, BN))
      return 0;
    if (tm.wT_lo == nullptr || !aligned16(tm.wT_lo) ||
        !make_map(&maps.bl[i], tm.wT_lo, (unsigned long long)tm.F, (unsigned long long)p.ncols, tm.wT_stride, BN))
      g.blo_tma = 0;
  }
  if (p.precise) {
    if (BN == 128) return launch_gemm<128, true>(t, p, maps, g, st);
    if (BN == 64) return launch_gemm<64, true>(t, p, maps, g, st);
    return launch_gemm<32, true>(t, p, maps, g, st);
  }
  if (BN == 256) return launch_gemm<256, false>(t, p, maps, g, st);
  if (BN == 128) return launch_gemm<128, false>(t, p, maps, g, st);
  if (BN == 64) return launch_gemm<64, false>(t, p, maps, g, st);
  return launch_gemm<32, false>(t, p, maps, g, st);
}

}  // namespace cape
