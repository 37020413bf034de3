// Dense weight-gradient contraction on TMA + tcgen05:  dW[f, c] = sum_r X[r, f] * G[r, c]   (3xTF32, fp32 in TMEM)
//
// Used whenever the basis operand of cape_cheb_dw is a plain tensor (identity operator: the T_0 term of un-pooled
// layers, the affine branch, 1x1 convs, and the "narrow side" form where the operators were applied to G first).
// Both operands are then dense row streams, and rows are the MMA K dimension, so both are MN-major in HBM already.
//
//   * one thread issues TMA box loads (32 floats x 32 rows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): each box lands in
//     shared memory as one MN block of the SWIZZLE_128B_BASE32B operand layout -- no register staging, no address
//     arithmetic, out-of-range rows/columns zero-filled by the TMA unit;
//   * tcgen05 kind::tf32 reads the top 19 bits of each fp32 word, so the raw tile IS the "hi" operand of the
//     3xTF32 scheme; eight converter warps only derive the "lo" tile (x - trunc_tf32(x)) from shared memory;
//   * one thread issues the MMAs (hi*hi + lo*hi + hi*lo per 8-row K group) into TMEM; the converter warps drain
//     TMEM at the end.  A CTA owns a 128-wide slice of f, all output columns and one split of the rows;
//     partial sums of the row splits go to the workspace (reduced deterministically by reduce_splits_kernel);
//   * the TMA ("hi") rings are deep (the loads have ~2 us to cover), the converter ("lo") rings are two stages:
//     ring depths are run-time parameters chosen on the host from the tile sizes (dd_plan).
#include "common.cuh"
#include "ellconv_params.cuh"
#include "tc_common.cuh"

namespace cape {

namespace {

using namespace tc;

constexpr int DD_CONV_WARPS = 8;
constexpr int DD_CONV_THREADS = DD_CONV_WARPS * 32;
constexpr int DD_THREADS = DD_CONV_THREADS + 64;    // + MMA warp + TMA warp
constexpr int DD_KCH = 32;                          // rows per pipeline stage = 4 MMAs of K = 8
constexpr int DD_A_TILE = 4 * 4096;                 // 128 f x 32 rows, hi or lo
constexpr int DD_MAX_STAGES = 8;
constexpr int DD_SMEM_BUDGET = 212 * 1024;          // rings only (barriers and alignment slack come on top)

struct DdParams {
  int ncols, F;
  long long total_rows, rows_per_split;
  float* out;          // dw (nsplit == 1) or workspace [nsplit, F, ncols]
  long long out_rs;
  int nsplit, accumulate, ovec, lo_mode;
  int sah, sal, sgh, sgl;      // ring depths: A hi (TMA), A lo (converters), G hi, G lo
  int budget;                  // bytes of shared memory the rings may use (the barriers sit right behind)
};

template <int BN>
struct DdCfg {
  static constexpr int G_TILE = (BN / 32) * 4096;
  static constexpr int SMEM_BYTES = 1024 + DD_SMEM_BUDGET + 1024;
};

// Ring depths for a problem with `nct` column sub-tiles: two lo stages each, the rest of the budget to the TMA rings
// with the A ring about as many CHUNKS deep as the G ring (one A tile is used by nct G tiles).
template <int BN>
void dd_plan(int nct, DdParams* p) {
  constexpr int GT = DdCfg<BN>::G_TILE;
  p->sal = 2; p->sgl = 2;
  int left = p->budget - 2 * DD_A_TILE - 2 * GT;
  int sah = nct >= 2 ? 3 : 2, sgh = 2;       // a chunk of G sub-tiles takes a while: keep two A tiles ahead
  left -= sah * DD_A_TILE + sgh * GT;
  while (true) {
    // next stage goes to the ring that currently looks ahead fewer chunks
    const bool to_g = (sgh < DD_MAX_STAGES) && ((long long)sgh < (long long)sah * nct || sah >= DD_MAX_STAGES);
    const int cost = to_g ? GT : DD_A_TILE;
    if ((to_g ? sgh : sah) >= DD_MAX_STAGES || cost > left) break;
    left -= cost;
    if (to_g) ++sgh; else ++sah;
  }
  p->sah = sah; p->sgh = sgh;
}

// lo = x - tf32(x); mode 0: tf32(x) = truncation (what the tensor core does with a raw fp32 word)
__device__ __forceinline__ float lo_part(float x, int mode) {
  float h;
  if (mode == 0) {
    h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  } else {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    h = __uint_as_float(u);
  }
  return x - h;
}

template <int BN>
__global__ void __launch_bounds__(DD_THREADS, 2)
dw_dense_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG,
                const __grid_constant__ DdParams p, int nct, int tmem_cols) {
  using Cfg = DdCfg<BN>;
  const int SAH = p.sah, SAL = p.sal, SGH = p.sgh, SGL = p.sgl;
  extern __shared__ uint8_t smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  char* ahi_ring = smem;
  char* alo_ring = ahi_ring + SAH * DD_A_TILE;
  char* ghi_ring = alo_ring + SAL * DD_A_TILE;
  char* glo_ring = ghi_ring + SGH * Cfg::G_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.budget);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 * DD_MAX_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ftile = blockIdx.x * 128;
  const long long rbeg = (long long)blockIdx.y * p.rows_per_split;
  const long long rend = min(p.total_rows, rbeg + p.rows_per_split);
  const long long nchunks = (rend - rbeg + DD_KCH - 1) / DD_KCH;
  // MN blocks of the A tile that hold real channels: the others are neither loaded nor converted -- whatever shared
  // memory holds there only reaches accumulator rows f >= F, which the epilogue never stores
  const int a_blocks = min(4, (p.F - ftile + 31) / 32);
  // per ring: "full" (TMA bytes landed / converters done) and "empty" (the MMAs that read the stage have retired)
  const uint32_t bar_ahf = smem_u32(bars), bar_ahe = smem_u32(bars + DD_MAX_STAGES);
  const uint32_t bar_alf = smem_u32(bars + 2 * DD_MAX_STAGES), bar_ale = smem_u32(bars + 3 * DD_MAX_STAGES);
  const uint32_t bar_ghf = smem_u32(bars + 4 * DD_MAX_STAGES), bar_ghe = smem_u32(bars + 5 * DD_MAX_STAGES);
  const uint32_t bar_glf = smem_u32(bars + 6 * DD_MAX_STAGES), bar_gle = smem_u32(bars + 7 * DD_MAX_STAGES);
  const uint32_t bar_accum = smem_u32(bars + 8 * DD_MAX_STAGES);

  if (warp == DD_CONV_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < DD_MAX_STAGES; ++s) {
        mbar_init(bar_ahf + 8 * s, 1); mbar_init(bar_ahe + 8 * s, 1);
        mbar_init(bar_alf + 8 * s, DD_CONV_WARPS); mbar_init(bar_ale + 8 * s, 1);
        mbar_init(bar_ghf + 8 * s, 1); mbar_init(bar_ghe + 8 * s, 1);
        mbar_init(bar_glf + 8 * s, DD_CONV_WARPS); mbar_init(bar_gle + 8 * s, 1);
      }
      mbar_init(bar_accum, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < DD_CONV_WARPS) {
    // =========================== converters: lo tiles from the TMA-written hi tiles ===========================
    int ah = 0, al = 0, gh = 0, gl = 0;
    uint32_t pah = 0, pal = 0, pgh = 0, pgl = 0;
    const int mode = p.lo_mode;
    for (long long kc = 0; kc < nchunks; ++kc) {
      {
        mbar_wait(bar_ahf + 8 * ah, pah);
        mbar_wait(bar_ale + 8 * al, pal ^ 1);
        const char* hi = ahi_ring + (size_t)ah * DD_A_TILE;
        char* lo = alo_ring + (size_t)al * DD_A_TILE;
#pragma unroll
        for (int i = 0; i < DD_A_TILE / 16 / DD_CONV_THREADS; ++i) {
          if (i >= a_blocks) break;                    // one MN block = 4096 B = one pass of the 256 converter threads
          const int off = (i * DD_CONV_THREADS + tid) * 16;
          const float4 v = *reinterpret_cast<const float4*>(hi + off);
          *reinterpret_cast<float4*>(lo + off) =
              make_float4(lo_part(v.x, mode), lo_part(v.y, mode), lo_part(v.z, mode), lo_part(v.w, mode));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_alf + 8 * al);
        if (++ah == SAH) { ah = 0; pah ^= 1; }
        if (++al == SAL) { al = 0; pal ^= 1; }
      }
      for (int cs = 0; cs < nct; ++cs) {
        mbar_wait(bar_ghf + 8 * gh, pgh);
        mbar_wait(bar_gle + 8 * gl, pgl ^ 1);
        const char* hi = ghi_ring + (size_t)gh * Cfg::G_TILE;
        char* lo = glo_ring + (size_t)gl * Cfg::G_TILE;
#pragma unroll
        for (int i = 0; i < Cfg::G_TILE / 16 / DD_CONV_THREADS; ++i) {
          const int off = (i * DD_CONV_THREADS + tid) * 16;
          const float4 v = *reinterpret_cast<const float4*>(hi + off);
          *reinterpret_cast<float4*>(lo + off) =
              make_float4(lo_part(v.x, mode), lo_part(v.y, mode), lo_part(v.z, mode), lo_part(v.w, mode));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_glf + 8 * gl);
        if (++gh == SGH) { gh = 0; pgh ^= 1; }
        if (++gl == SGL) { gl = 0; pgl ^= 1; }
      }
    }

    // =========================== epilogue: TMEM -> partial sums ===========================
    mbar_wait(bar_accum, 0);
    tc_fence_after();
    const int quad = warp & 3, half = warp >> 2;
    const int f = ftile + quad * 32 + lane;
    const int cpw = p.ncols >> 1;
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* orow = p.out + (p.nsplit > 1 ? (size_t)blockIdx.y * p.F * p.out_rs : 0) + (size_t)f * p.out_rs;
#pragma unroll 1
    for (int g = 0; g < cpw / 16; ++g) {
      const int c0 = half * cpw + g * 16;
      float v[16];
      tmem_ld16(taddr_row + (uint32_t)c0, v);
      if (f >= p.F) continue;
      if (nchunks == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (p.nsplit == 1 && p.accumulate) {
#pragma unroll
        for (int j = 0; j < 16; ++j) orow[c0 + j] += v[j];
      } else if (p.ovec) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) orow[c0 + j] = v[j];
      }
    }
    tc_fence_before();
  } else if (warp == DD_CONV_WARPS) {
    // =========================== MMA issuer (whole warp walks the loops, one elected lane issues) ===========================
    {
      // D=F32, A=B=TF32, A and B MN-major (bits 15,16), N=BN, M=128
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int ah = 0, al = 0, gh = 0, gl = 0;
      uint32_t pah = 0, pal = 0, pgh = 0, pgl = 0, acc_on = 0;
      for (long long kc = 0; kc < nchunks; ++kc) {
        mbar_wait(bar_ahf + 8 * ah, pah);
        mbar_wait(bar_alf + 8 * al, pal);
        const uint32_t ahi = smem_u32(ahi_ring + (size_t)ah * DD_A_TILE);
        const uint32_t alo = smem_u32(alo_ring + (size_t)al * DD_A_TILE);
        for (int cs = 0; cs < nct; ++cs) {
          mbar_wait(bar_ghf + 8 * gh, pgh);
          mbar_wait(bar_glf + 8 * gl, pgl);
          tc_fence_after();
          const uint32_t ghi = smem_u32(ghi_ring + (size_t)gh * Cfg::G_TILE);
          const uint32_t glo = smem_u32(glo_ring + (size_t)gl * Cfg::G_TILE);
          const uint32_t d = tmem_base + (uint32_t)(cs * BN);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < DD_KCH / 8; ++ks) {
              const uint64_t a_hi = make_desc_mn(ahi + ks * 1024), a_lo = make_desc_mn(alo + ks * 1024);
              const uint64_t g_hi = make_desc_mn(ghi + ks * 1024), g_lo = make_desc_mn(glo + ks * 1024);
              umma_tf32(d, a_hi, g_hi, idesc, ks == 0 ? acc_on : 1u);
              umma_tf32(d, a_lo, g_hi, idesc, 1);
              umma_tf32(d, a_hi, g_lo, idesc, 1);
            }
            umma_commit(bar_ghe + 8 * gh);
            umma_commit(bar_gle + 8 * gl);
            if (cs == nct - 1) {
              umma_commit(bar_ahe + 8 * ah);
              umma_commit(bar_ale + 8 * al);
            }
          }
          __syncwarp();
          if (++gh == SGH) { gh = 0; pgh ^= 1; }
          if (++gl == SGL) { gl = 0; pgl ^= 1; }
        }
        if (++ah == SAH) { ah = 0; pah ^= 1; }
        if (++al == SAL) { al = 0; pal ^= 1; }
        acc_on = 1;
      }
      if (elect_one()) umma_commit(bar_accum);          // also flips with no chunks at all (empty row range)
    }
    __syncwarp();
  } else {
    // =========================== TMA issuer (whole warp walks the loops, one elected lane issues) ===========================
    {
      if (lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmG);
      }
      __syncwarp();
      // The A ring runs ahead of the G ring by whole chunks, so the loop is over G sub-tiles in consumption order and
      // an A tile is issued as soon as its stage is free (never blocking the G stream behind a full A ring).
      int ah = 0, gh = 0;
      uint32_t pah = 0, pgh = 0;
      long long ka = 0;                                  // next A chunk to issue
      auto issue_a = [&](bool block) {
        while (ka < nchunks) {
          if (!block) {
            uint32_t ok;
            asm volatile(
                "{\n\t.reg .pred q;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, q;\n\t}"
                : "=r"(ok) : "r"(bar_ahe + 8 * ah), "r"(pah ^ 1) : "memory");
            if (!__all_sync(0xffffffffu, ok)) return;    // one answer for the whole warp
          } else {
            mbar_wait(bar_ahe + 8 * ah, pah ^ 1);
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_ahf + 8 * ah, (uint32_t)(a_blocks * 4096));
            const uint32_t adst = smem_u32(ahi_ring + (size_t)ah * DD_A_TILE);
            const int row0 = (int)(rbeg + ka * DD_KCH);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
              if (mb < a_blocks) tma_load_2d(adst + mb * 4096, &tmA, ftile + mb * 32, row0, bar_ahf + 8 * ah);
          }
          __syncwarp();
          if (++ah == SAH) { ah = 0; pah ^= 1; }
          ++ka;
          block = false;
        }
      };
      for (long long kc = 0; kc < nchunks; ++kc) {
        const int row0 = (int)(rbeg + kc * DD_KCH);
        issue_a(ka <= kc);                               // the A tile of this chunk must be on its way before its G tiles
        for (int cs = 0; cs < nct; ++cs) {
          mbar_wait(bar_ghe + 8 * gh, pgh ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_ghf + 8 * gh, Cfg::G_TILE);
            const uint32_t gdst = smem_u32(ghi_ring + (size_t)gh * Cfg::G_TILE);
#pragma unroll
            for (int mb = 0; mb < BN / 32; ++mb)
              tma_load_2d(gdst + mb * 4096, &tmG, cs * BN + mb * 32, row0, bar_ghf + 8 * gh);
          }
          __syncwarp();
          if (++gh == SGH) { gh = 0; pgh ^= 1; }
          issue_a(false);
        }
      }
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == DD_CONV_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}


// [rows, inner] fp32 row-major with `row_stride` floats between rows; boxes of 32 floats x 32 rows, MN-swizzled
bool make_map(CUtensorMap* m, const float* base, long long inner, long long rows, long long row_stride) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride * sizeof(float)};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
int launch_dd(const CUtensorMap& ma, const CUtensorMap& mg, const DdParams& p, int ftiles, cudaStream_t st) {
  using Cfg = DdCfg<BN>;
  static bool configured = false;
  if (!configured) {
    CAPE_CHECK_CUDA(cudaFuncSetAttribute(dw_dense_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int nct = (p.ncols + BN - 1) / BN;
  int tmem_cols = 32;
  while (tmem_cols < nct * BN) tmem_cols *= 2;
  DdParams q = p;
  // narrow problems (little MMA work per chunk, the pipeline hand-offs dominate): half the shared memory per CTA so
  // that two CTAs share an SM and overlap each other's stalls (measured: 118 -> 80 us for 441k rows x 64 x 64)
  q.budget = BN <= 64 ? 104 * 1024 : DD_SMEM_BUDGET;
  dd_plan<BN>(nct, &q);
  dim3 grid(ftiles, q.nsplit);
  dw_dense_kernel<BN><<<grid, DD_THREADS, 1024 + q.budget + 1024, st>>>(ma, mg, q, nct, tmem_cols);
  CAPE_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 1;
}

}  // namespace

// 1 = launched (partials in the workspace when *nsplit_out > 1), 0 = not eligible, <0 = error
int launch_dw_dense_tma(const cape_topology* t, const cape_dw_args* a, const OpView& op, int* nsplit_out,
                        cudaStream_t st) {
  if (!tensor_cores_enabled() || g_tuning[1] == 1) return 0;
  if (op.idx != nullptr || a->src_rows != a->rows_out) return 0;                 // dense basis operand only
  if (a->ncols % 32 != 0 || a->ncols < 32 || a->ncols > 512) return 0;
  if (a->F % 4 != 0 || a->F < 32 || a->src_stride % 4 != 0 || !aligned16(a->src) || !aligned16(a->g)) return 0;
  const long long total_rows = (long long)a->N * a->rows_out;
  if (total_rows < 4096 || total_rows >= (1LL << 31)) return 0;
  CUtensorMap ma, mg;
  if (!make_map(&ma, a->src, a->F, total_rows, a->src_stride)) return 0;
  if (!make_map(&mg, a->g, a->ncols, total_rows, a->ncols)) return 0;
  DdParams p{};
  p.ncols = a->ncols; p.F = a->F; p.total_rows = total_rows; p.lo_mode = g_tuning[2];
  const int ftiles = (a->F + 127) / 128;
  // one full wave of CTAs; two when the reduction per CTA would get long (fp32 accumulation error grows with it)
  const bool two_per_sm = a->ncols <= 64;                            // two CTAs share an SM: see launch_dd
  long long nsplit = (two_per_sm ? 2 : 1) * t->sm_count / ftiles;
  if (nsplit < 1) nsplit = 1;
  if (total_rows / nsplit > 4096) nsplit *= 2;
  const long long max_by_rows = (total_rows + 255) / 256;
  if (nsplit > max_by_rows) nsplit = max_by_rows;
  const long long per = (long long)a->F * a->ncols * (long long)sizeof(float);
  if (nsplit > 1 && nsplit * per > t->workspace_bytes) nsplit = t->workspace_bytes / per;
  if (nsplit < 1) nsplit = 1;
  long long rps = (total_rows + nsplit - 1) / nsplit;
  rps = (rps + DD_KCH - 1) / DD_KCH * DD_KCH;
  nsplit = (total_rows + rps - 1) / rps;
  p.rows_per_split = rps; p.nsplit = (int)nsplit; p.accumulate = a->accumulate;
  if (nsplit == 1) { p.out = a->dw; p.out_rs = a->dw_stride; }
  else { p.out = (float*)t->workspace; p.out_rs = a->ncols; }
  p.ovec = (p.out_rs % 4 == 0) && aligned16(p.out);
  *nsplit_out = (int)nsplit;
  if (a->ncols % 256 == 0 && g_tuning[3] != 2) return launch_dd<256>(ma, mg, p, ftiles, st);   // fewer, wider MMAs
  if (a->ncols > 64) return launch_dd<128>(ma, mg, p, ftiles, st);
  if (a->ncols > 32) return launch_dd<64>(ma, mg, p, ftiles, st);
  return launch_dd<32>(ma, mg, p, ftiles, st);
}

}  // namespace cape
