// Persistent stream-K variant of the swap-AB tcgen05 GEMM for decode-size token tiles (BN <= 64).
//
// The (weight tile, k-block) iteration space is cut into G equal contiguous ranges, one per
// persistent CTA (G = one CTA per SM), so every SM streams the same number of weight bytes no
// matter how badly the tile count divides the SM count (gate/up: 224 tiles on 148 SMs), the TMA
// ring is filled once per kernel instead of once per tile, and no cluster barriers are needed.
// A CTA whose range ends inside a tile owns that tile iff it holds the tile's first k-block; the
// other CTAs covering the tile write fp32 partial accumulators to an L2-resident workspace and
// bump a per-tile counter; the owner (which reaches the tile at the END of its range, when the
// partials of the later CTAs are long finished) adds them and runs the fused epilogue.
// Accumulators are double-buffered in TMEM so the epilogue of segment i overlaps the MMAs of i+1.
//
// Roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 TMEM + MMA issue, warps 2..5 epilogue.
#include "gemm_tc.cuh"

#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "gemm_epi.cuh"
#include "launch.cuh"

namespace b2b {

constexpr int SK_BM = 128;
constexpr int SK_ROW_BYTES = 128;
constexpr int SK_A_BYTES = SK_BM * SK_ROW_BYTES;

struct SkParams {
  float* ws;              // partial accumulators [tile][max_parts][BN][128]
  uint32_t* counters;     // per tile: number of partials delivered (self-resetting)
  int ipc;                // iterations (k-blocks) per CTA
  int nkb;                // k-blocks per tile
  int n_tiles_tok;        // token tiles
  int n_tiles;            // total tiles (tile = tile_n * n_tiles_tok + tile_tok)
  int max_parts;
};

template <int BN>
struct SkCfg {
  // 5 x 20 KB (+ exchange buffer) keeps a persistent CTA under half an SM's shared memory, so the NEXT
  // kernel's CTA can become resident early (PDL) and prefetch its weights while this one drains
  static constexpr int kStages = (BN <= 32) ? 5 : 4;
  static constexpr int kStageBytes = SK_A_BYTES + BN * SK_ROW_BYTES;
  static constexpr int kAccCols = BN < 32 ? 32 : BN;
  static constexpr int kTmemCols = 2 * kAccCols;
  static constexpr int kXchBytes = BN * 64 * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + kXchBytes + 1024 /*align*/ + 512 /*barriers*/ + BN * 4;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sk_epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct SkSeg {
  int tile, kb0, kb1;
};
// iterate the segments of CTA `cta`: contiguous k-block runs, one per touched tile
struct SkIter {
  long long it, it_end;
  int nkb;
  __device__ SkIter(int cta, const SkParams& sk) {
    const long long total = static_cast<long long>(sk.n_tiles) * sk.nkb;
    it = static_cast<long long>(cta) * sk.ipc;
    it_end = it + sk.ipc < total ? it + sk.ipc : total;
    nkb = sk.nkb;
  }
  __device__ bool next(SkSeg& s) {
    if (it >= it_end) return false;
    s.tile = static_cast<int>(it / nkb);
    s.kb0 = static_cast<int>(it % nkb);
    const long long room = it_end - it;
    s.kb1 = (nkb - s.kb0 < room) ? nkb : s.kb0 + static_cast<int>(room);
    it += s.kb1 - s.kb0;
    return true;
  }
};

template <int BN, int EPI, bool FP8>
__global__ void __launch_bounds__(192, 1) gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                         const __grid_constant__ CUtensorMap tmap_x,
                                                         const GemmParams p, const SkParams sk) {
  using Cfg = SkCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int STAGE_BYTES = Cfg::kStageBytes;
  constexpr int BKE = FP8 ? 128 : 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);                  // [BN][64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + Cfg::kXchBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* rstd_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + Cfg::kXchBytes + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 4);        // one arrival per epilogue warp
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::kTmemCols>(tmem_ptr_s);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t pol_w = l2_policy_evict_first();
      const uint64_t pol_x = l2_policy_evict_last();
      // Pass 1 (before the dependency wait): weight tiles for the first STAGES k-blocks of this CTA.
      int n_pre = 0;
      {
        SkIter iter(cta, sk);
        SkSeg sg;
        while (n_pre < STAGES && iter.next(sg)) {
          const int tile_n = sg.tile / sk.n_tiles_tok;
          for (int kb = sg.kb0; kb < sg.kb1 && n_pre < STAGES; ++kb, ++n_pre) {
            mbar_arrive_expect_tx(&full_bar[n_pre], STAGE_BYTES);
            tma_load_2d_hint(smem + n_pre * STAGE_BYTES, &tmap_w, &full_bar[n_pre], kb * BKE, tile_n * SK_BM, pol_w);
          }
        }
      }
      pdl_wait();
      if (p.wait_flag != nullptr) {
        wait_flag_ge(p.wait_flag, *reinterpret_cast<const volatile uint32_t*>(p.wait_epoch) + 1);
        fence_proxy_async_all();
      }
      // Pass 2: activations for the prefetched stages, then the steady-state ring.
      int kbi = 0;                       // running k-block index of this CTA
      SkIter iter(cta, sk);
      SkSeg sg;
      while (iter.next(sg)) {
        const int tile_n = sg.tile / sk.n_tiles_tok, tok0 = (sg.tile % sk.n_tiles_tok) * BN;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++kbi) {
          const int s = kbi % STAGES;
          if (kbi < n_pre) {
            tma_load_2d_hint(smem + s * STAGE_BYTES + SK_A_BYTES, &tmap_x, &full_bar[s], kb * BKE, tok0, pol_x);
          } else {
            const uint32_t ph = (kbi / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_2d_hint(smem + s * STAGE_BYTES, &tmap_w, &full_bar[s], kb * BKE, tile_n * SK_BM, pol_w);
            tma_load_2d_hint(smem + s * STAGE_BYTES + SK_A_BYTES, &tmap_x, &full_bar[s], kb * BKE, tok0, pol_x);
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = FP8 ? make_idesc_e4m3(SK_BM, BN) : make_idesc_bf16(SK_BM, BN);
    int kbi = 0, seg = 0;
    SkIter iter(cta, sk);
    SkSeg sg;
    while (iter.next(sg)) {
      const int acc = seg & 1;
      const uint32_t use = static_cast<uint32_t>(seg >> 1);          // how often this accumulator was used before
      mbar_wait(&tempty_bar[acc], (use & 1) ^ 1);                     // epilogue drained it
      tc_fence_after();
      const uint32_t tacc = tmem_base + static_cast<uint32_t>(acc * Cfg::kAccCols);
      for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++kbi) {
        const int s = kbi % STAGES;
        mbar_wait(&full_bar[s], (kbi / STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem + s * STAGE_BYTES));
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem + s * STAGE_BYTES + SK_A_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (kb > sg.kb0 || k > 0) ? 1u : 0u;
            if constexpr (FP8) umma_f8(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
            else umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
          }
          umma_commit(&empty_bar[s]);
          if (kb == sg.kb1 - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
      ++seg;
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    pdl_wait();
    int cur_tok_tile = -1, seg = 0;
    SkIter iter(cta, sk);
    SkSeg sg;
    while (iter.next(sg)) {
      const int acc = seg & 1;
      const uint32_t use = static_cast<uint32_t>(seg >> 1);
      const int tile_n = sg.tile / sk.n_tiles_tok, tok_tile = sg.tile % sk.n_tiles_tok, tok0 = tok_tile * BN;
      const bool full = (sg.kb0 == 0 && sg.kb1 == sk.nkb);
      const bool owner = (sg.kb0 == 0);
      // per-token input scale for the tiles this CTA finishes (computed while the MMAs run)
      if (owner && tok_tile != cur_tok_tile) {
        sk_epi_bar();                                  // previous users of rstd_s are done
        if (p.norm_src != nullptr) {
          if (p.wait_flag != nullptr)
            wait_flag_ge(p.wait_flag, *reinterpret_cast<const volatile uint32_t*>(p.wait_epoch) + 1);
          const int kv8 = p.k / 8;
          for (int t = warp - 2; t < BN; t += 16) {
            float ss[4] = {0.f, 0.f, 0.f, 0.f};
            const uint4* rowp[4];
            bool live[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int tok = tok0 + t + 4 * u;
              live[u] = (t + 4 * u < BN) && tok < p.m_tok;
              rowp[u] = reinterpret_cast<const uint4*>(p.norm_src + static_cast<size_t>(live[u] ? tok : tok0) * p.k);
            }
            for (int i = lane; i < kv8; i += 32) {
              uint4 v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) v[u] = live[u] ? rowp[u][i] : make_uint4(0, 0, 0, 0);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 f = __bfloat1622float2(h[j]);
                  ss[u] += f.x * f.x + f.y * f.y;
                }
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], o);
              if (lane == 0 && t + 4 * u < BN)
                rstd_s[t + 4 * u] = live[u] ? rsqrtf(ss[u] / static_cast<float>(p.k) + p.eps) : 0.f;
            }
          }
        } else {
          for (int t = et; t < BN; t += 128) {
            const int tok = tok0 + t;
            rstd_s[t] = (p.rstd != nullptr && tok < p.m_tok) ? p.rstd[tok] : 1.f;
          }
        }
        sk_epi_bar();
        cur_tok_tile = tok_tile;
      }

      // Owner of a tile that other CTAs contributed to: their partials were produced at the START of
      // those CTAs' ranges, i.e. long ago -- fetch and sum them into registers now, while this CTA's
      // own MMAs for the tile are still running, so the L2 round trips are off the critical path.
      float pacc[BN];
      int parts = 0;
      if (owner && !full) {
        const int last_cta = static_cast<int>((static_cast<long long>(sg.tile + 1) * sk.nkb - 1) / sk.ipc);
        parts = last_cta - cta;
        uint32_t spins = 0;
        while (ld_acquire_gpu(&sk.counters[sg.tile]) < static_cast<uint32_t>(parts)) {
          if (++spins > B2B_SPIN_LIMIT) { __trap(); }
        }
        const float* wbase = sk.ws + static_cast<size_t>(sg.tile) * sk.max_parts * BN * SK_BM + row;
#pragma unroll
        for (int c = 0; c < BN; ++c) pacc[c] = 0.f;
        for (int r = 0; r < parts; ++r) {
#pragma unroll
          for (int c = 0; c < BN; ++c) pacc[c] += __ldcg(wbase + (static_cast<size_t>(r) * BN + c) * SK_BM);
        }
      }

      mbar_wait(&tfull_bar[acc], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * Cfg::kAccCols) + (static_cast<uint32_t>(q * 32) << 16);

      if (!owner) {
        // ---- partial: accumulators -> workspace[tile][part][col][row], then count
        const int owner_cta = static_cast<int>((static_cast<long long>(sg.tile) * sk.nkb) / sk.ipc);
        const int part = cta - owner_cta - 1;
        float* w = sk.ws + (static_cast<size_t>(sg.tile) * sk.max_parts + part) * BN * SK_BM + row;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
          float v[16];
          tmem_ld16(taddr + c, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) w[static_cast<size_t>(c + i) * SK_BM] = v[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);          // accumulator free for segment seg+2
        __threadfence();
        sk_epi_bar();
        if (et == 0) atomicAdd(&sk.counters[sg.tile], 1u);
      } else {
        const EpiCtx ectx = epi_setup<EPI, FP8>(p, tile_n, row);
        if (p.free_flag != nullptr)
          wait_flag_ge(p.free_flag, *reinterpret_cast<const volatile uint32_t*>(p.signal_epoch));
        if constexpr (EPI == EPI_GLU) {
          if (row >= 64) {
#pragma unroll
            for (int c = 0; c < BN; c += 16) {
              float v[16];
              tmem_ld16(taddr + c, v);
              if (parts > 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += pacc[c + i];
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) xch[(c + i) * 64 + (row - 64)] = v[i];
            }
          }
          sk_epi_bar();
        }
#pragma unroll
        for (int c = 0; c < BN; c += 16) {
          if (EPI == EPI_GLU && row >= 64) break;
          if (tok0 + c >= p.m_tok) break;
          float v[16];
          tmem_ld16(taddr + c, v);
          if (parts > 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += pacc[c + i];
          }
          epi_apply16<EPI, FP8>(p, ectx, v, c, tok0, tile_n, row, lane, rstd_s, xch);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        const bool publish = (p.signal_flag != nullptr || p.bump_epoch != nullptr);
        if (publish) __threadfence_system();
        if (publish || !full || EPI == EPI_GLU) sk_epi_bar();     // all reads of partials / xch done, all stores fenced
        if (et == 0) {
          if (!full) sk.counters[sg.tile] = 0;                    // ready for the next launch
          if (publish) epi_publish_tile(p, static_cast<uint32_t>(sk.n_tiles));
        }
      }
      ++seg;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ============================================================================ host side
typedef CUresult (*PFN_encodeTiledSk)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_shared(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     int elt_bytes);   // gemm_tc.cu

struct SkDeviceState {
  float* ws = nullptr;
  uint32_t* counters = nullptr;
  size_t ws_bytes = 0;
  int sms = 0;
};
static SkDeviceState& sk_state() {
  static std::map<int, SkDeviceState> st;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  SkDeviceState& s = st[dev];
  if (s.ws == nullptr) {
    s.ws_bytes = 64u << 20;
    cudaMalloc(&s.ws, s.ws_bytes);
    cudaMalloc(&s.counters, 16384 * sizeof(uint32_t));
    cudaMemset(s.counters, 0, 16384 * sizeof(uint32_t));
    cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return s;
}

template <int BN, int EPI, bool FP8>
static int launch_sk_one(const GemmParams& p, const SkParams& sk, int grid, const CUtensorMap& tw,
                         const CUtensorMap& tx, cudaStream_t stream) {
  using Cfg = SkCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_sk_kernel<BN, EPI, FP8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  return static_cast<int>(launch_kernel(gemm_sk_kernel<BN, EPI, FP8>, dim3(grid), dim3(192), Cfg::kSmemBytes, stream,
                                        1u, tw, tx, p, sk));
}

template <int BN>
static int launch_sk_bn(const GemmParams& p, const SkParams& sk, int grid, const CUtensorMap& tw,
                        const CUtensorMap& tx, cudaStream_t stream) {
#define B2B_SK_CASE(E)                                                                                      \
  case E:                                                                                                   \
    return p.fp8 ? launch_sk_one<BN, E, true>(p, sk, grid, tw, tx, stream)                                  \
                 : launch_sk_one<BN, E, false>(p, sk, grid, tw, tx, stream);
  switch (p.epi) {
    B2B_SK_CASE(EPI_PLAIN)
    B2B_SK_CASE(EPI_RESIDUAL)
    B2B_SK_CASE(EPI_GLU)
    B2B_SK_CASE(EPI_QKV_ROPE)
    B2B_SK_CASE(EPI_GELU)
    default: return -4;
  }
#undef B2B_SK_CASE
}

template <int BN, int EPI>
static int sk_attr_pair() {
  int r = static_cast<int>(cudaFuncSetAttribute(gemm_sk_kernel<BN, EPI, false>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<BN>::kSmemBytes));
  if (r) return r;
  return static_cast<int>(cudaFuncSetAttribute(gemm_sk_kernel<BN, EPI, true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<BN>::kSmemBytes));
}
template <int BN>
static int sk_attr_bn() {
  int r = 0;
  if ((r = sk_attr_pair<BN, EPI_PLAIN>())) return r;
  if ((r = sk_attr_pair<BN, EPI_RESIDUAL>())) return r;
  if ((r = sk_attr_pair<BN, EPI_GLU>())) return r;
  if ((r = sk_attr_pair<BN, EPI_QKV_ROPE>())) return r;
  return sk_attr_pair<BN, EPI_GELU>();
}
int gemm_sk_init() {
  int r = 0;
  if ((r = sk_attr_bn<16>())) return r;
  if ((r = sk_attr_bn<32>())) return r;
  if ((r = sk_attr_bn<64>())) return r;
  sk_state();
  return 0;
}

bool gemm_sk_supported(const GemmParams& p, int bn) { return bn <= 64 && p.n_out % SK_BM == 0; }

int launch_gemm_sk(const GemmParams& p_in, const void* w, const void* x, int bn, cudaStream_t stream) {
  GemmParams p = p_in;
  const int elt = p.fp8 ? 1 : 2;
  const int bke = SK_ROW_BYTES / elt;
  if (p.n_out % SK_BM != 0 || p.k % bke != 0 || p.m_tok <= 0 || bn > 64) return -2;
  SkDeviceState& st = sk_state();
  SkParams sk;
  sk.nkb = p.k / bke;
  sk.n_tiles_tok = (p.m_tok + bn - 1) / bn;
  sk.n_tiles = (p.n_out / SK_BM) * sk.n_tiles_tok;
  const long long total = static_cast<long long>(sk.n_tiles) * sk.nkb;
  int grid = st.sms > 0 ? st.sms : 148;
  // at least 4 k-blocks per CTA: tiny problems use fewer CTAs rather than drowning in fix-ups
  if (total / 4 < grid) grid = static_cast<int>(total / 4 > 0 ? total / 4 : 1);
  sk.ipc = static_cast<int>((total + grid - 1) / grid);
  grid = static_cast<int>((total + sk.ipc - 1) / sk.ipc);
  sk.max_parts = (sk.nkb + sk.ipc - 1) / sk.ipc + 1;
  const size_t need = static_cast<size_t>(sk.n_tiles) * sk.max_parts * bn * SK_BM * sizeof(float);
  if (need > st.ws_bytes || sk.n_tiles > 16384) return -5;      // caller falls back to the cluster kernel
  sk.ws = st.ws;
  sk.counters = st.counters;
  p.splitk = 1;
  CUtensorMap tw, tx;
  int r = make_tmap_shared(&tw, w, p.n_out, p.k, p.k, SK_BM, elt);
  if (r) return r;
  r = make_tmap_shared(&tx, x, p.m_tok, p.k, p.k, bn, elt);
  if (r) return r;
  switch (bn) {
    case 16: return launch_sk_bn<16>(p, sk, grid, tw, tx, stream);
    case 32: return launch_sk_bn<32>(p, sk, grid, tw, tx, stream);
    case 64: return launch_sk_bn<64>(p, sk, grid, tw, tx, stream);
    default: return -3;
  }
}

}  // namespace b2b
