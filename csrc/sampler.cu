// Fused sampler for sm_100a: repetition penalty (seen-token bitmap) -> temperature ->
// exact top-p via two-level radix histogram of the probability bits -> multinomial draw
// (or argmax when temperature <= 0).  One CTA per sequence, logits are fp32 [B, ld] as
// written by the lm_head GEMM.  The sampled id is stored locally (token ring, history
// bitmap) and, on the last piece of a pipeline, straight into piece 0's token buffer on
// the peer GPU followed by a release flag (4 bytes/sequence over NVLink, no NCCL).
//
// Semantics follow the reference's generation defaults (bee2bee/hf.py:91-105):
// repetition_penalty 1.15, top_p 0.95, do_sample iff temperature > 0, greedy otherwise.
//
// Performance notes (profiles/launches_decode_step.md): every pass over the 128k logits is
// L2-latency bound, so logits are read 8 at a time per thread (2 x 128-bit loads in flight);
// histogram mass is accumulated in 32.32 fixed point with native 32-bit shared-memory atomics
// (float / 64-bit shared atomics are CAS loops).
#include "kernels.h"

#include "common.cuh"
#include "launch.cuh"

namespace b2b {

constexpr int SAMP_THREADS = 1024;
constexpr int NBINS = 4096;

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

struct SampleParams {
  const float* logits;      // [B, ld]
  uint32_t* seen;           // [rows, ceil(V/32)] bitmap of ids in the context, or null
  int* out_tokens;          // [rows] local
  int* peer_tokens;         // [rows] on piece 0 (may be == out_tokens / null)
  int* history;             // [rows, hist_stride] token ring, or null
  const int* hist_pos;      // [rows] write index into history
  int* hist_pos_out;        // [rows] incremented copy
  int hist_stride;
  int vocab;
  int ld;                   // row stride of logits (vocab padded to a GEMM tile)
  float softcap;            // final-logit soft-capping (Gemma-2), 0 = off
  const float* temperature; // [rows]
  const float* top_p;       // [rows]
  const float* rep_penalty; // [rows]
  const uint32_t* seeds;    // [rows]
  const uint32_t* step;     // device step counter (rng stream), may be null
  uint32_t* signal_flag;    // peer flag (token handoff) or null
  uint32_t* signal_epoch;   // local epoch for the flag
  uint32_t* done_counter;   // local, self-resetting
  const int* row_base;      // optional: per-sequence state lives at row *row_base + b
};

// logit -> (soft-cap) -> repetition penalty -> temperature
__device__ __forceinline__ float transform(float l, bool is_seen, float pen, float inv_temp, float cap) {
  if (cap > 0.f) l = cap * tanhf(l / cap);
  if (is_seen) l = l > 0.f ? l / pen : l * pen;
  return l * inv_temp;
}

// Visit every id in [lo, hi) (lo % 4 == 0): this thread takes quads first_quad, first_quad + quad_stride, ...
// and calls f(id, transformed_logit).  Two quads (8 logits) are loaded before either is processed.
template <typename F>
__device__ __forceinline__ void visit(const float* logits, const uint32_t* seen, int lo, int hi, int first_quad,
                                      int quad_stride, bool vec_ok, float pen, float inv_temp, float cap, F f) {
  const int nq = (hi - lo) / 4;
  for (int q = first_quad; q < nq; q += 2 * quad_stride) {
    const int q2 = q + quad_stride;
    const bool has2 = q2 < nq;
    const int i0 = lo + 4 * q, i1 = lo + 4 * q2;
    float a[4], b[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec_ok) {
      const float4 va = *reinterpret_cast<const float4*>(logits + i0);
      a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w;
      if (has2) {
        const float4 vb = *reinterpret_cast<const float4*>(logits + i1);
        b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = logits[i0 + k];
      if (has2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] = logits[i1 + k];
      }
    }
    const uint32_t sa = seen ? seen[i0 >> 5] : 0u;
    const uint32_t sb = (seen && has2) ? seen[i1 >> 5] : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) f(i0 + k, transform(a[k], (sa >> ((i0 + k) & 31)) & 1u, pen, inv_temp, cap));
    if (has2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) f(i1 + k, transform(b[k], (sb >> ((i1 + k) & 31)) & 1u, pen, inv_temp, cap));
    }
  }
  // tail (hi - lo not a multiple of 4), scalar
  for (int i = lo + 4 * nq + first_quad; i < hi; i += quad_stride)
    f(i, transform(logits[i], seen ? ((seen[i >> 5] >> (i & 31)) & 1u) : 0u, pen, inv_temp, cap));
}

__global__ void __launch_bounds__(SAMP_THREADS, 1) sample_kernel(const SampleParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red_f[32];
  __shared__ int red_i[32];
  __shared__ uint32_t hist_lo[NBINS];     // 32.32 fixed-point mass per bin: native ATOMS.ADD on the
  __shared__ uint32_t hist_hi[NBINS];     // low word, carries (rare) bump the high word
  __shared__ unsigned long long red_u[32];
  __shared__ unsigned long long s_u[2];
  __shared__ float s_bcast[4];
  __shared__ int s_ib[4];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bb = (p.row_base != nullptr ? p.row_base[0] : 0) + b;     // row of the per-sequence state
  const int V = p.vocab;
  const float* logits = p.logits + static_cast<size_t>(b) * p.ld;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  const float cap = p.softcap;
  uint32_t* seen = p.seen ? p.seen + static_cast<size_t>(bb) * ((V + 31) / 32) : nullptr;
  const float temp = p.temperature ? p.temperature[bb] : 0.f;
  const float pen = p.rep_penalty ? p.rep_penalty[bb] : 1.f;
  const float top_p = p.top_p ? p.top_p[bb] : 1.f;
  const bool greedy = !(temp > 0.f);
  const float inv_temp = greedy ? 1.f : 1.f / temp;
  const uint32_t* seen_r = (pen != 1.f) ? seen : nullptr;

  // ---- pass 1: max (+ argmax, lowest id on ties)
  float mx = -INFINITY; int amx = 0;
  visit(logits, seen_r, 0, V, tid, SAMP_THREADS, vec_ok, pen, inv_temp, cap,
        [&](int i, float l) { if (l > mx) { mx = l; amx = i; } });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, amx, o);
    if (om > mx || (om == mx && oi < amx)) { mx = om; amx = oi; }
  }
  if (lane == 0) { red_f[warp] = mx; red_i[warp] = amx; }
  __syncthreads();
  if (warp == 0) {
    mx = red_f[lane]; amx = red_i[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oi = __shfl_xor_sync(0xffffffffu, amx, o);
      if (om > mx || (om == mx && oi < amx)) { mx = om; amx = oi; }
    }
    if (lane == 0) { s_bcast[0] = mx; s_ib[0] = amx; }
  }
  __syncthreads();
  mx = s_bcast[0];
  int token = s_ib[0];

  if (!greedy) {
    auto fx = [](float e) { return static_cast<uint32_t>(fminf(e * 4294967296.0f, 4294967040.0f)); };
    auto hist_add = [&](int bin, uint32_t f) {
      const uint32_t old = atomicAdd(&hist_lo[bin], f);
      if (old + f < old) atomicAdd(&hist_hi[bin], 1u);
    };
    auto hist_get = [&](int bin) { return (static_cast<unsigned long long>(hist_hi[bin]) << 32) | hist_lo[bin]; };
    // top-down scan of the histogram (warp 0): highest bin whose suffix mass (plus carry) reaches `need`;
    // above = mass strictly above that bin, incl = mass including it.
    auto scan_down = [&](unsigned long long carry, unsigned long long need, int& found, unsigned long long& above,
                         unsigned long long& incl) {
      found = -1; above = carry; incl = carry;
      for (int base = NBINS - 32; base >= 0 && found < 0; base -= 32) {
        const unsigned long long v = hist_get(base + (31 - lane));     // lane 0 = highest bin of the chunk
        unsigned long long pre = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long n = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += n; }
        const unsigned ball = __ballot_sync(0xffffffffu, carry + pre >= need);
        if (ball) {
          const int l0 = __ffs(ball) - 1;
          found = base + (31 - l0);
          incl = carry + __shfl_sync(0xffffffffu, pre, l0);
          above = incl - __shfl_sync(0xffffffffu, v, l0);
        } else {
          carry += __shfl_sync(0xffffffffu, pre, 31);
          incl = carry;
        }
      }
    };

    // ---- pass 2: level-1 histogram of e = exp(l - max) in (0, 1]; key = top 12 bits below the sign
    for (int i = tid; i < NBINS; i += SAMP_THREADS) { hist_lo[i] = 0u; hist_hi[i] = 0u; }
    __syncthreads();
    unsigned long long zsum = 0ull;
    visit(logits, seen_r, 0, V, tid, SAMP_THREADS, vec_ok, pen, inv_temp, cap, [&](int, float l) {
      const float e = __expf(l - mx);
      const uint32_t f = fx(e);
      zsum += f;
      hist_add(__float_as_uint(e) >> 19, f);
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
    if (lane == 0) red_u[warp] = zsum;
    __syncthreads();
    if (warp == 0) {
      unsigned long long z = red_u[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
      if (lane == 0) s_u[0] = z;
    }
    __syncthreads();
    const unsigned long long Z = s_u[0];
    const unsigned long long need = static_cast<unsigned long long>(static_cast<double>(top_p) * static_cast<double>(Z));
    if (warp == 0) {
      int found; unsigned long long above, incl;
      scan_down(0ull, need, found, above, incl);
      if (lane == 0) { s_ib[1] = found < 0 ? 0 : found; s_u[1] = found < 0 ? 0ull : above; }
    }
    __syncthreads();
    const int B1 = s_ib[1];
    const unsigned long long above1 = s_u[1];       // mass strictly above the boundary bin

    // ---- pass 3: level-2 histogram inside the boundary bin (next 12 bits)
    for (int i = tid; i < NBINS; i += SAMP_THREADS) { hist_lo[i] = 0u; hist_hi[i] = 0u; }
    __syncthreads();
    visit(logits, seen_r, 0, V, tid, SAMP_THREADS, vec_ok, pen, inv_temp, cap, [&](int, float l) {
      const float e = __expf(l - mx);
      const uint32_t u = __float_as_uint(e);
      if (static_cast<int>(u >> 19) == B1) hist_add((u >> 7) & (NBINS - 1), fx(e));
    });
    __syncthreads();
    if (warp == 0) {
      int found; unsigned long long above, incl;
      scan_down(above1, need, found, above, incl);
      if (found < 0) found = 0;
      if (lane == 0) { s_ib[2] = found; s_bcast[3] = static_cast<float>(static_cast<double>(incl) * (1.0 / 4294967296.0)); }
    }
    __syncthreads();
    const uint32_t thr_bits = (static_cast<uint32_t>(B1) << 19) | (static_cast<uint32_t>(s_ib[2]) << 7);
    const float kept_mass = s_bcast[3];    // total mass of tokens with bits(e) >= thr_bits

    // ---- pass 4: multinomial draw over the kept set, in index order
    const uint32_t stepv = p.step ? *p.step : 0u;
    const uint32_t h = hash_u32((p.seeds ? p.seeds[bb] : 0x1234567u) ^ hash_u32(stepv * 0x9E3779B9u + b));
    const float u01 = (static_cast<float>(h >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float target = u01 * kept_mass;
    // one contiguous chunk per WARP (multiple of 128 ids), lanes own interleaved quads -> coalesced 128-bit reads
    const int cw = ((V + 31) / 32 + 127) & ~127;
    const int c0 = warp * cw, c1 = min(V, c0 + cw);
    float mine = 0.f;
    if (c0 < c1)
      visit(logits, seen_r, c0, c1, lane, 32, vec_ok, pen, inv_temp, cap, [&](int, float l) {
        const float e = __expf(l - mx);
        if (__float_as_uint(e) >= thr_bits) mine += e;
      });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);   // warp total
    if (lane == 0) red_f[warp] = mine;
    if (tid == 0) s_ib[3] = -1;
    __syncthreads();
    float wtot = red_f[lane], wpre = wtot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float n = __shfl_up_sync(0xffffffffu, wpre, o); if (lane >= o) wpre += n; }
    const float my_excl = __shfl_sync(0xffffffffu, wpre - wtot, warp);
    const float my_tot = __shfl_sync(0xffffffffu, wtot, warp);
    if (my_tot > 0.f && target >= my_excl && target < my_excl + my_tot) {
      // the selected warp walks its chunk in rows of 32 consecutive ids with a warp prefix sum
      float run = my_excl;
      int pick = -1, last_kept = -1;
      for (int base = c0; base < c1 && pick < 0; base += 32) {
        const int i = base + lane;
        float v = 0.f;
        if (i < c1) {
          const bool sn = seen_r ? ((seen_r[i >> 5] >> (i & 31)) & 1u) : false;
          const float e = __expf(transform(logits[i], sn, pen, inv_temp, cap) - mx);
          if (__float_as_uint(e) >= thr_bits) v = e;
        }
        float pre = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float n = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += n; }
        const unsigned kept = __ballot_sync(0xffffffffu, v > 0.f);
        if (kept) last_kept = base + (31 - __clz(kept));
        const unsigned hit = __ballot_sync(0xffffffffu, v > 0.f && run + pre > target);
        if (hit) pick = base + (__ffs(hit) - 1);
        run += __shfl_sync(0xffffffffu, pre, 31);
      }
      if (pick < 0) pick = last_kept;          // float round-off at the chunk edge
      if (lane == 0 && pick >= 0) atomicMax(&s_ib[3], pick);
    }
    __syncthreads();
    if (s_ib[3] >= 0) token = s_ib[3];   // else (round-off at the far edge): fall back to argmax
  }

  if (tid == 0) {
    p.out_tokens[bb] = token;
    if (seen != nullptr) atomicOr(&seen[token >> 5], 1u << (token & 31));
    if (p.history != nullptr) {
      const int pos = p.hist_pos[bb];
      if (pos < p.hist_stride) p.history[static_cast<size_t>(bb) * p.hist_stride + pos] = token;
      p.hist_pos_out[bb] = pos + 1;
    }
    if (p.peer_tokens != nullptr && p.peer_tokens != p.out_tokens) p.peer_tokens[bb] = token;
    if (p.signal_flag != nullptr) {
      __threadfence_system();
      const uint32_t prev = atomicAdd(p.done_counter, 1u);
      if (prev == gridDim.x - 1) {
        __threadfence_system();
        *p.done_counter = 0;
        const uint32_t e = *reinterpret_cast<volatile uint32_t*>(p.signal_epoch) + 1;
        *reinterpret_cast<volatile uint32_t*>(p.signal_epoch) = e;
        st_release_sys(p.signal_flag, e);
      }
    }
  }
}

// mark prompt tokens in the seen bitmap: ids [n], seq_of [n]
__global__ void mark_seen_kernel(const int* ids, const int* seq_of, uint32_t* seen, int n, int words, int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = ids[i];
  if (id < 0 || id >= vocab) return;
  atomicOr(&seen[static_cast<size_t>(seq_of[i]) * words + (id >> 5)], 1u << (id & 31));
}

int launch_sample(const float* logits, uint32_t* seen, int* out_tokens, int* peer_tokens, int* history,
                  const int* hist_pos, int* hist_pos_out, int hist_stride, int batch, int vocab, int ld, float softcap,
                  const float* temperature, const float* top_p, const float* rep_penalty, const uint32_t* seeds,
                  const uint32_t* step, uint32_t* signal_flag, uint32_t* signal_epoch, uint32_t* done_counter,
                  const int* row_base, cudaStream_t s) {
  SampleParams p;
  p.logits = logits; p.seen = seen; p.out_tokens = out_tokens; p.peer_tokens = peer_tokens; p.history = history;
  p.hist_pos = hist_pos; p.hist_pos_out = hist_pos_out; p.hist_stride = hist_stride; p.vocab = vocab; p.ld = ld; p.softcap = softcap;
  p.temperature = temperature; p.top_p = top_p; p.rep_penalty = rep_penalty; p.seeds = seeds; p.step = step;
  p.signal_flag = signal_flag; p.signal_epoch = signal_epoch; p.done_counter = done_counter; p.row_base = row_base;
  return static_cast<int>(launch_kernel(sample_kernel, dim3(batch), dim3(SAMP_THREADS), 0, s, 1, p));
}

int launch_mark_seen(const int* ids, const int* seq_of, uint32_t* seen, int n, int vocab, cudaStream_t s) {
  if (n <= 0) return 0;
  return static_cast<int>(launch_kernel(mark_seen_kernel, dim3((n + 255) / 256), dim3(256), 0, s, 1, ids, seq_of, seen, n,
                                        (vocab + 31) / 32, vocab));
}

}  // namespace b2b
