// Flash attention for sm_100a on tcgen05 / TMEM / TMA (head dim 64 or 128, causal or full, bf16).
//
// Replaces the score GEMM + fused_scale_(tril_)softmax(_mask)_dropout + context GEMM sequence of the
// reference (libai/layers/attention.py:211-253): no [b, a, s, s] tensor and no materialised mask.
// Q/K/V are read straight out of the packed QKV projection through strided 4-D TMA tensor maps
// ([B, A, S, D] views), O is written in [B, S, A, D] so the output projection consumes it as-is.
//
// Forward, one CTA per (batch, head, 128-query block), 192 threads:
//   warp 0        TMA producer : Q once, then K/V tiles of 128 keys (double buffered)
//   warp 1        MMA issuer   : S_j = Q·K_jᵀ (TMEM, double buffered)  and  O_j = P_j·V_j (TMEM)
//   warps 2..5    softmax      : thread r owns query row r: tcgen05.ld S → online softmax (exp2) →
//                                P (bf16) to swizzled smem as the A operand of the PV MMA; running
//                                output kept in registers: O ← (O + O_{j-1}) · α_j
// The QKᵀ MMA of block j+1 is issued before the softmax of block j, so tensor-core work hides behind
// the (MUFU-bound) softmax.
#include "common.cuh"

#include <type_traits>

#include <cstdlib>
#include <cstring>

namespace lb {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_THREADS = 192;
constexpr int ATT_BWD_THREADS = 320;
constexpr int ATT_FWD2_THREADS = 320; // forward v2: 8 softmax warps (row pairs exchange their block max through smem)  // TMA + MMA + 8 compute warps (two per TMEM lane quadrant: each takes half of the columns)

template <int D>
struct AttnCfg {
  static constexpr int QK_CHUNKS = D / 64;                 // 64-element (128B) K-chunks of Q and K tiles
  static constexpr int TILE_BYTES = ATT_BN * D * 2;        // one Q / K / V tile
  static constexpr int P_BYTES = ATT_BM * ATT_BN * 2;      // 32 KB
  static constexpr int P_STAGES = (D == 64) ? 2 : 1;
  static constexpr int SMEM_BYTES = TILE_BYTES /*Q*/ + 4 * TILE_BYTES /*K,V x2*/ + P_STAGES * P_BYTES + 1024 + 256;
  // TMEM columns: S0 [0,128) S1 [128,256) O [256, 256+D)
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t O_COL = 256;
};

struct AttnFwdParams {
  __nv_bfloat16* o;   // [B, S, A, D]
  float* lse;         // [B, A, S]
  int B, A, S;
  int causal;
  float scale_log2;   // scale * log2(e)
  float scale;
  const int* kv_lens; // optional [B]: number of valid keys per sample (right-padded batches)
  // additive score bias (v2 kernel): dense bf16 [.., S, S] with (batch, head, query) strides in elements (0 = broadcast;
  // T5 relative-position bias, Swin/BERT style masks) or ALiBi slopes per head (bias = slope * (key - query))
  const __nv_bfloat16* bias;
  long bias_strides[3];
  const float* alibi_slopes;
  // attention dropout on P (v2 kernel): keep <=> philox byte >= drop_thresh; O is scaled by inv_keep
  uint32_t drop_thresh;
  float inv_keep;
  RngArgs rng;
  long long* rng_out;   // [2]: the (seed, offset) pair used, for the backward
};

enum AttnBias : int { BIAS_NONE = 0, BIAS_DENSE = 1, BIAS_ALIBI = 2 };

// 16 random bytes for keys [16 g, 16 g + 16) of query row q_idx of (batch, head) slice bh
LB_DEVICE uint4 attn_dropout_bytes(uint32_t q_idx, uint32_t g, uint32_t bh, unsigned long long seed, unsigned long long offset) {
  return philox4x32_7(q_idx, g, bh ^ (static_cast<uint32_t>(offset >> 32) << 20), static_cast<uint32_t>(offset),
                      static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}
LB_DEVICE uint32_t rnd_byte(const uint4& r, int i) {   // i in [0, 16)
  const uint32_t w = (i < 4) ? r.x : (i < 8 ? r.y : (i < 12 ? r.z : r.w));
  return (w >> ((i & 3) * 8)) & 0xFFu;
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, AttnFwdParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int PST = Cfg::P_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::TILE_BYTES;            // 2 stages
  uint8_t* sV = sK + 2 * Cfg::TILE_BYTES;        // 2 stages
  uint8_t* sP = sV + 2 * Cfg::TILE_BYTES;        // PST stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + PST * Cfg::P_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* k_empty = bars + 3;       // 2
  uint64_t* v_full = bars + 5;        // 2
  uint64_t* v_empty = bars + 7;       // 2
  uint64_t* s_full = bars + 9;        // 2
  uint64_t* s_empty = bars + 11;      // 2
  uint64_t* p_full = bars + 13;       // 2
  uint64_t* p_empty = bars + 15;      // 2
  uint64_t* o_full = bars + 17;       // 1
  uint64_t* o_empty = bars + 18;      // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp_idx = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int q_blk = blockIdx.x;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q0 = q_blk * ATT_BM;
  const int kv_len = p.kv_lens != nullptr ? max(1, min(p.S, p.kv_lens[batch])) : p.S;
  int nkv = (kv_len + ATT_BN - 1) / ATT_BN;
  if (p.causal) nkv = min(nkv, q_blk + 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
      for (int c = 0; c < Cfg::QK_CHUNKS; ++c)
        tma_load_4d(sQ + c * (ATT_BM * 128), &tmap_q, q_full, c * 64, q0, head, batch);
      for (int j = 0; j < nkv; ++j) {
        const int b = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        mbar_wait(&k_empty[b], par);
        mbar_arrive_expect_tx(&k_full[b], Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::QK_CHUNKS; ++c)
          tma_load_4d(sK + b * Cfg::TILE_BYTES + c * (ATT_BN * 128), &tmap_k, &k_full[b], c * 64, j * ATT_BN, head, batch);
        mbar_wait(&v_empty[b], par);
        mbar_arrive_expect_tx(&v_full[b], Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::QK_CHUNKS; ++c)
          tma_load_4d(sV + b * Cfg::TILE_BYTES + c * (ATT_BN * 128), &tmap_v, &v_full[b], c * 64, j * ATT_BN, head, batch);
      }
    }
  } else if (warp_idx == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_qk = make_idesc_bf16(ATT_BM, ATT_BN, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, D, false, true);
    const uint32_t q_addr = smem_u32(sQ);
    auto issue_s = [&](int j) {
      const int b = j & 1;
      mbar_wait(&k_full[b], (j >> 1) & 1);
      mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t k_addr = smem_u32(sK + b * Cfg::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk / 4) * (ATT_BM * 128) + (kk % 4) * 32;
          umma_f16_ss(tmem_base + b * ATT_BN, make_smem_desc_sw128(q_addr + off, 0, 1024),
                      make_smem_desc_sw128(k_addr + off, 0, 1024), idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[b]);
        umma_commit(&k_empty[b]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) issue_s(j + 1);
      const int b = j & 1, pb = j % PST;
      mbar_wait(&p_full[pb], (j / PST) & 1);
      mbar_wait(&v_full[b], (j >> 1) & 1);
      mbar_wait(o_empty, (j & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t p_addr = smem_u32(sP + pb * Cfg::P_BYTES);
        const uint32_t v_addr = smem_u32(sV + b * Cfg::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          const uint64_t da = make_smem_desc_sw128(p_addr + (kk / 4) * (ATT_BM * 128) + (kk % 4) * 32, 0, 1024);
          const uint64_t db = make_smem_desc_sw128(v_addr + kk * (16 * 128), ATT_BN * 128, 1024);
          umma_f16_ss(tmem_base + Cfg::O_COL, da, db, idesc_pv, kk > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&v_empty[b]);
        umma_commit(&p_empty[pb]);
      }
      __syncwarp();
    }
  } else {
    // ================= softmax / output (128 threads, thread <-> query row) =================
    const int quad = warp_idx % 4;
    const int r = quad * 32 + lane;          // row inside the tile
    const int q_idx = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    float o_run[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o_run[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    auto add_o_block = [&](int jprev) {
      mbar_wait(o_full, jprev & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + Cfg::O_COL + c * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_run[c * 32 + i] += __uint_as_float(t[i]);
      }
      tc_fence_before_sync();
      mbar_arrive(o_empty);
    };

    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1, pb = j % PST;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t s_addr = tmem_base + lane_off + b * ATT_BN;
      const int k0 = j * ATT_BN;
      const bool need_mask = (p.causal && j == q_blk) || (k0 + ATT_BN > kv_len);
      // ---- pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(t[i]);
          if (need_mask) {
            const int kidx = k0 + c * 32 + i;
            if (kidx >= kv_len || (p.causal && kidx > q_idx)) x = -INFINITY;
          }
          mx = fmaxf(mx, x);
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m_run - m_use);   // m_run = -inf -> 0
      // ---- wait for the P buffer to be free, then pass 2: p = exp2(s*scale - m), write bf16 to swizzled smem
      mbar_wait(&p_empty[pb], ((j / PST) & 1) ^ 1);
      uint8_t* prow = sP + pb * Cfg::P_BYTES + r * 128;
      float rowsum = 0.f;
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, t);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0 = __uint_as_float(t[i]), x1 = __uint_as_float(t[i + 1]);
          float e0 = fast_exp2(x0 * p.scale_log2 - m_use), e1 = fast_exp2(x1 * p.scale_log2 - m_use);
          if (need_mask) {
            const int kidx = k0 + c * 32 + i;
            if (kidx >= kv_len || (p.causal && kidx > q_idx)) e0 = 0.f;
            if (kidx + 1 >= kv_len || (p.causal && kidx + 1 > q_idx)) e1 = 0.f;
          }
          rowsum += e0 + e1;
          packed[i / 2] = pack_bf16(e0, e1);
        }
        // 32 keys = 64 bytes = four 16-byte chunks; key chunk (c / 2) selects the 64-key half of the tile
        uint8_t* half_base = prow + (c / 2) * (ATT_BM * 128);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk16 = (c % 2) * 4 + q4;               // 16B chunk index inside the 128B row
          const int phys = chunk16 ^ (r & 7);                 // 128B swizzle
          *reinterpret_cast<uint4*>(half_base + phys * 16) =
              make_uint4(packed[q4 * 4], packed[q4 * 4 + 1], packed[q4 * 4 + 2], packed[q4 * 4 + 3]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&s_empty[b]);      // S buffer b may be overwritten by QKᵀ of block j+2
      fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&p_full[pb]);
      l_run = l_run * alpha + rowsum;
      // ---- fold the previous block's P·V into the running output, then rescale to the new max
      if (j > 0) add_o_block(j - 1);
      if (alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < D; ++i) o_run[i] *= alpha;
      }
      m_run = m_new;
    }
    add_o_block(nkv - 1);
    if (q_idx < p.S) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* orow = p.o + ((static_cast<size_t>(batch) * p.S + q_idx) * p.A + head) * D;
#pragma unroll
      for (int i = 0; i < D; i += 8) {
        *reinterpret_cast<uint4*>(orow + i) =
            make_uint4(pack_bf16(o_run[i] * inv, o_run[i + 1] * inv), pack_bf16(o_run[i + 2] * inv, o_run[i + 3] * inv),
                       pack_bf16(o_run[i + 4] * inv, o_run[i + 5] * inv), pack_bf16(o_run[i + 6] * inv, o_run[i + 7] * inv));
      }
      // natural-log LSE of the scaled scores
      p.lse[(static_cast<size_t>(batch) * p.A + head) * p.S + q_idx] =
          (m_run == -INFINITY) ? -INFINITY : (m_run + log2f(l_run)) * 0.6931471805599453f;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace lb


// =================================================================================================
// Forward v2: P and O stay in tensor memory.
//   * softmax threads write P (bf16) back into the S columns with tcgen05.st; the P·V MMA takes its A operand
//     from TMEM (tcgen05.mma [d], [a_tmem], b_desc) – no shared-memory round trip, no proxy fence
//   * O accumulates in TMEM across KV blocks; the running max is only advanced when it grows by more than 2^8
//     ("lazy rescale"), in which case the softmax threads rescale O in place (tcgen05.ld → mul → tcgen05.st)
//   * TMEM footprint 128 + D columns and 80 KB of shared memory at D = 64 → two CTAs per SM, whose
//     softmax / MMA / TMA phases overlap each other
// =================================================================================================
namespace lb {

// named barrier of the two warps that own the same TMEM lane quadrant (immediate ids: a register id would make ptxas
// reserve all 16 barriers of the CTA)
LB_DEVICE void pair_sync(int quad) {
  if (quad == 0) asm volatile("bar.sync 1, 64;\n" ::: "memory");
  else if (quad == 1) asm volatile("bar.sync 2, 64;\n" ::: "memory");
  else if (quad == 2) asm volatile("bar.sync 3, 64;\n" ::: "memory");
  else asm volatile("bar.sync 4, 64;\n" ::: "memory");
}

template <int D>
struct AttnCfg2 {
  static constexpr int QK_CHUNKS = D / 64;
  static constexpr int TILE_BYTES = ATT_BN * D * 2;
  static constexpr int XCHG_BYTES = 2 * 2 * 128 * 4;               // block-max exchange: [parity][half][row]
  static constexpr int SMEM_BYTES = 5 * TILE_BYTES + XCHG_BYTES + 1024 + 256;   // Q + 2 x (K, V)
  static constexpr uint32_t TMEM_COLS = (D == 64) ? 256 : 512;
  // S fp32 [0,128) | O fp32 [128,128+D) | P bf16 (64 columns) behind O: P no longer aliases S, so the two softmax
  // halves never race on it and Q·Kᵀ of the next block may be issued right behind P·V of the current one
  static constexpr uint32_t S_COL = 0, O_COL = 128, P_COL = 128 + D;
  static constexpr int MIN_CTAS = (D == 64) ? 2 : 1;
};

template <int D, int BIAS, bool DROP>
__global__ void __launch_bounds__(ATT_FWD2_THREADS, AttnCfg2<D>::MIN_CTAS)
attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, AttnFwdParams p) {
  using Cfg = AttnCfg2<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::TILE_BYTES;
  uint8_t* sV = sK + 2 * Cfg::TILE_BYTES;
  float* sXchg = reinterpret_cast<float*>(sV + 2 * Cfg::TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * Cfg::TILE_BYTES + Cfg::XCHG_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 11;
  uint64_t* s_free = bars + 12;      // the softmax warps hold S(j) in registers: Q·Kᵀ(j+1) may overwrite the columns
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp_idx = threadIdx.x / 32, lane = threadIdx.x % 32;
  // heavy (late) query blocks first: better tail behaviour under the causal mask
  const int q_blk = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q0 = q_blk * ATT_BM;
  const int kv_len = p.kv_lens != nullptr ? max(1, min(p.S, p.kv_lens[batch])) : p.S;
  int nkv = (kv_len + ATT_BN - 1) / ATT_BN;
  if (p.causal) nkv = min(nkv, q_blk + 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);        // one arrival per softmax warp
    mbar_init(o_full, 1);
    mbar_init(s_free, 8);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp_idx == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
      for (int c = 0; c < Cfg::QK_CHUNKS; ++c)
        tma_load_4d(sQ + c * (ATT_BM * 128), &tmap_q, q_full, c * 64, q0, head, batch);
      for (int j = 0; j < nkv; ++j) {
        const int b = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        mbar_wait(&k_empty[b], par);
        mbar_arrive_expect_tx(&k_full[b], Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::QK_CHUNKS; ++c)
          tma_load_4d(sK + b * Cfg::TILE_BYTES + c * (ATT_BN * 128), &tmap_k, &k_full[b], c * 64, j * ATT_BN, head, batch);
        mbar_wait(&v_empty[b], par);
        mbar_arrive_expect_tx(&v_full[b], Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::QK_CHUNKS; ++c)
          tma_load_4d(sV + b * Cfg::TILE_BYTES + c * (ATT_BN * 128), &tmap_v, &v_full[b], c * 64, j * ATT_BN, head, batch);
      }
    }
  } else if (warp_idx == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(ATT_BM, ATT_BN, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, D, false, true);
    const uint32_t q_addr = smem_u32(sQ);
    mbar_wait(q_full, 0);
    // S(j+1) = Q·K(j+1)ᵀ is issued as soon as the softmax warps have read S(j) for the last time (s_free), i.e. under the
    // exponentials / P store of block j and ahead of P(j)·V(j): the next block's scores are ready when the softmax warps
    // come back for them
    auto issue_qk = [&](int j) {
      const int b = j & 1;
      mbar_wait(&k_full[b], (j >> 1) & 1);
      if (j > 0) mbar_wait(s_free, (j - 1) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t k_addr = smem_u32(sK + b * Cfg::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk / 4) * (ATT_BM * 128) + (kk % 4) * 32;
          umma_f16_ss(tmem_base + Cfg::S_COL, make_smem_desc_sw128(q_addr + off, 0, 1024),
                      make_smem_desc_sw128(k_addr + off, 0, 1024), idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(&k_empty[b]);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1;
      if (j + 1 < nkv) issue_qk(j + 1);
      mbar_wait(p_full, j & 1);
      mbar_wait(&v_full[b], (j >> 1) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t v_addr = smem_u32(sV + b * Cfg::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          // A = P from tensor memory: 16 bf16 per row = 8 columns per k-step
          umma_f16_ts(tmem_base + Cfg::O_COL, tmem_base + Cfg::P_COL + kk * 8,
                      make_smem_desc_sw128(v_addr + kk * (16 * 128), ATT_BN * 128, 1024), idesc_pv,
                      (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&v_empty[b]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp_idx % 4;
    const int half = (warp_idx - 2) / 4;        // this warp's 64 of the 128 key columns
    const int r = quad * 32 + lane;
    const int q_idx = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + Cfg::S_COL + half * 64;
    const uint32_t p_addr = tmem_base + lane_off + Cfg::P_COL + half * 32;
    const uint32_t o_addr = tmem_base + lane_off + Cfg::O_COL + half * (D / 2);
    float m_run = -INFINITY, l_run = 0.f;       // l_run: partial row sum over this half's columns
    const int key_lim = p.causal ? min(kv_len, q_idx + 1) : kv_len;   // keys [0, key_lim) are visible to this query row
    constexpr float LOG2E = 1.4426950408889634f;
    unsigned long long rng_seed = 0, rng_offset = 0;
    if constexpr (DROP) {
      rng_resolve(p.rng, rng_seed, rng_offset);
      if (p.rng_out != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64) {
        p.rng_out[0] = static_cast<long long>(rng_seed);
        p.rng_out[1] = static_cast<long long>(rng_offset);
      }
    }
    const uint32_t bh = static_cast<uint32_t>(batch * p.A + head);
    const __nv_bfloat16* bias_row = nullptr;
    float slope2 = 0.f;
    if constexpr (BIAS == BIAS_DENSE)
      bias_row = p.bias + batch * p.bias_strides[0] + head * p.bias_strides[1] +
                 static_cast<long>(min(q_idx, p.S - 1)) * p.bias_strides[2];
    if constexpr (BIAS == BIAS_ALIBI) slope2 = p.alibi_slopes[head] * LOG2E;
    // score in the log2 domain: s * scale_log2 (+ bias * log2 e)
    auto bias2 = [&](const uint4 (&bq)[4], int i, int kidx) -> float {
      if constexpr (BIAS == BIAS_DENSE) {
        const uint32_t w = (&bq[i >> 3].x)[(i & 7) >> 1];
        const float2 f = unpack_bf16(w);
        return ((i & 1) ? f.y : f.x) * LOG2E;
      } else if constexpr (BIAS == BIAS_ALIBI) {
        return slope2 * static_cast<float>(kidx - q_idx);
      } else {
        return 0.f;
      }
    };
    auto load_bias = [&](uint4 (&bq)[4], int kbase) {
      if constexpr (BIAS == BIAS_DENSE) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          bq[v] = (kbase + v * 8 < p.S) ? *reinterpret_cast<const uint4*>(bias_row + kbase + v * 8) : make_uint4(0, 0, 0, 0);
      }
    };

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      const int k0 = j * ATT_BN + half * 64;
      const bool need_mask = (p.causal && j == q_blk) || (j * ATT_BN + ATT_BN > kv_len);
      // ---- pass 1: max over this half's columns, then exchange with the partner thread of the same row
      float mx = -INFINITY;
      auto pass1 = [&](auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t t[32];
          uint4 bq[4];
          tmem_ld_32x32b_x32(s_addr + c * 32, t);
          load_bias(bq, k0 + c * 32);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0 = __uint_as_float(t[i]), x1 = __uint_as_float(t[i + 1]);
            if constexpr (BIAS != BIAS_NONE) {   // with a bias the max is taken in the log2 domain
              x0 = fmaf(x0, p.scale_log2, bias2(bq, i, k0 + c * 32 + i));
              x1 = fmaf(x1, p.scale_log2, bias2(bq, i + 1, k0 + c * 32 + i + 1));
            }
            if constexpr (MASK) {
              const int kidx = k0 + c * 32 + i;
              x0 = kidx < key_lim ? x0 : -INFINITY;       // branch-free: one compare + select per element
              x1 = kidx + 1 < key_lim ? x1 : -INFINITY;
            }
            mx = fmaxf(mx, fmaxf(x0, x1));
          }
        }
      };
      if (need_mask) pass1(std::true_type{}); else pass1(std::false_type{});
      float* xch = sXchg + (j & 1) * 256;
      xch[half * 128 + r] = mx;
      pair_sync(quad);   // only the two warps that share these 32 rows
      mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
      // ---- lazy rescale: only move the reference max when it grows by more than 2^8
      const float m_cand = fmaxf(m_run, BIAS != BIAS_NONE ? mx : mx * p.scale_log2);
      const bool bump = (m_cand - m_run > 8.0f) || (m_run == -INFINITY && m_cand != -INFINITY);
      float alpha = 1.0f;
      if (bump) {
        alpha = fast_exp2(m_run - m_cand);   // 0 on the first block
        m_run = m_cand;
      }
      const bool bump_any = __any_sync(0xffffffffu, bump && j > 0);
      const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
      // ---- pass 2: P = exp2(s * scale - m) as bf16 into the P columns
      float rowsum = 0.f;
      auto pass2 = [&](auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t t[32];
          uint4 bq[4];
          tmem_ld_32x32b_x32(s_addr + c * 32, t);
          load_bias(bq, k0 + c * 32);
          uint4 rnd[2];
          if constexpr (DROP) {   // 32 keys = two 16-byte Philox results, generated while the TMEM load is in flight
            rnd[0] = attn_dropout_bytes(static_cast<uint32_t>(q_idx), static_cast<uint32_t>((k0 + c * 32) >> 4), bh, rng_seed, rng_offset);
            rnd[1] = attn_dropout_bytes(static_cast<uint32_t>(q_idx), static_cast<uint32_t>((k0 + c * 32) >> 4) + 1u, bh, rng_seed, rng_offset);
          }
          tmem_ld_wait();
          if (c == 1) {   // S(j) has been read for the last time
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
          }
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1;
            if constexpr (BIAS != BIAS_NONE) {
              e0 = fast_exp2(fmaf(__uint_as_float(t[i]), p.scale_log2, bias2(bq, i, k0 + c * 32 + i)) - m_use);
              e1 = fast_exp2(fmaf(__uint_as_float(t[i + 1]), p.scale_log2, bias2(bq, i + 1, k0 + c * 32 + i + 1)) - m_use);
            } else {
              e0 = fast_exp2(__uint_as_float(t[i]) * p.scale_log2 - m_use);
              e1 = fast_exp2(__uint_as_float(t[i + 1]) * p.scale_log2 - m_use);
            }
            if constexpr (MASK) {
              const int kidx = k0 + c * 32 + i;
              e0 = kidx < key_lim ? e0 : 0.f;
              e1 = kidx + 1 < key_lim ? e1 : 0.f;
            }
            rs0 += e0;   // the softmax normaliser is the sum of the UNdropped probabilities
            rs1 += e1;
            if constexpr (DROP) {
              if (rnd_byte(rnd[i >> 4], i & 15) < p.drop_thresh) e0 = 0.f;
              if (rnd_byte(rnd[i >> 4], (i & 15) + 1) < p.drop_thresh) e1 = 0.f;
            }
            packed[i / 2] = pack_bf16(e0, e1);
          }
          tmem_st_32x32b_x16(p_addr + c * 16, packed);
        }
        rowsum = rs0 + rs1;
      };
      if (need_mask) pass2(std::true_type{}); else pass2(std::false_type{});
      l_run = l_run * alpha + rowsum;
      // ---- O *= alpha (rare): wait until P·V of block j-1 has retired, then rescale this half's O columns
      if (bump_any) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
          uint32_t t[32];
          tmem_ld_32x32b_x32(o_addr + c * 32, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
          tmem_st_32x32b_x32(o_addr + c * 32, t);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: combine the two partial row sums, O / l  →  [B, S, A, D]
    float* xch = sXchg + (nkv & 1) * 256;
    xch[half * 128 + r] = l_run;
    pair_sync(quad);
    l_run += xch[(half ^ 1) * 128 + r];
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after_sync();
    const float inv = l_run > 0.f ? (DROP ? p.inv_keep : 1.0f) / l_run : 0.f;
    __nv_bfloat16* orow = p.o + ((static_cast<size_t>(batch) * p.S + q_idx) * p.A + head) * D + half * (D / 2);
#pragma unroll
    for (int c = 0; c < D / 64; ++c) {
      uint32_t t[32];
      tmem_ld_32x32b_x32(o_addr + c * 32, t);
      tmem_ld_wait();
      if (q_idx < p.S) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          *reinterpret_cast<uint4*>(orow + c * 32 + i) = make_uint4(
              pack_bf16(__uint_as_float(t[i]) * inv, __uint_as_float(t[i + 1]) * inv),
              pack_bf16(__uint_as_float(t[i + 2]) * inv, __uint_as_float(t[i + 3]) * inv),
              pack_bf16(__uint_as_float(t[i + 4]) * inv, __uint_as_float(t[i + 5]) * inv),
              pack_bf16(__uint_as_float(t[i + 6]) * inv, __uint_as_float(t[i + 7]) * inv));
        }
      }
    }
    if (half == 0 && q_idx < p.S)
      p.lse[(static_cast<size_t>(batch) * p.A + head) * p.S + q_idx] =
          (m_run == -INFINITY) ? -INFINITY : (m_run + log2f(l_run)) * 0.6931471805599453f;
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace lb

namespace {
bool make_qkv_tmap(CUtensorMap* m, const void* ptr, int B, int A, int S, int D, const long* st) {
  // dims innermost first: {D, S, A, B}; st = {batch, head, seq} strides in elements
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)S, (uint64_t)A, (uint64_t)B};
  uint64_t strides[4] = {2, (uint64_t)st[2] * 2, (uint64_t)st[1] * 2, (uint64_t)st[0] * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return lb_host::make_tmap_bf16(m, ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int D, int BIAS, bool DROP>
cudaError_t launch_fwd_v2_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                            const lb::AttnFwdParams& p, cudaStream_t s) {
  using Cfg = lb::AttnCfg2<D>;
  auto kern = lb::attn_fwd_v2_kernel<D, BIAS, DROP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.S + lb::ATT_BM - 1) / lb::ATT_BM, p.A, p.B);
  kern<<<grid, lb::ATT_FWD2_THREADS, Cfg::SMEM_BYTES, s>>>(tq, tk, tv, p);
  return cudaGetLastError();
}

template <int D>
cudaError_t launch_fwd_v2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                          const lb::AttnFwdParams& p, cudaStream_t s) {
  const int bias = p.bias != nullptr ? lb::BIAS_DENSE : (p.alibi_slopes != nullptr ? lb::BIAS_ALIBI : lb::BIAS_NONE);
  const bool drop = p.drop_thresh > 0;
  if (bias == lb::BIAS_NONE) return drop ? launch_fwd_v2_t<D, lb::BIAS_NONE, true>(tq, tk, tv, p, s) : launch_fwd_v2_t<D, lb::BIAS_NONE, false>(tq, tk, tv, p, s);
  if (bias == lb::BIAS_DENSE) return drop ? launch_fwd_v2_t<D, lb::BIAS_DENSE, true>(tq, tk, tv, p, s) : launch_fwd_v2_t<D, lb::BIAS_DENSE, false>(tq, tk, tv, p, s);
  return drop ? launch_fwd_v2_t<D, lb::BIAS_ALIBI, true>(tq, tk, tv, p, s) : launch_fwd_v2_t<D, lb::BIAS_ALIBI, false>(tq, tk, tv, p, s);
}

int attn_fwd_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LIBAI_B200_ATTN_FWD");
    v = (e != nullptr && e[0] == '1') ? 1 : 2;
  }
  return v;
}

template <int D>
cudaError_t launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const lb::AttnFwdParams& p,
                       cudaStream_t s) {
  if (attn_fwd_version() == 2 || p.bias != nullptr || p.alibi_slopes != nullptr || p.drop_thresh > 0)
    return launch_fwd_v2<D>(tq, tk, tv, p, s);   // (bias / dropout exist in the v2 kernel only)
  using Cfg = lb::AttnCfg<D>;
  auto kern = lb::attn_fwd_kernel<D>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.S + lb::ATT_BM - 1) / lb::ATT_BM, p.A, p.B);
  kern<<<grid, lb::ATT_THREADS, Cfg::SMEM_BYTES, s>>>(tq, tk, tv, p);
  return cudaGetLastError();
}
}  // namespace

// p_drop in [0, 1): attention dropout probability (quantised to 1/256 like the byte-threshold test of the kernel);
// bias (optional, bf16): dense additive score bias with (batch, head, query-row) strides in elements, key stride 1;
// alibi_slopes (optional, fp32 [A]).
extern "C" int lb_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int A, int S, int D,
                           const long* q_strides, const long* k_strides, const long* v_strides, int causal, float scale,
                           const int* kv_lens, const void* bias, const long* bias_strides, const float* alibi_slopes,
                           float p_drop, const lb::RngArgs* rng, long long* rng_out, cudaStream_t s) {
  if (D != 64 && D != 128) return -1;
  for (int i = 0; i < 3; ++i)
    if ((q_strides[i] % 8) || (k_strides[i] % 8) || (v_strides[i] % 8)) return -3;
  CUtensorMap tq, tk, tv;
  if (!make_qkv_tmap(&tq, q, B, A, S, D, q_strides)) return -2;
  if (!make_qkv_tmap(&tk, k, B, A, S, D, k_strides)) return -2;
  if (!make_qkv_tmap(&tv, v, B, A, S, D, v_strides)) return -2;
  lb::AttnFwdParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.B = B;
  p.A = A;
  p.S = S;
  p.causal = causal;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.kv_lens = kv_lens;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  for (int i = 0; i < 3; ++i) p.bias_strides[i] = (bias != nullptr && bias_strides != nullptr) ? bias_strides[i] : 0;
  if (bias != nullptr && ((p.bias_strides[0] | p.bias_strides[1] | p.bias_strides[2]) % 8 || (reinterpret_cast<uintptr_t>(bias) & 15)))
    return -3;
  p.alibi_slopes = bias != nullptr ? nullptr : alibi_slopes;
  uint32_t thr = (uint32_t)(p_drop * 256.0f + 0.5f);
  if (thr > 255u) thr = 255u;
  p.drop_thresh = (p_drop > 0.f && rng != nullptr) ? (thr == 0 ? 1u : thr) : 0u;
  p.inv_keep = 256.0f / (256.0f - (float)p.drop_thresh);
  memset(&p.rng, 0, sizeof(p.rng));
  if (rng != nullptr) p.rng = *rng;
  p.rng_out = rng_out;
  cudaError_t e = (D == 64) ? launch_fwd<64>(tq, tk, tv, p, s) : launch_fwd<128>(tq, tk, tv, p, s);
  return (int)e;
}

// =================================================================================================
// Backward.  One CTA per (batch, head, 128-key block); loops over the query blocks that attend to it.
//   S  = Q_i·Kᵀ , dP = dO_i·Vᵀ              (tcgen05, TMEM)
//   P  = exp(S·scale − lse) , dS = P ∘ (dP − δ) · scale     (softmax threads → bf16 tiles in smem)
//   dV += Pᵀ·dO_i , dK += dSᵀ·Q_i           (tcgen05, accumulators stay in TMEM for the whole CTA;
//                                            Pᵀ / dSᵀ are the same smem tiles read as MN-major A operands)
//   dQ_i += dS·K                            (tcgen05 → TMEM → fp32 red.add into dq_accum)
// δ = rowsum(dO ∘ O) is produced by attn_delta_kernel, dq_accum is converted to bf16 afterwards.
// =================================================================================================
namespace lb {

template <int D>
struct AttnBwdCfg {
  static constexpr int CHUNKS = D / 64;
  static constexpr int TILE_BYTES = 128 * D * 2;
  static constexpr int PS_BYTES = 128 * 128 * 2;
  static constexpr int QST = (D == 64) ? 2 : 1;
  static constexpr bool DQ_BULK = (D == 64);            // dQ via swizzled smem staging + TMA tensor reduce-add (fp32)
  // staging: per warp D/32 boxes of [32 rows x 32 floats] (128-byte rows, 128B swizzle = the layout of the fp32
  // dq_accum tensor map) -> one cp.reduce.async.bulk.tensor per box instead of one bulk reduce per thread
  static constexpr int DQ_BOX_BYTES = 32 * 128;
  static constexpr int DQ_STAGE_BYTES = DQ_BULK ? 4 * (D / 32) * DQ_BOX_BYTES : 0;
  static constexpr int SMEM_BYTES = 2 * TILE_BYTES /*K,V*/ + 2 * QST * TILE_BYTES /*Q,dO*/ + 2 * PS_BYTES + DQ_STAGE_BYTES + 1024 + 256;
  static constexpr uint32_t S_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 256 + D;
  static constexpr uint32_t DQ_COL = (D == 64) ? 384 : 0;  // D = 128: dQ aliases the S columns
};

struct AttnBwdParams {
  const float* lse;     // [B, A, S]
  const float* delta;   // [B, A, S]
  float* dq_accum;      // [B, A, S, D] fp32
  __nv_bfloat16* dk;    // strided [B, A, S, D] views
  __nv_bfloat16* dv;
  long g_strides[3];    // batch, head, seq strides (elements) of dk / dv
  int B, A, S;
  int causal;
  float scale, scale_log2;
  const int* kv_lens;   // optional [B]
  // same additive bias / dropout description as the forward; `rng_state` = the (seed, offset) pair the forward used
  const __nv_bfloat16* bias;
  long bias_strides[3];
  const float* alibi_slopes;
  float* dbias;         // optional fp32, same strides as `bias`: += P ∘ (dP − δ) (gradient of a learned bias)
  uint32_t drop_thresh;
  float inv_keep;
  const long long* rng_state;
  int debug_skip_dq;    // timing experiments only (LIBAI_B200_ATTN_DEBUG_SKIP_DQ=1): do not issue the dQ reductions
};

template <int D, int BIAS, bool DROP>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, AttnBwdParams p) {
  using Cfg = AttnBwdCfg<D>;
  constexpr int QST = Cfg::QST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::TILE_BYTES;
  uint8_t* sQ = sV + Cfg::TILE_BYTES;                 // QST stages
  uint8_t* sDO = sQ + QST * Cfg::TILE_BYTES;          // QST stages
  uint8_t* sP = sDO + QST * Cfg::TILE_BYTES;
  uint8_t* sDS = sP + Cfg::PS_BYTES;
  uint8_t* sDQ = sDS + Cfg::PS_BYTES;  // 1024-byte aligned (all tiles before it are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + Cfg::PS_BYTES + Cfg::DQ_STAGE_BYTES);
  uint64_t* kv_full = bars;          // 1
  uint64_t* qdo_full = bars + 1;     // 2
  uint64_t* qdo_empty = bars + 3;    // 2
  uint64_t* sdp_full = bars + 5;     // 1
  uint64_t* pds_full = bars + 6;     // 1
  uint64_t* pds_empty = bars + 7;    // 1
  uint64_t* dq_full = bars + 8;      // 1
  uint64_t* dq_empty = bars + 9;     // 1
  uint64_t* acc_full = bars + 10;    // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int kv_blk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
  const int k0 = kv_blk * 128;
  const int kv_len = p.kv_lens != nullptr ? max(1, min(p.S, p.kv_lens[batch])) : p.S;
  const int nq = (p.S + 127) / 128;
  const int i_begin = p.causal ? kv_blk : 0;
  const int iters = nq - i_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);     // one arrival per softmax warp
    mbar_init(pds_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    mbar_init(acc_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp_idx == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
      for (int c = 0; c < Cfg::CHUNKS; ++c) {
        tma_load_4d(sK + c * (128 * 128), &tmap_k, kv_full, c * 64, k0, head, batch);
        tma_load_4d(sV + c * (128 * 128), &tmap_v, kv_full, c * 64, k0, head, batch);
      }
      for (int it = 0; it < iters; ++it) {
        const int st = it % QST;
        const int q0 = (i_begin + it) * 128;
        mbar_wait(&qdo_empty[st], ((it / QST) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::CHUNKS; ++c) {
          tma_load_4d(sQ + st * Cfg::TILE_BYTES + c * (128 * 128), &tmap_q, &qdo_full[st], c * 64, q0, head, batch);
          tma_load_4d(sDO + st * Cfg::TILE_BYTES + c * (128 * 128), &tmap_do, &qdo_full[st], c * 64, q0, head, batch);
        }
      }
    }
  } else if (warp_idx == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);   // Q·Kᵀ, dO·Vᵀ
    constexpr uint32_t idesc_t = make_idesc_bf16(128, D, true, true);       // Pᵀ·dO, dSᵀ·Q
    constexpr uint32_t idesc_q = make_idesc_bf16(128, D, false, true);      // dS·K
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
    mbar_wait(kv_full, 0);
    // S / dP of query block `it`; issued one iteration ahead (right behind the dV/dK/dQ MMAs of block it-1) so the
    // tensor core computes them while the softmax warps are still draining dQ of the previous block
    auto issue_s_dp = [&](int it) {
      const int st = it % QST;
      mbar_wait(&qdo_full[st], (it / QST) & 1);
      if (D == 128) mbar_wait(dq_empty, (it & 1) ^ 1);   // dQ aliases S
      tc_fence_after_sync();
      const uint32_t q_addr = smem_u32(sQ + st * Cfg::TILE_BYTES);
      const uint32_t do_addr = smem_u32(sDO + st * Cfg::TILE_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk / 4) * (128 * 128) + (kk % 4) * 32;
          umma_f16_ss(tmem_base + Cfg::S_COL, make_smem_desc_sw128(q_addr + off, 0, 1024),
                      make_smem_desc_sw128(k_addr + off, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk / 4) * (128 * 128) + (kk % 4) * 32;
          umma_f16_ss(tmem_base + Cfg::DP_COL, make_smem_desc_sw128(do_addr + off, 0, 1024),
                      make_smem_desc_sw128(v_addr + off, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    issue_s_dp(0);
    for (int it = 0; it < iters; ++it) {
      const int st = it % QST;
      const uint32_t q_addr = smem_u32(sQ + st * Cfg::TILE_BYTES);
      const uint32_t do_addr = smem_u32(sDO + st * Cfg::TILE_BYTES);
      mbar_wait(pds_full, it & 1);   // P / dS tiles written; the softmax warps are done reading S / dP
      if (D == 64) mbar_wait(dq_empty, (it & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one()) {
        // K dimension of these three GEMMs = 128 (queries for dV/dK, keys for dQ): 8 steps of 16
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a_pt = make_smem_desc_sw128(p_addr + kk * 2048, 128 * 128, 1024);       // Pᵀ  (MN-major A)
          const uint64_t b_do = make_smem_desc_sw128(do_addr + kk * 2048, 128 * 128, 1024);      // dO  (MN-major B)
          umma_f16_ss(tmem_base + Cfg::DV_COL, a_pt, b_do, idesc_t, (it > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a_dst = make_smem_desc_sw128(ds_addr + kk * 2048, 128 * 128, 1024);     // dSᵀ (MN-major A)
          const uint64_t b_q = make_smem_desc_sw128(q_addr + kk * 2048, 128 * 128, 1024);        // Q   (MN-major B)
          umma_f16_ss(tmem_base + Cfg::DK_COL, a_dst, b_q, idesc_t, (it > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a_ds = make_smem_desc_sw128(ds_addr + (kk / 4) * (128 * 128) + (kk % 4) * 32, 0, 1024);  // dS (K-major A)
          const uint64_t b_k = make_smem_desc_sw128(k_addr + kk * 2048, 128 * 128, 1024);                          // K  (MN-major B)
          umma_f16_ss(tmem_base + Cfg::DQ_COL, a_ds, b_k, idesc_q, kk > 0 ? 1u : 0u);
        }
        umma_commit(dq_full);
        umma_commit(&qdo_empty[st]);
        umma_commit(pds_empty);
        if (it == iters - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (it + 1 < iters) issue_s_dp(it + 1);
    }
  } else {
    const int quad = warp_idx % 4;
    const int half = (warp_idx - 2) / 4;   // column half of S / dP (and of dQ, dK, dV) handled by this warp
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const size_t bh = static_cast<size_t>(batch) * p.A + head;
    constexpr float LOG2E = 1.4426950408889634f;
    unsigned long long rng_seed = 0, rng_offset = 0;
    if constexpr (DROP) {
      rng_seed = static_cast<unsigned long long>(p.rng_state[0]);
      rng_offset = static_cast<unsigned long long>(p.rng_state[1]);
    }
    float slope2 = 0.f;
    if constexpr (BIAS == BIAS_ALIBI) slope2 = p.alibi_slopes[head] * LOG2E;
    const float inv_scale = 1.0f / p.scale;
    for (int it = 0; it < iters; ++it) {
      const int q_blk = i_begin + it;
      const int q_idx = q_blk * 128 + r;
      const bool q_ok = q_idx < p.S;
      const long bias_off = batch * p.bias_strides[0] + head * p.bias_strides[1] +
                            static_cast<long>(min(q_idx, p.S - 1)) * p.bias_strides[2];
      const float lse2 = q_ok ? p.lse[bh * p.S + q_idx] * 1.4426950408889634f : 0.f;
      const float delta = q_ok ? p.delta[bh * p.S + q_idx] : 0.f;
      const bool need_mask = (p.causal && q_blk == kv_blk) || (k0 + 128 > kv_len) || (q_blk * 128 + 128 > p.S);
      const int key_lim = !q_ok ? 0 : (p.causal ? min(kv_len, q_idx + 1) : kv_len);   // keys [0, key_lim) are visible
      mbar_wait(sdp_full, it & 1);
      tc_fence_after_sync();
      mbar_wait(pds_empty, (it & 1) ^ 1);
      uint8_t* prow = sP + r * 128;
      uint8_t* dsrow = sDS + r * 128;
      const float delta_s = delta * p.scale;
      auto pds = [&](auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll
        for (int c = half * 2; c < half * 2 + 2; ++c) {
          uint32_t ts[32], td[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + Cfg::S_COL + c * 32, ts);
          tmem_ld_32x32b_x32(tmem_base + lane_off + Cfg::DP_COL + c * 32, td);
          uint4 bq[4];
          if constexpr (BIAS == BIAS_DENSE) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int kb = k0 + c * 32 + v * 8;
              bq[v] = kb < p.S ? *reinterpret_cast<const uint4*>(p.bias + bias_off + kb) : make_uint4(0, 0, 0, 0);
            }
          }
          uint4 rnd[2];
          if constexpr (DROP) {
            rnd[0] = attn_dropout_bytes(static_cast<uint32_t>(q_idx), static_cast<uint32_t>((k0 + c * 32) >> 4), static_cast<uint32_t>(bh), rng_seed, rng_offset);
            rnd[1] = attn_dropout_bytes(static_cast<uint32_t>(q_idx), static_cast<uint32_t>((k0 + c * 32) >> 4) + 1u, static_cast<uint32_t>(bh), rng_seed, rng_offset);
          }
          tmem_ld_wait();
          uint32_t pp[16], dd[16];
          float dsv[32];   // only materialised for the dbias path
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0, p1;
            if constexpr (BIAS == BIAS_DENSE) {
              const float2 bf = unpack_bf16((&bq[i >> 3].x)[(i & 7) >> 1]);
              p0 = fast_exp2(fmaf(__uint_as_float(ts[i]), p.scale_log2, bf.x * LOG2E) - lse2);
              p1 = fast_exp2(fmaf(__uint_as_float(ts[i + 1]), p.scale_log2, bf.y * LOG2E) - lse2);
            } else if constexpr (BIAS == BIAS_ALIBI) {
              const int kidx = k0 + c * 32 + i;
              p0 = fast_exp2(fmaf(__uint_as_float(ts[i]), p.scale_log2, slope2 * static_cast<float>(kidx - q_idx)) - lse2);
              p1 = fast_exp2(fmaf(__uint_as_float(ts[i + 1]), p.scale_log2, slope2 * static_cast<float>(kidx + 1 - q_idx)) - lse2);
            } else {
              p0 = fast_exp2(__uint_as_float(ts[i]) * p.scale_log2 - lse2);
              p1 = fast_exp2(__uint_as_float(ts[i + 1]) * p.scale_log2 - lse2);
            }
            if constexpr (MASK) {
              const int kidx = k0 + c * 32 + i;
              p0 = kidx < key_lim ? p0 : 0.f;           // branch-free: one compare + select per element
              p1 = kidx + 1 < key_lim ? p1 : 0.f;
            }
            float g0 = __uint_as_float(td[i]), g1 = __uint_as_float(td[i + 1]);   // dP w.r.t. the dropped, rescaled P
            float pd0 = p0, pd1 = p1;                                               // what multiplied V in the forward
            if constexpr (DROP) {
              const bool keep0 = rnd_byte(rnd[i >> 4], i & 15) >= p.drop_thresh;
              const bool keep1 = rnd_byte(rnd[i >> 4], (i & 15) + 1) >= p.drop_thresh;
              g0 = keep0 ? g0 * p.inv_keep : 0.f;
              g1 = keep1 ? g1 * p.inv_keep : 0.f;
              pd0 = keep0 ? p0 * p.inv_keep : 0.f;
              pd1 = keep1 ? p1 * p.inv_keep : 0.f;
            }
            const float d0 = p0 * fmaf(g0, p.scale, -delta_s);
            const float d1 = p1 * fmaf(g1, p.scale, -delta_s);
            pp[i / 2] = pack_bf16(pd0, pd1);
            dd[i / 2] = pack_bf16(d0, d1);
            if constexpr (BIAS == BIAS_DENSE) {
              dsv[i] = d0 * inv_scale;
              dsv[i + 1] = d1 * inv_scale;
            }
          }
          if constexpr (BIAS == BIAS_DENSE) {
            if (p.dbias != nullptr && q_ok) {   // d bias = P ∘ (dP − δ): fp32 reductions (broadcast dims sum over CTAs)
              float* drow = p.dbias + bias_off + k0 + c * 32;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                if (k0 + c * 32 + i < p.S)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(drow + i), "f"(dsv[i]), "f"(dsv[i + 1]),
                               "f"(dsv[i + 2]), "f"(dsv[i + 3])
                               : "memory");
              }
            }
          }
          const int half_off = (c / 2) * (128 * 128);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int phys = ((c % 2) * 4 + q4) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + half_off + phys * 16) =
                make_uint4(pp[q4 * 4], pp[q4 * 4 + 1], pp[q4 * 4 + 2], pp[q4 * 4 + 3]);
            *reinterpret_cast<uint4*>(dsrow + half_off + phys * 16) =
                make_uint4(dd[q4 * 4], dd[q4 * 4 + 1], dd[q4 * 4 + 2], dd[q4 * 4 + 3]);
          }
        }
      };
      if (need_mask) pds(std::true_type{}); else pds(std::false_type{});
      tc_fence_before_sync();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      // ---- dQ block: TMEM -> fp32 atomics
      mbar_wait(dq_full, it & 1);
      tc_fence_after_sync();
      if constexpr (Cfg::DQ_BULK) {
        // dQ rows of this warp -> swizzled [32 x 128 B] boxes in shared memory -> one TMA tensor reduce-add per box
        // into the fp32 dq_accum (rows beyond S are clipped by the tensor map; they hold zeros anyway)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // previous boxes consumed
        __syncwarp();
        uint8_t* wbox = sDQ + quad * (D / 32) * Cfg::DQ_BOX_BYTES;
        constexpr int DQ_PER_HALF = D / 64;  // 32-column chunks per warp
#pragma unroll
        for (int c = half * DQ_PER_HALF; c < (half + 1) * DQ_PER_HALF; ++c) {
          uint32_t t[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + Cfg::DQ_COL + c * 32, t);
          tmem_ld_wait();
          uint8_t* brow = wbox + c * Cfg::DQ_BOX_BYTES + lane * 128;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(brow + ((k ^ (lane & 7)) * 16)) =
                make_float4(__uint_as_float(t[k * 4]), __uint_as_float(t[k * 4 + 1]), __uint_as_float(t[k * 4 + 2]),
                            __uint_as_float(t[k * 4 + 3]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int c = half * DQ_PER_HALF; c < (half + 1) * DQ_PER_HALF; ++c) {
            asm volatile(
                "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n" ::"l"(
                    reinterpret_cast<uint64_t>(&tmap_dq)),
                "r"(c * 32), "r"(q_blk * 128 + quad * 32), "r"(static_cast<int>(bh)),
                "r"(smem_u32(wbox + c * Cfg::DQ_BOX_BYTES))
                : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
      } else {
      float* dq_row = p.dq_accum + (bh * p.S + q_idx) * D;
#pragma unroll
      for (int c = half * (D / 64); c < (half + 1) * (D / 64); ++c) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + Cfg::DQ_COL + c * 32, t);
        tmem_ld_wait();
        if (q_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dq_row + c * 32 + i),
                         "f"(__uint_as_float(t[i])), "f"(__uint_as_float(t[i + 1])), "f"(__uint_as_float(t[i + 2])),
                         "f"(__uint_as_float(t[i + 3]))
                         : "memory");
          }
        }
      }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
    }
    if (Cfg::DQ_BULK && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    // ---- dV / dK accumulators: thread r <-> key row r
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    const int k_idx = k0 + r;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
      __nv_bfloat16* row = base + batch * p.g_strides[0] + head * p.g_strides[1] + static_cast<long>(k_idx) * p.g_strides[2];
      const uint32_t col = which == 0 ? Cfg::DV_COL : Cfg::DK_COL;
#pragma unroll
      for (int c = half * (D / 64); c < (half + 1) * (D / 64); ++c) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + col + c * 32, t);
        tmem_ld_wait();
        if (k_idx < p.S) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            *reinterpret_cast<uint4*>(row + c * 32 + i) = make_uint4(
                pack_bf16(__uint_as_float(t[i]), __uint_as_float(t[i + 1])),
                pack_bf16(__uint_as_float(t[i + 2]), __uint_as_float(t[i + 3])),
                pack_bf16(__uint_as_float(t[i + 4]), __uint_as_float(t[i + 5])),
                pack_bf16(__uint_as_float(t[i + 6]), __uint_as_float(t[i + 7])));
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// =================================================================================================
// Backward, pipelined variant (D = 64).  Same maths and the same tiles as attn_bwd_kernel; what changes is who waits
// for whom.  ncu of the kernel above (profiles/r17_ncu_summary.md): tensor pipe 17 %, issue slots 18 % — the chain
// S,dP → softmax → dV,dK,dQ → dQ drain → next S,dP ran strictly in sequence.  Here
//   * 16 softmax warps (4 per TMEM lane quadrant, 32 key columns each): a thread holds its whole share of S and dP in
//     64 registers, so the S/dP columns are handed back to the MMA warp right after the two tcgen05.ld — the tensor core
//     computes S,dP of query block i+1 while the softmax of block i is still running;
//   * dQ is double-buffered in TMEM (S 128 + dP 128 + dV 64 + dK 64 + 2 x dQ 64 = 512 columns) and drained one
//     iteration late, after the P/dS tiles of the next block have been handed over: neither side waits for the drain;
//   * P/dS are packed in registers and written to shared memory at the end of the softmax, when the dV/dK/dQ MMAs of
//     the previous block (which read those tiles) have long retired;
//   * one mbarrier arrival per warp instead of one per thread; lse/δ of the next block are fetched a block ahead.
// =================================================================================================
constexpr int ATT_BWDP_CWARPS = 16;
constexpr int ATT_BWDP_QST = 3;
constexpr int ATT_BWDP_SMEM = AttnBwdCfg<64>::SMEM_BYTES + 2 * (ATT_BWDP_QST - AttnBwdCfg<64>::QST) * AttnBwdCfg<64>::TILE_BYTES;
static_assert(ATT_BWDP_SMEM <= 232448, "pipelined attention backward: shared memory");
// REGS: rebalance the register file with setmaxnreg.  18 warps leave 5 warps on some SM sub-partitions, i.e. 96 registers
// per thread — the softmax threads (64 registers of S/dP alone) then spill.  With REGS the CTA is 5 aligned warpgroups:
// one control group (TMA warp, MMA warp, two idle warps) that shrinks to 64 registers and four softmax groups that grow
// to 104 (per sub-partition: 5 x 32 x 96 at launch = 15360 of 16384; 64 + 4 x 112 per lane = 16384 afterwards).
template <bool REGS>
constexpr int att_bwdp_threads() { return (REGS ? 128 : 64) + ATT_BWDP_CWARPS * 32; }

template <int BIAS, bool DROP, bool REGS>
__global__ void __launch_bounds__(att_bwdp_threads<REGS>(), 1)
attn_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     const __grid_constant__ CUtensorMap tmap_dq, AttnBwdParams p) {
  constexpr int D = 64;
  using Cfg = AttnBwdCfg<D>;
  // three Q/dO stages: S,dP of block i+1 are issued while block i's softmax runs, i.e. before the dV/dK/dQ MMAs of
  // block i-1 (the last readers of a two-stage ring's slot) have even been issued
  constexpr int QST = ATT_BWDP_QST;
  static_assert(Cfg::DQ_BULK, "pipelined backward: D = 64 configuration");
  constexpr uint32_t DQ_COL0 = 384;   // two dQ accumulators: columns [384, 448) and [448, 512)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::TILE_BYTES;
  uint8_t* sQ = sV + Cfg::TILE_BYTES;
  uint8_t* sDO = sQ + QST * Cfg::TILE_BYTES;
  uint8_t* sP = sDO + QST * Cfg::TILE_BYTES;
  uint8_t* sDS = sP + Cfg::PS_BYTES;
  uint8_t* sDQ = sDS + Cfg::PS_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + Cfg::PS_BYTES + Cfg::DQ_STAGE_BYTES);
  uint64_t* kv_full = bars;          // 1
  uint64_t* qdo_full = bars + 1;     // 3
  uint64_t* qdo_empty = bars + 4;    // 3
  uint64_t* sdp_full = bars + 7;     // 1
  uint64_t* s_free = bars + 8;       // 1  (S is in the softmax warps' registers)
  uint64_t* dp_free = bars + 16;     // 1  (dP too)
  uint64_t* pds_full = bars + 9;     // 1
  uint64_t* pds_empty = bars + 10;   // 1
  uint64_t* dq_full = bars + 11;     // 2
  uint64_t* dq_empty = bars + 13;    // 2
  uint64_t* acc_full = bars + 15;    // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp_idx = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int kv_blk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
  const int k0 = kv_blk * 128;
  const int kv_len = p.kv_lens != nullptr ? max(1, min(p.S, p.kv_lens[batch])) : p.S;
  const int nq = (p.S + 127) / 128;
  const int i_begin = p.causal ? kv_blk : 0;
  const int iters = nq - i_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq);
    mbar_init(kv_full, 1);
    for (int i = 0; i < QST; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dq_full[i], 1);
      mbar_init(&dq_empty[i], ATT_BWDP_CWARPS / 2);
    }
    mbar_init(sdp_full, 1);
    mbar_init(s_free, ATT_BWDP_CWARPS);
    mbar_init(dp_free, ATT_BWDP_CWARPS);
    mbar_init(pds_full, ATT_BWDP_CWARPS);
    mbar_init(pds_empty, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp_idx == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr int CW0 = REGS ? 4 : 2;    // first softmax warp
  if constexpr (REGS) {
    if (warp_idx < CW0) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
  }

  if (warp_idx == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::TILE_BYTES);
      tma_load_4d(sK, &tmap_k, kv_full, 0, k0, head, batch);
      tma_load_4d(sV, &tmap_v, kv_full, 0, k0, head, batch);
      for (int it = 0; it < iters; ++it) {
        const int st = it % QST;
        const int q0 = (i_begin + it) * 128;
        mbar_wait(&qdo_empty[st], ((it / QST) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * Cfg::TILE_BYTES);
        tma_load_4d(sQ + st * Cfg::TILE_BYTES, &tmap_q, &qdo_full[st], 0, q0, head, batch);
        tma_load_4d(sDO + st * Cfg::TILE_BYTES, &tmap_do, &qdo_full[st], 0, q0, head, batch);
      }
    }
  } else if (warp_idx == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);   // Q·Kᵀ, dO·Vᵀ
    constexpr uint32_t idesc_t = make_idesc_bf16(128, D, true, true);       // Pᵀ·dO, dSᵀ·Q
    constexpr uint32_t idesc_q = make_idesc_bf16(128, D, false, true);      // dS·K
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
    mbar_wait(kv_full, 0);
    // S of query block `it` once the softmax warps have read S of block it-1 into registers, dP likewise
    auto issue_s_dp = [&](int it) {
      const int st = it % QST;
      mbar_wait(&qdo_full[st], (it / QST) & 1);
      const uint32_t q_addr = smem_u32(sQ + st * Cfg::TILE_BYTES);
      const uint32_t do_addr = smem_u32(sDO + st * Cfg::TILE_BYTES);
      if (it > 0) mbar_wait(s_free, (it - 1) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16_ss(tmem_base + Cfg::S_COL, make_smem_desc_sw128(q_addr + kk * 32, 0, 1024),
                      make_smem_desc_sw128(k_addr + kk * 32, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
      }
      __syncwarp();
      if (it > 0) mbar_wait(dp_free, (it - 1) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16_ss(tmem_base + Cfg::DP_COL, make_smem_desc_sw128(do_addr + kk * 32, 0, 1024),
                      make_smem_desc_sw128(v_addr + kk * 32, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    issue_s_dp(0);
    for (int it = 0; it < iters; ++it) {
      const int st = it % QST;
      const uint32_t q_addr = smem_u32(sQ + st * Cfg::TILE_BYTES);
      const uint32_t do_addr = smem_u32(sDO + st * Cfg::TILE_BYTES);
      if (it + 1 < iters) issue_s_dp(it + 1);   // the tensor core works on the next block under this block's softmax
      mbar_wait(pds_full, it & 1);       // P / dS tiles of block `it` written
      if (it >= 2) mbar_wait(&dq_empty[it & 1], ((it >> 1) & 1) ^ 1);   // dQ accumulator (it & 1): block it-2 drained
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t dq_col = DQ_COL0 + static_cast<uint32_t>(it & 1) * 64u;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // dQ first: its drain is the only consumer that is waited for
          const uint64_t a_ds = make_smem_desc_sw128(ds_addr + (kk / 4) * (128 * 128) + (kk % 4) * 32, 0, 1024);  // dS (K-major A)
          const uint64_t b_k = make_smem_desc_sw128(k_addr + kk * 2048, 128 * 128, 1024);                          // K  (MN-major B)
          umma_f16_ss(tmem_base + dq_col, a_ds, b_k, idesc_q, kk > 0 ? 1u : 0u);
        }
        umma_commit(&dq_full[it & 1]);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a_pt = make_smem_desc_sw128(p_addr + kk * 2048, 128 * 128, 1024);       // Pᵀ  (MN-major A)
          const uint64_t b_do = make_smem_desc_sw128(do_addr + kk * 2048, 128 * 128, 1024);      // dO  (MN-major B)
          umma_f16_ss(tmem_base + Cfg::DV_COL, a_pt, b_do, idesc_t, (it > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a_dst = make_smem_desc_sw128(ds_addr + kk * 2048, 128 * 128, 1024);     // dSᵀ (MN-major A)
          const uint64_t b_q = make_smem_desc_sw128(q_addr + kk * 2048, 128 * 128, 1024);        // Q   (MN-major B)
          umma_f16_ss(tmem_base + Cfg::DK_COL, a_dst, b_q, idesc_t, (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&qdo_empty[st]);
        umma_commit(pds_empty);
        if (it == iters - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (warp_idx >= CW0) {
    const int quad = warp_idx % 4;            // TMEM lane quadrant this warp may access
    const int quarter = (warp_idx - CW0) / 4; // 32-column slice of S / dP handled by this warp
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const size_t bh = static_cast<size_t>(batch) * p.A + head;
    constexpr float LOG2E = 1.4426950408889634f;
    unsigned long long rng_seed = 0, rng_offset = 0;
    if constexpr (DROP) {
      rng_seed = static_cast<unsigned long long>(p.rng_state[0]);
      rng_offset = static_cast<unsigned long long>(p.rng_state[1]);
    }
    float slope2 = 0.f;
    if constexpr (BIAS == BIAS_ALIBI) slope2 = p.alibi_slopes[head] * LOG2E;
    const float inv_scale = 1.0f / p.scale;
    const int c = quarter;
    // shared-memory addresses as 32-bit shared-window offsets (st.shared, not generic stores)
    const uint32_t p_row = smem_u32(sP) + static_cast<uint32_t>(r * 128 + (c / 2) * (128 * 128));
    const uint32_t ds_row = smem_u32(sDS) + static_cast<uint32_t>(r * 128 + (c / 2) * (128 * 128));
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    auto sts128 = [](uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    };

    // dQ of query block `j`: TMEM -> swizzled [32 x 128 B] boxes -> one TMA tensor reduce-add per box (fp32 dq_accum)
    auto drain_dq = [&](int j) {
      if (quarter >= 2) return;
      const int b = j & 1;
      mbar_wait(&dq_full[b], (j >> 1) & 1);
      tc_fence_after_sync();
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // this warp's box was consumed
      __syncwarp();
      const uint32_t wbox = smem_u32(sDQ) + static_cast<uint32_t>((quad * (D / 32) + quarter) * Cfg::DQ_BOX_BYTES);
      uint32_t t[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + DQ_COL0 + b * 64 + quarter * 32, t);
      tmem_ld_wait();
      tc_fence_before_sync();
      const uint32_t brow = wbox + static_cast<uint32_t>(lane * 128);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4)
        sts128(brow + ((static_cast<uint32_t>(k4) ^ (static_cast<uint32_t>(lane) & 7u)) * 16u), t[k4 * 4], t[k4 * 4 + 1],
               t[k4 * 4 + 2], t[k4 * 4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&dq_empty[b]);      // the accumulator is in shared memory now
        if (!p.debug_skip_dq) asm volatile(
            "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n" ::"l"(
                reinterpret_cast<uint64_t>(&tmap_dq)),
            "r"(quarter * 32), "r"((i_begin + j) * 128 + quad * 32), "r"(static_cast<int>(bh)), "r"(wbox)
            : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    };

    // row statistics of block `it` (raw values: they are consumed one block later, nothing may wait for the loads here)
    auto row_stats = [&](int it, float& lse_raw, float& delta_raw) {
      const int q_idx = (i_begin + it) * 128 + r;
      const bool ok = it < iters && q_idx < p.S;
      lse_raw = ok ? p.lse[bh * p.S + q_idx] : 0.f;
      delta_raw = ok ? p.delta[bh * p.S + q_idx] : 0.f;
    };
    float lse_raw, delta_raw;
    row_stats(0, lse_raw, delta_raw);

    for (int it = 0; it < iters; ++it) {
      const int q_blk = i_begin + it;
      const int q_idx = q_blk * 128 + r;
      const bool q_ok = q_idx < p.S;
      const bool need_mask = (p.causal && q_blk == kv_blk) || (k0 + 128 > kv_len) || (q_blk * 128 + 128 > p.S);
      const int key_lim = !q_ok ? 0 : (p.causal ? min(kv_len, q_idx + 1) : kv_len);   // keys [0, key_lim) are visible
      const float lse2 = lse_raw * LOG2E, delta_s = delta_raw * p.scale;
      row_stats(it + 1, lse_raw, delta_raw);     // next block's row statistics: in flight during this block's softmax
      [[maybe_unused]] long bias_off = 0;
      [[maybe_unused]] uint4 bq[4];
      if constexpr (BIAS == BIAS_DENSE) {
        bias_off = batch * p.bias_strides[0] + head * p.bias_strides[1] + static_cast<long>(min(q_idx, p.S - 1)) * p.bias_strides[2];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int kb = k0 + c * 32 + v * 8;
          bq[v] = kb < p.S ? *reinterpret_cast<const uint4*>(p.bias + bias_off + kb) : make_uint4(0, 0, 0, 0);
        }
      }
      [[maybe_unused]] uint4 rnd[2];
      if constexpr (DROP) {
        rnd[0] = attn_dropout_bytes(static_cast<uint32_t>(q_idx), static_cast<uint32_t>((k0 + c * 32) >> 4), static_cast<uint32_t>(bh), rng_seed, rng_offset);
        rnd[1] = attn_dropout_bytes(static_cast<uint32_t>(q_idx), static_cast<uint32_t>((k0 + c * 32) >> 4) + 1u, static_cast<uint32_t>(bh), rng_seed, rng_offset);
      }
      // ---- S -> P.  P is kept as packed bf16 pairs (what the dV MMA reads anyway); dS below is formed from these
      // rounded probabilities, which costs one extra bf16 rounding of a value that is rounded to bf16 again right after
      uint32_t pp[16];                       // P as it multiplied V in the forward (after dropout)
      [[maybe_unused]] uint32_t pu[16];      // DROP only: the undropped probabilities
      mbar_wait(sdp_full, it & 1);
      tc_fence_after_sync();
      {
        uint32_t ts[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + Cfg::S_COL + c * 32, ts);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        auto softmax = [&](auto mask_tag) {
          constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0, p1;
            if constexpr (BIAS == BIAS_DENSE) {
              const float2 bf = unpack_bf16((&bq[i >> 3].x)[(i & 7) >> 1]);
              p0 = fast_exp2(fmaf(__uint_as_float(ts[i]), p.scale_log2, bf.x * LOG2E) - lse2);
              p1 = fast_exp2(fmaf(__uint_as_float(ts[i + 1]), p.scale_log2, bf.y * LOG2E) - lse2);
            } else if constexpr (BIAS == BIAS_ALIBI) {
              const int kidx = k0 + c * 32 + i;
              p0 = fast_exp2(fmaf(__uint_as_float(ts[i]), p.scale_log2, slope2 * static_cast<float>(kidx - q_idx)) - lse2);
              p1 = fast_exp2(fmaf(__uint_as_float(ts[i + 1]), p.scale_log2, slope2 * static_cast<float>(kidx + 1 - q_idx)) - lse2);
            } else {
              p0 = fast_exp2(fmaf(__uint_as_float(ts[i]), p.scale_log2, -lse2));
              p1 = fast_exp2(fmaf(__uint_as_float(ts[i + 1]), p.scale_log2, -lse2));
            }
            if constexpr (MASK) {
              const int kidx = k0 + c * 32 + i;
              p0 = kidx < key_lim ? p0 : 0.f;             // branch-free: one compare + select per element
              p1 = kidx + 1 < key_lim ? p1 : 0.f;
            }
            if constexpr (DROP) {
              const bool keep0 = rnd_byte(rnd[i >> 4], i & 15) >= p.drop_thresh;
              const bool keep1 = rnd_byte(rnd[i >> 4], (i & 15) + 1) >= p.drop_thresh;
              pu[i / 2] = pack_bf16(p0, p1);
              pp[i / 2] = pack_bf16(keep0 ? p0 * p.inv_keep : 0.f, keep1 ? p1 * p.inv_keep : 0.f);
            } else {
              pp[i / 2] = pack_bf16(p0, p1);
            }
          }
        };
        if (need_mask) softmax(std::true_type{}); else softmax(std::false_type{});
      }
      // ---- dP -> dS = P ∘ (dP − δ) · scale
      uint32_t dd[16];
      {
        uint32_t td[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + Cfg::DP_COL + c * 32, td);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(dp_free);
        [[maybe_unused]] float dsv[32];   // only materialised for the dbias path
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float2 pr;
          if constexpr (DROP) pr = unpack_bf16(pu[i / 2]); else pr = unpack_bf16(pp[i / 2]);
          float g0 = __uint_as_float(td[i]), g1 = __uint_as_float(td[i + 1]);   // dP w.r.t. the dropped, rescaled P
          if constexpr (DROP) {
            g0 = rnd_byte(rnd[i >> 4], i & 15) >= p.drop_thresh ? g0 * p.inv_keep : 0.f;
            g1 = rnd_byte(rnd[i >> 4], (i & 15) + 1) >= p.drop_thresh ? g1 * p.inv_keep : 0.f;
          }
          const float d0 = pr.x * fmaf(g0, p.scale, -delta_s);
          const float d1 = pr.y * fmaf(g1, p.scale, -delta_s);
          dd[i / 2] = pack_bf16(d0, d1);
          if constexpr (BIAS == BIAS_DENSE) {
            dsv[i] = d0 * inv_scale;
            dsv[i + 1] = d1 * inv_scale;
          }
        }
        if constexpr (BIAS == BIAS_DENSE) {
          if (p.dbias != nullptr && q_ok) {   // d bias = P ∘ (dP − δ): fp32 reductions (broadcast dims sum over CTAs)
            float* drow = p.dbias + bias_off + k0 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (k0 + c * 32 + i < p.S)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(drow + i), "f"(dsv[i]), "f"(dsv[i + 1]),
                             "f"(dsv[i + 2]), "f"(dsv[i + 3])
                             : "memory");
            }
          }
        }
      }
      // the dV/dK/dQ MMAs of the previous block read the P/dS tiles: they were issued a whole softmax ago
      mbar_wait(pds_empty, (it & 1) ^ 1);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t phys = (static_cast<uint32_t>((c % 2) * 4 + q4) ^ sw) * 16u;
        sts128(p_row + phys, pp[q4 * 4], pp[q4 * 4 + 1], pp[q4 * 4 + 2], pp[q4 * 4 + 3]);
        sts128(ds_row + phys, dd[q4 * 4], dd[q4 * 4 + 1], dd[q4 * 4 + 2], dd[q4 * 4 + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (it > 0) drain_dq(it - 1);     // issued behind the previous block's P/dS hand-over: complete by now
    }
    drain_dq(iters - 1);
    // (the CTA may retire while the reductions are still in flight: only their shared-memory reads must be done)
    if (quarter < 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    // ---- dV / dK accumulators: thread r <-> key row r; warps of quarters 0,1 write dV, quarters 2,3 write dK
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    const int k_idx = k0 + r;
    {
      const int which = quarter / 2, cc = quarter % 2;
      __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
      __nv_bfloat16* row = base + batch * p.g_strides[0] + head * p.g_strides[1] + static_cast<long>(k_idx) * p.g_strides[2];
      const uint32_t col = which == 0 ? Cfg::DV_COL : Cfg::DK_COL;
      uint32_t t[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + col + cc * 32, t);
      tmem_ld_wait();
      if (k_idx < p.S) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          *reinterpret_cast<uint4*>(row + cc * 32 + i) = make_uint4(
              pack_bf16(__uint_as_float(t[i]), __uint_as_float(t[i + 1])),
              pack_bf16(__uint_as_float(t[i + 2]), __uint_as_float(t[i + 3])),
              pack_bf16(__uint_as_float(t[i + 4]), __uint_as_float(t[i + 5])),
              pack_bf16(__uint_as_float(t[i + 6]), __uint_as_float(t[i + 7])));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// delta[b, a, s] = sum_d dO[b, a, s, d] * O[b, a, s, d]
// D / 8 lanes per row, 16-byte loads; rows are visited in (b, s, a) order = memory order of the [b, s, a, d] tensors, so
// a warp reads 32 x 16 contiguous bytes per tensor, and every thread keeps 4 rows (8 loads) in flight.
template <int D>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ o, float* __restrict__ delta,
                  int B, int A, int S, long sb, long sa, long ss, float4* __restrict__ zero_fill, long zero_n4) {
  constexpr int LPR = D / 8;  // lanes per row
  const long total = static_cast<long>(B) * A * S;
  const long gid = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  // the fp32 dQ accumulator of the main kernel is cleared here (one launch and one pass less than a memset in front)
  for (long z = gid; z < zero_n4; z += static_cast<long>(gridDim.x) * blockDim.x) zero_fill[z] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long gstride = static_cast<long>(gridDim.x) * blockDim.x / LPR;
  const int sub = static_cast<int>(gid % LPR);
  const long warp_r0 = (gid - (threadIdx.x % 32)) / LPR;  // loop bound is warp-uniform (the shuffles below need all lanes)
  for (long base = 0; warp_r0 + base < total; base += 4 * gstride) {
    const long r0 = gid / LPR + base;
    uint4 x[4], y[4];
    long drow[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long r = r0 + u * gstride;
      const bool ok = r < total;
      const long rr = ok ? r : 0;
      const int a = static_cast<int>(rr % A), sq = static_cast<int>((rr / A) % S);
      const long b = rr / (static_cast<long>(A) * S);
      const long off = b * sb + a * sa + sq * ss + sub * 8;
      drow[u] = ok ? (b * A + a) * static_cast<long>(S) + sq : -1;
      x[u] = ok ? *reinterpret_cast<const uint4*>(dout + off) : make_uint4(0, 0, 0, 0);
      y[u] = ok ? *reinterpret_cast<const uint4*>(o + off) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 a0 = unpack_bf16(x[u].x), a1 = unpack_bf16(x[u].y), a2 = unpack_bf16(x[u].z), a3 = unpack_bf16(x[u].w);
      const float2 b0 = unpack_bf16(y[u].x), b1 = unpack_bf16(y[u].y), b2 = unpack_bf16(y[u].z), b3 = unpack_bf16(y[u].w);
      float acc = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
#pragma unroll
      for (int m = LPR / 2; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
      if (sub == 0 && drow[u] >= 0) delta[drow[u]] = acc;
    }
  }
}

// dq (strided bf16) = dq_accum (fp32 [B, A, S, D]); 8 elements per thread (two 16-byte loads, one 16-byte store)
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int B, int A, int S, int D, long sb,
                       long sa, long ss) {
  const long nvec = static_cast<long>(B) * A * S * D / 8;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i0 = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i0 < nvec; i0 += 2 * stride) {
    float4 lo[2], hi[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long i = i0 + u * stride;
      if (i < nvec) {
        lo[u] = reinterpret_cast<const float4*>(acc)[2 * i];
        hi[u] = reinterpret_cast<const float4*>(acc)[2 * i + 1];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long i = i0 + u * stride;
      if (i < nvec) {
        const long e = i * 8;
        const int d = static_cast<int>(e % D);
        const long row = e / D;
        const int sq = static_cast<int>(row % S), a = static_cast<int>((row / S) % A);
        const long b = row / (static_cast<long>(S) * A);
        *reinterpret_cast<uint4*>(dq + b * sb + a * sa + sq * ss + d) =
            make_uint4(pack_bf16(lo[u].x, lo[u].y), pack_bf16(lo[u].z, lo[u].w), pack_bf16(hi[u].x, hi[u].y), pack_bf16(hi[u].z, hi[u].w));
      }
    }
  }
}

}  // namespace lb

namespace {
template <int D, int BIAS, bool DROP>
cudaError_t launch_bwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                         const CUtensorMap& tdq, const lb::AttnBwdParams& p, cudaStream_t s) {
  using Cfg = lb::AttnBwdCfg<D>;
  auto kern = lb::attn_bwd_kernel<D, BIAS, DROP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.S + 127) / 128, p.A, p.B);
  kern<<<grid, lb::ATT_BWD_THREADS, Cfg::SMEM_BYTES, s>>>(tq, tk, tv, tdo, tdq, p);
  return cudaGetLastError();
}
template <int BIAS, bool DROP, bool REGS>
cudaError_t launch_bwd_pipe_r(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                              const CUtensorMap& tdq, const lb::AttnBwdParams& p, cudaStream_t s) {
  auto kern = lb::attn_bwd_pipe_kernel<BIAS, DROP, REGS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lb::ATT_BWDP_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.S + 127) / 128, p.A, p.B);
  kern<<<grid, lb::att_bwdp_threads<REGS>(), lb::ATT_BWDP_SMEM, s>>>(tq, tk, tv, tdo, tdq, p);
  return cudaGetLastError();
}
// LIBAI_B200_ATTN_BWD_PIPE: 0 = sequential kernel, 1 (default) = pipelined, 2 = pipelined + setmaxnreg (experimental:
// the first version, 64/112 registers = the whole register file, never got its TRY_ALLOC granted and hung)
static int attn_bwd_pipe_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LIBAI_B200_ATTN_BWD_PIPE");
    v = e == nullptr ? 1 : atoi(e);
  }
  return v;
}
template <int BIAS, bool DROP>
cudaError_t launch_bwd_pipe_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                              const CUtensorMap& tdq, const lb::AttnBwdParams& p, cudaStream_t s) {
  return attn_bwd_pipe_enabled() >= 2 ? launch_bwd_pipe_r<BIAS, DROP, true>(tq, tk, tv, tdo, tdq, p, s)
                                      : launch_bwd_pipe_r<BIAS, DROP, false>(tq, tk, tv, tdo, tdq, p, s);
}
template <int D>
cudaError_t launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                       const CUtensorMap& tdq, const lb::AttnBwdParams& p, cudaStream_t s) {
  const int bias = p.bias != nullptr ? lb::BIAS_DENSE : (p.alibi_slopes != nullptr ? lb::BIAS_ALIBI : lb::BIAS_NONE);
  const bool drop = p.drop_thresh > 0;
  if constexpr (D == 64) {
    if (attn_bwd_pipe_enabled()) {
      if (bias == lb::BIAS_NONE) return drop ? launch_bwd_pipe_t<lb::BIAS_NONE, true>(tq, tk, tv, tdo, tdq, p, s) : launch_bwd_pipe_t<lb::BIAS_NONE, false>(tq, tk, tv, tdo, tdq, p, s);
      if (bias == lb::BIAS_DENSE) return drop ? launch_bwd_pipe_t<lb::BIAS_DENSE, true>(tq, tk, tv, tdo, tdq, p, s) : launch_bwd_pipe_t<lb::BIAS_DENSE, false>(tq, tk, tv, tdo, tdq, p, s);
      return drop ? launch_bwd_pipe_t<lb::BIAS_ALIBI, true>(tq, tk, tv, tdo, tdq, p, s) : launch_bwd_pipe_t<lb::BIAS_ALIBI, false>(tq, tk, tv, tdo, tdq, p, s);
    }
  }
  if (bias == lb::BIAS_NONE) return drop ? launch_bwd_t<D, lb::BIAS_NONE, true>(tq, tk, tv, tdo, tdq, p, s) : launch_bwd_t<D, lb::BIAS_NONE, false>(tq, tk, tv, tdo, tdq, p, s);
  if (bias == lb::BIAS_DENSE) return drop ? launch_bwd_t<D, lb::BIAS_DENSE, true>(tq, tk, tv, tdo, tdq, p, s) : launch_bwd_t<D, lb::BIAS_DENSE, false>(tq, tk, tv, tdo, tdq, p, s);
  return drop ? launch_bwd_t<D, lb::BIAS_ALIBI, true>(tq, tk, tv, tdo, tdq, p, s) : launch_bwd_t<D, lb::BIAS_ALIBI, false>(tq, tk, tv, tdo, tdq, p, s);
}
}  // namespace

extern "C" int lb_attn_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o,
                           const float* lse, void* dq, void* dk, void* dv, float* delta, float* dq_accum, int B, int A,
                           int S, int D, const long* q_strides, const long* k_strides, const long* v_strides,
                           const long* do_strides, int causal, float scale, const int* kv_lens, const void* bias,
                           const long* bias_strides, const float* alibi_slopes, float* dbias, float p_drop,
                           const long long* rng_state, cudaStream_t s) {
  if (D != 64 && D != 128) return -1;
  for (int i = 0; i < 3; ++i)
    if ((q_strides[i] % 8) || (k_strides[i] % 8) || (v_strides[i] % 8) || (do_strides[i] % 8)) return -3;
  CUtensorMap tq, tk, tv, tdo;
  if (!make_qkv_tmap(&tq, q, B, A, S, D, q_strides)) return -2;
  if (!make_qkv_tmap(&tk, k, B, A, S, D, k_strides)) return -2;
  if (!make_qkv_tmap(&tv, v, B, A, S, D, v_strides)) return -2;
  if (!make_qkv_tmap(&tdo, dout, B, A, S, D, do_strides)) return -2;
  CUtensorMap tdq;
  {
    // fp32 dq_accum [B*A, S, D]: 32x32-float boxes (128-byte rows, 128B swizzle) for the TMA reduce-add
    uint64_t dims[3] = {(uint64_t)D, (uint64_t)S, (uint64_t)B * A};
    uint64_t strides[3] = {4, (uint64_t)D * 4, (uint64_t)S * D * 4};
    uint32_t box[3] = {32, 32, 1};
    if (!lb_host::make_tmap_typed(&tdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dq_accum, 3, dims, strides, box,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
      return -2;
  }
  const long rows = (long)B * A * S;
  static int skip_helpers = -1;   // timing experiments only: the main kernel alone (results are then meaningless)
  if (skip_helpers < 0) {
    const char* e = getenv("LIBAI_B200_ATTN_DEBUG_SKIP_HELPERS");
    skip_helpers = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  if (skip_helpers) cudaMemsetAsync(dq_accum, 0, sizeof(float) * rows * D, s);
  if (!skip_helpers) {
    // 4 rows per thread-group and iteration; cap the grid at a few waves
    const long groups = (rows + 3) / 4;
    const long lpr = D / 8;
    long blocks = (groups * lpr + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    if (D == 64)
      lb::attn_delta_kernel<64><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)o, delta, B,
                                                                   A, S, do_strides[0], do_strides[1], do_strides[2],
                                                                   reinterpret_cast<float4*>(dq_accum), rows * D / 4);
    else
      lb::attn_delta_kernel<128><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)o, delta, B,
                                                                    A, S, do_strides[0], do_strides[1], do_strides[2],
                                                                    reinterpret_cast<float4*>(dq_accum), rows * D / 4);
  }
  lb::AttnBwdParams p;
  p.lse = lse;
  p.delta = delta;
  p.dq_accum = dq_accum;
  p.dk = (__nv_bfloat16*)dk;
  p.dv = (__nv_bfloat16*)dv;
  // dq / dk / dv are views of one packed [B, S, A, 3, D] tensor: same strides as the packed q view
  p.g_strides[0] = (long)S * A * 3 * D;
  p.g_strides[1] = 3L * D;
  p.g_strides[2] = (long)A * 3 * D;
  p.B = B;
  p.A = A;
  p.S = S;
  p.causal = causal;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.kv_lens = kv_lens;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  for (int i = 0; i < 3; ++i) p.bias_strides[i] = (bias != nullptr && bias_strides != nullptr) ? bias_strides[i] : 0;
  p.alibi_slopes = bias != nullptr ? nullptr : alibi_slopes;
  p.dbias = bias != nullptr ? dbias : nullptr;
  uint32_t thr = (uint32_t)(p_drop * 256.0f + 0.5f);
  if (thr > 255u) thr = 255u;
  p.drop_thresh = (p_drop > 0.f && rng_state != nullptr) ? (thr == 0 ? 1u : thr) : 0u;
  p.inv_keep = 256.0f / (256.0f - (float)p.drop_thresh);
  p.rng_state = rng_state;
  {
    static int skip = -1;
    if (skip < 0) {
      const char* e = getenv("LIBAI_B200_ATTN_DEBUG_SKIP_DQ");
      skip = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    }
    p.debug_skip_dq = skip;
  }
  cudaError_t e = (D == 64) ? launch_bwd<64>(tq, tk, tv, tdo, tdq, p, s) : launch_bwd<128>(tq, tk, tv, tdo, tdq, p, s);
  if (e != cudaSuccess) return (int)e;
  if (skip_helpers) return 0;
  const long nvec = rows * D / 8;
  int blocks = (int)((nvec + 511) / 512);
  if (blocks > 148 * 8) blocks = 148 * 8;
  lb::attn_dq_convert_kernel<<<blocks, 256, 0, s>>>(dq_accum, (__nv_bfloat16*)dq, B, A, S, D, p.g_strides[0],
                                                    p.g_strides[1], p.g_strides[2]);
  return (int)cudaGetLastError();
}
