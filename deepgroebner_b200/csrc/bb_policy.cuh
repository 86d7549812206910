// bb_policy.cuh -- the pairs policy head fused with the environment: ParallelMultilayerPerceptron([H])
// (deepgroebner/networks.py:522-571 = ParallelEmbeddingLayer :49-95 with one Dense(H, relu) + ParallelDecidingLayer
// :414-460 with Dense(1) and log_softmax over the rows) followed by tf.random.categorical (pg.py:323-326), evaluated
// directly on the rows of the environment's state matrix (lead_monomials_vector, buchberger.cpp:354-370) without
// materialising the matrix.  fp32 arithmetic (the reference casts the int32 matrix to float32, networks.py:88).
//
//   logit[r] = b2 + sum_u w2[u] * relu(b1[u] + sum_c W1[c][u] * x[r][c]),   x[r] = lead(G[i_r]) || lead(G[j_r])
//   logp     = log_softmax(logit[0 .. |P|))          (padded rows get -1e9 in the reference, i.e. probability 0)
//   action   = inverse CDF of softmax(logit) at u * sum,  u = 24-bit uniform from a counter hash (documented in bbenv.h)
//
// One warp per environment.  The first layer is the one dense contraction on this path ([|P| rows] x [cols] x [H]), so
// it runs on the tensor cores, 16 rows at a time: mma.sync m16n8k8 TF32 (a warp-level MMA is the right size for a
// 16 x 16 x 128 tile per environment; tcgen05 tiles start at 64 rows of ONE matrix per CTA).  The A operand -- small
// non-negative integers, the exponents of the state matrix -- is exact in TF32; W1 is split as W_hi + W_lo (two TF32
// values, residual ~2^-22 |w|), two MMAs per tile, fp32 accumulation: fp32-grade logits (tests: rtol = atol = 1e-5
// against torch fp32).  Exponents >= 2048 (never seen on the BASELINE distributions) are split the same way.  The
// second layer (relu, dot with w2) is applied to the accumulator fragments in registers, a quad reduction gives one
// logit per row.  The weight image in shared memory is FRAGMENT-MAJOR: one 16-byte word per (k-step, n-tile, lane) holds
// that lane's B fragment of both halves {hi(k), hi(k+4), lo(k), lo(k+4)}, so a tile costs one conflict-free LDS.128.
// State matrices wider than 32 columns (KS > 4 k-steps) use the scalar path below: lane l owns hidden units l,
// l+32, ..., W1s[c][u] in shared memory, row features broadcast by shuffle.
#pragma once
#include "bb_device.cuh"

#define BB_POLICY_MAX_HIDDEN 256
#define BB_POLICY_MAX_COLS 64

struct BBPolicy {
  int hidden;            // H, multiple of 32, <= BB_POLICY_MAX_HIDDEN
  const float* W1;       // [cols][H]  (Keras Dense kernel layout: input-major)
  const float* b1;       // [H]
  const float* w2;       // [H]        (Dense(1) kernel)
  const float* b2;       // [1]
  unsigned long long seed;
  int greedy;            // 1: argmax instead of sampling (lowest row on ties)
};

// shared-memory image of the weights: W1s[cols*H], b1s[H], w2s[H], b2, pad to 4 floats, then (cols <= 32 only) the
// tensor-core image: uint4 Wf[KS][H/8][32] (B fragments, TF32 hi/lo, rows >= cols zero; KS = ceil(cols / 8)) and
// float4 bw[H/8][4] = {b1[8j+2t], b1[8j+2t+1], w2[8j+2t], w2[8j+2t+1]} (the accumulator fragment's two hidden units)
#define BB_POLICY_MMA_MAX_KS 4
__host__ __device__ __forceinline__ int policy_mma_offset(int cols, int H) { return (cols * H + 2 * H + 1 + 3) & ~3; }
__host__ __device__ __forceinline__ int policy_smem_floats(int cols, int H) {
  const int KS = (cols + 7) >> 3;
  return policy_mma_offset(cols, H) + (KS <= BB_POLICY_MMA_MAX_KS ? 4 * (KS * (H >> 3) * 32 + (H >> 3) * 4) : 0);
}
__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void policy_load_weights(const BBPolicy& W, int cols, float* sm) {
  const int H = W.hidden;
  for (int x = threadIdx.x; x < cols * H; x += blockDim.x) sm[x] = W.W1[x];
  for (int x = threadIdx.x; x < H; x += blockDim.x) { sm[cols * H + x] = W.b1[x]; sm[cols * H + H + x] = W.w2[x]; }
  if (threadIdx.x == 0) sm[cols * H + 2 * H] = W.b2[0];
  const int KS = (cols + 7) >> 3;
  if (KS <= BB_POLICY_MMA_MAX_KS) {
    const int NT = H >> 3;
    uint4* Wf = reinterpret_cast<uint4*>(sm + policy_mma_offset(cols, H));
    float4* bw = reinterpret_cast<float4*>(Wf + KS * NT * 32);
    for (int x = threadIdx.x; x < KS * NT * 32; x += blockDim.x) {
      const int l = x & 31, sj = x >> 5, j = sj % NT, ks = sj / NT;
      const int k0 = 8 * ks + (l & 3), k1 = k0 + 4, u = 8 * j + (l >> 2);
      const float w0 = k0 < cols ? W.W1[k0 * H + u] : 0.0f, w1 = k1 < cols ? W.W1[k1 * H + u] : 0.0f;
      const uint32_t h0 = f32_to_tf32(w0), h1 = f32_to_tf32(w1);
      Wf[x] = make_uint4(h0, h1, f32_to_tf32(w0 - __uint_as_float(h0)), f32_to_tf32(w1 - __uint_as_float(h1)));
    }
    for (int x = threadIdx.x; x < NT * 4; x += blockDim.x) {
      const int u = 8 * (x >> 2) + 2 * (x & 3);
      bw[x] = make_float4(W.b1[u], W.b1[u + 1], W.w2[u], W.w2[u + 1]);
    }
  }
  __syncthreads();
}

// 24-bit uniform in [0,1) for (environment stream id, step counter)
__device__ __forceinline__ float policy_uniform(unsigned long long seed, unsigned long long stream, unsigned long long counter) {
  return (float)(bb_hash_item_impl(seed + stream, counter) >> 40) * (1.0f / 16777216.0f);
}

// feature c of row `pr` for lane c < cols: exponent v of term t of G[i] (side 0) or G[j] (side 1), 0 if absent
template <int NV>
__device__ __forceinline__ float policy_feature(const BBParams& P, const Env& e, uint32_t pr, int c) {
  typedef KL<NV> K;
  const int half = NV * P.k;
  if (c >= 2 * half) return 0.0f;
  const int side = c >= half, cc = c - side * half;
  const int t = cc / NV, v = cc - t * NV;
  const GHeadMem* g = ENV_PTR(GHeadMem, e, P, o_ghead) + (side ? (pr >> 16) : (pr & 0xffffu));
  if (t >= (int)g->len) return 0.0f;
  const uint64_t key = t == 0 ? g->lm : (t == 1 ? g->k1 : ENV_PTR(uint64_t, e, P, o_tkey)[g->off + t]);
  return (float)K::exp(key, v);
}

// d += a * b, m16n8k8, A row-major (16 x 8), B column-major (8 x 8), TF32 operands, fp32 accumulators
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// logit[r] for every row of P on the tensor cores (cols <= 32).  Fragment coordinates: g = lane / 4, t = lane % 4;
// A: (row g | g+8, column 8s + t | 8s + t + 4); B: (k = 8s + t | 8s + t + 4, hidden unit 8j + g);
// C: (row g | g+8, hidden unit 8j + 2t | 8j + 2t + 1).
template <int NV>
__device__ __forceinline__ void policy_logits_mma(const BBParams& P, const Env& e, const float* sm, int H, float* logit) {
  const int lane = bb_lane(), g = lane >> 2, t = lane & 3;
  const int cols = P.cols, nP = e.nP, KS = (cols + 7) >> 3, NT = H >> 3;
  const float b2 = sm[cols * H + 2 * H];
  const uint4* Wf = reinterpret_cast<const uint4*>(sm + policy_mma_offset(cols, H)) + lane;
  const float4* bw = reinterpret_cast<const float4*>(Wf - lane + KS * NT * 32) + t;
  const uint32_t* pairs = ENV_PTR(uint32_t, e, P, o_pairs);
#pragma unroll 1
  for (int r0 = 0; r0 < nP; r0 += 16) {
    const int ra = r0 + g, rb = ra + 8;
    const uint32_t pa = ra < nP ? pairs[ra] : 0u, pb = rb < nP ? pairs[rb] : 0u;   // rows past |P| are computed and dropped
    uint32_t x[BB_POLICY_MMA_MAX_KS][4];   // A fragments as fp32 bit patterns (exact integers)
    bool big = false;
#pragma unroll
    for (int s = 0; s < BB_POLICY_MMA_MAX_KS; s++) {
      x[s][0] = x[s][1] = x[s][2] = x[s][3] = 0u;
      if (s < KS) {
        const float f0 = policy_feature<NV>(P, e, pa, 8 * s + t), f1 = policy_feature<NV>(P, e, pb, 8 * s + t);
        const float f2 = policy_feature<NV>(P, e, pa, 8 * s + t + 4), f3 = policy_feature<NV>(P, e, pb, 8 * s + t + 4);
        big |= f0 >= 2048.0f || f1 >= 2048.0f || f2 >= 2048.0f || f3 >= 2048.0f;
        x[s][0] = __float_as_uint(f0); x[s][1] = __float_as_uint(f1); x[s][2] = __float_as_uint(f2); x[s][3] = __float_as_uint(f3);
      }
    }
    float sa = 0.0f, sb = 0.0f;   // this lane's share of logit[ra], logit[rb]
    if (!__any_sync(BB_FULL, big)) {
#pragma unroll 2
      for (int j = 0; j < NT; j++) {
        const float4 q = bw[4 * j];
        float d[4] = {q.x, q.y, q.x, q.y};
#pragma unroll
        for (int s = 0; s < BB_POLICY_MMA_MAX_KS; s++) {
          if (s < KS) {
            const uint4 w = Wf[(s * NT + j) * 32];
            mma_tf32(d, x[s], w.x, w.y);
            mma_tf32(d, x[s], w.z, w.w);
          }
        }
        sa = fmaf(q.z, fmaxf(d[0], 0.0f), sa); sa = fmaf(q.w, fmaxf(d[1], 0.0f), sa);
        sb = fmaf(q.z, fmaxf(d[2], 0.0f), sb); sb = fmaf(q.w, fmaxf(d[3], 0.0f), sb);
      }
    } else {   // an exponent with more than 11 significant bits somewhere in the tile: x = x_hi + x_lo, four products
#pragma unroll 1
      for (int j = 0; j < NT; j++) {
        const float4 q = bw[4 * j];
        float d[4] = {q.x, q.y, q.x, q.y};
#pragma unroll 1
        for (int s = 0; s < KS; s++) {
          const uint4 w = Wf[(s * NT + j) * 32];
          uint32_t ah[4], al[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const float f = __uint_as_float(s == 0 ? x[0][c] : (s == 1 ? x[1][c] : (s == 2 ? x[2][c] : x[3][c])));
            const float hi = (float)((int)f & ~2047);
            ah[c] = __float_as_uint(hi); al[c] = __float_as_uint(f - hi);
          }
          mma_tf32(d, ah, w.x, w.y);
          mma_tf32(d, al, w.x, w.y);
          mma_tf32(d, ah, w.z, w.w);
          mma_tf32(d, al, w.z, w.w);
        }
        sa = fmaf(q.z, fmaxf(d[0], 0.0f), sa); sa = fmaf(q.w, fmaxf(d[1], 0.0f), sa);
        sb = fmaf(q.z, fmaxf(d[2], 0.0f), sb); sb = fmaf(q.w, fmaxf(d[3], 0.0f), sb);
      }
    }
    sa += __shfl_xor_sync(BB_FULL, sa, 1); sa += __shfl_xor_sync(BB_FULL, sa, 2);
    sb += __shfl_xor_sync(BB_FULL, sb, 1); sb += __shfl_xor_sync(BB_FULL, sb, 2);
    if (t == 0) {
      if (ra < nP) logit[ra] = sa + b2;
      if (rb < nP) logit[rb] = sb + b2;
    }
  }
}

// the same logits on the fp32 FMA pipe (state matrices wider than 32 columns)
template <int NV, int UPL>  // UPL = hidden units per lane = H / 32
__device__ __forceinline__ void policy_logits_fma(const BBParams& P, const Env& e, const float* sm, float* logit) {
  const int lane = bb_lane();
  const int cols = P.cols, H = UPL * 32, nP = e.nP;
  const float* W1s = sm; const float* b1s = sm + cols * H; const float* w2s = b1s + H;
  const float b2 = w2s[H];
  const uint32_t* pairs = ENV_PTR(uint32_t, e, P, o_pairs);
  float bias[UPL], w2r[UPL];
#pragma unroll
  for (int q = 0; q < UPL; q++) { bias[q] = b1s[q * 32 + lane]; w2r[q] = w2s[q * 32 + lane]; }
#pragma unroll 1
  for (int r = 0; r < nP; r++) {
    const uint32_t pr = pairs[r];
    const float x0 = policy_feature<NV>(P, e, pr, lane);  // lanes >= cols hold 0 and are never read
    const float x1 = cols > 32 ? policy_feature<NV>(P, e, pr, lane + 32) : 0.0f;
    float acc[UPL];
#pragma unroll
    for (int q = 0; q < UPL; q++) acc[q] = bias[q];
#pragma unroll 4
    for (int c = 0; c < cols; c++) {
      const float xc = __shfl_sync(BB_FULL, c < 32 ? x0 : x1, c & 31);
#pragma unroll
      for (int q = 0; q < UPL; q++) acc[q] = fmaf(W1s[c * H + q * 32 + lane], xc, acc[q]);
    }
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < UPL; q++) s = fmaf(w2r[q], fmaxf(acc[q], 0.0f), s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(BB_FULL, s, o);
    if (lane == 0) logit[r] = s + b2;
  }
}

// Evaluates the head on environment e (|P| >= 1).  logit scratch: the slot's o_logit array (float[max_pairs]).
// Returns the chosen row; logp_out = log pi(row).  If logits_out != nullptr the log-probabilities of all rows are
// written there ([pmax], rows beyond |P| untouched).
template <int NV, int UPL>
__device__ __forceinline__ int warp_policy(const BBParams& P, const Env& e, const float* sm, int greedy, float u,
                                           float& logp_out, float* logits_out, int pmax) {
  const int lane = bb_lane();
  const int nP = e.nP;
  float* logit = ENV_PTR(float, e, P, o_logit);
  if (((P.cols + 7) >> 3) <= BB_POLICY_MMA_MAX_KS) policy_logits_mma<NV>(P, e, sm, UPL * 32, logit);
  else policy_logits_fma<NV, UPL>(P, e, sm, logit);
  __syncwarp();
  float mx = -3.0e38f;
#pragma unroll 1
  for (int r = lane; r < nP; r += 32) mx = fmaxf(mx, logit[r]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(BB_FULL, mx, o));
  // softmax statistics
  float part = 0.0f;
#pragma unroll 1
  for (int r = lane; r < nP; r += 32) part += expf(logit[r] - mx);
  float total = part;
#pragma unroll
  for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(BB_FULL, total, o);
  const float lse = mx + logf(total);
  if (logits_out) {
#pragma unroll 1
    for (int r = lane; r < nP && r < pmax; r += 32) logits_out[r] = logit[r] - lse;
  }
  int action = nP - 1;
  if (greedy) {
    uint32_t best = 0xffffffffu;
#pragma unroll 1
    for (int r = lane; r < nP; r += 32) if (logit[r] == mx && (uint32_t)r < best) best = (uint32_t)r;
    action = (int)__reduce_min_sync(BB_FULL, best);
  } else {
    // inverse CDF in row order: first row whose inclusive prefix sum exceeds u * total
    const float target = u * total;
    float carry = 0.0f;
    bool hit = false;
#pragma unroll 1
    for (int b0 = 0; b0 < nP && !hit; b0 += 32) {
      const int r = b0 + lane;
      float v = r < nP ? expf(logit[r] - mx) : 0.0f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(BB_FULL, v, o); if (lane >= o) v += t; }
      v += carry;
      const uint32_t b = __ballot_sync(BB_FULL, r < nP && v > target);
      if (b) { action = b0 + __ffs(b) - 1; hit = true; }
      carry = __shfl_sync(BB_FULL, v, 31);
    }
  }
  logp_out = logit[action] - lse;
  __syncwarp();
  return action;
}

template <int NV>
__device__ __forceinline__ int warp_policy_dispatch(const BBParams& P, const Env& e, const float* sm, int hidden, int greedy,
                                                    float u, float& logp, float* logits_out, int pmax) {
  switch (hidden >> 5) {
    case 1: return warp_policy<NV, 1>(P, e, sm, greedy, u, logp, logits_out, pmax);
    case 2: return warp_policy<NV, 2>(P, e, sm, greedy, u, logp, logits_out, pmax);
    case 4: return warp_policy<NV, 4>(P, e, sm, greedy, u, logp, logits_out, pmax);
    default: return warp_policy<NV, 8>(P, e, sm, greedy, u, logp, logits_out, pmax);
  }
}
