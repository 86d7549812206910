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
// One warp per environment.  Lane l owns hidden units l, l+32, ... (H/32 of them); the first-layer weights of those
// units live in shared memory as W1s[c][u] so that a warp reads 32 consecutive floats per (c, unit group): no bank
// conflicts.  Row features are produced by the lanes c < cols (one exponent each) and broadcast by shuffle.
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

// shared-memory image of the weights: W1s[cols*H], b1s[H], w2s[H], b2
__device__ __forceinline__ int policy_smem_floats(int cols, int H) { return cols * H + 2 * H + 1; }
__device__ __forceinline__ void policy_load_weights(const BBPolicy& W, int cols, float* sm) {
  const int H = W.hidden;
  for (int x = threadIdx.x; x < cols * H; x += blockDim.x) sm[x] = W.W1[x];
  for (int x = threadIdx.x; x < H; x += blockDim.x) { sm[cols * H + x] = W.b1[x]; sm[cols * H + H + x] = W.w2[x]; }
  if (threadIdx.x == 0) sm[cols * H + 2 * H] = W.b2[0];
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

// Evaluates the head on environment e (|P| >= 1).  logit scratch: the slot's o_logit array (float[max_pairs]).
// Returns the chosen row; logp_out = log pi(row).  If logits_out != nullptr the log-probabilities of all rows are
// written there ([pmax], rows beyond |P| untouched).
template <int NV, int UPL>  // UPL = hidden units per lane = H / 32
__device__ __forceinline__ int warp_policy(const BBParams& P, const Env& e, const float* sm, int greedy, float u,
                                           float& logp_out, float* logits_out, int pmax) {
  const int lane = bb_lane();
  const int cols = P.cols, H = UPL * 32, nP = e.nP;
  const float* W1s = sm; const float* b1s = sm + cols * H; const float* w2s = b1s + H;
  const float b2 = w2s[H];
  const uint32_t* pairs = ENV_PTR(uint32_t, e, P, o_pairs);
  float* logit = ENV_PTR(float, e, P, o_logit);
  float bias[UPL], w2r[UPL];
#pragma unroll
  for (int q = 0; q < UPL; q++) { bias[q] = b1s[q * 32 + lane]; w2r[q] = w2s[q * 32 + lane]; }
  float mx = -3.0e38f;
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
    s += b2;
    if (lane == 0) logit[r] = s;
    mx = fmaxf(mx, s);
  }
  __syncwarp();
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
