// Text-decoder step kernels: embedding frontend, single-token self/cross attention over the KV cache, fused
// log-softmax statistics + top-K, and the device-resident beam-search bookkeeping.
//
// The reference runs beam search on the host (fairseq2 BeamSearchSeq2SeqGenerator; C++ mirror
// ggml/examples/unity/fairseq2.cpp:1371-1608) with one host round trip per step and an index_select copy of the
// whole KV cache per step (reorder_kv_cache, fairseq2.cpp:170-198).  Here the search state lives in device memory,
// no step synchronises with the host, and beam reordering is an ancestor-index table: row r at position t reads
// cache slot anc[r][t]; K/V are written once and never moved.
#include "common.cuh"

namespace sb {

constexpr int HD = 64;

// ------------------------------------------------------------------------------------------- embedding frontend
__global__ void embed_kernel(const int* __restrict__ ids, long long ids_ld, const int* __restrict__ step_ptr, int L,
                             const elem_t* __restrict__ embed, const float* __restrict__ pos, float scale,
                             elem_t* __restrict__ x, int dim) {
  // grid: (rows, L); x row = r*L + t.  With step_ptr the single column `*step_ptr` is embedded at position *step_ptr.
  const int r = blockIdx.x, t = blockIdx.y;
  const int id_col0 = step_ptr ? *step_ptr : 0, pos0 = id_col0;
  const int tok = ids[(long long)r * ids_ld + id_col0 + t];
  const elem_t* e = embed + (long long)tok * dim;
  const float* pp = pos + (long long)(pos0 + t) * dim;
  elem_t* xp = x + ((long long)r * L + t) * dim;
  for (int c = threadIdx.x * 2; c < dim; c += blockDim.x * 2) {
    float2 ev = __half22float2(*reinterpret_cast<const __half2*>(e + c));
    *reinterpret_cast<__half2*>(xp + c) = __floats2half2_rn(ev.x * scale + pp[c], ev.y * scale + pp[c + 1]);
  }
}

// ------------------------------------------------------------------------------------------- self attention (1 token)
// one warp per (row, head); 4 heads per CTA
__global__ void __launch_bounds__(128) decode_self_attn_kernel(const elem_t* __restrict__ qkv, elem_t* __restrict__ kcache,
                                                               elem_t* __restrict__ vcache, const int* __restrict__ anc,
                                                               int anc_ld, const int* __restrict__ step_ptr, int max_len,
                                                               elem_t* __restrict__ out, int rows, int heads) {
  extern __shared__ float sc_all[];  // [4][max_len]
  const int step = *step_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x;
  const int h = blockIdx.y * 4 + warp;
  if (h >= heads) return;
  const int dim = heads * HD;
  float* sc = sc_all + warp * max_len;
  const elem_t* qp = qkv + (long long)r * 3 * dim + h * HD;
  const elem_t* knew = qp + dim;
  const elem_t* vnew = qp + 2 * dim;
  // persist the new K/V (slot r, position step)
  {
    elem_t* kd = kcache + ((long long)step * rows + r) * dim + h * HD;
    elem_t* vd = vcache + ((long long)step * rows + r) * dim + h * HD;
    reinterpret_cast<__half2*>(kd)[lane] = reinterpret_cast<const __half2*>(knew)[lane];
    reinterpret_cast<__half2*>(vd)[lane] = reinterpret_cast<const __half2*>(vnew)[lane];
  }
  float q[HD];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    uint4 u = *reinterpret_cast<const uint4*>(qp + i * 8);
    const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(hh[e]);
      q[i * 8 + 2 * e] = f.x;
      q[i * 8 + 2 * e + 1] = f.y;
    }
  }
  float mx = -INFINITY;
  for (int t = lane; t <= step; t += 32) {
    const elem_t* kp = (t == step) ? knew : kcache + ((long long)t * rows + anc[(long long)r * anc_ld + t]) * dim + h * HD;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      uint4 u = *reinterpret_cast<const uint4*>(kp + i * 8);
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(hh[e]);
        acc += q[i * 8 + 2 * e] * f.x + q[i * 8 + 2 * e + 1] * f.y;
      }
    }
    acc *= 0.125f;
    sc[t] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t <= step; t += 32) {
    float p = __expf(sc[t] - mx);
    sc[t] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  for (int t = 0; t <= step; ++t) {
    const elem_t* vp = (t == step) ? vnew : vcache + ((long long)t * rows + anc[(long long)r * anc_ld + t]) * dim + h * HD;
    float2 f = __half22float2(reinterpret_cast<const __half2*>(vp)[lane]);
    const float p = sc[t];
    o0 += p * f.x;
    o1 += p * f.y;
  }
  const float inv = 1.f / sum;
  reinterpret_cast<__half2*>(out + (long long)r * dim + h * HD)[lane] = __floats2half2_rn(o0 * inv, o1 * inv);
}

// ------------------------------------------------------------------------------------------- cross attention (1 token)
__global__ void __launch_bounds__(128) decode_cross_attn_kernel(const elem_t* __restrict__ q, const elem_t* __restrict__ k,
                                                                const elem_t* __restrict__ v, long long kv_ld,
                                                                const int* __restrict__ enc_lens, int s_enc,
                                                                elem_t* __restrict__ out, int rows, int beam, int heads) {
  extern __shared__ float sc_all[];  // [4][s_enc]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x;
  const int h = blockIdx.y * 4 + warp;
  if (h >= heads) return;
  const int dim = heads * HD;
  const int b = r / beam;
  const int len = enc_lens ? min(enc_lens[b], s_enc) : s_enc;
  float* sc = sc_all + warp * s_enc;
  const elem_t* qp = q + (long long)r * dim + h * HD;
  float qv[HD];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    uint4 u = *reinterpret_cast<const uint4*>(qp + i * 8);
    const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(hh[e]);
      qv[i * 8 + 2 * e] = f.x;
      qv[i * 8 + 2 * e + 1] = f.y;
    }
  }
  const elem_t* kb = k + (long long)b * s_enc * kv_ld + h * HD;
  const elem_t* vb = v + (long long)b * s_enc * kv_ld + h * HD;
  float mx = -INFINITY;
  for (int t = lane; t < len; t += 32) {
    const elem_t* kp = kb + (long long)t * kv_ld;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      uint4 u = *reinterpret_cast<const uint4*>(kp + i * 8);
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(hh[e]);
        acc += qv[i * 8 + 2 * e] * f.x + qv[i * 8 + 2 * e + 1] * f.y;
      }
    }
    acc *= 0.125f;
    sc[t] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < len; t += 32) {
    float p = __expf(sc[t] - mx);
    sc[t] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  for (int t = 0; t < len; ++t) {
    float2 f = __half22float2(reinterpret_cast<const __half2*>(vb + (long long)t * kv_ld)[lane]);
    const float p = sc[t];
    o0 += p * f.x;
    o1 += p * f.y;
  }
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  reinterpret_cast<__half2*>(out + (long long)r * dim + h * HD)[lane] = __floats2half2_rn(o0 * inv, o1 * inv);
}

// ------------------------------------------------------------------------------------------- log-softmax stats + top-K
constexpr int TK_MAX = 16;
constexpr int TK_THREADS = 512;

__global__ void __launch_bounds__(TK_THREADS) logits_topk_kernel(const float* __restrict__ logits, long long ld, int vocab,
                                                                 int pad_idx, int eos_idx, int unk_idx, float unk_penalty, int K,
                                                                 float* __restrict__ cand_val, int* __restrict__ cand_idx,
                                                                 float* __restrict__ eos_lprob) {
  const int r = blockIdx.x;
  const float* row = logits + (long long)r * ld;
  float tv[TK_MAX];
  int ti[TK_MAX];
#pragma unroll
  for (int i = 0; i < TK_MAX; ++i) { tv[i] = -INFINITY; ti[i] = 0x7fffffff; }
  float mx = -INFINITY, sum = 0.f;
  for (int i = threadIdx.x; i < vocab; i += TK_THREADS) {
    const float x = row[i];
    // online softmax statistics over the raw logits
    if (x > mx) { sum = sum * __expf(mx - x) + 1.f; mx = x; }
    else sum += __expf(x - mx);
    float cv = x;
    if (i == pad_idx) cv = -INFINITY;
    else if (i == unk_idx) cv = x - unk_penalty;
    if (cv > tv[TK_MAX - 1] || (cv == tv[TK_MAX - 1] && i < ti[TK_MAX - 1])) {
      // insertion into the sorted (descending, ties by lower index) per-thread list (keeps its best TK_MAX >= K)
      float v = cv; int id = i;
#pragma unroll
      for (int j = 0; j < TK_MAX; ++j) {
        if (v > tv[j] || (v == tv[j] && id < ti[j])) {
          float t1 = tv[j]; int t2 = ti[j];
          tv[j] = v; ti[j] = id; v = t1; id = t2;
        }
      }
    }
  }
  // block reduction of (max, sum)
  __shared__ float s_m[TK_THREADS / 32], s_s[TK_THREADS / 32];
  __shared__ float s_bv[TK_THREADS / 32];
  __shared__ int s_bi[TK_THREADS / 32], s_bt[TK_THREADS / 32];
  __shared__ float s_lse;
  __shared__ int s_win_thread;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    float m2 = warp_max(mx);
    float s2 = warp_sum(sum * __expf(mx - m2));
    if (lane == 0) { s_m[warp] = m2; s_s[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float M = -INFINITY;
      for (int i = 0; i < TK_THREADS / 32; ++i) M = fmaxf(M, s_m[i]);
      float S = 0.f;
      for (int i = 0; i < TK_THREADS / 32; ++i) S += s_s[i] * __expf(s_m[i] - M);
      s_lse = M + logf(S);
      eos_lprob[r] = row[eos_idx] - s_lse;
    }
    __syncthreads();
  }
  const float lse = s_lse;
  // K rounds: pick the best head among all threads' sorted lists
  for (int round = 0; round < K; ++round) {
    float v = tv[0];
    int id = ti[0];
    int th = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float v2 = __shfl_xor_sync(0xffffffffu, v, o);
      int id2 = __shfl_xor_sync(0xffffffffu, id, o);
      int th2 = __shfl_xor_sync(0xffffffffu, th, o);
      if (v2 > v || (v2 == v && id2 < id)) { v = v2; id = id2; th = th2; }
    }
    if (lane == 0) { s_bv[warp] = v; s_bi[warp] = id; s_bt[warp] = th; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float bv = s_bv[0]; int bi = s_bi[0], bt = s_bt[0];
      for (int i = 1; i < TK_THREADS / 32; ++i)
        if (s_bv[i] > bv || (s_bv[i] == bv && s_bi[i] < bi)) { bv = s_bv[i]; bi = s_bi[i]; bt = s_bt[i]; }
      cand_val[(long long)r * K + round] = bv - lse;
      cand_idx[(long long)r * K + round] = bi;
      s_win_thread = bt;
    }
    __syncthreads();
    if (threadIdx.x == s_win_thread) {
      // pop the head of the winner's list (shift left)
#pragma unroll
      for (int j = 0; j < TK_MAX - 1; ++j) { tv[j] = tv[j + 1]; ti[j] = ti[j + 1]; }
      tv[TK_MAX - 1] = -INFINITY; ti[TK_MAX - 1] = 0x7fffffff;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------- beam bookkeeping
// one CTA per sentence; the step index is read from device memory so that one captured CUDA graph serves every step
__global__ void __launch_bounds__(128) beam_step_kernel(const sb_beam_t p) {
  const int b = blockIdx.x;
  const int beam = p.beam, K = p.K, ML = p.max_len;
  const int step = *p.step_ptr;
  __shared__ int s_parent[16];
  __shared__ int s_tok[16];
  __shared__ float s_score[16];
  __shared__ int s_active;
  if (step + 1 >= ML) return;  // search exhausted (graph replays beyond the last step are no-ops)
  if (threadIdx.x == 0) {
    int active = p.active[b];
    if (active) {
      const bool first_step = (step == p.prefix_len - 1);
      const int nsrc = first_step ? 1 : beam;  // first step: all beams are identical, use beam 0 (fairseq2.cpp:1499-1510)
      int used[16];
      bool forced_taken[16];
      for (int i = 0; i < nsrc; ++i) { used[i] = 0; forced_taken[i] = false; }
      const bool force_eos = (step == ML - 2);   // _tweak_lprobs fairseq2.cpp:1280-1291
      const bool block_eos = (step < p.min_len); // fairseq2.cpp:1273-1278
      int n_ongoing = 0, fin = p.fin_count[b];
      const int want = 2 * beam;
      for (int pick = 0; pick < want && n_ongoing < beam && active; ++pick) {
        float best = -INFINITY; int bb = -1, tok = -1;
        for (int i = 0; i < nsrc; ++i) {
          const int row = b * beam + i;
          const float base = p.scores[(long long)row * ML + step];
          float v; int t;
          if (force_eos) {
            if (forced_taken[i]) continue;
            v = p.eos_lprob[row] + base; t = p.eos_idx;
          } else {
            int u = used[i];
            while (u < K && block_eos && p.cand_idx[(long long)row * K + u] == p.eos_idx) ++u;
            used[i] = u;
            if (u >= K) continue;
            v = p.cand_val[(long long)row * K + u] + base; t = p.cand_idx[(long long)row * K + u];
          }
          // order: higher value first; ties by lower flat index (beam*V + token)
          if (bb < 0 || v > best || (v == best && ((long long)i * p.vocab + t) < ((long long)bb * p.vocab + tok))) {
            best = v; bb = i; tok = t;
          }
        }
        if (bb < 0) break;
        if (force_eos) forced_taken[bb] = true; else used[bb]++;
        if (tok == p.eos_idx && best != -INFINITY) {
          // _finalize_hypothesis (fairseq2.cpp:1310-1350)
          const int row = b * beam + bb;
          int* dst = p.fin_seqs + ((long long)b * beam + fin) * ML;
          for (int t = 0; t <= step; ++t) dst[t] = p.seqs[(long long)row * ML + t];
          dst[step + 1] = tok;
          p.fin_len[b * beam + fin] = step + 2;
          p.fin_score[b * beam + fin] = best / powf((float)(step + 1), p.len_penalty);
          ++fin;
          if (fin == beam) active = 0;
          continue;
        }
        s_parent[n_ongoing] = bb; s_tok[n_ongoing] = tok; s_score[n_ongoing] = best;
        ++n_ongoing;
      }
      p.fin_count[b] = fin;
      if (!active) {
        p.active[b] = 0;
        if (p.n_active) atomicSub(p.n_active, 1);
      } else {
        // fewer than `beam` continuations can only happen with -inf scores; pad by repeating the last one
        for (int i = n_ongoing; i < beam; ++i) {
          s_parent[i] = n_ongoing > 0 ? s_parent[n_ongoing - 1] : 0;
          s_tok[i] = n_ongoing > 0 ? s_tok[n_ongoing - 1] : p.eos_idx;
          s_score[i] = -INFINITY;
        }
      }
    }
    s_active = active;
  }
  __syncthreads();
  if (!s_active) return;
  // in-place beam reorder, column by column: thread t reads column t of every parent row, then writes it back
  for (int t = threadIdx.x; t <= step + 1; t += blockDim.x) {
    int sv[16]; float cv[16]; int av[16];
    for (int i = 0; i < beam; ++i) {
      const long long src = (long long)(b * beam + s_parent[i]) * ML + t;
      sv[i] = p.seqs[src]; cv[i] = p.scores[src]; av[i] = p.anc[src];
      if (t == step) av[i] = b * beam + s_parent[i];  // K/V of position `step` live in the parent's slot
      if (t == step + 1) { sv[i] = s_tok[i]; cv[i] = s_score[i]; }
    }
    for (int i = 0; i < beam; ++i) {
      const long long dst = (long long)(b * beam + i) * ML + t;
      p.seqs[dst] = sv[i]; p.scores[dst] = cv[i]; p.anc[dst] = av[i];
    }
  }
}

__global__ void step_advance_kernel(int* step_ptr) { *step_ptr += 1; }

}  // namespace sb

extern "C" int sb_embed_step(const int32_t* seqs, int32_t seq_ld, const int32_t* step_ptr, const void* embed,
                             const void* pos_table, float scale, void* x, int32_t rows, int32_t dim, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(seqs && step_ptr && embed && pos_table && x && rows > 0 && dim % 2 == 0, SB_EINVAL, "sb_embed_step: bad args");
  embed_kernel<<<dim3(rows, 1), 128, 0, (cudaStream_t)stream>>>(seqs, seq_ld, step_ptr, 1, (const elem_t*)embed,
                                                                (const float*)pos_table, scale, (elem_t*)x, dim);
  SB_LAUNCH_OK();
  return SB_OK;
}

extern "C" int sb_embed_seq(const int32_t* ids, int32_t ids_ld, int32_t L, const void* embed, const void* pos_table,
                            float scale, void* x, int32_t rows, int32_t dim, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(ids && embed && pos_table && x && rows > 0 && L > 0 && dim % 2 == 0, SB_EINVAL, "sb_embed_seq: bad args");
  embed_kernel<<<dim3(rows, L), 128, 0, (cudaStream_t)stream>>>(ids, ids_ld, nullptr, L, (const elem_t*)embed,
                                                                (const float*)pos_table, scale, (elem_t*)x, dim);
  SB_LAUNCH_OK();
  return SB_OK;
}

extern "C" int sb_decode_self_attn(const void* qkv, void* kcache, void* vcache, const int32_t* anc, int32_t anc_ld,
                                   const int32_t* step_ptr, int32_t max_len, void* out, int32_t rows, int32_t heads,
                                   sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(qkv && kcache && vcache && anc && out && step_ptr && rows > 0 && heads > 0 && max_len > 0, SB_EINVAL,
             "sb_decode_self_attn: bad args");
  size_t smem = (size_t)4 * max_len * sizeof(float);
  SB_REQUIRE(smem <= 48 * 1024, SB_ENOSUP, "sb_decode_self_attn: max_len %d too large", max_len);
  decode_self_attn_kernel<<<dim3(rows, (heads + 3) / 4), 128, smem, (cudaStream_t)stream>>>(
      (const elem_t*)qkv, (elem_t*)kcache, (elem_t*)vcache, anc, anc_ld, step_ptr, max_len, (elem_t*)out, rows, heads);
  SB_LAUNCH_OK();
  return SB_OK;
}

extern "C" int sb_decode_cross_attn(const void* q, const void* k, const void* v, int64_t kv_ld, const int32_t* enc_lens,
                                    int32_t s_enc, void* out, int32_t rows, int32_t beam, int32_t heads, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(q && k && v && out && rows > 0 && heads > 0 && s_enc > 0 && beam > 0, SB_EINVAL, "sb_decode_cross_attn: bad args");
  size_t smem = (size_t)4 * s_enc * sizeof(float);
  SB_REQUIRE(smem <= 48 * 1024, SB_ENOSUP, "sb_decode_cross_attn: s_enc %d too large", s_enc);
  decode_cross_attn_kernel<<<dim3(rows, (heads + 3) / 4), 128, smem, (cudaStream_t)stream>>>(
      (const elem_t*)q, (const elem_t*)k, (const elem_t*)v, kv_ld, enc_lens, s_enc, (elem_t*)out, rows, beam, heads);
  SB_LAUNCH_OK();
  return SB_OK;
}

extern "C" int sb_logits_topk(const float* logits, int64_t ld, int32_t rows, int32_t vocab, int32_t pad_idx, int32_t eos_idx,
                              int32_t unk_idx, float unk_penalty, int32_t K, float* cand_val, int32_t* cand_idx,
                              float* eos_lprob, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(logits && cand_val && cand_idx && eos_lprob && rows > 0 && vocab > 0, SB_EINVAL, "sb_logits_topk: bad args");
  SB_REQUIRE(K > 0 && K <= TK_MAX, SB_ENOSUP, "sb_logits_topk: K=%d unsupported (<= %d)", K, TK_MAX);
  logits_topk_kernel<<<rows, TK_THREADS, 0, (cudaStream_t)stream>>>(logits, ld, vocab, pad_idx, eos_idx, unk_idx, unk_penalty, K,
                                                                    cand_val, cand_idx, eos_lprob);
  SB_LAUNCH_OK();
  return SB_OK;
}

extern "C" int sb_beam_step(const sb_beam_t* p, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(p && p->batch > 0 && p->beam > 0 && p->beam <= 16 && p->K > 0, SB_EINVAL, "sb_beam_step: bad args");
  SB_REQUIRE(p->step_ptr != nullptr, SB_EINVAL, "sb_beam_step: step_ptr is null");
  beam_step_kernel<<<p->batch, 128, 0, (cudaStream_t)stream>>>(*p);
  SB_LAUNCH_OK();
  return SB_OK;
}

extern "C" int sb_step_advance(int32_t* step_ptr, sb_stream_t stream) {
  SB_REQUIRE(step_ptr != nullptr, SB_EINVAL, "sb_step_advance: null");
  sb::step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_ptr);
  SB_LAUNCH_OK();
  return SB_OK;
}
