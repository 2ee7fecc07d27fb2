// genrec_b200 - trie-constrained beam step of TIGER's generate() on the device (genrec/models/tiger.py:312-452).
//
// The reference masks the logits with a Python double loop over (batch, beam) walking a dict trie, and picks the next beams with a
// per-batch Python loop calling .item() on every candidate: the decode is host-bound.  Here the trie is a CSR over node ids
// (child_off [n_nodes + 1], child_tok / child_node [n_edges], children sorted by token, root = node 0, dead = -1) and a decode step is
// two launches:
//   trie_log_softmax_kernel   legal-token mask from the node's children (or the step's vocabulary range without a trie),
//                             masked_fill(-1e32) / temperature, softmax AND log_softmax                 (tiger.py:364-384)
//   beam_select_kernel        total = beam_logp + cand_logp, descending sort, first K candidates whose token sequence is new,
//                             -1e32 / zero-sequence / root-node fillers, trie descent of the survivors   (tiger.py:386-441)
// Candidate sampling stays torch.multinomial on the probabilities produced here (the reference's RNG stream is part of its output).
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"   // exp_accurate

namespace grb {

struct TrieCsr {
    const int* child_off;    // [n_nodes + 1]
    const int* child_tok;    // [n_edges] raw token ids (0 .. num_embeddings-1), ascending inside a node
    const int* child_node;   // [n_edges]
    int n_nodes;
};

// one CTA per beam row.  node < 0 (dead) or a leaf: no legal token -> every entry is -1e32 / T, i.e. the uniform distribution, exactly
// as the reference's masked_fill produces it.  use_trie = 0: legal = [vocab_offset, vocab_offset + num_emb), the rest is -inf.
__global__ void __launch_bounds__(256) trie_log_softmax_kernel(const float* __restrict__ logits, int V, const int* __restrict__ node,
                                                               TrieCsr trie, int use_trie, int vocab_offset, int num_emb, float temperature,
                                                               float* __restrict__ probs, float* __restrict__ logp) {
    pdl_wait();
    extern __shared__ unsigned beam_smem[];
    unsigned* legal = beam_smem;                       // bitmap, ceil(V / 32) words
    __shared__ float red[8];
    __shared__ float bc[2];
    const int row = blockIdx.x, tid = threadIdx.x;
    const int words = (V + 31) >> 5;
    for (int w = tid; w < words; w += 256) legal[w] = 0u;
    __syncthreads();
    if (use_trie) {
        const int nd = node[row];
        if (nd >= 0 && nd < trie.n_nodes) {
            const int c0 = trie.child_off[nd], c1 = trie.child_off[nd + 1];
            for (int c = c0 + tid; c < c1; c += 256) {
                const int v = vocab_offset + trie.child_tok[c];
                if (v >= 0 && v < V) atomicOr(&legal[v >> 5], 1u << (v & 31));
            }
        }
    } else {
        for (int v = vocab_offset + tid; v < vocab_offset + num_emb && v < V; v += 256)
            if (v >= 0) atomicOr(&legal[v >> 5], 1u << (v & 31));
    }
    __syncthreads();
    const float* x = logits + (size_t)row * V;
    const float fill = use_trie ? -1e32f : -INFINITY;
    // logits / temperature exactly as the reference rounds it (a true division, and no contraction of the later "- max" into an FMA:
    // a dead node's row is -1e32 / T everywhere and must come out as EXACTLY equal entries, i.e. the uniform distribution)
    auto val = [&](int v) { return __fdiv_rn(((legal[v >> 5] >> (v & 31)) & 1u) ? x[v] : fill, temperature); };
    float m = -INFINITY;
    for (int v = tid; v < V; v += 256) m = fmaxf(m, val(v));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) { float t = red[0]; for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]); bc[0] = t; }
    __syncthreads();
    m = bc[0];
    float s = 0.f;
    for (int v = tid; v < V; v += 256) s += exp_accurate(__fsub_rn(val(v), m));
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; bc[1] = t; }
    __syncthreads();
    const float sum = bc[1];
    const float lsum = 0.6931471805599453f * __log2f(sum);
    for (int v = tid; v < V; v += 256) {
        const float d = __fsub_rn(val(v), m);
        probs[(size_t)row * V + v] = __fdiv_rn(exp_accurate(d), sum);
        logp[(size_t)row * V + v] = d - lsum;
    }
}

struct BeamSelectArgs {
    const long long* beam_seqs;   // [B, K, S]
    const float* beam_logps;      // [B, K]
    const long long* cand_tok;    // [B, K, KK]  raw token ids (vocabulary index - vocab_offset)
    const float* cand_logp;       // [B, K, KK]
    const int* nodes;             // [B, K] (nullable: no trie)
    TrieCsr trie;
    int K, KK, S;
    long long* new_seqs;          // [B, K, S + 1]
    float* new_logps;             // [B, K]
    int* new_nodes;               // [B, K] (nullable)
};
constexpr int BEAM_MAX_CAND = 1024;
// one CTA (1024 threads) per batch row; K <= 32, K * KK <= 1024.  Order of equal totals: lower flat candidate index first.
__global__ void __launch_bounds__(BEAM_MAX_CAND) beam_select_kernel(BeamSelectArgs a) {
    pdl_wait();
    __shared__ float s_key[BEAM_MAX_CAND];
    __shared__ int s_idx[BEAM_MAX_CAND];
    __shared__ int s_cls[32];
    __shared__ int s_pick[32];
    __shared__ int s_npick;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = a.K * a.KK;
    const long long* seqs = a.beam_seqs + (size_t)b * a.K * a.S;
    // parents whose token sequences are identical produce identical children: class = the smallest such parent
    if (tid < a.K) {
        int c = tid;
        for (int p = 0; p < tid; ++p) {
            bool same = true;
            for (int t = 0; t < a.S; ++t) same = same && seqs[(size_t)p * a.S + t] == seqs[(size_t)tid * a.S + t];
            if (same) { c = p; break; }
        }
        s_cls[tid] = c;
    }
    if (tid < n) {
        const int p = tid / a.KK;
        s_key[tid] = a.beam_logps[(size_t)b * a.K + p] + a.cand_logp[(size_t)b * n + tid];    // (tiger.py:393)
        s_idx[tid] = tid;
    } else {
        s_key[tid] = -INFINITY;
        s_idx[tid] = 0x7fffffff;
    }
    __syncthreads();
    // bitonic sort, descending by key, ascending by index among equal keys (NaN never occurs: log-probabilities)
    for (int k = 2; k <= BEAM_MAX_CAND; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = tid ^ j;
            if (o > tid) {
                const float ka = s_key[tid], kb = s_key[o];
                const int ia = s_idx[tid], ib = s_idx[o];
                const bool a_first = ka > kb || (ka == kb && ia < ib);   // a belongs before b in the final order
                const bool up = (tid & k) == 0;
                if (up ? !a_first : a_first) { s_key[tid] = kb; s_key[o] = ka; s_idx[tid] = ib; s_idx[o] = ia; }
            }
            __syncthreads();
        }
    }
    // greedy scan by one warp: lane l remembers the l-th pick
    if (tid < 32) {
        int my_cls = -1; long long my_tok = -1;
        int npick = 0;
        for (int j = 0; j < n && npick < a.K; ++j) {
            const int ci = s_idx[j];
            const int p = ci / a.KK;
            const long long t = a.cand_tok[(size_t)b * n + ci];
            const int c = s_cls[p];
            const bool dup = tid < npick && my_cls == c && my_tok == t;
            if (__ballot_sync(0xffffffffu, dup) == 0u) {
                if (tid == npick) { my_cls = c; my_tok = t; s_pick[npick] = j; }
                ++npick;
            }
        }
        if (tid == 0) s_npick = npick;
    }
    __syncthreads();
    const int npick = s_npick;
    if (tid < a.K) {
        const int k = tid;
        long long* out = a.new_seqs + ((size_t)b * a.K + k) * (a.S + 1);
        if (k < npick) {
            const int j = s_pick[k];
            const int ci = s_idx[j];
            const int p = ci / a.KK;
            const long long t = a.cand_tok[(size_t)b * n + ci];
            for (int s = 0; s < a.S; ++s) out[s] = seqs[(size_t)p * a.S + s];
            out[a.S] = t;
            a.new_logps[(size_t)b * a.K + k] = s_key[j];
            if (a.new_nodes) {
                int nd = a.nodes ? a.nodes[(size_t)b * a.K + p] : -1, child = -1;
                if (nd >= 0 && nd < a.trie.n_nodes) {
                    int lo = a.trie.child_off[nd], hi = a.trie.child_off[nd + 1];
                    while (lo < hi) {                       // children are sorted by token
                        const int mid = (lo + hi) >> 1;
                        const int tk = a.trie.child_tok[mid];
                        if (tk == t) { child = a.trie.child_node[mid]; break; }
                        if (tk < t) lo = mid + 1; else hi = mid;
                    }
                }
                a.new_nodes[(size_t)b * a.K + k] = child;   // parent_node.get(tid, DEAD_NODE)            (tiger.py:419-421)
            }
        } else {                                            // fewer than K distinct candidates           (tiger.py:423-429)
            for (int s = 0; s <= a.S; ++s) out[s] = 0;
            a.new_logps[(size_t)b * a.K + k] = -1e32f;
            if (a.new_nodes) a.new_nodes[(size_t)b * a.K + k] = 0;
        }
    }
}

}  // namespace grb
