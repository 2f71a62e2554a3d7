// MusicGen LM decode step for B200 (sm_100a).
//
// One decode step = LMModel.forward on one token per row (audiocraft/models/lm.py:221-268) + CFG mix + sampling
// (lm.py:393-418) + the delay-pattern write-back (lm.py:553-562), replayed as ONE CUDA graph per step with every
// step-dependent quantity (position, tokens) resident on the device.
//
// The step is HBM-bound: all layer weights (fp16, 3.65 GB for medium) and the KV cache are read once per step and
// arithmetic intensity is ~rows FLOP/B.  Design consequences:
//   * weights stay in the reference's [out][in] fp16 layout and stream straight from HBM into tensor-core
//     fragments (ld.global.nc.L1::no_allocate 128-bit), never through shared memory: a 16x32 block of W is the
//     A operand of two m16n8k16 MMAs, the activations (a few KB, L1/L2 resident) are the B operand, so the tile
//     is 16 output features x (8*NT) rows and nothing is wasted on padding rows up to 128.
//   * every GEMM spreads its weight matrix over >= 2 CTAs per SM; small-N GEMMs split K across CTAs and the
//     partial sums are reduced (in a fixed order: bit-reproducible) by the consumer kernel, which is the residual
//     add + LayerNorm, so that reduction costs no extra pass.
//   * K/V go from the QKV GEMM epilogue straight into the cache; cross-attention K/V are computed once per
//     generate() instead of every step (the reference recomputes them, transformer.py:355-357).
//   * attention for one query token: one CTA per (row, head) streaming K then V with 128-bit loads.
#include "common.cuh"
#include <math.h>
#include <new>

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float half_round(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

__device__ __forceinline__ float block_sum(float v, float* red) {  // red: >= 33 floats
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// ------------------------------------------------------------------------------------------------ embed + sin pos
// x[r] = sum_k emb_k[seq[b,k,pos]] + pos_scale * [cos(pos/f_i), sin(pos/f_i)]   (lm.py:244, transformer.py:70-89,701-705)
__global__ void __launch_bounds__(256) lm_embed_kernel(const __half* __restrict__ emb, const float* __restrict__ inv_freq,
                                                       const int64_t* __restrict__ seq, const int* __restrict__ P,
                                                       float* __restrict__ x, int d, int n_q, int card, int max_seq,
                                                       int batch, float pos_scale) {
    const int r = blockIdx.x, b = r % batch, pos = P[0];
    __shared__ int tok[16];
    if (threadIdx.x < n_q) {
        long long t = seq[((size_t)b * n_q + threadIdx.x) * max_seq + pos];
        tok[threadIdx.x] = (int)(t < 0 ? card : (t > card ? card : t));
    }
    __syncthreads();
    const int half_d = d >> 1;
    for (int i = threadIdx.x; i < d; i += 256) {
        float v = 0.f;
        for (int k = 0; k < n_q; ++k) v += __half2float(emb[((size_t)k * (card + 1) + tok[k]) * d + i]);
        const int j = i < half_d ? i : i - half_d;
        const float phase = (float)pos / inv_freq[j];
        v += pos_scale * (i < half_d ? cosf(phase) : sinf(phase));
        x[(size_t)r * d + i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ residual + LN
// x[r] += sum_s part[s][r] (fixed order), then h16[r] = LayerNorm(x[r]) * gamma + beta  (eps 1e-5, fp32 statistics).
constexpr int LN_MAX_PER_THREAD = 16;  // d <= 4096
__global__ void __launch_bounds__(256) lm_ln_kernel(float* __restrict__ x, const float* __restrict__ part, int nsplit,
                                                    size_t split_stride, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, __half* __restrict__ out, int d) {
    __shared__ float red[33];
    const int r = blockIdx.x;
    float v[LN_MAX_PER_THREAD];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_PER_THREAD; ++j) {
        const int i = threadIdx.x + j * 256;
        v[j] = 0.f;
        if (i < d) {
            float a = x[(size_t)r * d + i];
            for (int sp = 0; sp < nsplit; ++sp) a += part[sp * split_stride + (size_t)r * d + i];
            if (nsplit) x[(size_t)r * d + i] = a;
            v[j] = a;
            s += a;
        }
    }
    const float mean = block_sum(s, red) / d;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_PER_THREAD; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < d) { float c = v[j] - mean; q = fmaf(c, c, q); }
    }
    const float rstd = 1.f / sqrtf(block_sum(q, red) / d + 1e-5f);
#pragma unroll
    for (int j = 0; j < LN_MAX_PER_THREAD; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < d) out[(size_t)r * d + i] = __float2half_rn((v[j] - mean) * rstd * gamma[i] + beta[i]);
    }
}

// ------------------------------------------------------------------------------------------------ skinny GEMM
enum { EPI_PARTIAL = 0, EPI_QKV = 1, EPI_GELU = 2, EPI_F32 = 3, EPI_CROSSKV = 4 };

struct GemmParams {
    const __half* W;  // [N][K] fp16, reference layout
    const __half* X;  // [8*NT][K] fp16, rows >= `rows` are zero
    int N, K, rows, kb_per_warp;
    float* out_f32; int ld_out; size_t split_stride;  // PARTIAL / F32
    __half* out_f16;                                   // GELU
    float* q32; __half* kc; __half* vc; int d, H, cache_len; const int* pos;  // QKV / CROSSKV
    int text_len, row0;                                                      // CROSSKV
};

template <int NT, int EPI>
__global__ void __launch_bounds__(256) lm_gemm_kernel(GemmParams p) {
    constexpr int U = NT <= 2 ? 4 : (NT <= 4 ? 2 : 1);
    constexpr int RP = 8 * NT + 1;
    __shared__ float red[8 * 16 * RP];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c4 = lane & 3;
    const int f0 = blockIdx.x * 16;
    const int nkb = p.K >> 5;
    const int kb0 = min(nkb, ((int)blockIdx.y * 8 + warp) * p.kb_per_warp);
    const int kb1 = min(nkb, kb0 + p.kb_per_warp);

    float c[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;

    const __half* w0 = p.W + (size_t)(f0 + g) * p.K + 8 * c4;
    const __half* w1 = w0 + (size_t)8 * p.K;
    const __half* xr = p.X + (size_t)g * p.K + 8 * c4;

    for (int kb = kb0; kb < kb1; kb += U) {
        uint4 wa[U], wb[U], xv[U][NT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (kb + u < kb1) {
                wa[u] = ld_stream_u4(w0 + (size_t)(kb + u) * 32);
                wb[u] = ld_stream_u4(w1 + (size_t)(kb + u) * 32);
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    xv[u][j] = *reinterpret_cast<const uint4*>(xr + (size_t)(8 * j) * p.K + (size_t)(kb + u) * 32);
            } else {
                wa[u] = wb[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int j = 0; j < NT; ++j) xv[u][j] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                mma16816(c[j], wa[u].x, wb[u].x, wa[u].y, wb[u].y, xv[u][j].x, xv[u][j].y);
                mma16816(c[j], wa[u].z, wb[u].z, wa[u].w, wb[u].w, xv[u][j].z, xv[u][j].w);
            }
    }
    // cross-warp (split-K inside the CTA) reduction in a fixed order
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        red[(warp * 16 + g) * RP + 8 * j + 2 * c4] = c[j][0];
        red[(warp * 16 + g) * RP + 8 * j + 2 * c4 + 1] = c[j][1];
        red[(warp * 16 + g + 8) * RP + 8 * j + 2 * c4] = c[j][2];
        red[(warp * 16 + g + 8) * RP + 8 * j + 2 * c4 + 1] = c[j][3];
    }
    __syncthreads();
    for (int idx = tid; idx < 16 * 8 * NT; idx += 256) {
        const int row = idx >> 4, feat = idx & 15;
        if (row >= p.rows) continue;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[(w * 16 + feat) * RP + row];
        const int n = f0 + feat;
        if (EPI == EPI_PARTIAL) {
            p.out_f32[blockIdx.y * p.split_stride + (size_t)row * p.ld_out + n] = v;
        } else if (EPI == EPI_F32) {
            p.out_f32[(size_t)row * p.ld_out + n] = v;
        } else if (EPI == EPI_GELU) {
            p.out_f16[(size_t)row * p.ld_out + n] = __float2half_rn(gelu_erf(half_round(v)));
        } else if (EPI == EPI_QKV) {
            if (n < p.d) {
                p.q32[(size_t)row * p.d + n] = v;
            } else {
                const int which = (n - p.d) / p.d, nn = n % p.d, h = nn >> 6, dd = nn & 63;
                __half* cache = which ? p.vc : p.kc;
                cache[(((size_t)row * p.H + h) * p.cache_len + p.pos[0]) * 64 + dd] = __float2half_rn(v);
            }
        } else {  // EPI_CROSSKV: GEMM rows are (row, text position) pairs
            const int R = p.row0 + row, r = R / p.text_len, tc = R % p.text_len;
            const int which = n / p.d, nn = n % p.d, h = nn >> 6, dd = nn & 63;
            __half* cache = which ? p.vc : p.kc;
            cache[(((size_t)r * p.H + h) * p.cache_len + tc) * 64 + dd] = __float2half_rn(v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ attention (1 query)
struct AttnParams {
    const float* q; int q_nsplit; size_t q_split_stride;  // q[s][row][d] fp32 partial sums
    const __half* kc; const __half* vc; __half* out;
    int H, d, cache_len; const int* pos; int fixed_len; float scale;
};

__global__ void __launch_bounds__(128) lm_attn_kernel(AttnParams p) {
    extern __shared__ float sc[];  // [len] scores, then probabilities
    __shared__ float red[33];
    __shared__ float osm[4][64];
    const int h = blockIdx.x, row = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sl = lane & 7, pg = lane >> 3;
    const int n = p.fixed_len > 0 ? p.fixed_len : p.pos[0] + 1;

    float q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float a = 0.f;
        for (int s = 0; s < p.q_nsplit; ++s) a += p.q[s * p.q_split_stride + (size_t)row * p.d + h * 64 + sl * 8 + e];
        q[e] = half_round(a);
    }
    const size_t base = ((size_t)row * p.H + h) * p.cache_len * 64 + sl * 8;
    const __half* kb = p.kc + base;
    const __half* vb = p.vc + base;

    float lmax = -INFINITY;
    for (int p0 = warp * 4 + pg; p0 < n; p0 += 64) {
        uint4 kv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            kv[u] = (p0 + 16 * u < n) ? ld_stream_u4(kb + (size_t)(p0 + 16 * u) * 64) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const __half2* k2 = reinterpret_cast<const __half2*>(&kv[u]);
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(k2[e]);
                s = fmaf(q[2 * e], f.x, s);
                s = fmaf(q[2 * e + 1], f.y, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            const int pp = p0 + 16 * u;
            if (pp < n) {
                s *= p.scale;
                if (sl == 0) sc[pp] = s;
                lmax = fmaxf(lmax, s);
            }
        }
    }
    const float m = block_max(lmax, red);  // (block_max's barriers also publish sc[])
    float lsum = 0.f;
    for (int i = tid; i < n; i += 128) {
        float e = expf(sc[i] - m);
        sc[i] = e;
        lsum += e;
    }
    const float denom = block_sum(lsum, red);

    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int p0 = warp * 4 + pg; p0 < n; p0 += 64) {
        uint4 vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            vv[u] = (p0 + 16 * u < n) ? ld_stream_u4(vb + (size_t)(p0 + 16 * u) * 64) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pp = p0 + 16 * u;
            const float w = pp < n ? sc[pp] : 0.f;
            const __half2* v2 = reinterpret_cast<const __half2*>(&vv[u]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(v2[e]);
                acc[2 * e] = fmaf(w, f.x, acc[2 * e]);
                acc[2 * e + 1] = fmaf(w, f.y, acc[2 * e + 1]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (pg == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) osm[warp][sl * 8 + e] = acc[e];
    }
    __syncthreads();
    if (tid < 64) {
        float o = (osm[0][tid] + osm[1][tid]) + (osm[2][tid] + osm[3][tid]);
        p.out[(size_t)row * p.d + h * 64 + tid] = __float2half_rn(o / denom);
    }
}

// ------------------------------------------------------------------------------------------------ sampling
// Philox4x32-10 (counter-based; one 4-word block per 4 candidates) -> 24-bit uniforms -> Exponential(1).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float exp1_noise(uint64_t seed, uint32_t step, uint32_t stream, uint32_t i) {
    uint4 r = philox4x32_10(make_uint4(i >> 2, stream, step, 0x5a17u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint32_t w = (i & 3) == 0 ? r.x : ((i & 3) == 1 ? r.y : ((i & 3) == 2 ? r.z : r.w));
    float u = ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1), never 0 or 1
    return -logf(u);
}

struct SampleParams {
    const float* logits;  // [rows][n_q*card]
    const float* noise;   // [batch][n_q][card] or NULL
    float* logits_out;    // [batch][n_q][card] CFG-mixed logits or NULL
    int64_t* seq; const uint8_t* seq_mask; const int* pos; int max_seq;  // in-loop write-back (seq may be NULL)
    int64_t* tokens;      // [batch][n_q] stand-alone output (may be NULL)
    int batch, rows, n_q, card, NP;
    int use_sampling, top_k; float temp, top_p, cfg_coef; uint64_t seed; uint32_t step;
};

// descending order, ties by ascending index
__device__ __forceinline__ bool before(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

__global__ void __launch_bounds__(1024) lm_sample_kernel(SampleParams p) {
    extern __shared__ float sm[];
    float* pr = sm;                 // [card] logits -> probabilities
    float* sv = pr + p.card;        // [NP] sort values
    int* si = (int*)(sv + p.NP);    // [NP] sort indices
    __shared__ float red[33];
    __shared__ float bestv[32];
    __shared__ int besti[32];
    __shared__ float s_scalar;
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const int card = p.card;
    const bool cfg = p.rows == 2 * p.batch;
    const float* lc = p.logits + ((size_t)b * p.n_q + k) * card;
    const float* lu = p.logits + ((size_t)(p.batch + b) * p.n_q + k) * card;
    const uint32_t step = p.pos ? (uint32_t)p.pos[0] : p.step;

    for (int i = tid; i < card; i += nt) {
        float l = lc[i];
        if (cfg) { float u = lu[i]; l = u + (l - u) * p.cfg_coef; }   // lm.py:399
        pr[i] = l;
        if (p.logits_out) p.logits_out[((size_t)b * p.n_q + k) * card + i] = l;
    }
    __syncthreads();

    const bool sampling = p.use_sampling && p.temp > 0.f;
    bool sorted_space = false;
    if (sampling) {
        float lm = -INFINITY;
        for (int i = tid; i < card; i += nt) { float l = pr[i] / p.temp; pr[i] = l; lm = fmaxf(lm, l); }
        const float m = block_max(lm, red);
        float ls = 0.f;
        for (int i = tid; i < card; i += nt) { float e = expf(pr[i] - m); pr[i] = e; ls += e; }
        const float s = block_sum(ls, red);
        for (int i = tid; i < card; i += nt) pr[i] = pr[i] / s;
        __syncthreads();
        const int kk = p.top_k > card ? card : p.top_k;
        if (p.top_p > 0.f || kk > 0) {
            for (int i = tid; i < p.NP; i += nt) { sv[i] = i < card ? pr[i] : -INFINITY; si[i] = i; }
            __syncthreads();
            for (int size = 2; size <= p.NP; size <<= 1)
                for (int j = size >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < p.NP; i += nt) {
                        const int l = i ^ j;
                        if (l > i) {
                            const bool fwd = (i & size) == 0;
                            float va = sv[i], vb = sv[l];
                            int ia = si[i], ib = si[l];
                            const bool ok = before(va, ia, vb, ib);
                            if (fwd ? !ok : ok) { sv[i] = vb; sv[l] = va; si[i] = ib; si[l] = ia; }
                        }
                    }
                    __syncthreads();
                }
            if (p.top_p > 0.f) {
                // utils.sample_top_p (utils/utils.py:125-141): sequential cumsum like torch's CPU kernel
                if (tid == 0) {
                    float cum = 0.f;
                    for (int i = 0; i < card; ++i) {
                        const float v = sv[i];
                        cum += v;
                        if (cum - v > p.top_p) sv[i] = 0.f;
                    }
                }
                __syncthreads();
                float ls2 = 0.f;
                for (int i = tid; i < card; i += nt) ls2 += sv[i];
                const float s2 = block_sum(ls2, red);
                for (int i = tid; i < card; i += nt) pr[i] = sv[i] / s2;  // pr now lives in sorted space
                sorted_space = true;
                __syncthreads();
            } else {
                // utils.sample_top_k (utils/utils.py:108-122): keep p >= k-th largest, renormalise
                if (tid == 0) s_scalar = sv[kk - 1];
                __syncthreads();
                const float kth = s_scalar;
                float ls2 = 0.f;
                for (int i = tid; i < card; i += nt) { float v = pr[i] >= kth ? pr[i] : 0.f; pr[i] = v; ls2 += v; }
                const float s2 = block_sum(ls2, red);
                for (int i = tid; i < card; i += nt) pr[i] = pr[i] / s2;
                __syncthreads();
            }
        }
        // torch.multinomial(num_samples=1): argmax_i p_i / q_i, q ~ Exponential(1)
        for (int i = tid; i < card; i += nt) {
            const float qn = p.noise ? p.noise[((size_t)b * p.n_q + k) * card + i]
                                     : exp1_noise(p.seed, step, (uint32_t)(b * p.n_q + k), (uint32_t)i);
            pr[i] = pr[i] / qn;
        }
        __syncthreads();
    }
    // first-max argmax over pr
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < card; i += nt) {
        const float v = pr[i];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { bestv[tid >> 5] = bv; besti[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (nt >> 5); ++w)
            if (bestv[w] > bv || (bestv[w] == bv && besti[w] < bi)) { bv = bestv[w]; bi = besti[w]; }
        if (bi == 0x7fffffff) bi = 0;
        int tok = sorted_space ? si[bi] : bi;
        if (p.tokens) p.tokens[(size_t)b * p.n_q + k] = tok;
        if (p.seq) {
            const int off = p.pos[0] + 1;
            if (off < p.max_seq) {
                if (!p.seq_mask[(size_t)k * p.max_seq + off]) tok = card;            // lm.py:555-556
                int64_t* dst = p.seq + ((size_t)b * p.n_q + k) * p.max_seq + off;
                if (*dst == -1) *dst = tok;                                           // lm.py:559-562
            }
        }
    }
}

__global__ void lm_advance_kernel(int* P) { P[0] += 1; }

__global__ void lm_f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n_valid, size_t n_total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) dst[i] = __float2half_rn(i < n_valid ? src[i] : 0.f);
}

// ------------------------------------------------------------------------------------------------ host side
struct acb_lm {
    acb_lm_config cfg;
    acb_lm_weights w;
    acb_lm_buffers buf;
    acb_lm_sampling samp;
    cudaStream_t capture_stream = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int batch = 0, rows = 0, rows_pad = 0, text_len = 0, seq_len = 0, sms = 148;
    int launches = 0;
    bool has_cross = false;
};

static int nt_for_rows(int rows) { return rows <= 8 ? 1 : (rows <= 16 ? 2 : (rows <= 32 ? 4 : 8)); }

template <int EPI>
static int launch_gemm(int nt, const GemmParams& p, int nsplit, cudaStream_t s) {
    dim3 grid(p.N / 16, nsplit);
    switch (nt) {
        case 1: lm_gemm_kernel<1, EPI><<<grid, 256, 0, s>>>(p); break;
        case 2: lm_gemm_kernel<2, EPI><<<grid, 256, 0, s>>>(p); break;
        case 4: lm_gemm_kernel<4, EPI><<<grid, 256, 0, s>>>(p); break;
        default: lm_gemm_kernel<8, EPI><<<grid, 256, 0, s>>>(p); break;
    }
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}

// number of K splits so that the weight matrix is spread over >= 2 CTAs per SM (no empty splits)
static int pick_split(int N, int K, int sms, bool allow) {
    if (!allow) return 1;
    int tiles = N / 16, nkb = K / 32;
    int s = acb_ceil_div(2 * sms, tiles);
    s = max(1, min(s, ACB_LM_MAX_SPLIT));
    s = min(s, max(1, nkb / 8));
    int kbpw = acb_ceil_div(nkb, 8 * s);
    return acb_ceil_div(nkb, 8 * kbpw);
}

static GemmParams base_gemm(const void* W, const void* X, int N, int K, int rows, int nsplit) {
    GemmParams p{};
    p.W = (const __half*)W;
    p.X = (const __half*)X;
    p.N = N; p.K = K; p.rows = rows;
    p.kb_per_warp = acb_ceil_div(K / 32, 8 * nsplit);
    return p;
}

#define ACB_TRY(expr) do { int rc_ = (expr); if (rc_ != ACB_OK) return rc_; } while (0)

static int enqueue_step(acb_lm* lm, cudaStream_t s, float* logits_out, int* n_launch, bool gemms_only = false) {
    const acb_lm_config& c = lm->cfg;
    const acb_lm_buffers& B = lm->buf;
    const int d = c.dim, ffn = c.ffn_dim, L = c.num_layers, H = c.num_heads, rows = lm->rows, nt = nt_for_rows(rows);
    const size_t part_stride = (size_t)lm->rows_pad * d;
    const size_t kv_layer = (size_t)c.max_rows * H * c.max_seq * 64;
    const size_t ckv_layer = (size_t)c.max_rows * H * c.max_text * 64;
    const float scale = 1.0f / sqrtf(64.f);
    int nl = 0;

    if (!gemms_only) {
        lm_embed_kernel<<<rows, 256, 0, s>>>((const __half*)lm->w.emb, lm->w.inv_freq, B.seq, B.pos, B.x, d, c.n_q, c.card,
                                             c.max_seq, lm->batch, c.pos_scale);
        ACB_LAUNCH_CHECK(); ++nl;
    }
    int pending = 0;  // split-K partial sums waiting to be folded into x by the next LN
    for (int l = 0; l < L; ++l) {
        const float* ln = lm->w.ln + (size_t)l * 6 * d;
        // --- self attention
        if (!gemms_only) {
            lm_ln_kernel<<<rows, 256, 0, s>>>(B.x, B.part, pending, part_stride, ln, ln + d, (__half*)B.h16, d);
            ACB_LAUNCH_CHECK(); ++nl;
        }
        {
            GemmParams p = base_gemm((const __half*)lm->w.w_qkv + (size_t)l * 3 * d * d, B.h16, 3 * d, d, rows, 1);
            p.q32 = B.q32; p.kc = (__half*)B.k_cache + l * kv_layer; p.vc = (__half*)B.v_cache + l * kv_layer;
            p.d = d; p.H = H; p.cache_len = c.max_seq; p.pos = B.pos;
            ACB_TRY(launch_gemm<EPI_QKV>(nt, p, 1, s)); ++nl;
        }
        if (!gemms_only) {
            AttnParams a{B.q32, 1, 0, (__half*)B.k_cache + l * kv_layer, (__half*)B.v_cache + l * kv_layer, (__half*)B.a16,
                         H, d, c.max_seq, B.pos, 0, scale};
            lm_attn_kernel<<<dim3(H, rows), 128, (size_t)c.max_seq * sizeof(float), s>>>(a);
            ACB_LAUNCH_CHECK(); ++nl;
        }
        {
            const int ns = pick_split(d, d, lm->sms, true);
            GemmParams p = base_gemm((const __half*)lm->w.w_o + (size_t)l * d * d, B.a16, d, d, rows, ns);
            p.out_f32 = B.part; p.ld_out = d; p.split_stride = part_stride;
            ACB_TRY(launch_gemm<EPI_PARTIAL>(nt, p, ns, s)); ++nl;
            pending = ns;
        }
        // --- cross attention
        if (lm->has_cross) {
            if (!gemms_only) {
                lm_ln_kernel<<<rows, 256, 0, s>>>(B.x, B.part, pending, part_stride, ln + 2 * d, ln + 3 * d, (__half*)B.h16, d);
                ACB_LAUNCH_CHECK(); ++nl;
            }
            const int nsq = pick_split(d, d, lm->sms, true);
            {
                GemmParams p = base_gemm((const __half*)lm->w.w_cq + (size_t)l * d * d, B.h16, d, d, rows, nsq);
                p.out_f32 = B.part; p.ld_out = d; p.split_stride = part_stride;
                ACB_TRY(launch_gemm<EPI_PARTIAL>(nt, p, nsq, s)); ++nl;
            }
            if (!gemms_only) {
                AttnParams a{B.part, nsq, part_stride, (__half*)B.ck_cache + l * ckv_layer,
                             (__half*)B.cv_cache + l * ckv_layer, (__half*)B.a16, H, d, c.max_text, B.pos, lm->text_len,
                             scale};
                lm_attn_kernel<<<dim3(H, rows), 128, (size_t)c.max_text * sizeof(float), s>>>(a);
                ACB_LAUNCH_CHECK(); ++nl;
            }
            {
                const int ns = pick_split(d, d, lm->sms, true);
                GemmParams p = base_gemm((const __half*)lm->w.w_co + (size_t)l * d * d, B.a16, d, d, rows, ns);
                p.out_f32 = B.part; p.ld_out = d; p.split_stride = part_stride;
                ACB_TRY(launch_gemm<EPI_PARTIAL>(nt, p, ns, s)); ++nl;
                pending = ns;
            }
        }
        // --- feed forward
        if (!gemms_only) {
            lm_ln_kernel<<<rows, 256, 0, s>>>(B.x, B.part, pending, part_stride, ln + 4 * d, ln + 5 * d, (__half*)B.h16, d);
            ACB_LAUNCH_CHECK(); ++nl;
        }
        {
            GemmParams p = base_gemm((const __half*)lm->w.w_ff1 + (size_t)l * ffn * d, B.h16, ffn, d, rows, 1);
            p.out_f16 = (__half*)B.f16; p.ld_out = ffn;
            ACB_TRY(launch_gemm<EPI_GELU>(nt, p, 1, s)); ++nl;
        }
        {
            const int ns = pick_split(d, ffn, lm->sms, true);
            GemmParams p = base_gemm((const __half*)lm->w.w_ff2 + (size_t)l * d * ffn, B.f16, d, ffn, rows, ns);
            p.out_f32 = B.part; p.ld_out = d; p.split_stride = part_stride;
            ACB_TRY(launch_gemm<EPI_PARTIAL>(nt, p, ns, s)); ++nl;
            pending = ns;
        }
    }
    if (!gemms_only) {
        lm_ln_kernel<<<rows, 256, 0, s>>>(B.x, B.part, pending, part_stride, lm->w.out_norm, lm->w.out_norm + d, (__half*)B.h16, d);
        ACB_LAUNCH_CHECK(); ++nl;
    }
    {
        const int N = c.n_q * c.card;
        GemmParams p = base_gemm(lm->w.heads, B.h16, N, d, rows, 1);
        p.out_f32 = B.logits; p.ld_out = N;
        ACB_TRY(launch_gemm<EPI_F32>(nt, p, 1, s)); ++nl;
    }
    if (!gemms_only) {
        int NP = 1;
        while (NP < c.card) NP <<= 1;
        SampleParams sp{B.logits, lm->samp.noise_from_buffer ? B.noise : nullptr, logits_out, B.seq, B.seq_mask, B.pos, c.max_seq, nullptr, lm->batch, rows,
                        c.n_q, c.card, NP, lm->samp.use_sampling, lm->samp.top_k, lm->samp.temp, lm->samp.top_p,
                        lm->samp.cfg_coef, lm->samp.seed, 0};
        size_t smem = ((size_t)c.card + 2 * (size_t)NP) * sizeof(float);
        lm_sample_kernel<<<dim3(c.n_q, lm->batch), 1024, smem, s>>>(sp);
        ACB_LAUNCH_CHECK(); ++nl;
    }
    if (!gemms_only) {
        lm_advance_kernel<<<1, 1, 0, s>>>(B.pos);
        ACB_LAUNCH_CHECK(); ++nl;
    }
    if (n_launch) *n_launch = nl;
    return ACB_OK;
}

extern "C" int acb_lm_create(const acb_lm_config* cfg, const acb_lm_weights* w, const acb_lm_buffers* buf, acb_lm_t** out) {
    ACB_REQUIRE(cfg && w && buf && out, "acb_lm_create: null argument");
    ACB_REQUIRE(cfg->dim % 64 == 0 && cfg->dim == cfg->num_heads * 64, "acb_lm_create: head_dim must be 64 (dim=%d heads=%d)",
                cfg->dim, cfg->num_heads);
    ACB_REQUIRE(cfg->dim <= 256 * LN_MAX_PER_THREAD, "acb_lm_create: dim %d too large", cfg->dim);
    ACB_REQUIRE(cfg->ffn_dim % 32 == 0 && cfg->card % 16 == 0 && cfg->n_q >= 1 && cfg->n_q <= 16, "acb_lm_create: bad ffn/card/n_q");
    ACB_REQUIRE(cfg->card <= 4096, "acb_lm_create: card %d > 4096 not built", cfg->card);
    ACB_REQUIRE(cfg->max_rows >= 1 && cfg->max_rows <= 64, "acb_lm_create: max_rows %d not in [1,64]", cfg->max_rows);
    ACB_REQUIRE(cfg->max_seq >= 2 && cfg->max_seq <= 12000, "acb_lm_create: max_seq %d out of range", cfg->max_seq);
    acb_lm* lm = new (std::nothrow) acb_lm();
    ACB_REQUIRE(lm, "acb_lm_create: out of host memory");
    lm->cfg = *cfg; lm->w = *w; lm->buf = *buf;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&lm->sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaStreamCreateWithFlags(&lm->capture_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete lm; acb_set_error("acb_lm_create: cudaStreamCreate: %s", cudaGetErrorString(e)); return ACB_ERR_CUDA; }
    // sampling kernel needs > 48 KB only for card > ~4000; attention scores for max_seq > 12288
    *out = lm;
    return ACB_OK;
}

static void drop_graph(acb_lm* lm) {
    if (lm->exec) { cudaGraphExecDestroy(lm->exec); lm->exec = nullptr; }
    if (lm->graph) { cudaGraphDestroy(lm->graph); lm->graph = nullptr; }
}

extern "C" int acb_lm_destroy(acb_lm_t* lm) {
    if (!lm) return ACB_OK;
    drop_graph(lm);
    if (lm->capture_stream) cudaStreamDestroy(lm->capture_stream);
    delete lm;
    return ACB_OK;
}

extern "C" int acb_lm_begin(acb_lm_t* lm, const float* cross, int batch, int rows, int text_len, int seq_len,
                            const acb_lm_sampling* sampling, void* stream) {
    ACB_REQUIRE(lm && sampling, "acb_lm_begin: null argument");
    const acb_lm_config& c = lm->cfg;
    ACB_REQUIRE(batch >= 1 && (rows == batch || rows == 2 * batch), "acb_lm_begin: rows must be batch or 2*batch");
    ACB_REQUIRE(rows <= c.max_rows, "acb_lm_begin: rows %d > max_rows %d", rows, c.max_rows);
    ACB_REQUIRE(seq_len >= 2 && seq_len <= c.max_seq, "acb_lm_begin: seq_len %d > max_seq %d", seq_len, c.max_seq);
    ACB_REQUIRE(!c.cross_attention || cross, "acb_lm_begin: the model has cross attention, a condition tensor is required"
                " (the reference asserts the same, transformer.py:553-556)");
    ACB_REQUIRE(!cross || (text_len >= 1 && text_len <= c.max_text), "acb_lm_begin: text_len %d out of range", text_len);
    cudaStream_t s = (cudaStream_t)stream;
    lm->batch = batch; lm->rows = rows; lm->rows_pad = 8 * nt_for_rows(rows); lm->text_len = text_len; lm->seq_len = seq_len;
    lm->samp = *sampling;
    lm->has_cross = c.cross_attention && cross;
    const int d = c.dim, H = c.num_heads;
    // zero the padded activation rows once; kernels only ever write rows < `rows`
    ACB_CHECK_CUDA(cudaMemsetAsync(lm->buf.h16, 0, (size_t)lm->rows_pad * d * sizeof(__half), s));
    ACB_CHECK_CUDA(cudaMemsetAsync(lm->buf.a16, 0, (size_t)lm->rows_pad * d * sizeof(__half), s));
    ACB_CHECK_CUDA(cudaMemsetAsync(lm->buf.f16, 0, (size_t)lm->rows_pad * c.ffn_dim * sizeof(__half), s));
    int hp[4] = {0, rows, batch, text_len};
    ACB_CHECK_CUDA(cudaMemcpyAsync(lm->buf.pos, hp, sizeof(hp), cudaMemcpyHostToDevice, s));
    if (lm->has_cross) {
        const size_t M = (size_t)rows * text_len, Mpad = (M + 63) / 64 * 64;
        lm_f32_to_f16_kernel<<<(unsigned)((Mpad * d + 255) / 256), 256, 0, s>>>(cross, (__half*)lm->buf.cross16, M * d, Mpad * d);
        ACB_LAUNCH_CHECK();
        const size_t ckv_layer = (size_t)c.max_rows * H * c.max_text * 64;
        for (int l = 0; l < c.num_layers; ++l)
            for (size_t r0 = 0; r0 < M; r0 += 64) {
                GemmParams p = base_gemm((const __half*)lm->w.w_ckv + (size_t)l * 2 * d * d,
                                         (const __half*)lm->buf.cross16 + r0 * d, 2 * d, d, (int)min((size_t)64, M - r0), 1);
                p.kc = (__half*)lm->buf.ck_cache + l * ckv_layer; p.vc = (__half*)lm->buf.cv_cache + l * ckv_layer;
                p.d = d; p.H = H; p.cache_len = c.max_text; p.text_len = text_len; p.row0 = (int)r0;
                ACB_TRY(launch_gemm<EPI_CROSSKV>(8, p, 1, s));
            }
    }
    // opt in to large dynamic shared memory where needed
    {
        int NP = 1;
        while (NP < c.card) NP <<= 1;
        size_t smem = ((size_t)c.card + 2 * (size_t)NP) * sizeof(float);
        if (smem > 48 * 1024)
            ACB_CHECK_CUDA(cudaFuncSetAttribute(lm_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        size_t asm_ = (size_t)max(c.max_seq, c.max_text) * sizeof(float);
        if (asm_ > 48 * 1024)
            ACB_CHECK_CUDA(cudaFuncSetAttribute(lm_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asm_));
    }
    // capture one decode step
    drop_graph(lm);
    ACB_CHECK_CUDA(cudaStreamBeginCapture(lm->capture_stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_step(lm, lm->capture_stream, nullptr, &lm->launches);
    cudaError_t e = cudaStreamEndCapture(lm->capture_stream, &lm->graph);
    if (rc != ACB_OK) { drop_graph(lm); return rc; }
    if (e != cudaSuccess) { acb_set_error("acb_lm_begin: graph capture failed: %s", cudaGetErrorString(e)); drop_graph(lm); return ACB_ERR_CUDA; }
    ACB_CHECK_CUDA(cudaGraphInstantiate(&lm->exec, lm->graph, 0));
    return ACB_OK;
}

extern "C" int acb_lm_steps(acb_lm_t* lm, int n_steps, void* stream) {
    ACB_REQUIRE(lm && lm->exec, "acb_lm_steps: call acb_lm_begin first");
    ACB_REQUIRE(n_steps >= 0, "acb_lm_steps: negative step count");
    for (int i = 0; i < n_steps; ++i) ACB_CHECK_CUDA(cudaGraphLaunch(lm->exec, (cudaStream_t)stream));
    return ACB_OK;
}

extern "C" int acb_lm_step_logits(acb_lm_t* lm, float* logits_out, void* stream) {
    ACB_REQUIRE(lm && lm->rows > 0, "acb_lm_step_logits: call acb_lm_begin first");
    return enqueue_step(lm, (cudaStream_t)stream, logits_out, nullptr);
}

extern "C" int acb_lm_debug_gemms(acb_lm_t* lm, void* stream, int* n_launches) {
    ACB_REQUIRE(lm && lm->rows > 0, "acb_lm_debug_gemms: call acb_lm_begin first");
    return enqueue_step(lm, (cudaStream_t)stream, nullptr, n_launches, true);
}

extern "C" int acb_lm_rows_pad(int rows) { return 8 * nt_for_rows(rows); }

extern "C" int acb_lm_launches_per_step(const acb_lm_t* lm) { return lm ? lm->launches : 0; }

extern "C" int acb_sample(const float* logits, const float* noise, int64_t* tokens, int batch, int rows, int n_q, int card,
                          const acb_lm_sampling* sampling, uint64_t step, void* stream) {
    ACB_REQUIRE(logits && tokens && sampling, "acb_sample: null argument");
    ACB_REQUIRE(batch >= 1 && (rows == batch || rows == 2 * batch) && n_q >= 1 && card >= 2 && card <= 4096, "acb_sample: bad shape");
    int NP = 1;
    while (NP < card) NP <<= 1;
    SampleParams sp{logits, sampling->noise_from_buffer ? noise : nullptr, nullptr, nullptr, nullptr, nullptr, 0, tokens, batch, rows, n_q, card, NP,
                    sampling->use_sampling, sampling->top_k, sampling->temp, sampling->top_p, sampling->cfg_coef,
                    sampling->seed, (uint32_t)step};
    size_t smem = ((size_t)card + 2 * (size_t)NP) * sizeof(float);
    if (smem > 48 * 1024)
        ACB_CHECK_CUDA(cudaFuncSetAttribute(lm_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lm_sample_kernel<<<dim3(n_q, batch), 1024, smem, (cudaStream_t)stream>>>(sp);
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}
