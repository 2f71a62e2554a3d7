// MusicGen LM decode step as ONE persistent kernel (sm_100a): every layer of LMModel.forward on one token per row
// (audiocraft/models/lm.py:221-268, modules/transformer.py:550-574, 693-713) runs inside a single cooperative launch of
// one CTA per SM; the CFG mix + sampler (lm.cu: lm_sample_kernel) follows as the second and last kernel of the step.
//
// Why one kernel.  A decode step is a strict chain of ~12 small dependent phases per layer.  As separate kernels each
// phase paid ~3.9 us (round 1: 532 launches, 2.07 ms per step at KV length 1) while its HBM traffic needs 0.1-0.3 us.
// Here the phases are separated by a grid barrier (csrc/gridbar.cuh, < 1 us) and, because weights are static, a
// dedicated producer warp streams them with TMA bulk copies through a deep shared-memory ring that runs AHEAD of the
// phase the compute warps are in: the HBM weight stream never stops at a phase boundary.
//
// Work decomposition.  Every GEMM y[rows][N] = act[rows][K] . W[N][K]^T is cut into items = (128-feature tile, K slice)
// so that ~all 148 SMs hold one item; the 128 x 64 fp16 weight tiles are pre-packed (acb_lm_pack_weight) in the
// canonical K-major UMMA layout, so a tile is ONE contiguous 16 KB bulk copy and is consumed straight from shared memory
// by tcgen05.mma (swap-AB: features on M = 128, rows on N = R in {16,32,48,64}; accumulator in TMEM).  No thread ever
// touches a weight fragment.  Items write fp32 partial sums; the consumer of a GEMM reduces them in a fixed order
// (bit-reproducible): attention sums the 4 QKV partials of its head, the residual phases sum the O / FFN2 partials
// into x and emit LayerNorm statistics per d/8 columns, so the next GEMM normalises its activations while it stages them.
//
// Phases per layer (grid barrier after each):
//   QKV gemm (LN1 on load) | self-attention (+KV append) | O gemm | residual+stats |
//   CQ gemm (LNc on load) | cross-attention | CO gemm | residual+stats | FF1 gemm (LN2 on load) | gelu+reduce |
//   FF2 gemm | residual+stats          then: heads gemm (out_norm on load) | logits reduce.
#include "lm_step.cuh"
#include "gridbar.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int NCW = 8;                  // compute warps
constexpr int CT = NCW * 32;            // compute threads
constexpr int BLOCK = CT + 32;          // + the TMA producer warp
constexpr int TILE_BYTES = 16384;       // 128 features x 64 K, fp16
constexpr int TILE_HALVES = 8192;
constexpr int MAX_STAGE = 14;
constexpr int SCRATCH_BYTES = 8192;     // mbarriers + small per-phase scratch behind the ring and the activation tile
constexpr int ATT_UNROLL = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// barrier among the compute warps only (the producer warp never joins)
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }

// UMMA shared-memory descriptor, canonical K-major layout without swizzle (cute/arch/mma_sm100_desc.hpp; the same
// encoding conv1d_t5_kernel in encodec.cu runs on hardware): element (row, 16-byte k-chunk c) lives at
//   c * LBO + (row / 8) * 128 + (row % 8) * 16      start [0,14) >> 4, LBO [16,30) >> 4, SBO [32,46) >> 4 = 128 B, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float half_round(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// sum over the CT compute threads; `red` holds NCW floats and is reusable after the NEXT cbar()
__device__ __forceinline__ float cw_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    cbar();
    float t = lane < NCW ? red[lane] : 0.f;
    return warp_sum(t);
}

struct Smem {
    unsigned char* ring; unsigned char* act;
    uint64_t* full; uint64_t* empty; uint64_t* accf; uint32_t* tslot;
    float* mean; float* rstd;       // [64] LayerNorm statistics of the rows for the GEMM being staged
    float* red;                     // [4][NCW]
    float* wm; float* wl; float* wacc;   // attention merge: [NCW], [NCW], [NCW][64]
    float* sq; float* sk; float* sv;     // [64] each: this step's q / k / v of the (row, head) task
    float* qs;                      // [NCW][64] cross-attention queries, one per warp
};

__device__ __forceinline__ Smem carve(unsigned char* sm, const StepParams& p) {
    Smem s;
    s.ring = sm;
    s.act = sm + (size_t)p.n_stage * TILE_BYTES;
    unsigned char* q = s.act + p.act_bytes;
    s.full = reinterpret_cast<uint64_t*>(q);
    s.empty = s.full + MAX_STAGE;
    s.accf = s.empty + MAX_STAGE;
    s.tslot = reinterpret_cast<uint32_t*>(s.accf + 2);
    float* f = reinterpret_cast<float*>(s.tslot + 4);
    s.mean = f; f += 64;
    s.rstd = f; f += 64;
    s.red = f; f += 4 * NCW;
    s.wm = f; f += NCW;
    s.wl = f; f += NCW;
    s.wacc = f; f += NCW * 64;
    s.sq = f; f += 64;
    s.sk = f; f += 64;
    s.sv = f; f += 64;
    s.qs = f; f += NCW * 64;
    return s;
}

// Which CTA runs item i of GEMM gi in layer `layer`: rotate so that the few SMs without an item differ per GEMM.
__device__ __forceinline__ int cta_rank(int cta, int gi, int layer, int n_cta) {
    return (cta + gi * 53 + layer * 17) % n_cta;
}

// ------------------------------------------------------------------------------------------------ producer warp
// Walks the step's GEMMs in execution order and keeps the ring full: stage s of use n is handed over on full[s] with
// parity (n & 1) and taken back on empty[s] (armed by the tcgen05.commit that follows the MMAs reading it).
__device__ void producer_loop(const StepParams& p, const Smem& s, int cta, int n_cta) {
    uint32_t it = 0;
    int n_gemm = 0;
    auto stream_gemm = [&](int gi, int layer) {
        if (n_gemm++ >= p.max_gemms) return;   // debug stop (ACB_LM_STEP_STOP): the compute warps leave before this GEMM
        const StepGemm& G = p.g[gi];
        const __half* base = G.wp + (size_t)layer * G.layer_stride;
        for (int item = cta_rank(cta, gi, layer, n_cta); item < G.n_items; item += n_cta) {
            const int nt = item / G.ksplit, ks = item - nt * G.ksplit;
            const __half* src = base + ((size_t)nt * G.nkb + (size_t)ks * G.kb_per) * TILE_HALVES;
            for (int kb = 0; kb < G.kb_per; ++kb, ++it) {
                const uint32_t st = it % (uint32_t)p.n_stage, par = (it / (uint32_t)p.n_stage) & 1u;
                mbar_wait(s.empty + st, par ^ 1u);
                mbar_expect_tx(s.full + st, TILE_BYTES);
                bulk_g2s(s.ring + (size_t)st * TILE_BYTES, src + (size_t)kb * TILE_HALVES, TILE_BYTES, s.full + st);
            }
        }
    };
    for (int l = 0; l < p.L; ++l) {
        stream_gemm(SG_QKV, l);
        stream_gemm(SG_O, l);
        if (p.has_cross) { stream_gemm(SG_CQ, l); stream_gemm(SG_CO, l); }
        stream_gemm(SG_FF1, l);
        stream_gemm(SG_FF2, l);
    }
    stream_gemm(SG_HEADS, 0);
}

// ------------------------------------------------------------------------------------------------ compute side
struct Cons {
    uint32_t it;          // ring position (meaningful in thread 0, kept in step by every compute thread)
    uint32_t acc_use[2];  // uses of each TMEM accumulator (parity of accf)
    uint32_t n_item;      // items this CTA has run (selects the accumulator)
    uint32_t tmem;
    unsigned nbar;
};

enum { ALOAD_LN = 0, ALOAD_F16 = 1 };

// LayerNorm statistics of every row from the per-chunk (mean, M2) records (Chan's merge, equal counts).
__device__ __forceinline__ void row_stats(const StepParams& p, const Smem& s) {
    const int tid = threadIdx.x;
    if (tid < p.rows) {
        float mc[ACB_STEP_STAT_CHUNKS], m2 = 0.f, mean = 0.f;
#pragma unroll
        for (int c = 0; c < ACB_STEP_STAT_CHUNKS; ++c) {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(p.stats + ((size_t)c * p.R + tid) * 2));
            mc[c] = v.x; m2 += v.y; mean += v.x;
        }
        mean *= 1.f / ACB_STEP_STAT_CHUNKS;
        float dev = 0.f;
#pragma unroll
        for (int c = 0; c < ACB_STEP_STAT_CHUNKS; ++c) { const float dlt = mc[c] - mean; dev = fmaf(dlt, dlt, dev); }
        m2 += dev * (float)(p.d / ACB_STEP_STAT_CHUNKS);
        s.mean[tid] = mean;
        s.rstd[tid] = 1.f / sqrtf(m2 / (float)p.d + 1e-5f);
    }
    cbar();
}

__device__ void gemm_phase(const StepParams& p, const Smem& s, Cons& c, int gi, int layer, int aload, const float* gamma,
                           const float* beta, const __half* src16, int ld16, int cta, int n_cta) {
    const StepGemm& G = p.g[gi];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = p.R, rows = p.rows;
    const int first = cta_rank(cta, gi, layer, n_cta);
    if (first >= G.n_items) return;
    if (aload == ALOAD_LN) row_stats(p, s);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(R >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int item = first; item < G.n_items; item += n_cta) {
        const int nt = item / G.ksplit, ks = item - nt * G.ksplit;
        const int k0 = ks * G.kb_per * 64, nch = G.kb_per * 8;
        // ---- stage the activations [rows][k0 .. k0 + 64*kb_per) as the B operand: chunk-major, 16 bytes per (chunk, row)
        for (int idx = tid; idx < rows * nch; idx += CT) {
            const int ch = idx / rows, r = idx - ch * rows;
            const int k = k0 + ch * 8;
            uint4 pk;
            if (aload == ALOAD_LN) {
                const float4 a = ldcg4(p.x + (size_t)r * p.d + k), b = ldcg4(p.x + (size_t)r * p.d + k + 4);
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + k)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + k + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + k)), b1 = __ldg(reinterpret_cast<const float4*>(beta + k + 4));
                const float mu = s.mean[r], rs = s.rstd[r];
                __half2 h0 = __floats2half2_rn((a.x - mu) * rs * g0.x + b0.x, (a.y - mu) * rs * g0.y + b0.y);
                __half2 h1 = __floats2half2_rn((a.z - mu) * rs * g0.z + b0.z, (a.w - mu) * rs * g0.w + b0.w);
                __half2 h2 = __floats2half2_rn((b.x - mu) * rs * g1.x + b1.x, (b.y - mu) * rs * g1.y + b1.y);
                __half2 h3 = __floats2half2_rn((b.z - mu) * rs * g1.z + b1.z, (b.w - mu) * rs * g1.w + b1.w);
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            } else {
                pk = __ldcg(reinterpret_cast<const uint4*>(src16 + (size_t)r * ld16 + k));
            }
            *reinterpret_cast<uint4*>(s.act + ((size_t)ch * R + r) * 16) = pk;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        cbar();
        const uint32_t a_sel = c.n_item & 1u;
        if (warp == 0) {
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = c.tmem + a_sel * (uint32_t)R;
                const uint32_t act_s = smem_u32(s.act), lbo_b = (uint32_t)R * 16u;
                uint32_t it = c.it;
                for (int kb = 0; kb < G.kb_per; ++kb, ++it) {
                    const uint32_t st = it % (uint32_t)p.n_stage, par = (it / (uint32_t)p.n_stage) & 1u;
                    mbar_wait(s.full + st, par);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_s = smem_u32(s.ring + (size_t)st * TILE_BYTES);
                    const uint32_t b_s = act_s + (uint32_t)kb * 8u * lbo_b;
#pragma unroll
                    for (int j = 0; j < 4; ++j)   // one instruction = 16 K elements = two 16-byte chunks of A and of B
                        umma_f16(acc, umma_desc(a_s + j * 4096u, 2048u), umma_desc(b_s + j * 2u * lbo_b, lbo_b), idesc,
                                 (kb | j) ? 1u : 0u);
                    umma_commit(s.empty + st);     // stage free again once these MMAs have read it
                }
                umma_commit(s.accf + a_sel);
            }
            __syncwarp();
        }
        c.it += (uint32_t)G.kb_per;
        mbar_wait(s.accf + a_sel, c.acc_use[a_sel] & 1u);
        ++c.acc_use[a_sel];
        ++c.n_item;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {   // epilogue: TMEM lane = feature, column = row.  warp w reads lanes [32 (w % 4), +32), columns of half (w / 4)
            const int q = warp & 3, hf = warp >> 2, cpw = R >> 1;
            float* out = p.part + ((size_t)ks * R) * G.N + (size_t)nt * 128 + q * 32 + lane;
            for (int c0 = hf * cpw; c0 < (hf + 1) * cpw; c0 += 8) {
                uint32_t v[8];
                const uint32_t taddr = c.tmem + a_sel * (uint32_t)R + (uint32_t)c0 + ((uint32_t)(q * 32) << 16);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c0 + j < rows) out[(size_t)(c0 + j) * G.N] = __uint_as_float(v[j]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cbar();   // activation tile and accumulator reusable
    }
}

// x[r][cols of chunk c] (+)= sum of split-K partials, and the chunk's LayerNorm record (mean, centred sum of squares).
// EMBED: x = sum_k emb_k[token] + sinusoidal position (lm.py:244, transformer.py:70-89, 701-705) instead.
template <bool EMBED>
__device__ void residual_phase(const StepParams& p, const Smem& s, int ksplit, int pos, int cta, int n_cta) {
    const int tid = threadIdx.x, d = p.d, len = d / ACB_STEP_STAT_CHUNKS, tasks = p.rows * ACB_STEP_STAT_CHUNKS;
    for (int t = cta; t < tasks; t += n_cta) {
        const int c = t / p.rows, r = t - c * p.rows;
        const bool live = tid < len;
        const int col = c * len + tid;
        float v = 0.f;
        if (live) {
            if (EMBED) {
                const int b = r % p.batch, half_d = d >> 1;
                for (int k = 0; k < p.n_q; ++k) {
                    long long tk = p.seq[((size_t)b * p.n_q + k) * p.max_seq + pos];
                    const int tok = (int)(tk < 0 ? p.card : (tk > p.card ? p.card : tk));
                    v += __half2float(p.emb[((size_t)k * (p.card + 1) + tok) * d + col]);
                }
                if (p.sin_pos) {
                    const int j = col < half_d ? col : col - half_d;
                    const float phase = (float)pos / p.inv_freq[j];
                    v += p.pos_scale * (col < half_d ? cosf(phase) : sinf(phase));
                }
            } else {
                v = __ldcg(p.x + (size_t)r * d + col);
                for (int ks = 0; ks < ksplit; ++ks) v += __ldcg(p.part + ((size_t)ks * p.R + r) * d + col);   // fixed order
            }
            p.x[(size_t)r * d + col] = v;
        }
        const float mean = cw_sum(live ? v : 0.f, s.red) / (float)len;
        const float dv = live ? v - mean : 0.f;
        const float m2 = cw_sum(dv * dv, s.red + NCW);
        if (tid == 0) *reinterpret_cast<float2*>(p.stats + ((size_t)c * p.R + r) * 2) = make_float2(mean, m2);
        cbar();
    }
}

// h16[r][n] = gelu(fp16(sum of the FF1 partials))   (linear1 output is fp16 under autocast, then F.gelu: transformer.py:569)
__device__ void gelu_phase(const StepParams& p, int ksplit, int cta, int n_cta) {
    const int n4 = p.ffn >> 2, total = p.rows * n4;
    for (int g = cta * CT + threadIdx.x; g < total; g += n_cta * CT) {
        const int r = g / n4, c4 = g - r * n4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ks = 0; ks < ksplit; ++ks) {
            const float4 t = ldcg4(p.part + ((size_t)ks * p.R + r) * p.ffn + c4 * 4);
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        __half2 lo = __floats2half2_rn(gelu_erf(half_round(a.x)), gelu_erf(half_round(a.y)));
        __half2 hi = __floats2half2_rn(gelu_erf(half_round(a.z)), gelu_erf(half_round(a.w)));
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p.h16 + (size_t)r * p.ffn + c4 * 4) = pk;
    }
}

__device__ void logits_phase(const StepParams& p, int ksplit, int cta, int n_cta) {
    const int N = p.n_q * p.card, n4 = N >> 2, total = p.rows * n4;
    for (int g = cta * CT + threadIdx.x; g < total; g += n_cta * CT) {
        const int r = g / n4, c4 = g - r * n4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ks = 0; ks < ksplit; ++ks) {
            const float4 t = ldcg4(p.part + ((size_t)ks * p.R + r) * N + c4 * 4);
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        *reinterpret_cast<float4*>(p.logits + (size_t)r * N + c4 * 4) = a;
    }
}

struct OnlineSM { float m, l, acc[8]; };
__device__ __forceinline__ void osm_merge(OnlineSM& a, float m2, float l2, const float (&acc2)[8]) {
    const float mn = fmaxf(a.m, m2);
    const float ca = a.m == -INFINITY ? 0.f : __expf(a.m - mn), cb = m2 == -INFINITY ? 0.f : __expf(m2 - mn);
    a.l = a.l * ca + l2 * cb;
#pragma unroll
    for (int e = 0; e < 8; ++e) a.acc[e] = a.acc[e] * ca + acc2[e] * cb;
    a.m = mn;
}
__device__ __forceinline__ void osm_step(OnlineSM& st, float sc, const float (&v)[8]) {
    const float mn = fmaxf(st.m, sc);
    const float corr = __expf(st.m - mn);   // exp(-inf) = 0 on the first position
    const float pw = __expf(sc - mn);
    st.l = st.l * corr + pw;
#pragma unroll
    for (int e = 0; e < 8; ++e) st.acc[e] = fmaf(pw, v[e], st.acc[e] * corr);
    st.m = mn;
}

// Self-attention of the step's single query per (row, head): sums the QKV partials of the head, appends k / v to the
// cache (fp16, like the reference's cached fp16 keys), then one online-softmax pass over the cache
// (StreamingMultiheadAttention.forward, transformer.py:315-451 with the `b h t d` cache of :266-298).
__device__ void self_attn_phase(const StepParams& p, const Smem& s, int layer, int pos, int cta, int n_cta) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sl = lane & 7, pg = lane >> 3;
    const int d = p.d, H = p.H, R = p.R, N3 = 3 * d, ksplit = p.g[SG_QKV].ksplit;
    const size_t kv_layer = (size_t)p.max_rows * H * p.max_seq * 64;
    __half* kc = p.kc + (size_t)layer * kv_layer;
    __half* vc = p.vc + (size_t)layer * kv_layer;
    const int tasks = p.rows * H;
    for (int t = (cta + layer * 29) % n_cta; t < tasks; t += n_cta) {
        const int row = t / H, h = t - row * H;
        const size_t base = ((size_t)row * H + h) * p.max_seq * 64;
        if (tid < 192) {
            const int which = tid >> 6, dd = tid & 63;
            const float* src = p.part + (size_t)row * N3 + which * d + h * 64 + dd;
            float v = 0.f;
            for (int ks = 0; ks < ksplit; ++ks) v += __ldcg(src + (size_t)ks * R * N3);   // fixed order
            if (p.rope && which < 2) {
                // RotaryEmbedding.rotate_qk (modules/rope.py:84-125) on the fp16 q / k of this position: the head dim is 32
                // complex pairs (2i, 2i+1), rotated by pos * max_period^(-2i/64) in fp32, mixed with `scale`, cast back.
                // (warps 0-3 hold q and k entirely, so the pair exchange is warp-uniform)
                const float vh = half_round(v), other = __shfl_xor_sync(0xffffffffu, vh, 1);
                const bool even = (dd & 1) == 0;
                const float re = even ? vh : other, im = even ? other : vh;
                const float ang = (float)pos * p.rope_freq[dd >> 1];
                float sn, cs;
                sincosf(ang, &sn, &cs);
                const float rr = cs * p.pos_scale + (1.f - p.pos_scale), ri = sn * p.pos_scale;
                v = even ? re * rr - im * ri : re * ri + im * rr;
            }
            if (which == 0) {
                s.sq[dd] = half_round(v) * p.attn_scale;
            } else {
                const __half hv = __float2half_rn(v);
                (which == 1 ? kc : vc)[base + (size_t)pos * 64 + dd] = hv;
                (which == 1 ? s.sk : s.sv)[dd] = __half2float(hv);
            }
        }
        cbar();
        float q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) q[e] = s.sq[sl * 8 + e];
        OnlineSM st;
        st.m = -INFINITY; st.l = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) st.acc[e] = 0.f;
        const __half* kb = kc + base + sl * 8;
        const __half* vb = vc + base + sl * 8;
        // cached positions [0, pos): a warp instruction reads 4 consecutive positions (512 contiguous bytes)
        for (int pb = warp * 4; pb < pos; pb += NCW * 4 * ATT_UNROLL) {
            uint4 kv[ATT_UNROLL], vv[ATT_UNROLL];
#pragma unroll
            for (int u = 0; u < ATT_UNROLL; ++u) {
                const int pp = pb + u * NCW * 4 + pg;
                if (pp < pos) {
                    kv[u] = ld_stream_u4(kb + (size_t)pp * 64);
                    vv[u] = ld_stream_u4(vb + (size_t)pp * 64);
                } else {
                    kv[u] = vv[u] = make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < ATT_UNROLL; ++u) {
                const int pp = pb + u * NCW * 4 + pg;
                const __half2* k2 = reinterpret_cast<const __half2*>(&kv[u]);
                float sc = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(k2[e]);
                    sc = fmaf(q[2 * e], f.x, sc);
                    sc = fmaf(q[2 * e + 1], f.y, sc);
                }
                sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                sc += __shfl_xor_sync(0xffffffffu, sc, 2);
                sc += __shfl_xor_sync(0xffffffffu, sc, 4);
                if (pp < pos) {
                    const __half2* v2 = reinterpret_cast<const __half2*>(&vv[u]);
                    float vf[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(v2[e]); vf[2 * e] = f.x; vf[2 * e + 1] = f.y; }
                    osm_step(st, sc, vf);
                }
            }
        }
        if (warp == 0) {   // the position appended by this step, from shared memory
            float sc = 0.f, vf[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { sc = fmaf(q[e], s.sk[sl * 8 + e], sc); vf[e] = s.sv[sl * 8 + e]; }
            sc += __shfl_xor_sync(0xffffffffu, sc, 1);
            sc += __shfl_xor_sync(0xffffffffu, sc, 2);
            sc += __shfl_xor_sync(0xffffffffu, sc, 4);
            if (pg == 0) osm_step(st, sc, vf);
        }
        // merge the 4 position groups of the warp, then the warps
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, st.m, o), l2 = __shfl_xor_sync(0xffffffffu, st.l, o);
            float a2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a2[e] = __shfl_xor_sync(0xffffffffu, st.acc[e], o);
            osm_merge(st, m2, l2, a2);
        }
        if (pg == 0) {
            if (sl == 0) { s.wm[warp] = st.m; s.wl[warp] = st.l; }
#pragma unroll
            for (int e = 0; e < 8; ++e) s.wacc[warp * 64 + sl * 8 + e] = st.acc[e];
        }
        cbar();
        if (tid < 64) {
            float mx = s.wm[0];
#pragma unroll
            for (int w = 1; w < NCW; ++w) mx = fmaxf(mx, s.wm[w]);
            float l = 0.f, o = 0.f;
#pragma unroll
            for (int w = 0; w < NCW; ++w) {
                const float cw = s.wm[w] == -INFINITY ? 0.f : __expf(s.wm[w] - mx);
                l = fmaf(s.wl[w], cw, l);
                o = fmaf(s.wacc[w * 64 + tid], cw, o);
            }
            p.a16[(size_t)row * d + h * 64 + tid] = __float2half_rn(o / l);
        }
        cbar();
    }
}

// Cross-attention over the cached text keys / values (computed once per generate): one warp per (row, head); the padded
// / null text positions are zero keys that still take part in the softmax (conditioners.py:1731-1746, transformer.py:343).
__device__ void cross_attn_phase(const StepParams& p, const Smem& s, int layer, int cta, int n_cta) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = p.d, H = p.H, R = p.R, ksplit = p.g[SG_CQ].ksplit, n = p.text_len;
    const size_t ckv_layer = (size_t)p.max_rows * H * p.max_text * 64;
    const __half* kc = p.ckc + (size_t)layer * ckv_layer;
    const __half* vc = p.cvc + (size_t)layer * ckv_layer;
    float* qs = s.qs + warp * 64;
    const int tasks = p.rows * H;
    for (int t = (cta + layer * 31) % n_cta + n_cta * warp; t < tasks; t += n_cta * NCW) {
        const int row = t / H, h = t - row * H;
        {
            const float* qp = p.part + (size_t)row * d + h * 64 + lane * 2;
            float a0 = 0.f, a1 = 0.f;
            for (int ks = 0; ks < ksplit; ++ks) {   // fixed order
                const float2 v = __ldcg(reinterpret_cast<const float2*>(qp + (size_t)ks * R * d));
                a0 += v.x; a1 += v.y;
            }
            qs[lane * 2] = half_round(a0) * p.attn_scale;
            qs[lane * 2 + 1] = half_round(a1) * p.attn_scale;
        }
        __syncwarp();
        const size_t base = ((size_t)row * H + h) * p.max_text * 64;
        float mx = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
        for (int t0 = 0; t0 < n; t0 += 32) {
            const int tt = t0 + lane;
            float sc = -INFINITY;
            if (tt < n) {
                const uint4* kr = reinterpret_cast<const uint4*>(kc + base + (size_t)tt * 64);
                sc = 0.f;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const uint4 kk = kr[c8];
                    const __half2* k2 = reinterpret_cast<const __half2*>(&kk);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(k2[e]);
                        sc = fmaf(qs[c8 * 8 + 2 * e], f.x, sc);
                        sc = fmaf(qs[c8 * 8 + 2 * e + 1], f.y, sc);
                    }
                }
            }
            const float cm = fmaxf(mx, warp_max(sc));
            const float corr = mx == -INFINITY ? 0.f : __expf(mx - cm);
            const float pw = tt < n ? __expf(sc - cm) : 0.f;
            l = l * corr + warp_sum(pw);
            o0 *= corr; o1 *= corr;
            const int cnt = min(32, n - t0);
            for (int j = 0; j < cnt; ++j) {
                const float wj = __shfl_sync(0xffffffffu, pw, j);
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(vc + base + (size_t)(t0 + j) * 64 + lane * 2));
                o0 = fmaf(wj, f.x, o0);
                o1 = fmaf(wj, f.y, o1);
            }
            mx = cm;
        }
        *reinterpret_cast<__half2*>(p.a16 + (size_t)row * d + h * 64 + lane * 2) = __floats2half2_rn(o0 / l, o1 / l);
        __syncwarp();
    }
}

__device__ __forceinline__ void grid_sync(const StepParams& p, Cons& c, int cta, int n_cta) {
    cbar();
    ++c.nbar;
    if (threadIdx.x == 0) {
        gridbar_arrive(p.bar);
        gridbar_wait(p.bar, c.nbar * (unsigned)n_cta);
        if (p.trace && cta == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            p.trace[c.nbar] = now;
        }
    }
    cbar();
}

// grid barrier at the end of a phase; true = the debug stop point (p.stop_after barriers) has been reached
__device__ __forceinline__ bool phase_end(const StepParams& p, Cons& c, int cta, int n_cta) {
    grid_sync(p, c, cta, n_cta);
    return (int)c.nbar >= p.stop_after;
}

__device__ void consumer_body(const StepParams& p, const Smem& s, int cta, int n_cta) {
    Cons c;
    c.it = 0; c.acc_use[0] = c.acc_use[1] = 0; c.n_item = 0; c.tmem = *s.tslot; c.nbar = 0;
    const int pos = p.pos[0];
    const int d = p.d;
#define PHASE_END() do { if (phase_end(p, c, cta, n_cta)) return; } while (0)
    residual_phase<true>(p, s, 0, pos, cta, n_cta);
    PHASE_END();
    for (int l = 0; l < p.L; ++l) {
        const float* ln = p.ln + (size_t)l * 6 * d;
        gemm_phase(p, s, c, SG_QKV, l, ALOAD_LN, ln, ln + d, nullptr, 0, cta, n_cta);
        PHASE_END();
        self_attn_phase(p, s, l, pos, cta, n_cta);
        PHASE_END();
        gemm_phase(p, s, c, SG_O, l, ALOAD_F16, nullptr, nullptr, p.a16, d, cta, n_cta);
        PHASE_END();
        residual_phase<false>(p, s, p.g[SG_O].ksplit, pos, cta, n_cta);
        PHASE_END();
        if (p.has_cross) {
            gemm_phase(p, s, c, SG_CQ, l, ALOAD_LN, ln + 2 * d, ln + 3 * d, nullptr, 0, cta, n_cta);
            PHASE_END();
            cross_attn_phase(p, s, l, cta, n_cta);
            PHASE_END();
            gemm_phase(p, s, c, SG_CO, l, ALOAD_F16, nullptr, nullptr, p.a16, d, cta, n_cta);
            PHASE_END();
            residual_phase<false>(p, s, p.g[SG_CO].ksplit, pos, cta, n_cta);
            PHASE_END();
        }
        gemm_phase(p, s, c, SG_FF1, l, ALOAD_LN, ln + 4 * d, ln + 5 * d, nullptr, 0, cta, n_cta);
        PHASE_END();
        gelu_phase(p, p.g[SG_FF1].ksplit, cta, n_cta);
        PHASE_END();
        gemm_phase(p, s, c, SG_FF2, l, ALOAD_F16, nullptr, nullptr, p.h16, p.ffn, cta, n_cta);
        PHASE_END();
        residual_phase<false>(p, s, p.g[SG_FF2].ksplit, pos, cta, n_cta);
        PHASE_END();
    }
    gemm_phase(p, s, c, SG_HEADS, 0, ALOAD_LN, p.out_norm, p.out_norm + d, nullptr, 0, cta, n_cta);
    PHASE_END();
#undef PHASE_END
    logits_phase(p, p.g[SG_HEADS].ksplit, cta, n_cta);
}

__global__ void __launch_bounds__(BLOCK, 1) lm_step_kernel(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(1024) unsigned char step_sm[];
    const Smem s = carve(step_sm, p);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, n_cta = gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < p.n_stage; ++i) { mbar_init(s.full + i, 1); mbar_init(s.empty + i, 1); }
        mbar_init(s.accf, 1); mbar_init(s.accf + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (p.trace && cta == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            p.trace[0] = now;
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s.tslot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < p.act_bytes / 16; i += BLOCK) reinterpret_cast<uint4*>(s.act)[i] = make_uint4(0, 0, 0, 0);   // padded rows stay 0
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == NCW) {
        if (lane == 0) producer_loop(p, s, cta, n_cta);
    } else {
        consumer_body(p, s, cta, n_cta);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*s.tslot), "r"((uint32_t)p.tmem_cols) : "memory");
}

// [N][K] row-major fp16 -> tiles of 128 features x 64 K in the canonical K-major UMMA layout:
//   tile (nt, kb) at ((nt * nkb + kb) * 8192) halves; inside: [k-chunk c (8)][feature f (128)][8 halves]
__global__ void lm_pack_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int N, int K) {
    const int nkb = K >> 6;
    const size_t total = (size_t)N * K / 8;   // 16-byte chunks
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(i & 127), c = (int)((i >> 7) & 7);
        const size_t tile = i >> 10;
        const int kb = (int)(tile % nkb), nt = (int)(tile / nkb);
        const uint4 v = *reinterpret_cast<const uint4*>(src + ((size_t)nt * 128 + f) * K + (size_t)kb * 64 + c * 8);
        reinterpret_cast<uint4*>(dst)[i] = v;
    }
}

}  // namespace

extern "C" int acb_lm_pack_weight(const void* w, void* wp, int n, int k, void* stream) {
    ACB_REQUIRE(w && wp && n > 0 && k > 0 && n % 128 == 0 && k % 64 == 0, "acb_lm_pack_weight: N %% 128 and K %% 64 must be 0 (N=%d K=%d)", n, k);
    const size_t total = (size_t)n * k / 8;
    const int blocks = (int)((total + 255) / 256 < 65535 * 16 ? (total + 255) / 256 : 65535 * 16);
    lm_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)w, (__half*)wp, n, k);
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}

// K-slices per GEMM: minimise (k-blocks the busiest CTA multiplies) + a charge per split (partial traffic, reduction).
static void plan_gemm(StepGemm& G, const void* wp, size_t layer_stride, int N, int K, int sms, int R) {
    G.wp = (const __half*)wp; G.layer_stride = layer_stride; G.N = N; G.K = K;
    G.n_tiles = N / 128; G.nkb = K / 64;
    float best = 1e30f;
    int best_ks = 1;
    for (int ks = 1; ks <= ACB_STEP_MAX_SPLIT && ks <= G.nkb; ++ks) {
        if (G.nkb % ks) continue;
        const int per = G.nkb / ks;
        if ((size_t)per * 64 * R * 2 > 96 * 1024) continue;   // activation tile must fit beside the ring
        const int items = G.n_tiles * ks;
        const float cost = (float)((items + sms - 1) / sms) * per + 0.35f * ks;
        if (cost < best) { best = cost; best_ks = ks; }
    }
    G.ksplit = best_ks; G.kb_per = G.nkb / best_ks; G.n_items = G.n_tiles * best_ks;
}

int lm_step_prepare(const acb_lm_config& c, const acb_lm_weights& w, const acb_lm_buffers& b, int rows, int batch, int text_len,
                    bool has_cross, int sms, StepLaunch* out) {
    const int d = c.dim, ffn = c.ffn_dim, NH = c.n_q * c.card;
    ACB_REQUIRE(w.wp_qkv && w.wp_o && w.wp_ff1 && w.wp_ff2 && w.wp_heads && (!c.cross_attention || (w.wp_cq && w.wp_co)),
                "fused step: packed weights missing");
    ACB_REQUIRE(b.stats && b.part && b.bar, "fused step: stats / part / bar buffers missing");
    ACB_REQUIRE(d % 128 == 0 && ffn % 128 == 0 && NH % 128 == 0, "fused step: dim, ffn and n_q*card must be multiples of 128");
    ACB_REQUIRE(d / ACB_STEP_STAT_CHUNKS <= CT && (d / ACB_STEP_STAT_CHUNKS) >= 1 && d % (8 * ACB_STEP_STAT_CHUNKS) == 0, "fused step: dim %d not supported", d);
    ACB_REQUIRE(rows >= 1 && rows <= 64, "fused step: rows %d not in [1,64]", rows);
    StepLaunch L{};
    StepParams& p = L.p;
    p.d = d; p.H = c.num_heads; p.L = c.num_layers; p.ffn = ffn; p.n_q = c.n_q; p.card = c.card;
    p.rows = rows; p.R = (rows + 15) / 16 * 16; p.batch = batch; p.has_cross = has_cross ? 1 : 0; p.text_len = text_len;
    p.max_seq = c.max_seq; p.max_text = c.max_text; p.max_rows = c.max_rows;
    p.pos_scale = c.pos_scale; p.attn_scale = 1.0f / sqrtf(64.f);
    p.emb = (const __half*)w.emb; p.inv_freq = w.inv_freq; p.ln = w.ln; p.out_norm = w.out_norm;
    plan_gemm(p.g[SG_QKV], w.wp_qkv, (size_t)3 * d * d, 3 * d, d, sms, p.R);
    plan_gemm(p.g[SG_O], w.wp_o, (size_t)d * d, d, d, sms, p.R);
    plan_gemm(p.g[SG_CQ], w.wp_cq, (size_t)d * d, d, d, sms, p.R);
    plan_gemm(p.g[SG_CO], w.wp_co, (size_t)d * d, d, d, sms, p.R);
    plan_gemm(p.g[SG_FF1], w.wp_ff1, (size_t)ffn * d, ffn, d, sms, p.R);
    plan_gemm(p.g[SG_FF2], w.wp_ff2, (size_t)d * ffn, d, ffn, sms, p.R);
    plan_gemm(p.g[SG_HEADS], w.wp_heads, 0, NH, d, sms, p.R);
    int kb_max = 1;
    for (int i = 0; i < ACB_STEP_GEMMS; ++i) kb_max = kb_max > p.g[i].kb_per ? kb_max : p.g[i].kb_per;
    p.act_bytes = kb_max * 8 * p.R * 16;
    int dev = 0, max_smem = 0;
    ACB_CHECK_CUDA(cudaGetDevice(&dev));
    ACB_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int ns = (max_smem - p.act_bytes - SCRATCH_BYTES) / TILE_BYTES;
    if (ns > MAX_STAGE) ns = MAX_STAGE;
    const char* e = getenv("ACB_LM_STAGES");
    if (e && atoi(e) >= 2 && atoi(e) < ns) ns = atoi(e);
    ACB_REQUIRE(ns >= 2, "fused step: not enough shared memory for the weight ring (%d B available)", max_smem);
    p.n_stage = ns;
    int cols = 32;
    while (cols < 2 * p.R) cols <<= 1;
    p.tmem_cols = cols;
    p.x = b.x; p.part = b.part; p.stats = b.stats; p.a16 = (__half*)b.a16; p.h16 = (__half*)b.f16; p.logits = b.logits;
    p.kc = (__half*)b.k_cache; p.vc = (__half*)b.v_cache; p.ckc = (const __half*)b.ck_cache; p.cvc = (const __half*)b.cv_cache;
    p.seq = b.seq; p.pos = b.pos; p.bar = (unsigned*)b.bar; p.trace = nullptr;
    p.sin_pos = c.positional_embedding != 1; p.rope = c.positional_embedding >= 1; p.rope_freq = w.rope_freq;
    ACB_REQUIRE(!p.rope || w.rope_freq, "fused step: rope_freq table missing");
    p.stop_after = 1 << 30; p.max_gemms = 1 << 30;
    if (const char* es = getenv("ACB_LM_STEP_STOP")) {   // bring-up aid: leave after this many grid barriers
        const int stop = atoi(es);
        if (stop >= 1) {
            // phase kinds in execution order: embed, per layer [G A G R (G A G R) G E G R], heads G; count the GEMM
            // phases among the first `stop` phases
            int n = 1, gem = 0;   // phase 0 = embed
            const int per = has_cross ? 12 : 8;
            static const int is_gemm_c[12] = {1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0};
            static const int is_gemm_n[8] = {1, 0, 1, 0, 1, 0, 1, 0};
            for (int l = 0; l < c.num_layers && n < stop; ++l)
                for (int k = 0; k < per && n < stop; ++k, ++n) gem += has_cross ? is_gemm_c[k] : is_gemm_n[k];
            if (n < stop) { ++gem; ++n; }   // heads
            p.stop_after = stop; p.max_gemms = gem;
        }
    }
    L.grid = sms; L.block = BLOCK;
    L.smem = (size_t)ns * TILE_BYTES + p.act_bytes + SCRATCH_BYTES;
    L.n_phases = 1 + c.num_layers * (has_cross ? 12 : 8) + 2;
    L.cooperative = true;
    ACB_CHECK_CUDA(cudaFuncSetAttribute(lm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    int per_sm = 0;
    ACB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lm_step_kernel, BLOCK, L.smem));
    ACB_REQUIRE(per_sm >= 1, "fused step: the step kernel does not fit an SM (%zu B shared memory)", L.smem);
    *out = L;
    return ACB_OK;
}

int lm_step_launch(const StepLaunch& L, cudaStream_t s) {
    ACB_CHECK_CUDA(cudaMemsetAsync(L.p.bar, 0, 128, s));   // grid-barrier counter (a memset node inside the step graph)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(L.block); cfg.dynamicSmemBytes = L.smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = L.cooperative ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ACB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lm_step_kernel, L.p));
    return ACB_OK;
}
