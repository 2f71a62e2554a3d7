// EnCodec hot path for B200 (sm_100a): SEANet convolutions, LSTM recurrence, residual VQ.
//
// fp32 CUDA-core arithmetic on purpose: RVQ code indices must match the fp32 reference bit for bit on the
// same latent, and a TF32/BF16 tensor-core conv would move the latents by ~1e-3 and flip near-tie codes
// (SURVEY.md section 7 "hard parts").  What is B200-specific here is the data movement: padding, ELU,
// bias, residual add and the transposed-conv trim are folded into the conv kernels (one read + one write
// of every activation per layer), weight-norm is folded once at load, the [frames x bins] VQ distance
// matrix never leaves the SM, and the LSTM keeps W_hh resident in the 227 KB shared memory of 128 SMs.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// weight-norm fold: w[g][:] = g[g] * v[g][:] / ||v[g]||   (audiocraft/modules/conv.py:21-30)
// ------------------------------------------------------------------------------------------------
__global__ void weight_norm_fold_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                        float* __restrict__ w, int groups, int inner) {
    int grp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (grp >= groups) return;
    const float* vr = v + (size_t)grp * inner;
    float ss = 0.f;
    for (int i = lane; i < inner; i += 32) ss = fmaf(vr[i], vr[i], ss);
    ss = warp_sum(ss);
    float scale = g[grp] / sqrtf(ss);
    for (int i = lane; i < inner; i += 32) w[(size_t)grp * inner + i] = vr[i] * scale;
}

extern "C" int acb_weight_norm_fold(const float* v, const float* g, float* w, int groups, int inner, void* stream) {
    ACB_REQUIRE(v && g && w && groups > 0 && inner > 0, "acb_weight_norm_fold: bad arguments");
    int wpb = 4;
    weight_norm_fold_kernel<<<acb_ceil_div(groups, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(v, g, w, groups, inner);
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}

// ------------------------------------------------------------------------------------------------
// conv1d: implicit GEMM on fp32 FMA.  CTA tile = (8*CPT output channels) x (32*TPT output steps) of one
// batch item; warp = channel group (weights are warp-broadcast from smem), lane = time (x reads and y
// writes are unit-stride).  The input slab is staged once per input-channel chunk with reflect / zero
// padding and ELU applied on the way in, de-interleaved by stride phase so strided convs read smem
// conflict-free.
// ------------------------------------------------------------------------------------------------
struct ConvParams {
    const float* x; const float* w; const float* bias; const float* res; float* y;
    int c_in, c_out, t_in, t_virt, t_out, K, stride, dil, pad_left, reflect, elu, ci_chunk, span, PL;
};

__device__ __forceinline__ float conv_fetch(const float* __restrict__ xr, int g, int t_in, int t_virt, int reflect,
                                            int elu) {
    if (reflect) {
        if (g < 0) g = -g;
        if (g >= t_virt) g = 2 * (t_virt - 1) - g;
    }
    float v = (g >= 0 && g < t_in) ? xr[g] : 0.f;
    return elu ? acb_elu(v) : v;
}

template <int CPT, int TPT>
__global__ void __launch_bounds__(256) conv1d_kernel(ConvParams p) {
    constexpr int BM = 8 * CPT, BN = 32 * TPT;
    extern __shared__ float smem[];
    const int XS = p.stride * p.PL;               // floats per staged input channel
    float* ws = smem;                             // [ci_chunk*K][BM]
    float* xs = ws + p.ci_chunk * p.K * BM;       // [ci_chunk][XS]
    int* koff = (int*)(xs + p.ci_chunk * XS);     // [K] tap offset inside a staged channel

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.x * BN, co0 = blockIdx.y * BM, b = blockIdx.z;
    const float* xb = p.x + (size_t)b * p.c_in * p.t_in;

    if (tid < p.K) {
        int kd = tid * p.dil;
        koff[tid] = (kd % p.stride) * p.PL + kd / p.stride;
    }

    float acc[CPT][TPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int j = 0; j < TPT; ++j) acc[i][j] = 0.f;

    const int g0 = t0 * p.stride - p.pad_left;
    for (int ci0 = 0; ci0 < p.c_in; ci0 += p.ci_chunk) {
        const int nci = min(p.ci_chunk, p.c_in - ci0);
        __syncthreads();  // previous chunk fully consumed (also publishes koff on the first pass)
        for (int idx = tid; idx < nci * p.K * BM; idx += 256) {
            int r = idx / BM, c = idx - r * BM;
            ws[idx] = (co0 + c < p.c_out) ? p.w[((size_t)ci0 * p.K + r) * p.c_out + co0 + c] : 0.f;
        }
        for (int idx = tid; idx < nci * p.span; idx += 256) {
            int cl = idx / p.span, j = idx - cl * p.span;
            float v = conv_fetch(xb + (size_t)(ci0 + cl) * p.t_in, g0 + j, p.t_in, p.t_virt, p.reflect, p.elu);
            xs[cl * XS + (j % p.stride) * p.PL + j / p.stride] = v;
        }
        __syncthreads();
        for (int cl = 0; cl < nci; ++cl) {
            const float* xc = xs + cl * XS + lane;
            const float* wc = ws + cl * p.K * BM + warp * CPT;
            for (int k = 0; k < p.K; ++k) {
                float wv[CPT];
#pragma unroll
                for (int i = 0; i < CPT; ++i) wv[i] = wc[k * BM + i];
                const float* xk = xc + koff[k];
#pragma unroll
                for (int j = 0; j < TPT; ++j) {
                    float xv = xk[32 * j];
#pragma unroll
                    for (int i = 0; i < CPT; ++i) acc[i][j] = fmaf(wv[i], xv, acc[i][j]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        int co = co0 + warp * CPT + i;
        if (co >= p.c_out) continue;
        float bv = p.bias ? p.bias[co] : 0.f;
        size_t row = ((size_t)b * p.c_out + co) * p.t_out;
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
            int t = t0 + lane + 32 * j;
            if (t < p.t_out) {
                float v = acc[i][j] + bv;
                if (p.res) v += p.res[row + t];
                p.y[row + t] = v;
            }
        }
    }
}

template <int CPT>
static int launch_conv1d(const ConvParams& p, int batch, cudaStream_t s) {
    constexpr int TPT = 4, BM = 8 * CPT, BN = 32 * TPT;
    ConvParams q = p;
    q.ci_chunk = max(1, min(p.c_in, 32 / p.K));
    q.span = (BN - 1) * p.stride + (p.K - 1) * p.dil + 1;
    q.PL = acb_ceil_div(q.span, p.stride);
    size_t smem = ((size_t)q.ci_chunk * p.K * BM + (size_t)q.ci_chunk * p.stride * q.PL) * sizeof(float) +
                  p.K * sizeof(int);
    ACB_REQUIRE(smem <= 200 * 1024, "acb_conv1d: tile needs %zu B of shared memory", smem);
    if (smem > 48 * 1024)
        ACB_CHECK_CUDA(cudaFuncSetAttribute(conv1d_kernel<CPT, TPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
    dim3 grid(acb_ceil_div(p.t_out, BN), acb_ceil_div(p.c_out, BM), batch);
    conv1d_kernel<CPT, TPT><<<grid, 256, smem, s>>>(q);
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}

extern "C" int acb_conv1d(const float* x, const float* w_packed, const float* bias, const float* residual, float* y,
                          int batch, int c_in, int c_out, int t_in, int t_virtual, int t_out, int kernel, int stride,
                          int dilation, int pad_left, int reflect, int elu_in, void* stream) {
    ACB_REQUIRE(x && w_packed && y, "acb_conv1d: null pointer");
    ACB_REQUIRE(batch > 0 && c_in > 0 && c_out > 0 && t_in > 0 && t_out > 0, "acb_conv1d: empty shape");
    ACB_REQUIRE(kernel >= 1 && kernel <= 64 && stride >= 1 && dilation >= 1 && pad_left >= 0, "acb_conv1d: bad taps");
    ACB_REQUIRE(t_virtual >= t_in, "acb_conv1d: t_virtual < t_in");
    ACB_REQUIRE(batch <= 65535, "acb_conv1d: batch > 65535");
    ConvParams p{x, w_packed, bias, residual, y, c_in, c_out, t_in, t_virtual, t_out, kernel, stride, dilation,
                 pad_left, reflect, elu_in, 0, 0, 0};
    cudaStream_t s = (cudaStream_t)stream;
    if (c_out >= 64) return launch_conv1d<8>(p, batch, s);
    if (c_out >= 32) return launch_conv1d<4>(p, batch, s);
    if (c_out >= 16) return launch_conv1d<2>(p, batch, s);
    return launch_conv1d<1>(p, batch, s);
}

// ------------------------------------------------------------------------------------------------
// transposed conv1d (kernel = 2*stride) + trim, as S interleaved 2-tap convolutions: lane = input step ti,
// every thread produces the S consecutive outputs u = ti*S + p for CPT channels, so x is read once per
// input channel and reused over all phases.
// ------------------------------------------------------------------------------------------------
struct ConvTrParams {
    const float* x; const float* w; const float* bias; float* y;
    int c_in, c_out, t_in, t_out, trim_left, elu, ci_chunk;
};

template <int S, int CPT>
__global__ void __launch_bounds__(256) convtr1d_kernel(ConvTrParams p) {
    constexpr int K = 2 * S, BM = 8 * CPT;
    extern __shared__ float smem[];
    float* ws = smem;                        // [ci_chunk][K][BM]
    float* xs = ws + p.ci_chunk * K * BM;    // [ci_chunk][33]: slot l+1 <-> ti0+l, slot 0 <-> ti0-1

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ti0 = blockIdx.x * 32, co0 = blockIdx.y * BM, b = blockIdx.z;
    const float* xb = p.x + (size_t)b * p.c_in * p.t_in;

    float acc[CPT][S];
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int q = 0; q < S; ++q) acc[i][q] = 0.f;

    for (int ci0 = 0; ci0 < p.c_in; ci0 += p.ci_chunk) {
        const int nci = min(p.ci_chunk, p.c_in - ci0);
        __syncthreads();
        for (int idx = tid; idx < nci * K * BM; idx += 256) {
            int r = idx / BM, c = idx - r * BM;
            ws[idx] = (co0 + c < p.c_out) ? p.w[((size_t)ci0 * K + r) * p.c_out + co0 + c] : 0.f;
        }
        for (int idx = tid; idx < nci * 33; idx += 256) {
            int cl = idx / 33, l = idx - cl * 33;
            int ti = ti0 + l - 1;
            float v = (ti >= 0 && ti < p.t_in) ? xb[(size_t)(ci0 + cl) * p.t_in + ti] : 0.f;
            xs[idx] = p.elu ? acb_elu(v) : v;
        }
        __syncthreads();
        for (int cl = 0; cl < nci; ++cl) {
            const float xa = xs[cl * 33 + lane + 1], xp = xs[cl * 33 + lane];
            const float* wc = ws + cl * K * BM + warp * CPT;
#pragma unroll
            for (int q = 0; q < S; ++q) {
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    acc[i][q] = fmaf(xa, wc[q * BM + i], acc[i][q]);
                    acc[i][q] = fmaf(xp, wc[(q + S) * BM + i], acc[i][q]);
                }
            }
        }
    }
    const int ti = ti0 + lane;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        int co = co0 + warp * CPT + i;
        if (co >= p.c_out) continue;
        float bv = p.bias ? p.bias[co] : 0.f;
        size_t row = ((size_t)b * p.c_out + co) * p.t_out;
#pragma unroll
        for (int q = 0; q < S; ++q) {
            int o = ti * S + q - p.trim_left;
            if (o >= 0 && o < p.t_out) p.y[row + o] = acc[i][q] + bv;
        }
    }
}

template <int S>
static int launch_convtr(const ConvTrParams& p, int batch, cudaStream_t s) {
    constexpr int CPT = (S >= 8) ? 4 : 8, BM = 8 * CPT, K = 2 * S;
    ConvTrParams q = p;
    q.ci_chunk = max(1, min(p.c_in, 64 / K));
    size_t smem = ((size_t)q.ci_chunk * K * BM + (size_t)q.ci_chunk * 33) * sizeof(float);
    int n_ti = acb_ceil_div(p.t_out + p.trim_left, S);
    dim3 grid(acb_ceil_div(n_ti, 32), acb_ceil_div(p.c_out, BM), batch);
    convtr1d_kernel<S, CPT><<<grid, 256, smem, s>>>(q);
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}

extern "C" int acb_convtr1d(const float* x, const float* w_packed, const float* bias, float* y, int batch, int c_in,
                            int c_out, int t_in, int t_out, int kernel, int stride, int trim_left, int elu_in,
                            void* stream) {
    ACB_REQUIRE(x && w_packed && y, "acb_convtr1d: null pointer");
    ACB_REQUIRE(batch > 0 && batch <= 65535 && c_in > 0 && c_out > 0 && t_in > 0 && t_out > 0, "acb_convtr1d: bad shape");
    ACB_REQUIRE(kernel == 2 * stride, "acb_convtr1d: only kernel == 2*stride is built (got k=%d s=%d)", kernel, stride);
    ACB_REQUIRE(trim_left >= 0 && t_out + trim_left <= (t_in + 1) * stride, "acb_convtr1d: trim out of range");
    ConvTrParams p{x, w_packed, bias, y, c_in, c_out, t_in, t_out, trim_left, elu_in, 0};
    cudaStream_t s = (cudaStream_t)stream;
    switch (stride) {
        case 2: return launch_convtr<2>(p, batch, s);
        case 3: return launch_convtr<3>(p, batch, s);
        case 4: return launch_convtr<4>(p, batch, s);
        case 5: return launch_convtr<5>(p, batch, s);
        case 8: return launch_convtr<8>(p, batch, s);
    }
    acb_set_error("acb_convtr1d: stride %d not built (2,3,4,5,8)", stride);
    return ACB_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// LSTM recurrence: one persistent cooperative kernel for all T steps.  CTA c owns hidden units
// [c*U, (c+1)*U) i.e. 4U rows of W_hh, held in shared memory for the whole sequence; per step it reads
// h_{t-1} (all units) from L2, computes its 4U x B gate pre-activations, applies the cell update for its
// units and publishes h_t; a grid-wide barrier separates the steps.
// ------------------------------------------------------------------------------------------------
struct LstmParams {
    const float* gx; const float* whh; const float* skip; float* y; float* hbuf; unsigned* bar;
    int B, H, T, U;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

constexpr int LSTM_BC = 8;  // batch items per matvec pass

__global__ void __launch_bounds__(256) lstm_kernel(LstmParams p) {
    extern __shared__ float smem[];
    const int H = p.H, U = p.U, R = 4 * U;
    float* wsm = smem;                      // [R][H]
    float* hs = wsm + (size_t)R * H;        // [LSTM_BC][H]
    float* gs = hs + (size_t)LSTM_BC * H;   // [R][LSTM_BC]
    float* cs = gs + R * LSTM_BC;           // [U][B] cell state

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int unit0 = blockIdx.x * U;
    const unsigned ncta = gridDim.x;

    for (int idx = tid; idx < R * (H / 4); idx += 256) {
        int r = idx / (H / 4), c4 = idx - r * (H / 4);
        int gate = r / U, u = r - gate * U;
        reinterpret_cast<float4*>(wsm)[idx] =
            reinterpret_cast<const float4*>(p.whh + ((size_t)gate * H + unit0 + u) * H)[c4];
    }
    for (int idx = tid; idx < U * p.B; idx += 256) cs[idx] = 0.f;
    __syncthreads();

    for (int t = 0; t < p.T; ++t) {
        const float* hprev = p.hbuf + (size_t)(t & 1) * p.B * H;
        float* hnext = p.hbuf + (size_t)((t + 1) & 1) * p.B * H;
        for (int b0 = 0; b0 < p.B; b0 += LSTM_BC) {
            const int nb = min(LSTM_BC, p.B - b0);
            // gate inputs for this CTA's (unit, batch) pairs: issued early, consumed after the matvec
            float gxr[4] = {0.f, 0.f, 0.f, 0.f};
            const int pu = tid / LSTM_BC, pb = tid % LSTM_BC;
            const bool pw = (tid < U * LSTM_BC) && (pb < nb);
            if (pw) {
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    gxr[g] = __ldg(p.gx + ((size_t)(b0 + pb) * 4 * H + (size_t)g * H + unit0 + pu) * p.T + t);
            }
            if (t > 0) {
                for (int idx = tid; idx < LSTM_BC * (H / 4); idx += 256) {
                    int bb = idx / (H / 4), c4 = idx - bb * (H / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bb < nb) v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)(b0 + bb) * H) + c4);
                    reinterpret_cast<float4*>(hs)[idx] = v;
                }
                __syncthreads();
                for (int r0 = warp * 4; r0 < R; r0 += 32) {
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                    for (int kk = lane * 4; kk < H; kk += 128) {
                        float4 w4[4], h4[LSTM_BC];
#pragma unroll
                        for (int r = 0; r < 4; ++r) w4[r] = *reinterpret_cast<const float4*>(wsm + (size_t)(r0 + r) * H + kk);
#pragma unroll
                        for (int bb = 0; bb < LSTM_BC; ++bb) h4[bb] = *reinterpret_cast<const float4*>(hs + (size_t)bb * H + kk);
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int bb = 0; bb < LSTM_BC; ++bb) {
                                float a = acc[r * LSTM_BC + bb];
                                a = fmaf(w4[r].x, h4[bb].x, a);
                                a = fmaf(w4[r].y, h4[bb].y, a);
                                a = fmaf(w4[r].z, h4[bb].z, a);
                                a = fmaf(w4[r].w, h4[bb].w, a);
                                acc[r * LSTM_BC + bb] = a;
                            }
                    }
                    // transpose-reduce: afterwards lane l holds the warp-wide sum of acc[l]
#pragma unroll
                    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
                        const bool upper = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < n; ++i) {
                            float send = upper ? acc[i] : acc[i + n];
                            float keep = upper ? acc[i + n] : acc[i];
                            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    gs[(r0 + lane / LSTM_BC) * LSTM_BC + (lane % LSTM_BC)] = acc[0];
                }
                __syncthreads();
            }
            if (pw) {
                float gi = gxr[0], gf = gxr[1], gg = gxr[2], go = gxr[3];
                if (t > 0) {
                    gi += gs[(0 * U + pu) * LSTM_BC + pb];
                    gf += gs[(1 * U + pu) * LSTM_BC + pb];
                    gg += gs[(2 * U + pu) * LSTM_BC + pb];
                    go += gs[(3 * U + pu) * LSTM_BC + pb];
                }
                float c = cs[pu * p.B + b0 + pb];
                c = sigmoidf_(gf) * c + sigmoidf_(gi) * tanhf(gg);
                float h = sigmoidf_(go) * tanhf(c);
                cs[pu * p.B + b0 + pb] = c;
                __stcg(hnext + (size_t)(b0 + pb) * H + unit0 + pu, h);
                size_t yo = ((size_t)(b0 + pb) * H + unit0 + pu) * p.T + t;
                p.y[yo] = p.skip ? h + p.skip[yo] : h;
            }
            __syncthreads();  // gs / hs reused by the next batch chunk
        }
        // grid barrier: everyone has published h_t before anyone reads it
        if (tid == 0) {
            __threadfence();
            atomicAdd(p.bar, 1u);
            const unsigned target = ncta * (unsigned)(t + 1);
            while (ld_acquire_u32(p.bar) < target) { }
        }
        __syncthreads();
    }
}

extern "C" int64_t acb_lstm_state_bytes(int batch, int hidden) {
    return ((int64_t)2 * batch * hidden + 64) * (int64_t)sizeof(float);
}

extern "C" int acb_lstm_recurrent(const float* gates_x, const float* w_hh, const float* skip, float* y,
                                  float* state_ws, int batch, int hidden, int t_len, void* stream) {
    ACB_REQUIRE(gates_x && w_hh && y && state_ws, "acb_lstm_recurrent: null pointer");
    ACB_REQUIRE(batch > 0 && hidden > 0 && t_len > 0, "acb_lstm_recurrent: empty shape");
    ACB_REQUIRE(hidden % 4 == 0, "acb_lstm_recurrent: hidden must be a multiple of 4");
    int U = hidden >= 128 ? hidden / 128 : 1;
    ACB_REQUIRE(hidden % U == 0, "acb_lstm_recurrent: hidden %d not divisible by %d units per CTA", hidden, U);
    int ncta = hidden / U;
    size_t smem = ((size_t)4 * U * hidden + (size_t)LSTM_BC * hidden + (size_t)4 * U * LSTM_BC + (size_t)U * batch) *
                  sizeof(float);
    ACB_REQUIRE(smem <= 227 * 1024, "acb_lstm_recurrent: hidden=%d batch=%d needs %zu B smem per CTA", hidden, batch, smem);
    ACB_REQUIRE(U * LSTM_BC <= 256, "acb_lstm_recurrent: too many units per CTA");
    cudaStream_t s = (cudaStream_t)stream;
    ACB_CHECK_CUDA(cudaFuncSetAttribute(lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0;
    ACB_CHECK_CUDA(cudaGetDevice(&dev));
    ACB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ACB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_kernel, 256, smem));
    ACB_REQUIRE(per_sm * sms >= ncta, "acb_lstm_recurrent: %d CTAs cannot be co-resident (%d SMs x %d)", ncta, sms, per_sm);
    size_t hbytes = (size_t)2 * batch * hidden * sizeof(float);
    ACB_CHECK_CUDA(cudaMemsetAsync(state_ws, 0, hbytes + 64 * sizeof(float), s));
    LstmParams p{gates_x, w_hh, skip, y, state_ws, (unsigned*)(state_ws + (size_t)2 * batch * hidden), batch, hidden,
                 t_len, U};
    void* args[] = {&p};
    ACB_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)lstm_kernel, dim3(ncta), dim3(256), args, smem, s));
    return ACB_OK;
}

// ------------------------------------------------------------------------------------------------
// RVQ encode: CTA = 32 frames; residuals live in smem for all n_q rounds; the codebook streams through
// smem in tiles of 64 codes (from L2: 4 x 1 MB codebooks stay resident there).  Thread (warp w, lane f)
// scores frame f against 8 codes of every tile.
// ------------------------------------------------------------------------------------------------
constexpr int RVQ_F = 32, RVQ_TILE = 64;

__global__ void __launch_bounds__(256) rvq_encode_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                         const float* __restrict__ cbn, int64_t* __restrict__ codes,
                                                         int B, int D, int T, int n_q, int bins) {
    extern __shared__ float smem[];
    float* rs = smem;                      // [D][32] residual, frame fastest
    float* cs = rs + D * RVQ_F;            // [RVQ_TILE][D+1]
    float* bs = cs + RVQ_TILE * (D + 1);   // [8][32] best score per warp
    int* bi = (int*)(bs + 8 * RVQ_F);      // [8][32] best index per warp
    int* sel = bi + 8 * RVQ_F;             // [32] chosen code per frame

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n0 = (long long)blockIdx.x * RVQ_F;
    const long long N = (long long)B * T;
    const long long n = n0 + lane;
    const bool live = n < N;
    const int fb = live ? (int)(n / T) : 0, ft = live ? (int)(n % T) : 0;

    for (int d = warp; d < D; d += 8) rs[d * RVQ_F + lane] = live ? z[((size_t)fb * D + d) * T + ft] : 0.f;
    __syncthreads();

    for (int q = 0; q < n_q; ++q) {
        const float* cbq = cb + (size_t)q * bins * D;
        float xx = 0.f;
        for (int d = 0; d < D; ++d) { float r = rs[d * RVQ_F + lane]; xx = fmaf(r, r, xx); }
        float best = -INFINITY;
        int besti = 0;
        for (int c0 = 0; c0 < bins; c0 += RVQ_TILE) {
            __syncthreads();
            for (int idx = tid; idx < RVQ_TILE * D; idx += 256) {
                int c = idx / D, d = idx - c * D;
                cs[c * (D + 1) + d] = (c0 + c < bins) ? cbq[(size_t)(c0 + c) * D + d] : 0.f;
            }
            __syncthreads();
            float dot[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) dot[i] = 0.f;
            const float* cw = cs + (warp * 8) * (D + 1);
            for (int d = 0; d < D; ++d) {
                float r = rs[d * RVQ_F + lane];
#pragma unroll
                for (int i = 0; i < 8; ++i) dot[i] = fmaf(r, cw[i * (D + 1) + d], dot[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int c = c0 + warp * 8 + i;
                if (c < bins) {
                    // same association as core_vq.py:166-170: -((|x|^2 - 2 x.e) + |e|^2)
                    float sc = -((xx - 2.f * dot[i]) + cbn[(size_t)q * bins + c]);
                    if (sc > best) { best = sc; besti = c; }
                }
            }
        }
        bs[warp * RVQ_F + lane] = best;
        bi[warp * RVQ_F + lane] = besti;
        __syncthreads();
        if (warp == 0) {
            float m = bs[lane];
            int mi = bi[lane];
            for (int w = 1; w < 8; ++w) {
                float v = bs[w * RVQ_F + lane];
                int vi = bi[w * RVQ_F + lane];
                if (v > m || (v == m && vi < mi)) { m = v; mi = vi; }  // first max wins (torch.max)
            }
            sel[lane] = mi;
            if (live) codes[((size_t)fb * n_q + q) * T + ft] = mi;
        }
        __syncthreads();
        const int mine = sel[lane];
        for (int d = warp; d < D; d += 8) rs[d * RVQ_F + lane] -= cbq[(size_t)mine * D + d];
        __syncthreads();
    }
}

extern "C" int acb_rvq_encode(const float* latent, const float* codebooks, const float* cb_sqnorm, int64_t* codes,
                              int batch, int dim, int t_len, int n_q, int bins, void* stream) {
    ACB_REQUIRE(latent && codebooks && cb_sqnorm && codes, "acb_rvq_encode: null pointer");
    ACB_REQUIRE(batch > 0 && t_len > 0 && n_q > 0 && bins > 0, "acb_rvq_encode: empty shape");
    ACB_REQUIRE(dim > 0 && dim <= 512, "acb_rvq_encode: dim %d out of range (<=512)", dim);
    size_t smem = ((size_t)dim * RVQ_F + (size_t)RVQ_TILE * (dim + 1) + 8 * RVQ_F) * sizeof(float) +
                  (8 * RVQ_F + RVQ_F) * sizeof(int);
    if (smem > 48 * 1024)
        ACB_CHECK_CUDA(cudaFuncSetAttribute(rvq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long N = (long long)batch * t_len;
    rvq_encode_kernel<<<(unsigned)((N + RVQ_F - 1) / RVQ_F), 256, smem, (cudaStream_t)stream>>>(
        latent, codebooks, cb_sqnorm, codes, batch, dim, t_len, n_q, bins);
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}

// RVQ decode: latent[b][d][t] = sum_k E_k[codes[b][k][t]][d]; 32 frames per CTA, transposed through smem so
// both the codebook row reads and the [B][D][T] writes are coalesced.
__global__ void __launch_bounds__(256) rvq_decode_kernel(const int64_t* __restrict__ codes, const float* __restrict__ cb,
                                                         float* __restrict__ out, int B, int D, int T, int n_q,
                                                         int bins) {
    extern __shared__ float smem[];  // [32][D+1]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n0 = (long long)blockIdx.x * 32, N = (long long)B * T;
    for (int f = warp; f < 32; f += 8) {
        long long n = n0 + f;
        if (n >= N) continue;
        int b = (int)(n / T), t = (int)(n % T);
        for (int d = lane; d < D; d += 32) {
            float acc = 0.f;
            for (int q = 0; q < n_q; ++q) {
                long long c = codes[((size_t)b * n_q + q) * T + t];
                c = c < 0 ? 0 : (c >= bins ? bins - 1 : c);
                acc += cb[((size_t)q * bins + c) * D + d];
            }
            smem[f * (D + 1) + d] = acc;
        }
    }
    __syncthreads();
    long long n = n0 + lane;
    if (n < N) {
        int b = (int)(n / T), t = (int)(n % T);
        for (int d = warp; d < D; d += 8) out[((size_t)b * D + d) * T + t] = smem[lane * (D + 1) + d];
    }
}

extern "C" int acb_rvq_decode(const int64_t* codes, const float* codebooks, float* latent, int batch, int dim, int t_len,
                              int n_q, int bins, void* stream) {
    ACB_REQUIRE(codes && codebooks && latent, "acb_rvq_decode: null pointer");
    ACB_REQUIRE(batch > 0 && dim > 0 && t_len > 0 && n_q > 0 && bins > 0, "acb_rvq_decode: empty shape");
    size_t smem = (size_t)32 * (dim + 1) * sizeof(float);
    ACB_REQUIRE(smem <= 48 * 1024, "acb_rvq_decode: dim %d too large", dim);
    long long N = (long long)batch * t_len;
    rvq_decode_kernel<<<(unsigned)((N + 31) / 32), 256, smem, (cudaStream_t)stream>>>(codes, codebooks, latent, batch, dim,
                                                                                     t_len, n_q, bins);
    ACB_LAUNCH_CHECK();
    return ACB_OK;
}
