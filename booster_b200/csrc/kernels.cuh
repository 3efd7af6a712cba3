// kernels.cuh — hand-written sm_100a kernels for the quantized LLaMA decode path.
//
// Arithmetic contract (what "parity" means; the oracle is the reference CPU path):
//   * activations are quantized exactly as the CPU reference does before every quantized matmul:
//     Q8_K (256-wide, fp32 scale, -127/max sign trick, round-half-even) for Q4_K/Q5_K/Q6_K weights
//     (cpp/ggml/src/ggml-quants.c:3593-3630) and Q8_0 (32-wide, fp16 scale) for Q8_0 weights (:936-1000);
//     NOT the Q8_1 scheme of the reference's CUDA backend (cpp/ggml/src/ggml-cuda/quantize.cu:4-38).
//   * inside one 256-weight super-block all integer sums are exact, identical to
//     ggml_vec_dot_q{4,5,6}_K_q8_K (cpp/ggml/src/ggml-quants.c:6832,7400,8037); only the fp32 summation
//     order across super-blocks differs (parallel tree instead of 8 AVX lanes).
//   * RMSNorm accumulates x*x in double like ggml_compute_forward_rms_norm_f32 (cpp/ggml/src/ggml.c:11850).
//   * RoPE takes cos/sin from a host-built table that follows ggml_rope_cache_init (cpp/ggml/src/ggml.c:14017).
//   * attention follows the DEFAULT (non-flash) route of llm_build_kqv at batch 1
//     (cpp/src/llama.cpp:8248-8297): f16 K/V widened to f32, f32 q, f32 softmax.
//
// Weight layout in HBM ("planes"): every matrix keeps its GGUF block bytes but split per field so that
// each plane is a dense 16-byte-aligned stream per row (Q4_K's 144 B and Q6_K's 210 B blocks are not):
//   Q4_K: p0 qs[n][nb*128]  p1 scales[n][nb*12]  p2 dm(half2)[n][nb]
//   Q5_K: p0 qs[n][nb*128]  p1 scales[n][nb*12]  p2 dm(half2)[n][nb]   p3 qh[n][nb*32]
//   Q6_K: p0 ql[n][nb*128]  p1 qh[n][nb*64]      p2 scales(i8)[n][nb*16] p3 d(half)[n][nb]
//   Q8_0: p0 qs[n][k]       p1 d(half)[n][k/32]
// Total bytes are exactly the GGUF tensor bytes (the algorithmic bytes of SURVEY.md §8d).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum { T_F32 = 0, T_F16 = 1, T_Q8_0 = 8, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14 };

struct QMat {
    int type = 0;
    int n_rows = 0;
    int k = 0;
    const uint8_t * p0 = nullptr;
    const uint8_t * p1 = nullptr;
    const uint8_t * p2 = nullptr;
    const uint8_t * p3 = nullptr;
};

// per-token scalars, device-resident so that one CUDA graph serves every token
struct DecodeState {
    int32_t token;     // input token id of this step
    int32_t pos;       // its position == KV slot it is written to
    int32_t round_q;   // 1: round q to f16 before K.q (reference behaviour for batch > 1, ggml.c:12345-12371)
    int32_t step;      // greedy loop: index into out_tokens
};

enum { EPI_STORE = 0, EPI_RESID = 1, EPI_QKV = 2, EPI_SILU = 3 };
enum { PAIR_ADJACENT = 0, PAIR_ZIP = 1 };

struct MatvecArgs {
    QMat seg[3];
    int n_seg;
    int pair_mode;
    int n_pairs;
    int k;
    // prologue: x f32[k]; optional RMSNorm (norm_w != nullptr) then activation quantization into smem
    const float * x;
    const float * norm_w;
    float eps;
    int act_q8_0;              // 0: Q8_K activations, 1: Q8_0 activations
    // epilogue
    float * out;
    const float * resid;
    // EPI_QKV
    float * q_out;
    __half * k_cache;          // this layer's [n_ctx][kv_dim]
    __half * v_cache;
    int n_q, n_k, head_dim, kv_dim;
    const float2 * rope;       // [n_ctx][head_dim/2] (cos, sin)
    const DecodeState * st;
};

static constexpr int MV_THREADS = 256;
static constexpr int MV_WARPS   = MV_THREADS / 32;

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_v4(const void * p) {
    // streaming 16-byte load: read-only path, do not pollute L1 (weights are touched once per token)
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_u32(const void * p) { return __ldg((const uint32_t *) p); }
__device__ __forceinline__ uint32_t ldg_u16(const void * p) { return __ldg((const uint16_t *) p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int dot16(const uint4 & w, const int4 & a, int acc) {
    acc = __dp4a((int) w.x, a.x, acc);
    acc = __dp4a((int) w.y, a.y, acc);
    acc = __dp4a((int) w.z, a.z, acc);
    acc = __dp4a((int) w.w, a.w, acc);
    return acc;
}
__device__ __forceinline__ uint4 and4(const uint4 & v, uint32_t m) { return make_uint4(v.x & m, v.y & m, v.z & m, v.w & m); }
__device__ __forceinline__ uint4 shr4(const uint4 & v, int s)      { return make_uint4(v.x >> s, v.y >> s, v.z >> s, v.w >> s); }
__device__ __forceinline__ uint4 shl4(const uint4 & v, int s)      { return make_uint4(v.x << s, v.y << s, v.z << s, v.w << s); }
__device__ __forceinline__ uint4 or4(const uint4 & a, const uint4 & b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }

// ------------------------------------------------------------------------------------------------------------
// activation quantization (block-cooperative, result in shared memory)
//   smem layout: int8 q[k] | float dx[k/256 or k/32] | int16 bsums[k/16] (Q8_K only)
// ------------------------------------------------------------------------------------------------------------
struct ActSmem {
    int8_t  * q;
    float   * dx;
    int16_t * bsums;
};
__host__ __device__ __forceinline__ size_t act_smem_bytes(int k, int act_q8_0) {
    size_t n = (size_t) k;                                    // q
    n = (n + 15) / 16 * 16;
    n += (size_t) (act_q8_0 ? k / 32 : k / 256) * 4;         // dx
    n = (n + 15) / 16 * 16;
    if (!act_q8_0) n += (size_t) (k / 16) * 2;               // bsums
    return (n + 15) / 16 * 16;
}
__device__ __forceinline__ ActSmem act_smem_carve(uint8_t * base, int k, int act_q8_0) {
    ActSmem a;
    a.q = (int8_t *) base;
    size_t off = ((size_t) k + 15) / 16 * 16;
    a.dx = (float *) (base + off);
    off += (size_t) (act_q8_0 ? k / 32 : k / 256) * 4;
    off = (off + 15) / 16 * 16;
    a.bsums = (int16_t *) (base + off);
    return a;
}

// One warp quantizes 256 consecutive values (8 per lane) to Q8_K. Restates quantize_row_q8_K_ref
// (cpp/ggml/src/ggml-quants.c:3593-3630): `max` is the FIRST element of largest magnitude (strict >),
// iscale = -127/max, q = min(127, round_half_even(iscale*x)), d = 1/iscale, bsums over groups of 16.
__device__ __forceinline__ void q8k_block_warp(const float (&v)[8], int lane, int8_t * q_out /* 256 */,
                                               float * d_out, int16_t * bsums_out /* 16 */) {
    float amax = 0.f, mval = 0.f;
    int   midx = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float ax = fabsf(v[i]);
        if (ax > amax) { amax = ax; mval = v[i]; midx = lane * 8 + i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float oa = __shfl_xor_sync(0xffffffffu, amax, o);
        const float ov = __shfl_xor_sync(0xffffffffu, mval, o);
        const int   oi = __shfl_xor_sync(0xffffffffu, midx, o);
        if (oa > amax || (oa == amax && oi < midx)) { amax = oa; mval = ov; midx = oi; }
    }
    uint32_t w0 = 0, w1 = 0;
    int s8 = 0;
    float d = 0.f;
    if (amax != 0.f) {
        const float iscale = __fdiv_rn(-127.f, mval);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int qi = __float2int_rn(__fmul_rn(iscale, v[i]));
            qi = min(127, qi);
            s8 += qi;
            if (i < 4) w0 |= ((uint32_t) (qi & 0xff)) << (8 * i);
            else       w1 |= ((uint32_t) (qi & 0xff)) << (8 * (i - 4));
        }
        d = __fdiv_rn(1.f, iscale);
    }
    *reinterpret_cast<uint2 *>(q_out + lane * 8) = make_uint2(w0, w1);
    const int s16 = s8 + __shfl_xor_sync(0xffffffffu, s8, 1);
    if ((lane & 1) == 0) bsums_out[lane >> 1] = (int16_t) s16;
    if (lane == 0) *d_out = d;
}

// One warp quantizes 256 consecutive values = 8 Q8_0 blocks of 32 (4 lanes each). Restates the AVX path of
// quantize_row_q8_0 (cpp/ggml/src/ggml-quants.c:936-1000): d = amax/127 stored as fp16, id = 127/amax,
// q = round_half_even(x*id). The dot product later uses the fp16-rounded d.
__device__ __forceinline__ void q80_blocks_warp(const float (&v)[8], int lane, int8_t * q_out /* 256 */,
                                                float * d_out /* 8 */) {
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    const float d  = __fdiv_rn(amax, 127.f);
    const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
    uint32_t w0 = 0, w1 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int qi = __float2int_rn(__fmul_rn(v[i], id));
        if (i < 4) w0 |= ((uint32_t) (qi & 0xff)) << (8 * i);
        else       w1 |= ((uint32_t) (qi & 0xff)) << (8 * (i - 4));
    }
    *reinterpret_cast<uint2 *>(q_out + lane * 8) = make_uint2(w0, w1);
    if ((lane & 3) == 0) d_out[lane >> 2] = __half2float(__float2half_rn(d));
}

// Block-cooperative prologue: optional RMSNorm(+weight) then activation quantization into shared memory.
// RMSNorm restates ggml_compute_forward_rms_norm_f32 (cpp/ggml/src/ggml.c:11850-11896: double accumulation
// of float x*x, scale = 1/sqrtf(mean+eps), y = x*scale) followed by the separate ggml_mul with the norm
// weight (cpp/src/llama.cpp:7928-7958).
__device__ __forceinline__ void prologue_quantize(const float * __restrict__ x, const float * __restrict__ norm_w,
                                                  float eps, int k, int act_q8_0, const ActSmem & A,
                                                  float * red_smem /* >= 16 doubles */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarp = blockDim.x >> 5;
    float scale = 1.f;
    if (norm_w != nullptr) {
        double s = 0.0;
        for (int i = tid * 4; i < k; i += blockDim.x * 4) {
            const float4 v = *reinterpret_cast<const float4 *>(x + i);
            s += (double) __fmul_rn(v.x, v.x);
            s += (double) __fmul_rn(v.y, v.y);
            s += (double) __fmul_rn(v.z, v.z);
            s += (double) __fmul_rn(v.w, v.w);
        }
        s = warp_sum_d(s);
        double * red = reinterpret_cast<double *>(red_smem);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < nwarp; w++) tot += red[w];
        const float mean = (float) (tot / (double) k);
        scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    }
    const int n256 = k / 256;
    for (int b = warp; b < n256; b += nwarp) {
        const float * xb = x + b * 256 + lane * 8;
        const float4 a0 = *reinterpret_cast<const float4 *>(xb);
        const float4 a1 = *reinterpret_cast<const float4 *>(xb + 4);
        float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (norm_w != nullptr) {
            const float4 w0 = *reinterpret_cast<const float4 *>(norm_w + b * 256 + lane * 8);
            const float4 w1 = *reinterpret_cast<const float4 *>(norm_w + b * 256 + lane * 8 + 4);
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fmul_rn(v[i], scale), ww[i]);
        }
        if (act_q8_0) q80_blocks_warp(v, lane, A.q + b * 256, A.dx + b * 8);
        else          q8k_block_warp(v, lane, A.q + b * 256, A.dx + b, A.bsums + b * 16);
    }
}

// ------------------------------------------------------------------------------------------------------------
// per-type dot products: R rows at a time against the quantized activation vector in shared memory.
// Each returns per-lane partial sums in acc[]; the caller warp-reduces.
// ------------------------------------------------------------------------------------------------------------
struct RowPtrs { const uint8_t * p0; const uint8_t * p1; const uint8_t * p2; const uint8_t * p3; };

// scale / min bytes of sub-blocks 2j and 2j+1 from the 12 packed bytes (format: get_scale_min_k4,
// cpp/ggml/src/ggml-quants.c:1891-1898); returns sc_lo | sc_hi<<8 in .x and m_lo | m_hi<<8 in .y
__device__ __forceinline__ uint2 k4_scales_pair(uint32_t s0, uint32_t s1, uint32_t s2, int j) {
    const uint32_t sc_a = s0 & 0x3f3f3f3fu;
    const uint32_t m_a  = s1 & 0x3f3f3f3fu;
    const uint32_t sc_b = (s2 & 0x0f0f0f0fu) | (((s0 >> 6) & 0x03030303u) << 4);
    const uint32_t m_b  = ((s2 >> 4) & 0x0f0f0f0fu) | (((s1 >> 6) & 0x03030303u) << 4);
    const uint32_t scw = j < 2 ? sc_a : sc_b;
    const uint32_t mw  = j < 2 ? m_a : m_b;
    const int sh = (j & 1) * 16;
    return make_uint2((scw >> sh) & 0xffffu, (mw >> sh) & 0xffffu);
}

template <int R, int UNR, bool Q5>
__device__ __forceinline__ void dot_q45k(const RowPtrs (&rp)[R], int nb, int lane, const ActSmem & A, float (&acc)[R]) {
    // unit = one 16-byte chunk of qs: sub-block pair j, half h -> 16 low-nibble + 16 high-nibble weights
    const int n_units = nb * 8;
    const int cc = lane & 7, j = cc >> 1, h = cc & 1;
    const int a_off = 64 * j + 16 * h;
    for (int u0 = lane; u0 < n_units; u0 += 32 * UNR) {
        uint4    w[UNR][R];
        uint4    qh[UNR][R];
        uint32_t s0[UNR][R], s1[UNR][R], s2[UNR][R], dmv[UNR][R];
#pragma unroll
        for (int t = 0; t < UNR; t++) {
            const int u = u0 + 32 * t;
            if (u < n_units) {
                const int b = u >> 3;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    w[t][r]   = ldg_stream_v4(rp[r].p0 + (size_t) u * 16);
                    s0[t][r]  = ldg_u32(rp[r].p1 + b * 12);
                    s1[t][r]  = ldg_u32(rp[r].p1 + b * 12 + 4);
                    s2[t][r]  = ldg_u32(rp[r].p1 + b * 12 + 8);
                    dmv[t][r] = ldg_u32(rp[r].p2 + b * 4);
                    if (Q5) qh[t][r] = ldg_stream_v4(rp[r].p3 + b * 32 + 16 * h);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < UNR; t++) {
            const int u = u0 + 32 * t;
            if (u < n_units) {
                const int b = u >> 3;
                const int4 alo = *reinterpret_cast<const int4 *>(A.q + b * 256 + a_off);
                const int4 ahi = *reinterpret_cast<const int4 *>(A.q + b * 256 + a_off + 32);
                const float dx = A.dx[b];
                const int bs_lo = A.bsums[b * 16 + 4 * j + h];
                const int bs_hi = A.bsums[b * 16 + 4 * j + 2 + h];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint4 lo = and4(w[t][r], 0x0f0f0f0fu);
                    uint4 hi = and4(shr4(w[t][r], 4), 0x0f0f0f0fu);
                    if (Q5) {
                        // qh bit 2j -> +16 on the low-nibble weights, bit 2j+1 -> +16 on the high-nibble ones
                        // (dequantize_row_q5_K, cpp/ggml/src/ggml-quants.c:2756-2782)
                        lo = or4(lo, shl4(and4(shr4(qh[t][r], 2 * j), 0x01010101u), 4));
                        hi = or4(hi, shl4(and4(shr4(qh[t][r], 2 * j + 1), 0x01010101u), 4));
                    }
                    const int isum_lo = dot16(lo, alo, 0);
                    const int isum_hi = dot16(hi, ahi, 0);
                    const uint2 sm = k4_scales_pair(s0[t][r], s1[t][r], s2[t][r], j);
                    const int isc = (int) (sm.x & 0xff) * isum_lo + (int) (sm.x >> 8) * isum_hi;
                    const int ism = (int) (sm.y & 0xff) * bs_lo + (int) (sm.y >> 8) * bs_hi;
                    const __half2 dmh = *reinterpret_cast<const __half2 *>(&dmv[t][r]);
                    const float d    = __low2float(dmh) * dx;
                    const float dmin = __high2float(dmh) * dx;
                    acc[r] = fmaf(d, (float) isc, acc[r]);
                    acc[r] = fmaf(-dmin, (float) ism, acc[r]);
                }
            }
        }
    }
}

template <int R, int UNR>
__device__ __forceinline__ void dot_q6k(const RowPtrs (&rp)[R], int nb, int lane, const ActSmem & A, float (&acc)[R]) {
    // unit = 64 weights: half n of the super-block, 16-lane slice uu: ql[64n+16uu], ql[64n+32+16uu], qh[32n+16uu]
    // (layout: dequantize_row_q6_K, cpp/ggml/src/ggml-quants.c:2970-3000)
    const int n_units = nb * 4;
    const int n = (lane >> 1) & 1, uu = lane & 1;
    for (int u0 = lane; u0 < n_units; u0 += 32 * UNR) {
        uint4    ql0[UNR][R], ql1[UNR][R], qh[UNR][R];
        uint32_t sc4[UNR][R][2];
        uint32_t dv[UNR][R];
#pragma unroll
        for (int t = 0; t < UNR; t++) {
            const int u = u0 + 32 * t;
            if (u < n_units) {
                const int b = u >> 2;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    ql0[t][r] = ldg_stream_v4(rp[r].p0 + (size_t) b * 128 + 64 * n + 16 * uu);
                    ql1[t][r] = ldg_stream_v4(rp[r].p0 + (size_t) b * 128 + 64 * n + 32 + 16 * uu);
                    qh[t][r]  = ldg_stream_v4(rp[r].p1 + (size_t) b * 64 + 32 * n + 16 * uu);
                    // the 8 int8 scales of half n (two aligned words); group g uses byte 2g+uu
                    sc4[t][r][0] = ldg_u32(rp[r].p2 + b * 16 + 8 * n);
                    sc4[t][r][1] = ldg_u32(rp[r].p2 + b * 16 + 8 * n + 4);
                    dv[t][r]     = ldg_u16(rp[r].p3 + b * 2);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < UNR; t++) {
            const int u = u0 + 32 * t;
            if (u < n_units) {
                const int b = u >> 2;
                const int8_t * ab = A.q + b * 256 + 128 * n + 16 * uu;
                const int4 a0 = *reinterpret_cast<const int4 *>(ab);
                const int4 a1 = *reinterpret_cast<const int4 *>(ab + 32);
                const int4 a2 = *reinterpret_cast<const int4 *>(ab + 64);
                const int4 a3 = *reinterpret_cast<const int4 *>(ab + 96);
                const float dx = A.dx[b];
                const int16_t * bs = A.bsums + b * 16 + 8 * n + uu;
                const int bs0 = bs[0], bs1 = bs[2], bs2 = bs[4], bs3 = bs[6];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint4 h = qh[t][r];
                    const uint4 q0 = or4(and4(ql0[t][r], 0x0f0f0f0fu), and4(shl4(h, 4), 0x30303030u));
                    const uint4 q1 = or4(and4(ql1[t][r], 0x0f0f0f0fu), and4(shl4(h, 2), 0x30303030u));
                    const uint4 q2 = or4(and4(shr4(ql0[t][r], 4), 0x0f0f0f0fu), and4(h, 0x30303030u));
                    const uint4 q3 = or4(and4(shr4(ql1[t][r], 4), 0x0f0f0f0fu), and4(shr4(h, 2), 0x30303030u));
                    const int d0 = dot16(q0, a0, 0) - 32 * bs0;
                    const int d1 = dot16(q1, a1, 0) - 32 * bs1;
                    const int d2 = dot16(q2, a2, 0) - 32 * bs2;
                    const int d3 = dot16(q3, a3, 0) - 32 * bs3;
                    const uint32_t sa = sc4[t][r][0], sb = sc4[t][r][1];
                    const int sh = 8 * uu;
                    const int c0 = (int) (int8_t) ((sa >> sh) & 0xff);
                    const int c1 = (int) (int8_t) ((sa >> (sh + 16)) & 0xff);
                    const int c2 = (int) (int8_t) ((sb >> sh) & 0xff);
                    const int c3 = (int) (int8_t) ((sb >> (sh + 16)) & 0xff);
                    const int isum = c0 * d0 + c1 * d1 + c2 * d2 + c3 * d3;
                    const float d = __half2float(__ushort_as_half((unsigned short) dv[t][r])) * dx;
                    acc[r] = fmaf(d, (float) isum, acc[r]);
                }
            }
        }
    }
}

template <int R, int UNR>
__device__ __forceinline__ void dot_q80(const RowPtrs (&rp)[R], int nb32, int lane, const ActSmem & A, float (&acc)[R]) {
    // unit = 16 int8 weights (half a Q8_0 block); ggml_vec_dot_q8_0_q8_0 (cpp/ggml/src/ggml-quants.c:5227)
    const int n_units = nb32 * 2;
    for (int u0 = lane; u0 < n_units; u0 += 32 * UNR) {
        uint4    w[UNR][R];
        uint32_t dv[UNR][R];
#pragma unroll
        for (int t = 0; t < UNR; t++) {
            const int u = u0 + 32 * t;
            if (u < n_units) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    w[t][r]  = ldg_stream_v4(rp[r].p0 + (size_t) u * 16);
                    dv[t][r] = ldg_u16(rp[r].p1 + (u >> 1) * 2);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < UNR; t++) {
            const int u = u0 + 32 * t;
            if (u < n_units) {
                const int4 a = *reinterpret_cast<const int4 *>(A.q + u * 16);
                const float dx = A.dx[u >> 1];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int isum = dot16(w[t][r], a, 0);
                    const float d = __half2float(__ushort_as_half((unsigned short) dv[t][r])) * dx;
                    acc[r] = fmaf(d, (float) isum, acc[r]);
                }
            }
        }
    }
}

__device__ __forceinline__ RowPtrs row_ptrs(int type, int kk, const uint8_t * p0, const uint8_t * p1, const uint8_t * p2,
                                            const uint8_t * p3, int row) {
    RowPtrs r;
    const size_t k = (size_t) kk;
    const size_t nb = k / 256;
    switch (type) {
        case T_Q4_K: r.p0 = p0 + row * nb * 128; r.p1 = p1 + row * nb * 12; r.p2 = p2 + row * nb * 4;  r.p3 = nullptr; break;
        case T_Q5_K: r.p0 = p0 + row * nb * 128; r.p1 = p1 + row * nb * 12; r.p2 = p2 + row * nb * 4;  r.p3 = p3 + row * nb * 32; break;
        case T_Q6_K: r.p0 = p0 + row * nb * 128; r.p1 = p1 + row * nb * 64; r.p2 = p2 + row * nb * 16; r.p3 = p3 + row * nb * 2; break;
        default:     r.p0 = p0 + row * k;        r.p1 = p1 + row * (k / 32) * 2; r.p2 = nullptr; r.p3 = nullptr; break;   // T_Q8_0
    }
    return r;
}

__device__ __forceinline__ float silu_f32(float x) { return x / (1.0f + expf(-x)); }   // ggml_silu_f32, ggml.c:2393

// ------------------------------------------------------------------------------------------------------------
// The fused quantized mat-vec: [RMSNorm] + activation quant (prologue) -> W.x over row pairs -> epilogue.
// Grid-stride over row PAIRS, one warp per pair per pass.
//   PAIR_ADJACENT: pair p = rows (2p, 2p+1) of the virtual concatenation seg[0] ++ seg[1] ++ seg[2]
//                  (the RoPE partner of row 2i is row 2i+1: NORM mode, cpp/ggml/src/ggml.c:14121-14135)
//   PAIR_ZIP     : pair p = (seg[0] row p, seg[1] row p)  (gate/up, cpp/src/llama.cpp:8875-8880)
// ------------------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(MV_THREADS, 2) k_matvec(const MatvecArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ double red_smem[MV_WARPS];
    const ActSmem A = act_smem_carve(smem_raw, a.k, a.act_q8_0);

    prologue_quantize(a.x, a.norm_w, a.eps, a.k, a.act_q8_0, A, reinterpret_cast<float *>(red_smem));
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * MV_WARPS + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * MV_WARPS;

    for (int pair = warp_global; pair < a.n_pairs; pair += n_warps) {
        int si0 = 0, r0 = 0, si1 = 0, r1 = 0;
        if (a.pair_mode == PAIR_ZIP) {
            si0 = 0; r0 = pair; si1 = 1; r1 = pair;
        } else {
            int row = 2 * pair;
            int s = 0;
            while (s < a.n_seg - 1 && row >= a.seg[s].n_rows) { row -= a.seg[s].n_rows; s++; }
            si0 = si1 = s; r0 = row; r1 = row + 1;
        }
        // select the segment with branches (dynamic indexing of the kernel-parameter struct would force a local copy)
        const QMat & m0 = si0 == 0 ? a.seg[0] : (si0 == 1 ? a.seg[1] : a.seg[2]);
        const QMat & m1 = si1 == 0 ? a.seg[0] : (si1 == 1 ? a.seg[1] : a.seg[2]);
        const int type = m0.type;
        const RowPtrs rp[2] = { row_ptrs(type, m0.k, m0.p0, m0.p1, m0.p2, m0.p3, r0), row_ptrs(type, m1.k, m1.p0, m1.p1, m1.p2, m1.p3, r1) };
        float acc[2] = {0.f, 0.f};
        switch (type) {
            case T_Q4_K: dot_q45k<2, 2, false>(rp, a.k / 256, lane, A, acc); break;
            case T_Q5_K: dot_q45k<2, 1, true>(rp, a.k / 256, lane, A, acc); break;
            case T_Q6_K: dot_q6k<2, 1>(rp, a.k / 256, lane, A, acc); break;
            default:     dot_q80<2, 2>(rp, a.k / 32, lane, A, acc); break;
        }
        const float v0 = warp_sum(acc[0]);
        const float v1 = warp_sum(acc[1]);
        if (lane == 0) {
            if (EPI == EPI_STORE) {
                a.out[2 * pair]     = v0;
                a.out[2 * pair + 1] = v1;
            } else if (EPI == EPI_RESID) {
                // ggml_add(cur, inpSA) / ggml_add(cur, ffn_inp): cpp/src/llama.cpp:8865, 8901
                a.out[2 * pair]     = __fadd_rn(v0, a.resid[2 * pair]);
                a.out[2 * pair + 1] = __fadd_rn(v1, a.resid[2 * pair + 1]);
            } else if (EPI == EPI_SILU) {
                // silu(gate) * up: cpp/src/llama.cpp:7960-8085 (LLM_FFN_SILU, LLM_FFN_PAR)
                a.out[pair] = __fmul_rn(silu_f32(v0), v1);
            } else {  // EPI_QKV
                const int row = 2 * pair;
                const int pos = a.st->pos;
                if (row < a.n_q + a.n_k) {
                    // RoPE NORM mode on the pair (x0, x1): cpp/ggml/src/ggml.c:14121-14135
                    const int i0 = row % a.head_dim;
                    const float2 cs = a.rope[(size_t) pos * (a.head_dim / 2) + (i0 >> 1)];
                    const float y0 = __fsub_rn(__fmul_rn(v0, cs.x), __fmul_rn(v1, cs.y));
                    const float y1 = __fadd_rn(__fmul_rn(v0, cs.y), __fmul_rn(v1, cs.x));
                    if (row < a.n_q) {
                        a.q_out[row]     = y0;
                        a.q_out[row + 1] = y1;
                    } else {
                        // K stored post-RoPE as f16 at slot `pos`: llm_build_kv_store, cpp/src/llama.cpp:7849-7853
                        __half2 * dst = reinterpret_cast<__half2 *>(a.k_cache + (size_t) pos * a.kv_dim + (row - a.n_q));
                        *dst = __halves2half2(__float2half_rn(y0), __float2half_rn(y1));
                    }
                } else {
                    __half2 * dst = reinterpret_cast<__half2 *>(a.v_cache + (size_t) pos * a.kv_dim + (row - a.n_q - a.n_k));
                    *dst = __halves2half2(__float2half_rn(v0), __float2half_rn(v1));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// stand-alone activation quantization kernels (operator-level tests; same device functions as the prologue)
// out layouts are the ggml block structs: block_q8_K {float d; int8 qs[256]; int16 bsums[16]} (292 B),
// block_q8_0 {half d; int8 qs[32]} (34 B)  (cpp/ggml/src/ggml-common.h:311-315, 186-190)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_quantize_export(const float * __restrict__ x, int k, int act_q8_0, uint8_t * __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ double red_smem[MV_WARPS];
    const ActSmem A = act_smem_carve(smem_raw, k, act_q8_0);
    prologue_quantize(x, nullptr, 0.f, k, act_q8_0, A, reinterpret_cast<float *>(red_smem));
    __syncthreads();
    if (!act_q8_0) {
        for (int b = 0; b < k / 256; b++) {
            uint8_t * o = out + (size_t) b * 292;
            if (threadIdx.x == 0) *reinterpret_cast<float *>(o) = A.dx[b];
            for (int i = threadIdx.x; i < 256; i += blockDim.x) o[4 + i] = (uint8_t) A.q[b * 256 + i];
            for (int i = threadIdx.x; i < 16; i += blockDim.x) {
                const int16_t s = A.bsums[b * 16 + i];
                o[260 + 2 * i] = (uint8_t) (s & 0xff);
                o[261 + 2 * i] = (uint8_t) ((s >> 8) & 0xff);
            }
        }
    } else {
        for (int b = threadIdx.x; b < k / 32; b += blockDim.x) {
            uint8_t * o = out + (size_t) b * 34;
            const __half hd = __float2half_rn(A.dx[b]);     // dx is already an exact f16 value
            *reinterpret_cast<__half *>(o) = hd;              // 34*b is even: 2-byte aligned
            for (int i = 0; i < 32; i++) o[2 + i] = (uint8_t) A.q[b * 32 + i];
        }
    }
}

// y = rms_norm(x) * w  (operator-level test of the prologue arithmetic, no quantization)
__global__ void k_rms_norm(const float * __restrict__ x, const float * __restrict__ w, int k, float eps, float * __restrict__ y) {
    __shared__ double red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    double s = 0.0;
    for (int i = tid; i < k; i += blockDim.x) s += (double) __fmul_rn(x[i], x[i]);
    s = warp_sum_d(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < nwarp; i++) tot += red[i];
    const float mean  = (float) (tot / (double) k);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    for (int i = tid; i < k; i += blockDim.x) {
        const float v = __fmul_rn(x[i], scale);
        y[i] = w ? __fmul_rn(v, w[i]) : v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// embedding row dequantization == ggml_compute_forward_get_rows_q -> dequantize_row_* (cpp/ggml/src/ggml.c:13186,
// cpp/ggml/src/ggml-quants.c:1609,2548,2756,2970). `rows` is the tensor in its ORIGINAL ggml block layout.
// The float ops are written un-fused (mul then sub) because the CPU build does not contract them (-std=c11).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float half_at(const uint8_t * p) {
    return __half2float(__ushort_as_half((unsigned short) (p[0] | (p[1] << 8))));
}
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t * q, int & sc, int & m) {
    if (j < 4) { sc = q[j] & 63; m = q[j + 4] & 63; }
    else       { sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}
__device__ __forceinline__ float dequant_elem(int type, const uint8_t * row, int i) {
    switch (type) {
        case T_F32: return reinterpret_cast<const float *>(row)[i];
        case T_F16: return __half2float(reinterpret_cast<const __half *>(row)[i]);
        case T_Q8_0: {
            const uint8_t * b = row + (size_t) (i / 32) * 34;
            return __fmul_rn((float) (int8_t) b[2 + (i & 31)], half_at(b));
        }
        case T_Q4_K: {
            const uint8_t * b = row + (size_t) (i / 256) * 144;
            const int e = i & 255, j64 = e >> 6, l = e & 31, hi = (e >> 5) & 1;
            int sc, m; k4_scale_min(2 * j64 + hi, b + 4, sc, m);
            const float d1 = __fmul_rn(half_at(b), (float) sc), m1 = __fmul_rn(half_at(b + 2), (float) m);
            const uint8_t qb = b[16 + 32 * j64 + l];
            const int q = hi ? (qb >> 4) : (qb & 0xF);
            return __fsub_rn(__fmul_rn(d1, (float) q), m1);
        }
        case T_Q5_K: {
            const uint8_t * b = row + (size_t) (i / 256) * 176;
            const int e = i & 255, j64 = e >> 6, l = e & 31, hi = (e >> 5) & 1;
            int sc, m; k4_scale_min(2 * j64 + hi, b + 4, sc, m);
            const float d1 = __fmul_rn(half_at(b), (float) sc), m1 = __fmul_rn(half_at(b + 2), (float) m);
            const uint8_t qb = b[48 + 32 * j64 + l];
            const int hb = (b[16 + l] >> (2 * j64 + hi)) & 1;
            const int q = (hi ? (qb >> 4) : (qb & 0xF)) + 16 * hb;
            return __fsub_rn(__fmul_rn(d1, (float) q), m1);
        }
        case T_Q6_K: {
            const uint8_t * b = row + (size_t) (i / 256) * 210;
            const int e = i & 255, n = e >> 7, g = (e >> 5) & 3, l = e & 31;
            const uint8_t qlb = b[64 * n + 32 * (g & 1) + l];
            const int lo = (g >> 1) ? (qlb >> 4) : (qlb & 0xF);
            const int h2 = (b[128 + 32 * n + l] >> (2 * g)) & 3;
            const int q = (int) (int8_t) (lo | (h2 << 4)) - 32;
            const int sc = (int) (int8_t) b[192 + 8 * n + 2 * g + (l >> 4)];
            // y = d * sc * q evaluated left to right (cpp/ggml/src/ggml-quants.c:2988-2991)
            return __fmul_rn(__fmul_rn(half_at(b + 208), (float) sc), (float) q);
        }
        default: return 0.f;
    }
}
__global__ void k_embed(int type, const uint8_t * __restrict__ rows, size_t row_bytes, int k,
                        const DecodeState * __restrict__ st, int token_override, float * __restrict__ out) {
    const int token = st ? st->token : token_override;
    const uint8_t * row = rows + (size_t) token * row_bytes;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) out[i] = dequant_elem(type, row, i);
}

// ------------------------------------------------------------------------------------------------------------
// decode attention, default (non-flash) route at batch 1 (cpp/src/llama.cpp:8248-8297):
//   kq = K(f16->f32) . q(f32)  [tinyBLAS F16xF32, cpp/ggml/src/ggml.c:12325-12341]
//   softmax(kq*scale + mask)   [cpp/ggml/src/ggml.c:13682-13778; causality from `pos`, no mask tensor]
//   kqv = V(f16->f32) . p      [same tinyBLAS route]
// Split-KV flash-decode, one launch: grid (n_head_kv, ATT_SPLITS), 256 threads. A CTA serves the `GQA` query
// heads that share its KV head (K and V rows are read once per group, cf. the broadcast in
// cpp/ggml/src/ggml.c:12207-12243) over 64-position tiles t = split, split + ATT_SPLITS, ...
// Latency design (HBM-bound, 8.4 MB per layer at ctx 2048 => every byte must be in flight at once):
//   * all K loads (4 x 16 B per thread: 4 lanes per key row) AND all V loads (16 x 4 B per thread) of a tile are
//     issued before any arithmetic, so a tile costs one DRAM round trip;
//   * scores: 32-dim partial dots per lane, 2 shuffles per head; softmax state per head kept in shared memory;
//   * P.V: thread = (dim pair, 16-position group), partial sums merged through shared memory;
//   * the last CTA of a KV head (atomic ticket) merges the splits — no separate combine launch.
// ------------------------------------------------------------------------------------------------------------
static constexpr int ATT_THREADS = 256;
static constexpr int ATT_TILE    = 64;
static constexpr int ATT_SPLITS  = 32;
static constexpr int ATT_MAX_GQA = 8;

struct AttnArgs {
    const float * q;          // [n_head][hd] post-RoPE
    const __half * k_cache;   // [n_ctx][kv_dim]
    const __half * v_cache;
    float * part_o;           // [n_head][ATT_SPLITS][hd]  un-normalised
    float * part_ml;          // [n_head][ATT_SPLITS][2]   (running max, sum)
    unsigned int * tickets;   // [n_head_kv] zero-initialised, self-resetting
    float * out;              // [n_head*hd]  (kqv_merged_cont)
    int n_head, n_head_kv, head_dim, kv_dim;
    float scale;
    const DecodeState * st;
    int n_kv_override;        // >0: use instead of st->pos+1 (operator-level test)
};

template <int GQA>
__global__ void __launch_bounds__(ATT_THREADS) k_attn(const AttnArgs a) {
    constexpr int HD = 128;                                   // every LLaMA/Mistral config on this path
    __shared__ __align__(16) float qs[GQA][HD];
    __shared__ __align__(16) float sc[GQA][ATT_TILE];         // scores, then probabilities
    __shared__ float run_m[GQA], run_l[GQA], resc[GQA];
    __shared__ __align__(16) float red_o[4][GQA][HD];         // P.V partials of the 4 position groups
    __shared__ float wsplit[GQA][ATT_SPLITS];
    __shared__ unsigned int s_ticket;

    const int g = blockIdx.x, split = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_kv = a.n_kv_override > 0 ? a.n_kv_override : a.st->pos + 1;
    const int round_q = a.st ? a.st->round_q : 0;
    const int n_tiles = (n_kv + ATT_TILE - 1) / ATT_TILE;

    for (int i = tid; i < GQA * HD; i += ATT_THREADS) {
        float v = a.q[(size_t) (g * GQA) * HD + i];
        if (round_q) v = __half2float(__float2half_rn(v));    // batch > 1: q is rounded to f16 (ggml.c:12345-12371)
        (&qs[0][0])[i] = v;
    }
    if (tid < GQA) { run_m[tid] = -INFINITY; run_l[tid] = 0.f; }

    // thread roles
    const int kp = tid >> 2, kq4 = tid & 3;                   // scores: key position in tile, 32-dim quarter
    const int dp = tid & 63, pg = tid >> 6;                   // P.V: dim pair, position group (16 positions)
    float o[GQA][2];
#pragma unroll
    for (int h = 0; h < GQA; h++) { o[h][0] = 0.f; o[h][1] = 0.f; }
    __syncthreads();

    for (int tile = split; tile < n_tiles; tile += ATT_SPLITS) {
        const int p0 = tile * ATT_TILE;
        // ---- issue every load of the tile
        uint4 kreg[4];
        const bool kvalid = p0 + kp < n_kv;
        if (kvalid) {
            const uint4 * kr = reinterpret_cast<const uint4 *>(a.k_cache + (size_t) (p0 + kp) * a.kv_dim + g * HD + kq4 * 32);
#pragma unroll
            for (int i = 0; i < 4; i++) kreg[i] = ldg_stream_v4(kr + i);
        }
        uint32_t vreg[16];
        const __half * vbase = a.v_cache + (size_t) (p0 + pg * 16) * a.kv_dim + g * HD + 2 * dp;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            vreg[i] = 0u;
            if (p0 + pg * 16 + i < n_kv) vreg[i] = __ldg(reinterpret_cast<const uint32_t *>(vbase + (size_t) i * a.kv_dim));
        }
        // ---- scores
        float acc[GQA];
#pragma unroll
        for (int h = 0; h < GQA; h++) acc[h] = 0.f;
        if (kvalid) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const __half2 * k2 = reinterpret_cast<const __half2 *>(&kreg[i]);
                float kf[8];
#pragma unroll
                for (int e = 0; e < 4; e++) { const float2 f = __half22float2(k2[e]); kf[2 * e] = f.x; kf[2 * e + 1] = f.y; }
#pragma unroll
                for (int h = 0; h < GQA; h++) {
                    const float4 qa = *reinterpret_cast<const float4 *>(&qs[h][kq4 * 32 + i * 8]);
                    const float4 qb = *reinterpret_cast<const float4 *>(&qs[h][kq4 * 32 + i * 8 + 4]);
                    acc[h] = fmaf(kf[0], qa.x, acc[h]); acc[h] = fmaf(kf[1], qa.y, acc[h]);
                    acc[h] = fmaf(kf[2], qa.z, acc[h]); acc[h] = fmaf(kf[3], qa.w, acc[h]);
                    acc[h] = fmaf(kf[4], qb.x, acc[h]); acc[h] = fmaf(kf[5], qb.y, acc[h]);
                    acc[h] = fmaf(kf[6], qb.z, acc[h]); acc[h] = fmaf(kf[7], qb.w, acc[h]);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < GQA; h++) {
            float s = acc[h];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (kq4 == 0) sc[h][kp] = kvalid ? __fmul_rn(s, a.scale) : -INFINITY;
        }
        __syncthreads();
        // ---- online softmax state: warp h owns head h (64 scores = 2 per lane)
        if (warp < GQA) {
            const float s0 = sc[warp][lane], s1 = sc[warp][lane + 32];
            const float mt = warp_max(fmaxf(s0, s1));
            const float m_old = run_m[warp];
            const float m_new = fmaxf(m_old, mt);
            const float p0v = expf(s0 - m_new), p1v = expf(s1 - m_new);     // exp(-inf) = 0 for masked slots
            const float lt = warp_sum(p0v + p1v);
            sc[warp][lane] = p0v; sc[warp][lane + 32] = p1v;
            if (lane == 0) {
                const float r = m_old == -INFINITY ? 0.f : expf(m_old - m_new);
                resc[warp] = r;
                run_l[warp] = run_l[warp] * r + lt;
                run_m[warp] = m_new;
            }
        }
        __syncthreads();
        // ---- P.V for this thread's 16 positions x 2 dims
#pragma unroll
        for (int h = 0; h < GQA; h++) {
            const float r = resc[h];
            float o0 = o[h][0] * r, o1 = o[h][1] * r;
#pragma unroll
            for (int i4 = 0; i4 < 4; i4++) {
                const float4 pv = *reinterpret_cast<const float4 *>(&sc[h][pg * 16 + i4 * 4]);
                const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 vf = __half22float2(*reinterpret_cast<const __half2 *>(&vreg[i4 * 4 + e]));
                    o0 = fmaf(pp[e], vf.x, o0);
                    o1 = fmaf(pp[e], vf.y, o1);
                }
            }
            o[h][0] = o0; o[h][1] = o1;
        }
        __syncthreads();     // sc / resc are rewritten by the next tile
    }

    // ---- merge the 4 position groups, publish this split's partial
#pragma unroll
    for (int h = 0; h < GQA; h++) { red_o[pg][h][2 * dp] = o[h][0]; red_o[pg][h][2 * dp + 1] = o[h][1]; }
    __syncthreads();
    for (int i = tid; i < GQA * HD; i += ATT_THREADS) {
        const int h = i / HD, d = i % HD;
        const float v = red_o[0][h][d] + red_o[1][h][d] + red_o[2][h][d] + red_o[3][h][d];
        a.part_o[((size_t) (g * GQA + h) * ATT_SPLITS + split) * HD + d] = v;
    }
    if (tid < GQA) {
        const size_t oi = ((size_t) (g * GQA + tid) * ATT_SPLITS + split) * 2;
        a.part_ml[oi] = run_m[tid]; a.part_ml[oi + 1] = run_l[tid];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&a.tickets[g], 1u);
    __syncthreads();
    if (s_ticket != ATT_SPLITS - 1) return;
    // ---- last CTA of this KV head: combine the splits
    __threadfence();
    if (tid == 0) a.tickets[g] = 0u;                          // self-reset for the next launch
    for (int i = tid; i < GQA * ATT_SPLITS; i += ATT_THREADS) {
        const int h = i / ATT_SPLITS, s = i % ATT_SPLITS;
        (&wsplit[0][0])[i] = __ldcg(&a.part_ml[((size_t) (g * GQA + h) * ATT_SPLITS + s) * 2]);   // m_s for now
    }
    __syncthreads();
    if (tid < GQA) {
        float M = -INFINITY;
        for (int s = 0; s < ATT_SPLITS; s++) M = fmaxf(M, wsplit[tid][s]);
        float L = 0.f;
        for (int s = 0; s < ATT_SPLITS; s++) {
            const float ms = wsplit[tid][s];
            const float w = ms == -INFINITY ? 0.f : expf(ms - M);
            L += w * __ldcg(&a.part_ml[((size_t) (g * GQA + tid) * ATT_SPLITS + s) * 2 + 1]);
            wsplit[tid][s] = w;
        }
        run_l[tid] = 1.0f / L;
    }
    __syncthreads();
    for (int i = tid; i < GQA * HD; i += ATT_THREADS) {
        const int h = i / HD, d = i % HD;
        float v = 0.f;
        for (int s = 0; s < ATT_SPLITS; s++) {
            const float w = wsplit[h][s];
            if (w != 0.f) v = fmaf(w, __ldcg(&a.part_o[((size_t) (g * GQA + h) * ATT_SPLITS + s) * HD + d]), v);
        }
        a.out[(size_t) (g * GQA + h) * HD + d] = v * run_l[h];   // kqv_merged_cont layout: [n_head*hd]
    }
}

// ------------------------------------------------------------------------------------------------------------
// greedy sampling on device: arg-max over the logits with the lowest index winning ties (a strictly-greater scan
// like sample_top_token, cpp/bridge.cpp:962-981), two tiny launches: per-CTA reduction + one 64-bit atomicMax on
// a packed (orderable value, ~index) key, then a 1-thread finish that hands the token to the next step.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);            // monotone map float -> uint
    return ((unsigned long long) u << 32) | (unsigned long long) (0xffffffffu - (uint32_t) idx);
}
__global__ void k_argmax_partial(const float * __restrict__ logits, int n, unsigned long long * key) {
    __shared__ unsigned long long sk[32];
    unsigned long long best = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long kk = argmax_key(logits[i], i);
        best = kk > best ? kk : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, o);
        best = ok > best ? ok : best;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sk[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int) (blockDim.x >> 5); w++) best = sk[w] > best ? sk[w] : best;
        atomicMax(key, best);
    }
}
// advance != 0: greedy loop (token -> next step's DecodeState, out_tokens[step]); advance == 0: out_tokens[0] only
__global__ void k_argmax_finish(unsigned long long * key, DecodeState * st, int32_t * out_tokens, int advance) {
    if (threadIdx.x != 0) return;
    const int idx = (int) (0xffffffffu - (uint32_t) (*key & 0xffffffffull));
    *key = 0ull;
    if (advance) {
        if (out_tokens) out_tokens[st->step] = idx;
        st->token = idx; st->pos += 1; st->step += 1;
    } else {
        out_tokens[0] = idx;
    }
}

__global__ void k_set_state(DecodeState * st, const DecodeState v) { *st = v; }

// advance (pos, step) on stages that do not sample (pipeline stages other than the last)
__global__ void k_advance(DecodeState * st) {
    if (threadIdx.x == 0) { st->pos += 1; st->step += 1; }
}

// RoPE on a [n_heads][head_dim] f32 buffer in place (operator-level test; engine fuses it into EPI_QKV)
__global__ void k_rope(float * x, int n_heads, int head_dim, const float2 * __restrict__ rope_row) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half_dim = head_dim / 2;
    if (idx >= n_heads * half_dim) return;
    const int hh = idx / half_dim, i = idx % half_dim;
    const float2 cs = rope_row[i];
    float * p = x + (size_t) hh * head_dim + 2 * i;
    const float x0 = p[0], x1 = p[1];
    p[0] = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
    p[1] = __fadd_rn(__fmul_rn(x0, cs.y), __fmul_rn(x1, cs.x));
}

}  // namespace b200
