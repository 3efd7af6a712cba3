// prefill.cuh — prompt batches (llama_decode with n_tokens > 1): the same arithmetic as the per-token kernels, with every
// weight tile fetched ONCE for a chunk of 64 tokens instead of once per token.
//
// Why not tensor cores (ggml-cuda's mul_mat_q): the reference's CPU result keeps EIGHT separate fp32
// accumulators per output element — AVX2 lane m owns bytes 4m..4m+3 of every 32-byte group, and acc[m] = fma(d_block,
// (float) isum[m], acc[m]) runs block after block before the final hsum (cpp/ggml/src/ggml-quants.c:5361-5382, 6914-6977).
// Bit-exact parity therefore needs the integer sum of every 4-byte group of every (row, token) SEPARATELY, scaled and
// chained in fp32 in block order. An int8 MMA (mma.sync k = 16/32, tcgen05 kind::i8 k = 32) sums at least 16 products per
// output; the 4-byte granularity is exactly dp4a's. A tensor-core prefill is possible only by giving up bit-exactness,
// and with re-quantized activations between mat-muls that costs ~1e-2 on the logits, outside the 1e-3 bar.
// So the batch path is a dp4a kernel whose gain is data reuse: weights are read from HBM/L2 once per 64 tokens and the
// per-token work (integers + chains) runs out of shared memory — it is issue-bound, not HBM-bound.
//
//   k_embed_batch   token_embd rows of T tokens -> X[T][n_embd]
//   k_quant_batch   per token: [RMSNorm * w] + Q8_K / Q8_0 quantization (prologue_quantize, the decode path's own code) into
//                   RECORDS in global memory, laid out [chunk of 64 tokens][256-weight block][token][record] so that the
//                   blocks of one K step of a whole chunk are one contiguous bulk copy
//   k_matmul_batch  CTA = (32-row unit, 64-token chunk), 16 warps x 4 tokens; per K step one TMA bulk copy of the weight
//                   tile(s) + one of the chunk's activation records into a 2-stage shared-memory ring (full / empty
//                   mbarriers); every warp computes tile_ints for its 4 tokens and advances 4 x 12 fp32 chains in
//                   registers; epilogues as in the per-token kernel (RoPE + KV-cache rows, +residual, SiLU*up, store)
// Attention of a batch runs the per-token attention kernels with the token index in blockIdx.z (kernels.cuh).
#pragma once
#include "kernels.cuh"

namespace b200 {

static constexpr int PB_CHUNK = 64;            // tokens per CTA of the batched mat-mul
static constexpr int PB_WARPS = 16;
static constexpr int PB_TOK = PB_CHUNK / PB_WARPS;   // tokens per warp (chains in registers: 4 x 12)
static constexpr int PB_STAGES = 2;

// activation record of one token for one 256-weight block: q[256] | dx | bp[8] | as[64] (Q8_K; `as` only when a Q6_K
// matrix consumes the vector) or q[256] | dx[8] (Q8_0: eight 32-weight blocks)
__host__ __device__ __forceinline__ int pb_record_bytes(int act_q8_0, int with_as) {
    return act_q8_0 ? 256 + 32 : (with_as ? 256 + 16 + 32 + 256 : 256 + 16 + 32);   // dx padded to 16 bytes: bp / as stay 16-byte aligned
}

__global__ void k_embed_batch(int type, const uint8_t * __restrict__ rows, size_t row_bytes, int k, const int32_t * __restrict__ tokens,
                              float * __restrict__ X) {
    const uint8_t * row = rows + (size_t) tokens[blockIdx.y] * row_bytes;
    float * out = X + (size_t) blockIdx.y * k;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) out[i] = dequant_elem(type, row, i);
}

struct QuantBatchArgs {
    const float * X;          // [T][k]
    int k, T;
    const float * norm_w;     // nullptr: no norm
    float eps;
    double inv_k;
    int act_q8_0, with_as;
    uint8_t * rec;            // [ceil(T/64)][k/256][64][record]
    int layout;               // 0: records above (k_matmul_batch), 1: k_mma_batch blocks, 2: k_umma_batch blocks
};

// one CTA per token: the decode path's prologue (RMSNorm + quantization into shared memory), then the image is written out
// as records
__global__ void __launch_bounds__(512) k_quant_batch(const QuantBatchArgs a) {
    extern __shared__ __align__(16) uint8_t qb_smem[];
    __shared__ double red_smem[MV_MAX_WARPS];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const ActSmem A = act_smem_carve(qb_smem, a.k, a.act_q8_0);
    const bool norm = a.norm_w != nullptr;
    float ww[PRO_U][8] = {};
    if (norm) {
#pragma unroll
        for (int u = 0; u < PRO_U; u++) {
            const int b = warp + u * W;
            if (b < a.k / 256) ldg8(a.norm_w + b * 256 + lane * 8, ww[u]);
        }
    }
    prologue_quantize<false, true>(a.X + (size_t) t * a.k, norm, a.eps, a.k, a.inv_k, a.act_q8_0, A, red_smem, ww, []() {}, W);
    __syncthreads();
    const int n256 = a.k / 256, rb = pb_record_bytes(a.act_q8_0, a.with_as);
    const int chunk = t / PB_CHUNK, j = t % PB_CHUNK;
    for (int b = warp; b < n256; b += W) {
        uint8_t * r = a.rec + (((size_t) chunk * n256 + b) * PB_CHUNK + j) * rb;
        *reinterpret_cast<uint2 *>(r + lane * 8) = *reinterpret_cast<const uint2 *>(A.q + (size_t) b * 256 + lane * 8);
        if (a.act_q8_0) {
            if (lane < 8) *reinterpret_cast<float *>(r + 256 + lane * 4) = A.dx[b * 8 + lane];
        } else {
            if (lane == 0) *reinterpret_cast<float *>(r + 256) = A.dx[b];
            if (lane < 8) *reinterpret_cast<int *>(r + 272 + lane * 4) = A.bp[(size_t) b * 8 + lane];
            if (a.with_as) *reinterpret_cast<int2 *>(r + 304 + lane * 8) = *reinterpret_cast<const int2 *>(A.as + (size_t) b * 64 + lane * 2);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// attention of a prompt batch: k_attn_scores (kernels.cuh, token in blockIdx.z) writes the raw score rows, then
//   k_attn_softmax_rows  one warp per (token, head) row, in place: k_attn_softmax's arithmetic (max, ggml_v_expf, per-16
//                        _mm512_reduce_add_ps sums accumulated in double, * (float)(1/sum)), each row normalised ONCE
//   k_attn_pv_batch      CTA = (KV head, 16 output dims, 32 / GQA tokens): every V chunk is staged once for the 32 (token,
//                        head) rows of the CTA; thread (chain c, dim) advances the 32 rows' tinyBLAS chains t = c, c+16, ...
//                        A token's probabilities end at its own padded length; beyond it p = 0 and fma(v, 0, acc) == acc,
//                        which is what the reference's masked (-inf -> 0) columns contribute (cpp/src/llama.cpp:14132-14200).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_attn_softmax_rows(const AttnArgs a, int nz) {
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nz * a.n_head) return;
    const int z = row / a.n_head, h = row - z * a.n_head;
    const int n_pad = (a.n_kv_override + z + 31) / 32 * 32, n16 = n_pad / 16;
    float * S = a.S + (size_t) z * a.zs + (size_t) h * a.s_stride;
    float mx = -INFINITY;
    for (int gi = lane; gi < n16; gi += 32) {
        const float4 * r4 = reinterpret_cast<const float4 *>(S + 16 * gi);
#pragma unroll
        for (int q = 0; q < 4; q++) { const float4 v = r4[q]; mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w)); }
    }
    mx = warp_max(mx);
    double part = 0.0;
    for (int gi = lane; gi < n16; gi += 32) {
        float4 * r4 = reinterpret_cast<float4 *>(S + 16 * gi);
        float4 t[2];
#pragma unroll 1
        for (int pq = 0; pq < 2; pq++) {                       // float4 pairs {0, 2} and {1, 3}: a[8+i] + a[i]
            float4 lo = r4[pq], hi = r4[pq + 2];
            lo.x = v_expf(__fsub_rn(lo.x, mx)); lo.y = v_expf(__fsub_rn(lo.y, mx)); lo.z = v_expf(__fsub_rn(lo.z, mx)); lo.w = v_expf(__fsub_rn(lo.w, mx));
            hi.x = v_expf(__fsub_rn(hi.x, mx)); hi.y = v_expf(__fsub_rn(hi.y, mx)); hi.z = v_expf(__fsub_rn(hi.z, mx)); hi.w = v_expf(__fsub_rn(hi.w, mx));
            r4[pq] = lo; r4[pq + 2] = hi;
            const float4 tt = make_float4(__fadd_rn(hi.x, lo.x), __fadd_rn(hi.y, lo.y), __fadd_rn(hi.z, lo.z), __fadd_rn(hi.w, lo.w));
            if (pq) t[1] = tt; else t[0] = tt;
        }
        const float u0 = __fadd_rn(t[1].x, t[0].x), u1 = __fadd_rn(t[1].y, t[0].y), u2 = __fadd_rn(t[1].z, t[0].z), u3 = __fadd_rn(t[1].w, t[0].w);
        part += (double) __fadd_rn(__fadd_rn(u0, u2), __fadd_rn(u1, u3));
    }
    const double sum = warp_sum_d(part);
    const float inv = (float) (1.0 / sum);
    for (int gi = lane; gi < n16; gi += 32) {
        float4 * r4 = reinterpret_cast<float4 *>(S + 16 * gi);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float4 v = r4[q];
            v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv); v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
            r4[q] = v;
        }
    }
}

// scores of a prompt batch: CTA = (KV head, 64 keys, 8 tokens). The K rows of the tile are loaded ONCE into registers (4 lanes per
// key row, as in k_attn_scores) and every (token, head) of the CTA runs ggml_vec_dot_f16 against them: q rounded to f16, 4
// accumulators x 16 lanes over i in {0, 64}, (0+2)+(1+3), _mm512_reduce_add_ps (cpp/ggml/src/ggml.c:12345-12371) — the batch > 1
// branch of k_attn_scores, token by token. A token's row ends at its own padded length (-inf from its n_kv to there).
static constexpr int SCB_TQ = 8;
template <int GQA>
__global__ void __launch_bounds__(ATT_THREADS) k_attn_scores_batch(const AttnArgs a, int nz) {
    constexpr int HD = 128;
    __shared__ __align__(16) float qs[SCB_TQ][GQA][HD];
    const int g = blockIdx.x, tile = blockIdx.y, z0 = blockIdx.z * SCB_TQ, tid = threadIdx.x;
    const int ntok = min(SCB_TQ, nz - z0);
    const int n_kv_last = a.n_kv_override + z0 + ntok - 1;
    if (tile * ATT_TILE >= (n_kv_last + 31) / 32 * 32) return;
    const int t = tile * ATT_TILE + (tid >> 2), c4 = tid & 3;
    uint2 kv[8];
#pragma unroll
    for (int s = 0; s < 8; s++) kv[s] = make_uint2(0u, 0u);
    if (t < n_kv_last) {
        const uint2 * kr = reinterpret_cast<const uint2 *>(a.k_cache + (size_t) t * a.kv_dim + g * HD + 4 * c4);
#pragma unroll
        for (int s = 0; s < 8; s++) kv[s] = kr[s * 4];                        // 4 halfs at element 16s + 4c4
    }
    for (int i = tid; i < ntok * GQA * HD; i += ATT_THREADS) {
        const int tok = i / (GQA * HD), rem = i - tok * (GQA * HD);
        const float v = a.q[(size_t) (z0 + tok) * a.zq + (size_t) (g * GQA) * HD + rem];
        (&qs[0][0][0])[i] = __half2float(__float2half_rn(v));                 // src1 converted to the vec_dot_type F16
    }
    float kf[8][4];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[s].x));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[s].y));
        kf[s][0] = f0.x; kf[s][1] = f0.y; kf[s][2] = f1.x; kf[s][3] = f1.y;
    }
    __syncthreads();
#pragma unroll 1
    for (int tok = 0; tok < ntok; tok++) {
        const int n_kv = a.n_kv_override + z0 + tok, n_pad = (n_kv + 31) / 32 * 32;
        if (t >= n_pad) continue;                                  // n_pad % 32 == 0 and a warp holds 8 consecutive keys: warp-uniform
        float * Srow = a.S + (size_t) (z0 + tok) * a.zs + (size_t) (g * GQA) * a.s_stride + t;
#pragma unroll 1
        for (int h = 0; h < GQA; h++) {
            float aj[4][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 q0 = *reinterpret_cast<const float4 *>(&qs[tok][h][16 * j + 4 * c4]);
                const float4 q1 = *reinterpret_cast<const float4 *>(&qs[tok][h][64 + 16 * j + 4 * c4]);
                aj[j][0] = __fmaf_rn(kf[4 + j][0], q1.x, __fmul_rn(kf[j][0], q0.x));
                aj[j][1] = __fmaf_rn(kf[4 + j][1], q1.y, __fmul_rn(kf[j][1], q0.y));
                aj[j][2] = __fmaf_rn(kf[4 + j][2], q1.z, __fmul_rn(kf[j][2], q0.z));
                aj[j][3] = __fmaf_rn(kf[4 + j][3], q1.w, __fmul_rn(kf[j][3], q0.w));
            }
            float ch[4], t3[4], t6[4];
#pragma unroll
            for (int e = 0; e < 4; e++) ch[e] = __fadd_rn(__fadd_rn(aj[0][e], aj[2][e]), __fadd_rn(aj[1][e], aj[3][e]));
#pragma unroll
            for (int e = 0; e < 4; e++) t3[e] = __fadd_rn(__shfl_xor_sync(0xffffffffu, ch[e], 2), ch[e]);
#pragma unroll
            for (int e = 0; e < 4; e++) t6[e] = __fadd_rn(__shfl_xor_sync(0xffffffffu, t3[e], 1), t3[e]);
            const float res = __fadd_rn(__fadd_rn(t6[0], t6[2]), __fadd_rn(t6[1], t6[3]));
            if (c4 == 0) Srow[(size_t) h * a.s_stride] = t < n_kv ? __fmul_rn(res, a.scale) : -INFINITY;
        }
    }
}

static constexpr int PVB_ROWS = 32;            // (token, head) rows per CTA: 32 / GQA tokens x GQA heads
static constexpr int PVB_CHUNK = 256;          // positions staged at a time
static constexpr int PVB_DIMS = 16;
__host__ __device__ constexpr size_t pvb_smem_bytes() {
    return (size_t) PVB_ROWS * PVB_CHUNK * 4 + (size_t) PVB_CHUNK * PVB_DIMS * 2 > (size_t) PVB_ROWS * 16 * (PVB_DIMS + 1) * 4
               ? (size_t) PVB_ROWS * PVB_CHUNK * 4 + (size_t) PVB_CHUNK * PVB_DIMS * 2 : (size_t) PVB_ROWS * 16 * (PVB_DIMS + 1) * 4;
}
template <int GQA>
__global__ void __launch_bounds__(256) k_attn_pv_batch(const AttnArgs a, int nz) {
    constexpr int HD = 128, TQ = PVB_ROWS / GQA;
    extern __shared__ __align__(16) uint8_t pvb_dyn[];
    float * ps = reinterpret_cast<float *>(pvb_dyn);                                        // [32 rows][PVB_CHUNK]
    __half (*vs)[PVB_DIMS] = reinterpret_cast<__half (*)[PVB_DIMS]>(pvb_dyn + (size_t) PVB_ROWS * PVB_CHUNK * 4);
    const int g = blockIdx.x, slice = blockIdx.y, z0 = blockIdx.z * TQ, tid = threadIdx.x;
    const int c = tid / PVB_DIMS, dl = tid % PVB_DIMS;
    const int ntok = min(TQ, nz - z0);
    const int n_kv_last = a.n_kv_override + z0 + ntok - 1;     // cells attended by the CTA's last token
    const int n_max = (n_kv_last + 31) / 32 * 32;
    const __half * vbase = a.v_cache + g * HD + slice * PVB_DIMS;
    float acc[PVB_ROWS];
#pragma unroll
    for (int r = 0; r < PVB_ROWS; r++) acc[r] = 0.f;
    for (int t0 = 0; t0 < n_max; t0 += PVB_CHUNK) {
        const int len = min(PVB_CHUNK, n_max - t0);
        if (t0) __syncthreads();
        for (int i = tid; i < len * 2; i += 256) {             // V rows: 2 x 16 bytes; rows beyond the last cell are zero
            const int t = t0 + (i >> 1);
            if (t < n_kv_last) cp_async16(&vs[i >> 1][(i & 1) * 8], vbase + (size_t) t * a.kv_dim + (i & 1) * 8);
            else *reinterpret_cast<uint4 *>(&vs[i >> 1][(i & 1) * 8]) = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int i = tid; i < PVB_ROWS * (len / 4); i += 256) {
            const int r = i / (len / 4), j = i - r * (len / 4), tok = r / GQA, h = r - tok * GQA, t = t0 + 4 * j;
            const int n_pad_z = (a.n_kv_override + z0 + tok + 31) / 32 * 32;
            float * dst = ps + (size_t) r * PVB_CHUNK + 4 * j;
            if (tok < ntok && t < n_pad_z) cp_async16(dst, a.S + (size_t) (z0 + tok) * a.zs + (size_t) (g * GQA + h) * a.s_stride + t);
            else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        const int steps = len / 16;
#pragma unroll 2
        for (int s = 0; s < steps; s++) {
            const int tt = 16 * s + c;
            const float v = __half2float(vs[tt][dl]);
#pragma unroll
            for (int r = 0; r < PVB_ROWS; r++) acc[r] = __fmaf_rn(v, ps[(size_t) r * PVB_CHUNK + tt], acc[r]);
        }
    }
    __syncthreads();
    float (*red)[16][PVB_DIMS + 1] = reinterpret_cast<float (*)[16][PVB_DIMS + 1]>(pvb_dyn);
#pragma unroll
    for (int r = 0; r < PVB_ROWS; r++) red[r][c][dl] = acc[r];
    __syncthreads();
    for (int i = tid; i < PVB_ROWS * PVB_DIMS; i += 256) {
        const int r = i / PVB_DIMS, dd = i % PVB_DIMS, tok = r / GQA, h = r - tok * GQA;
        if (tok >= ntok) continue;
        float t3[8], t6[4];                                   // _mm512_reduce_add_ps over the 16 chains
#pragma unroll
        for (int j = 0; j < 8; j++) t3[j] = __fadd_rn(red[r][8 + j][dd], red[r][j][dd]);
#pragma unroll
        for (int j = 0; j < 4; j++) t6[j] = __fadd_rn(t3[4 + j], t3[j]);
        a.out[(size_t) (z0 + tok) * a.zq + (size_t) (g * GQA + h) * HD + slice * PVB_DIMS + dd] =
            __fadd_rn(__fadd_rn(t6[0], t6[2]), __fadd_rn(t6[1], t6[3]));
    }
}

struct MatmulBatchArgs {
    TMat seg[3];
    int n_seg, n_units, k, tiles_unit;
    int act_q8_0, with_as;
    const uint8_t * rec;      // activation records (k_quant_batch)
    int T;
    int epi;
    float * out; int out_stride;          // [T][out_stride]
    const float * resid; int resid_stride;
    // EPI_QKV
    float * q_out; int q_stride;
    __half * k_cache; __half * v_cache;
    int n_q, n_k, head_dim, kv_dim;
    const float2 * rope;
    int pos0;                 // position (== KV cell) of token 0 of the batch
    // k_mma_batch (prefill_mma.cuh): shared-memory layout chosen on the host
    uint32_t mb_a_bytes, mb_raw_stride, mb_stage_bytes, mb_rec_copy;
    int mb_stages;
};

// resolve a unit against the segments (same rule as describe_unit)
__device__ __forceinline__ UnitDesc pb_describe_unit(const MatmulBatchArgs & a, int unit) {
    int si = 0, u = unit, row_base = 0;
    if (a.n_seg > 1 && u >= a.seg[0].n_units) { u -= a.seg[0].n_units; row_base += a.seg[0].n_rows; si = 1;
        if (a.n_seg > 2 && u >= a.seg[1].n_units) { u -= a.seg[1].n_units; row_base += a.seg[1].n_rows; si = 2; } }
    UnitDesc d;
    d.type  = si == 0 ? a.seg[0].type : (si == 1 ? a.seg[1].type : a.seg[2].type);
    d.bytes = (uint32_t) (si == 0 ? a.seg[0].tile_bytes : (si == 1 ? a.seg[1].tile_bytes : a.seg[2].tile_bytes));
    const uint8_t * base = si == 0 ? a.seg[0].p0 : (si == 1 ? a.seg[1].p0 : a.seg[2].p0);
    d.tiles = base + (size_t) u * a.tiles_unit * d.bytes;
    d.row0  = row_base + u * 32;
    return d;
}

// the row's epilogue for token t (matvec_epilogue with per-token buffers)
__device__ __forceinline__ void pb_epilogue(const MatmulBatchArgs & a, float val, int row, int lane, int t) {
    const float oth = __shfl_xor_sync(0xffffffffu, val, 1);    // partner row (2i <-> 2i+1)
    if (t >= a.T) return;
    if (a.epi == EPI_STORE) {
        a.out[(size_t) t * a.out_stride + row] = val;
    } else if (a.epi == EPI_RESID) {
        a.out[(size_t) t * a.out_stride + row] = __fadd_rn(val, a.resid[(size_t) t * a.resid_stride + row]);
    } else if (a.epi == EPI_SILU) {
        if ((lane & 1) == 0) a.out[(size_t) t * a.out_stride + (row >> 1)] = __fmul_rn(silu_exact(val), oth);
    } else {  // EPI_QKV
        const int pos = a.pos0 + t;
        const float v0 = (lane & 1) ? oth : val, v1 = (lane & 1) ? val : oth;
        if (row < a.n_q + a.n_k) {
            const float2 cs = a.rope[(size_t) pos * (a.head_dim / 2) + (((row & ~1) & (a.head_dim - 1)) >> 1)];
            const float y = (lane & 1) ? __fadd_rn(__fmul_rn(v0, cs.y), __fmul_rn(v1, cs.x))
                                       : __fsub_rn(__fmul_rn(v0, cs.x), __fmul_rn(v1, cs.y));
            if (row < a.n_q) a.q_out[(size_t) t * a.q_stride + row] = y;
            else a.k_cache[(size_t) pos * a.kv_dim + (row - a.n_q)] = __float2half_rn(y);
        } else {
            a.v_cache[(size_t) pos * a.kv_dim + (row - a.n_q - a.n_k)] = __float2half_rn(val);
        }
    }
}

// shared memory of a stage: weight tiles of the step (8 consecutive 32-weight tiles for Q8_0, one tile otherwise) | records
__host__ __device__ __forceinline__ int pb_step_tiles(int act_q8_0) { return act_q8_0 ? 8 : 1; }
__host__ __device__ __forceinline__ size_t pb_stage_bytes(int max_tile_bytes, int act_q8_0, int with_as) {
    const size_t w = ((size_t) max_tile_bytes * pb_step_tiles(act_q8_0) + 127) / 128 * 128;
    return w + (size_t) PB_CHUNK * pb_record_bytes(act_q8_0, with_as);
}

__global__ void __launch_bounds__(PB_WARPS * 32, 1) k_matmul_batch(const __grid_constant__ MatmulBatchArgs a, int max_tile_bytes) {
    extern __shared__ __align__(128) uint8_t pb_smem[];
    __shared__ __align__(8) uint64_t bars[2 * PB_STAGES];      // full[stage], empty[stage]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, unit = blockIdx.y;
    const int rb = pb_record_bytes(a.act_q8_0, a.with_as), st_tiles = pb_step_tiles(a.act_q8_0);
    const size_t w_bytes = ((size_t) max_tile_bytes * st_tiles + 127) / 128 * 128;
    const size_t stage_bytes = w_bytes + (size_t) PB_CHUNK * rb;
    const int n_steps = a.tiles_unit / st_tiles;                // K steps: 256 weights each
    const UnitDesc ud = pb_describe_unit(a, unit);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[PB_STAGES]);
    if (tid == 0) {
        for (int s = 0; s < PB_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, PB_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint8_t * rec_chunk = a.rec + (size_t) chunk * (a.k / 256) * PB_CHUNK * rb;
    auto issue = [&](int step) {                               // thread 0: weights + the chunk's records of K step `step`
        const int s = step % PB_STAGES;
        const uint32_t dst = smem_u32(pb_smem + (size_t) s * stage_bytes);
        const uint32_t wb = ud.bytes * (uint32_t) st_tiles, ab = (uint32_t) (PB_CHUNK * rb);
        mbar_expect_tx(full0 + 8 * s, wb + ab);
        bulk_g2s(dst, ud.tiles + (size_t) step * wb, wb, full0 + 8 * s);
        bulk_g2s(dst + (uint32_t) w_bytes, rec_chunk + (size_t) step * PB_CHUNK * rb, ab, full0 + 8 * s);
    };
    if (tid == 0) { for (int s = 0; s < PB_STAGES && s < n_steps; s++) issue(s); }

    auto body = [&](auto tag) {
        constexpr int TYPE = decltype(tag)::value;
        float acc[PB_TOK][12];
#pragma unroll
        for (int j = 0; j < PB_TOK; j++) {
#pragma unroll
            for (int c = 0; c < 12; c++) acc[j][c] = 0.f;
        }
        for (int step = 0; step < n_steps; step++) {
            const int s = step % PB_STAGES;
            const uint32_t par = (uint32_t) ((step / PB_STAGES) & 1);
            mbar_wait(full0 + 8 * s, par);
            const uint8_t * stage = pb_smem + (size_t) s * stage_bytes;
            if (TYPE == T_Q8_0) {
                // eight 32-weight blocks per step; a block's weights are read once for the warp's four tokens. The lane sums
                // come out of dp4a already as floats: the accumulator input 0x4B400000 is the bit pattern of 1.5 * 2^23, and
                // adding an integer |v| < 2^22 to it gives the bit pattern of 12582912.0 + v — one exact FADD instead of an
                // I2F on the quarter-rate conversion pipe (|v| <= 4 * 127 * 127)
#pragma unroll 1
                for (int tt = 0; tt < 8; tt++) {
                    const uint8_t * tile = stage + (size_t) tt * ud.bytes, * sl = tile + lane * 16;
                    const uint4 w0 = lds_u4(sl), w1 = lds_u4(sl + 512);
                    const float dw = __half2float(*reinterpret_cast<const __half *>(tile + 1024 + lane * 2));
#pragma unroll
                    for (int j = 0; j < PB_TOK; j++) {
                        const uint8_t * r = stage + w_bytes + (size_t) (warp * PB_TOK + j) * rb;
                        const int4 a0 = *reinterpret_cast<const int4 *>(r + tt * 32), a1 = *reinterpret_cast<const int4 *>(r + tt * 32 + 16);
                        const float d = __fmul_rn(dw, *reinterpret_cast<const float *>(r + 256 + tt * 4));      // fp16(x.d) * fp16(y.d)
#pragma unroll
                        for (int wi = 0; wi < 4; wi++) {
                            const float f0 = __fsub_rn(__int_as_float(__dp4a((int) word_of(w0, wi), word_of(a0, wi), 0x4B400000)), 12582912.f);
                            const float f1 = __fsub_rn(__int_as_float(__dp4a((int) word_of(w1, wi), word_of(a1, wi), 0x4B400000)), 12582912.f);
                            acc[j][wi]     = __fmaf_rn(d, f0, acc[j][wi]);
                            acc[j][4 + wi] = __fmaf_rn(d, f1, acc[j][4 + wi]);
                        }
                    }
                }
            } else {
#pragma unroll
            for (int j = 0; j < PB_TOK; j++) {                 // (unrolled: acc[j] stays in registers)
                const uint8_t * r = stage + w_bytes + (size_t) (warp * PB_TOK + j) * rb;
                ActSmem A;
                A.q = (int8_t *) r; A.dx = (float *) (r + 256); A.bp = (int *) (r + 272); A.as = (int *) (r + 304);
                {
                    BlockInts bi;
                    tile_ints<TYPE>(stage, lane, 0, A, bi);
                    // chain step in registers, strictly in block order (the per-token kernel's CHAIN_REGS arithmetic)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[j][c] = __fmaf_rn(bi.d, (float) bi.s[c], acc[j][c]);
                    if (TYPE == T_Q4_K) {
#pragma unroll
                        for (int l = 0; l < 4; l++) acc[j][8 + l] = __fmaf_rn(bi.dmin, (float) bi.p[l], acc[j][8 + l]);
                    } else if (TYPE == T_Q5_K) {
                        acc[j][8] = __fadd_rn(acc[j][8], __fmul_rn(bi.dmin, (float) (bi.p[0] + bi.p[1] + bi.p[2] + bi.p[3])));
                    }
                }
            }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);        // this warp is done with the stage
            if (tid == 0 && step + PB_STAGES < n_steps) {
                mbar_wait(empty0 + 8 * s, par);                // every warp is: refill it
                issue(step + PB_STAGES);
            }
        }
#pragma unroll
        for (int j = 0; j < PB_TOK; j++) {
            const float val = finish_row<TYPE>(acc[j]);
            pb_epilogue(a, val, ud.row0 + lane, lane, chunk * PB_CHUNK + warp * PB_TOK + j);
        }
    };
    switch (ud.type) {
        case T_Q4_K: body(TypeTag<T_Q4_K>{}); break;
        case T_Q5_K: body(TypeTag<T_Q5_K>{}); break;
        case T_Q6_K: body(TypeTag<T_Q6_K>{}); break;
        default:     body(TypeTag<T_Q8_0>{}); break;
    }
}

}  // namespace b200
