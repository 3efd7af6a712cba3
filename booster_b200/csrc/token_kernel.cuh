// token_kernel.cuh — ONE persistent kernel per decoded token.
//
// The per-token work of llama_decode_internal (cpp/src/llama.cpp:14537-14840, graph build_llama :8781-8925) is a fixed list
// of phases — per layer: QKV mat-vec, attention scores, soft-max + P.V, wo, gate|up, down; then the head — and every phase
// needs the complete result of the one before it. Launched as separate kernels (kernels.cuh, the first design) a layer
// costs 58 us against 23 us of HBM time: each launch pays a kernel boundary, a cold instruction path and a serial prologue
// while HBM idles. Here one co-resident grid (one 512-thread CTA per SM) walks the phase list:
//   * phases are joined by a GRID BARRIER (one red.release per CTA + an ld.acquire spin on a monotonic 64-bit counter)
//     instead of a kernel boundary;
//   * a phase is split in two halves: `begin` — everything that does not depend on the previous phase's result (barrier
//     init, the first weight tiles of the mat-vec's TMA ring, norm weights, K rows and V slices of EARLIER positions) — runs
//     BEFORE the CTA waits at the barrier, so HBM keeps streaming through it; `run` is the dependent part;
//   * the code is executed 32 times per token and stays in the instruction caches (the separate kernels start cold);
//   * phase descriptors (the same MatvecArgs the stand-alone kernel takes) live in global memory and are copied into shared
//     memory one phase ahead.
// The arithmetic is the SAME device code as the stand-alone kernels (mv_begin / mv_run, the attention steps below follow
// k_attn_scores / k_attn_softmax_pv), so every result is bit-identical to them and to the reference.
#pragma once
#include "kernels.cuh"
#include <cstdio>

namespace b200 {

enum { PH_MATVEC = 0, PH_SCORES = 1, PH_SOFTMAX_PV = 2 };
static constexpr int TK_THREADS = 512;
static constexpr int TK_SC_TILE = TK_THREADS / 4;      // key positions per scores block (4 lanes per key row)
static constexpr int TK_PV_DIMS = 8;                   // output dims per soft-max + P.V block (one 16-byte piece of a V row)

struct AttnPhase {
    const float * q;          // [n_head][128] post-RoPE
    const __half * k_cache;   // this layer's [n_ctx][kv_dim]
    const __half * v_cache;
    float * S;                // scores, TRANSPOSED per head: element (h, t) at (h*16 + (t & 15)) * rs + (t >> 4), so that the
                              // 16 tinyBLAS chains of P.V (positions t = c mod 16) each read a contiguous row
    int rs;                   // row stride: n_ctx / 16 rounded up to 4
    float * out;              // [n_head*128]  (kqv_merged_cont)
    int n_head_kv, kv_dim;
    float scale;
    const DecodeState * st;
    const int32_t * cell_pos; // position held by each KV cell (read only when st->managed)
    int v_chunk;              // positions of V staged in shared memory at a time (multiple of 64)
};

struct alignas(16) Phase {
    int kind;
    int pad_[3];
    MatvecArgs mv;
    AttnPhase at;
};

// ------------------------------------------------------------------------------------------------------------
// grid barrier: a monotonic counter; every CTA adds 1 per barrier, barrier i of a launch completes at base + (i+1)*n_cta.
// The waiter's fence (fence.acq_rel.gpu -> the SM's L1 is invalidated) makes the other CTAs' global writes visible to the
// plain loads of ALL threads of the CTA that follow the CTA barrier after it (the cooperative-groups grid.sync() pattern).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// one thread, after a CTA-wide barrier: the release is cumulative over the CTA's writes that the barrier ordered before it
__device__ __forceinline__ void grid_arrive(unsigned long long * ctr) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" :: "l"(ctr), "l"(1ull) : "memory");
}
// A grid that never gathers (a CTA that died, a launch that was not co-resident) must not hang the process: after
// GRID_WAIT_NS of spinning the waiter reports where it stands and traps — the host sees a launch failure.
static constexpr unsigned long long GRID_WAIT_NS = 4000000000ull;
__device__ __noinline__ void grid_wait_timeout(const unsigned long long * ctr, unsigned long long target, int phase) {
    printf("booster_b200: grid barrier timeout: CTA %d phase %d counter %llu target %llu\n", (int) blockIdx.x, phase, ld_acquire_u64(ctr), target);
    __trap();
}
__device__ __forceinline__ void grid_wait(const unsigned long long * ctr, unsigned long long target, int phase) {
    // relaxed polls (no fence per poll), ONE acquire fence once the grid has gathered
    if (ld_relaxed_u64(ctr) < target) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        while (ld_relaxed_u64(ctr) < target) {
            if ((++spins & 1023u) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > GRID_WAIT_NS) grid_wait_timeout(ctr, target, phase);
            }
        }
    }
    __threadfence();
}
__device__ __forceinline__ void cp_async16_cg(void * smem_dst, const void * gsrc) { cp_async16(smem_dst, gsrc); }

// ------------------------------------------------------------------------------------------------------------
// attention scores: S[h][t] = (K[t] . q[h]) * scale for t <= pos, -inf for the padded tail — k_attn_scores' arithmetic
// (4 lanes per key row, lane c4 owns the tinyBLAS chains 4c4..4c4+3). A block = (KV head g, 128 positions); blocks are
// dealt round-robin to the CTAs. K rows of EARLIER positions are requested in `begin` (registers), before the barrier.
// ------------------------------------------------------------------------------------------------------------
struct ScState { uint2 kv[8]; };

__device__ __forceinline__ void sc_load_k(const AttnPhase & a, int vb, int n_blocks, int n_kv, int tid, uint2 (&kv)[8]) {
#pragma unroll
    for (int s = 0; s < 8; s++) kv[s] = make_uint2(0u, 0u);
    if (vb >= n_blocks) return;
    const int g = vb % a.n_head_kv, tile = vb / a.n_head_kv;
    const int t = tile * TK_SC_TILE + (tid >> 2), c4 = tid & 3;
    if (t < n_kv && t != a.st->cell) {
        const uint2 * kr = reinterpret_cast<const uint2 *>(a.k_cache + (size_t) t * a.kv_dim + g * 128 + 4 * c4);
#pragma unroll
        for (int s = 0; s < 8; s++) kv[s] = __ldg(kr + s * 4);            // 4 halfs at element 16s + 4c4 (row of an earlier token)
    }
}
__device__ __forceinline__ void scores_begin(const AttnPhase & a, int cta, ScState & s) {
    const int n_kv = a.st->n_kv;                               // DecodeState is not written during the token
    const int n_pad = (n_kv + 31) / 32 * 32;
    const int n_blocks = a.n_head_kv * ((n_pad + TK_SC_TILE - 1) / TK_SC_TILE);
    sc_load_k(a, cta, n_blocks, n_kv, threadIdx.x, s.kv);
}
template <int GQA>
__device__ __forceinline__ void scores_run(const AttnPhase & a, int cta, int n_cta, float (*qs)[128], ScState & s) {
    constexpr int HD = 128;
    const int tid = threadIdx.x, tl = tid >> 2, c4 = tid & 3;
    const int n_kv = a.st->n_kv;
    const int n_pad = (n_kv + 31) / 32 * 32;
    const int n_blocks = a.n_head_kv * ((n_pad + TK_SC_TILE - 1) / TK_SC_TILE);
    const int round_q = a.st->round_q;
    int g_staged = -1;
    for (int vb = cta; vb < n_blocks; vb += n_cta) {
        const int g = vb % a.n_head_kv, tile = vb / a.n_head_kv;
        const int t = tile * TK_SC_TILE + tl;
        if (g != g_staged) {
            if (g_staged >= 0) __syncthreads();                // the previous block's reads of qs are done
            for (int i = tid; i < GQA * HD; i += TK_THREADS) {
                float v = a.q[(size_t) (g * GQA) * HD + i];
                if (round_q) v = __half2float(__float2half_rn(v));    // src1 converted to the vec_dot_type F16 (ggml.c:12345-12371)
                (&qs[0][0])[i] = v;
            }
            g_staged = g;
            __syncthreads();
        }
        const bool visible = attn_cell_visible(a.st, a.cell_pos, t, n_kv);
        if (t == a.st->cell) {                                 // this token's row: written by the QKV phase (plain loads)
            const uint2 * kr = reinterpret_cast<const uint2 *>(a.k_cache + (size_t) t * a.kv_dim + g * HD + 4 * c4);
#pragma unroll
            for (int e = 0; e < 8; e++) s.kv[e] = kr[e * 4];
        }
        float kf[8][4];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&s.kv[e].x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&s.kv[e].y));
            kf[e][0] = f0.x; kf[e][1] = f0.y; kf[e][2] = f1.x; kf[e][3] = f1.y;
        }
        sc_load_k(a, vb + n_cta, n_blocks, n_kv, tid, s.kv);  // the next block's rows are requested before this block's arithmetic
        if (t < n_pad) {                                       // n_pad % 32 == 0: whole warps take this branch together
#pragma unroll 1
            for (int h = 0; h < GQA; h++) {                    // rolled: instruction footprint
                float ch[4];
                if (!round_q) {
                    // tinyBLAS<16>: lane c: acc = fma(k[16s+c], q[16s+c], acc), s = 0..7
#pragma unroll
                    for (int e = 0; e < 4; e++) ch[e] = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const float4 qv = *reinterpret_cast<const float4 *>(&qs[h][16 * e + 4 * c4]);
                        ch[0] = __fmaf_rn(kf[e][0], qv.x, ch[0]); ch[1] = __fmaf_rn(kf[e][1], qv.y, ch[1]);
                        ch[2] = __fmaf_rn(kf[e][2], qv.z, ch[2]); ch[3] = __fmaf_rn(kf[e][3], qv.w, ch[3]);
                    }
                } else {
                    // ggml_vec_dot_f16: sum[j][c] over i in {0, 64}: element i + 16j + c; then (0+2)+(1+3)
                    float aj[4][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float4 q0 = *reinterpret_cast<const float4 *>(&qs[h][16 * j + 4 * c4]);
                        const float4 q1 = *reinterpret_cast<const float4 *>(&qs[h][64 + 16 * j + 4 * c4]);
                        aj[j][0] = __fmaf_rn(kf[4 + j][0], q1.x, __fmul_rn(kf[j][0], q0.x));
                        aj[j][1] = __fmaf_rn(kf[4 + j][1], q1.y, __fmul_rn(kf[j][1], q0.y));
                        aj[j][2] = __fmaf_rn(kf[4 + j][2], q1.z, __fmul_rn(kf[j][2], q0.z));
                        aj[j][3] = __fmaf_rn(kf[4 + j][3], q1.w, __fmul_rn(kf[j][3], q0.w));
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) ch[e] = __fadd_rn(__fadd_rn(aj[0][e], aj[2][e]), __fadd_rn(aj[1][e], aj[3][e]));
                }
                // _mm512_reduce_add_ps over the 16 chains: lanes c4=0..3 hold chains 4c4..4c4+3
                float t3[4], t6[4];
#pragma unroll
                for (int e = 0; e < 4; e++) t3[e] = __fadd_rn(__shfl_xor_sync(0xffffffffu, ch[e], 2), ch[e]);   // a[8+i] + a[i] (valid in c4 = 0,1)
#pragma unroll
                for (int e = 0; e < 4; e++) t6[e] = __fadd_rn(__shfl_xor_sync(0xffffffffu, t3[e], 1), t3[e]);   // t3[4+i] + t3[i] (valid in c4 = 0)
                const float res = __fadd_rn(__fadd_rn(t6[0], t6[2]), __fadd_rn(t6[1], t6[3]));
                if (c4 == 0)
                    a.S[((size_t) (g * GQA + h) * 16 + (t & 15)) * a.rs + (t >> 4)] = visible ? __fmul_rn(res, a.scale) : -INFINITY;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// soft_max_ext + P.V: a block = (KV head g, 8 output dims) and ALL positions (the 16 chains of an output element run over
// t in order). Every block of a KV head normalises the GQA score rows itself (k_attn_softmax_pv's arithmetic: a thread
// owns whole 16-position vectors, so the _mm512_reduce_add_ps tree of the exponentials is register arithmetic; per-vector
// float sums accumulate in double). The rows sit in shared memory TRANSPOSED ([head][chain][t / 16]); P.V then runs as ONE
// WARP PER HEAD: lane = (chain c, dim quad): acc[i] = fma(V[16s + c][4dq + i], p[16s + c], acc[i]) over s — one 8-byte V
// load per step and one 16-byte load of four consecutive p per four steps (the stand-alone kernel spends two shared-memory
// loads per fma and is LSU-bound).
//   shared: ps f32 [GQA][16][rs] | vs f16 [v_chunk][8]
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pv_stage_v(const AttnPhase & a, int g, int slice, int t0, int n_kv, int n_pad, bool cur_row_too, __half (*vs)[TK_PV_DIMS]) {
    const __half * vbase = a.v_cache + g * 128 + slice * TK_PV_DIMS;
    const int len = min(a.v_chunk, n_pad - t0);
    for (int i = threadIdx.x; i < len; i += TK_THREADS) {
        const int t = t0 + i;
        if (t < n_kv && (cur_row_too || t != a.st->cell)) cp_async16(&vs[i][0], vbase + (size_t) t * a.kv_dim);
        else if (t >= n_kv) *reinterpret_cast<uint4 *>(&vs[i][0]) = make_uint4(0u, 0u, 0u, 0u);   // p == 0 there: the product must be 0, never NaN
    }
    cp_async_commit();
}
template <int GQA>
__device__ __forceinline__ void pv_begin(const AttnPhase & a, int cta, uint8_t * dyn) {
    const int n_kv = a.st->n_kv;
    const int n_pad = (n_kv + 31) / 32 * 32;
    if (cta >= a.n_head_kv * (128 / TK_PV_DIMS)) return;
    __half (*vs)[TK_PV_DIMS] = reinterpret_cast<__half (*)[TK_PV_DIMS]>(dyn + (size_t) GQA * 16 * a.rs * 4);
    pv_stage_v(a, cta % a.n_head_kv, cta / a.n_head_kv, 0, n_kv, n_pad, false, vs);   // rows of earlier tokens only
}
template <int GQA>
__device__ __forceinline__ void pv_run(const AttnPhase & a, int cta, int n_cta, uint8_t * dyn,
                                       float (*redf)[16], double (*redd)[16], float (*red)[16][TK_PV_DIMS + 1]) {
    constexpr int HD = 128;
    constexpr int TH = TK_THREADS / GQA;                       // threads per head in the soft-max
    constexpr int NW = TH / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = tid / TH, ht = tid % TH, w = ht >> 5;
    const int n_kv = a.st->n_kv;
    const int n_pad = (n_kv + 31) / 32 * 32;
    const int n16 = n_pad / 16, rs = a.rs;
    float * ps = reinterpret_cast<float *>(dyn);
    __half (*vs)[TK_PV_DIMS] = reinterpret_cast<__half (*)[TK_PV_DIMS]>(dyn + (size_t) GQA * 16 * rs * 4);
    const int n_blocks = a.n_head_kv * (HD / TK_PV_DIMS);
    for (int vb = cta; vb < n_blocks; vb += n_cta) {
        const int g = vb % a.n_head_kv, slice = vb / a.n_head_kv;
        if (vb != cta) {
            __syncthreads();                                   // the previous block is done with ps / vs
            pv_stage_v(a, g, slice, 0, n_kv, n_pad, true, vs);
        } else if (a.st->cell < a.v_chunk && tid == 0) {
            // the first block's earlier rows were requested in pv_begin; this token's row is written by the QKV phase
            cp_async16(&vs[a.st->cell][0], a.v_cache + g * HD + slice * TK_PV_DIMS + (size_t) a.st->cell * a.kv_dim);
        }
        // the GQA x 16 score rows of the KV head, n16 floats each (16-byte copies: the rows are padded to 4 floats)
        {
            const float * Sg = a.S + (size_t) (g * GQA) * 16 * rs;
            const int n4 = (n16 + 3) / 4;
            for (int row = warp; row < GQA * 16; row += TK_THREADS / 32)
                for (int j = lane; j < n4; j += 32) cp_async16(ps + (size_t) row * rs + 4 * j, Sg + (size_t) row * rs + 4 * j);
            cp_async_commit();
            cp_async_wait<0>();                                // (also completes this thread's V copies)
        }
        __syncthreads();
        // soft_max_ext of the head's row (cpp/ggml/src/ggml.c:13682-13778): vector s = positions 16s .. 16s+15 = element s
        // of the 16 chain rows
        float * row = ps + (size_t) h * 16 * rs;
        float mx = -INFINITY;
        for (int s = ht; s < n16; s += TH) {
#pragma unroll
            for (int c = 0; c < 16; c++) mx = fmaxf(mx, row[c * rs + s]);
        }
        mx = warp_max(mx);
        if (lane == 0) redf[h][w] = mx;
        asm volatile("bar.sync %0, %1;" :: "r"(1 + h), "r"(TH) : "memory");
        mx = redf[h][0];
#pragma unroll
        for (int j = 1; j < NW; j++) mx = fmaxf(mx, redf[h][j]);
        double part = 0.0;
        for (int s = ht; s < n16; s += TH) {
            // _mm512_reduce_add_ps of the vector's 16 exponentials: (a[8+i] + a[i]) pairs chain i with i+8, so the vector is
            // walked as two such groups of 4 + 4 in a ROLLED loop (8 inlined ggml_v_expf instead of 16: instruction footprint)
            float4 tq[2];
#pragma unroll 1
            for (int pq = 0; pq < 2; pq++) {
                float * lo = row + (size_t) (4 * pq) * rs + s, * hi = lo + (size_t) 8 * rs;
                float4 l, hgh;
                l.x = v_expf(__fsub_rn(lo[0], mx)); l.y = v_expf(__fsub_rn(lo[rs], mx)); l.z = v_expf(__fsub_rn(lo[2 * rs], mx)); l.w = v_expf(__fsub_rn(lo[3 * rs], mx));
                hgh.x = v_expf(__fsub_rn(hi[0], mx)); hgh.y = v_expf(__fsub_rn(hi[rs], mx)); hgh.z = v_expf(__fsub_rn(hi[2 * rs], mx)); hgh.w = v_expf(__fsub_rn(hi[3 * rs], mx));
                lo[0] = l.x; lo[rs] = l.y; lo[2 * rs] = l.z; lo[3 * rs] = l.w;
                hi[0] = hgh.x; hi[rs] = hgh.y; hi[2 * rs] = hgh.z; hi[3 * rs] = hgh.w;
                const float4 tt = make_float4(__fadd_rn(hgh.x, l.x), __fadd_rn(hgh.y, l.y), __fadd_rn(hgh.z, l.z), __fadd_rn(hgh.w, l.w));
                if (pq) tq[1] = tt; else tq[0] = tt;           // chains 0-3|8-11 are t[0..3] of the tree, 4-7|12-15 are t[4..7]
            }
            const float u0 = __fadd_rn(tq[1].x, tq[0].x), u1 = __fadd_rn(tq[1].y, tq[0].y), u2 = __fadd_rn(tq[1].z, tq[0].z), u3 = __fadd_rn(tq[1].w, tq[0].w);
            part += (double) __fadd_rn(__fadd_rn(u0, u2), __fadd_rn(u1, u3));
        }
        part = warp_sum_d(part);
        if (lane == 0) redd[h][w] = part;
        __syncthreads();                                       // every head's exponentials and partial sums, all V copies
        // ---- P.V: warp = head, lane = (chain c, dim quad dq)
        if (warp < GQA) {
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < NW; j++) sum += redd[warp][j];
            const float inv = (float) (1.0 / sum);
            const int c = lane >> 1, dq = lane & 1;
            const float * pr = ps + ((size_t) warp * 16 + c) * rs;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int t0 = 0; t0 < n_pad; t0 += a.v_chunk) {
                const int steps = min(a.v_chunk, n_pad - t0) / 16, s0 = t0 >> 4;
                const __half * vr = &vs[c][4 * dq];
                int s = 0;
                for (; s + 4 <= steps; s += 4) {
                    const float4 p4 = *reinterpret_cast<const float4 *>(pr + s0 + s);
                    const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint2 vv = *reinterpret_cast<const uint2 *>(vr + (size_t) (s + u) * 16 * TK_PV_DIMS);
                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&vv.x));
                        const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&vv.y));
                        const float p = __fmul_rn(pp[u], inv);   // the normalised probability (ggml_vec_scale_f32 by 1/sum)
                        acc[0] = __fmaf_rn(f0.x, p, acc[0]); acc[1] = __fmaf_rn(f0.y, p, acc[1]);
                        acc[2] = __fmaf_rn(f1.x, p, acc[2]); acc[3] = __fmaf_rn(f1.y, p, acc[3]);
                    }
                }
                for (; s < steps; s++) {
                    const uint2 vv = *reinterpret_cast<const uint2 *>(vr + (size_t) s * 16 * TK_PV_DIMS);
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&vv.x));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&vv.y));
                    const float p = __fmul_rn(pr[s0 + s], inv);
                    acc[0] = __fmaf_rn(f0.x, p, acc[0]); acc[1] = __fmaf_rn(f0.y, p, acc[1]);
                    acc[2] = __fmaf_rn(f1.x, p, acc[2]); acc[3] = __fmaf_rn(f1.y, p, acc[3]);
                }
                if (t0 + a.v_chunk < n_pad) {                  // contexts longer than the V stage: next chunk (all GQA warps agree)
                    asm volatile("bar.sync 14, %0;" :: "r"(GQA * 32) : "memory");
                    const __half * vbase = a.v_cache + g * HD + slice * TK_PV_DIMS;
                    const int t1 = t0 + a.v_chunk, len = min(a.v_chunk, n_pad - t1);
                    for (int i = tid; i < len; i += GQA * 32) {
                        if (t1 + i < n_kv) cp_async16(&vs[i][0], vbase + (size_t) (t1 + i) * a.kv_dim);
                        else *reinterpret_cast<uint4 *>(&vs[i][0]) = make_uint4(0u, 0u, 0u, 0u);
                    }
                    cp_async_commit();
                    cp_async_wait<0>();
                    asm volatile("bar.sync 14, %0;" :: "r"(GQA * 32) : "memory");
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) red[warp][c][4 * dq + i] = acc[i];
        }
        __syncthreads();
        if (tid < GQA * TK_PV_DIMS) {
            const int hh = tid / TK_PV_DIMS, dd = tid % TK_PV_DIMS;
            float t3[8], t6[4];                                // _mm512_reduce_add_ps over the 16 chains
#pragma unroll
            for (int j = 0; j < 8; j++) t3[j] = __fadd_rn(red[hh][8 + j][dd], red[hh][j][dd]);
#pragma unroll
            for (int j = 0; j < 4; j++) t6[j] = __fadd_rn(t3[4 + j], t3[j]);
            a.out[(size_t) (g * GQA + hh) * HD + slice * TK_PV_DIMS + dd] =
                __fadd_rn(__fadd_rn(t6[0], t6[2]), __fadd_rn(t6[1], t6[3]));     // kqv_merged_cont layout: [n_head*hd]
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// the per-token kernel. trace (TR): [phase][cta][4] globaltimer stamps: 0 run starts | 1 run done | 2 arrived + next phase's
// begin done | 3 barrier passed.
// ------------------------------------------------------------------------------------------------------------
static constexpr int TK_TRACE_SLOTS = 4;
template <int GQA, bool TR>
__global__ void __launch_bounds__(TK_THREADS, 1) k_token(const Phase * __restrict__ plan, int n_phases, unsigned long long * bar,
                                                         unsigned long long * trace) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ Phase ph[2];
    __shared__ double red_smem[MV_MAX_WARPS];
    __shared__ __align__(16) float qs[GQA][128];
    __shared__ float  redf[GQA][16];
    __shared__ double redd[GQA][16];
    __shared__ float  red[GQA][16][TK_PV_DIMS + 1];
    __shared__ unsigned long long bar_base;
    const int tid = threadIdx.x, warp = tid >> 5, cta = blockIdx.x, n_cta = gridDim.x;
    constexpr int PH_CHUNKS = (int) (sizeof(Phase) / 16);

    for (int i = tid; i < PH_CHUNKS; i += TK_THREADS)
        reinterpret_cast<uint4 *>(&ph[0])[i] = reinterpret_cast<const uint4 *>(plan)[i];
    if (tid == 0) {
        // every launch adds exactly (n_phases - 1) * n_cta to the counter, so its value at launch is a multiple of that: a CTA
        // that reads it after faster CTAs have already arrived at the first barrier still finds the same base
        const unsigned long long per = (unsigned long long) (n_phases > 1 ? n_phases - 1 : 1) * (unsigned long long) n_cta;
        const unsigned long long c0 = ld_acquire_u64(bar);
        bar_base = c0 - c0 % per;
    }
    __syncthreads();

    MvState ms;
    ScState ss;
    auto begin = [&](const Phase & P) {
        if (P.kind == PH_MATVEC) { if (warp < P.mv.warps) mv_begin<false>(P.mv, smem_raw, cta, n_cta, P.mv.prefill, ms); }
        else if (P.kind == PH_SCORES) scores_begin(P.at, cta, ss);
        else pv_begin<GQA>(P.at, cta, smem_raw);
    };
    auto stamp = [&](int p, int slot) {
        if (TR) {
            if (trace != nullptr && tid == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                trace[((size_t) p * n_cta + cta) * TK_TRACE_SLOTS + slot] = t;
            }
        }
    };
    begin(ph[0]);
    for (int p = 0; p < n_phases; p++) {
        const Phase & P = ph[p & 1];
        if (p + 1 < n_phases) {                                // the next descriptor, one phase ahead
            for (int i = tid; i < PH_CHUNKS; i += TK_THREADS)
                cp_async16(&reinterpret_cast<uint4 *>(&ph[(p + 1) & 1])[i], &reinterpret_cast<const uint4 *>(plan + p + 1)[i]);
            cp_async_commit();
        }
        stamp(p, 0);
#if B200_HANG_DEBUG
        if (tid == 0) g_dbg_phase[cta & 255] = p;
#endif
        if (P.kind == PH_MATVEC) { if (warp < P.mv.warps) mv_run<false>(P.mv, red_smem, P.mv.prefill, ms); }
        else if (P.kind == PH_SCORES) scores_run<GQA>(P.at, cta, n_cta, qs, ss);
        else pv_run<GQA>(P.at, cta, n_cta, smem_raw, redf, redd, red);
        cp_async_wait<0>();
        __syncthreads();                                       // the phase's shared memory is free, its global writes are issued
        if (P.kind == PH_MATVEC && warp < P.mv.warps) mv_end(P.mv, ms);
        stamp(p, 1);
        if (p + 1 < n_phases) {
            if (tid == 0) grid_arrive(bar);
            begin(ph[(p + 1) & 1]);                            // HBM streams while the grid gathers
            stamp(p, 2);
            if (tid == 0) grid_wait(bar, bar_base + (unsigned long long) (p + 1) * (unsigned long long) n_cta, p);
            __syncthreads();
            stamp(p, 3);
        }
    }
}

}  // namespace b200
