// kernels.cuh — hand-written sm_100a kernels for the quantized LLaMA decode path, BIT-EXACT with the
// reference's CPU arithmetic (its AVX-512 "native" build; oracle/oracle_port.c is the scalar statement).
//
// Why bit-exact and not "within tolerance": activations are re-quantized to int8 (Q8_K / Q8_0) before every
// mat-mul. A 1-ulp difference in any upstream fp32 value eventually flips one int8, which moves that layer's
// output by ~1e-3, which flips dozens of int8 in the next quantization, and within two or three layers the
// deviation saturates at the quantization-noise floor (~1e-2 of the logits) and stays there through the KV cache.
// (Measured: DESIGN.md "why bit-exact".) The north-star tolerance of 1e-3 is therefore only reachable by
// reproducing every fp32 operation of the reference in its exact order:
//   * quantized dots: integer sums per super-block are exact; the fp32 part follows ggml_vec_dot_q{4,5,6}_K_q8_K's
//     AVX2 code (cpp/ggml/src/ggml-quants.c:6914-6977, 7487-7560, 8145-8220): 8 lanes, lane m owns bytes 4m..4m+3
//     of every 32-byte group, acc[m] = fma(d_b, (float) sumi_b[m], acc[m]) sequentially over super-blocks b,
//     then hsum_float_8 (:47-53). Q8_0 follows ggml_vec_dot_q8_0_q8_0 / tinyBLAS_Q0_AVX (:5361-5382).
//   * RMSNorm: double accumulation (cpp/ggml/src/ggml.c:11850-11896).
//   * RoPE: cos/sin from a host-built table (ggml_rope_cache_init, cpp/ggml/src/ggml.c:14017-14031).
//   * attention (default route, cpp/src/llama.cpp:8248-8297): K.q and V.p as tinyBLAS<16> computes them
//     (cpp/ggml/src/llamafile/sgemm.cpp:408-430: 16 lanes, fma chain over k, _mm512_reduce_add_ps), batch>1 K.q as
//     ggml_vec_dot_f16 (cpp/ggml/src/ggml.c:2038-2075), softmax with the ggml_v_expf polynomial (:2447-2472).
//   * SiLU: ggml_v_silu (cpp/ggml/src/ggml.c:2475-2482).
//
// HBM layout ("tiles"): a matrix keeps exactly its GGUF bytes, re-tiled once at load. Tile T = (row/32)*nb + b
// holds block b of 32 consecutive rows, ONE ROW PER LANE, every field transposed so that a warp-wide load of one
// field is a dense 128/512-byte stream. A lane therefore walks the blocks of its own row in order and its fp32
// fma chains never leave its registers:
// fma chains never leave its registers. A tile is ONE contiguous chunk (a single TMA bulk copy):
//   Q4_K (4608 B): qs uint4[8][32] | u32[4][32] = 12 scale bytes + (d,dmin)
//   Q5_K (5632 B): same | qh uint4[2][32]
//   Q6_K (6720 B): ql uint4[8][32] | scales u32[4][32] | qh uint4[4][32] | d u16[32]
//   Q8_0 (1088 B): qs uint4[2][32] | d u16[32]                 (blocks of 32 weights)
// ffn_gate and ffn_up are interleaved row by row into one virtual matrix (row 2r = gate r, 2r+1 = up r).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

// build-time A/B switches (scripts/gpu_variants.sh builds one library per setting)
#ifndef B200_CHAIN_UNROLL
#define B200_CHAIN_UNROLL 1
#endif
#ifndef B200_LOOKAHEAD
#define B200_LOOKAHEAD 0     // L2 look-ahead prefetches (PfRange): measured neutral-to-negative on B200, see DESIGN.md
#endif
#ifndef B200_WARM
#define B200_WARM 0          // dry run of the prologue before griddepcontrol.wait (instruction-cache warm-up)
#endif

namespace b200 {

enum { T_F32 = 0, T_F16 = 1, T_Q8_0 = 8, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14 };

struct TMat {
    int type = 0;
    int n_rows = 0;       // (virtual) rows
    int nb = 0;           // blocks per row: 256-weight super-blocks, or 32-weight blocks for Q8_0
    int rows_unit = 32;   // a work unit = 32 rows (one per lane) x all nb blocks
    int tiles_unit = 0;   // = nb
    int n_units = 0;      // = n_rows/32
    int tile_bytes = 0;   // = tile_bytes_of(type)
    const uint8_t * p0 = nullptr;   // tiles, tile_bytes_of(type) each
};
__host__ __device__ __forceinline__ int tile_bytes_of(int type) {
    return type == T_Q4_K ? 4608 : type == T_Q5_K ? 5632 : type == T_Q6_K ? 6720 : 1088;
}

// per-token scalars, device-resident so that one CUDA graph serves every token
struct alignas(16) DecodeState {
    int32_t token;     // input token id of this step
    int32_t pos;       // its position
    int32_t round_q;   // 1: batch > 1 arithmetic for K.q (q rounded to f16, ggml_vec_dot_f16 order)
    int32_t step;      // greedy loop: index into out_tokens
    // KV cells (struct llama_kv_cache, cpp/src/llama.cpp:2495-2539). Until a context shift cell == pos and n_kv == pos + 1
    // (what llama_kv_cache_find_slot gives a sequence that only ever grows); after llama_kv_cache_seq_rm / seq_add the cells
    // keep their places, freed cells are re-used by new tokens, and attention masks by the cells' positions (managed != 0).
    int32_t cell;      // KV cell this token's K / V rows are written to
    int32_t n_kv;      // cells attended: highest used cell + 1 (the kernels pad it to 32 like the reference, :14698)
    int32_t managed;   // 0: cell t holds position t for t < n_kv; 1: positions come from cell_pos[]
    int32_t pad_;
};

enum { EPI_STORE = 0, EPI_RESID = 1, EPI_QKV = 2, EPI_SILU = 3 };

// L2 look-ahead: a byte range of HBM that a LATER kernel of the token will stream (its weight tiles, this layer's K/V
// rows). Every kernel of the forward pass carries up to PF_RANGES of them and turns them into bulk L2 prefetches in its
// data-independent prologue, so that HBM keeps streaming while the (short, latency-bound) kernels in between wait on
// each other; the consumer then finds its tiles in the 126 MB L2. per_pos != 0: the range grows with the position
// (K/V rows 0..pos): bytes = min(bytes, (pos + 1) * per_pos).
static constexpr int PF_RANGES = 3;
static constexpr uint32_t PF_CHUNK = 8192;
struct PfRange { const uint8_t * p; uint32_t bytes; uint32_t per_pos; };

struct MatvecArgs {
    TMat seg[3];
    int n_seg;
    int n_units;               // total over segments
    int k;
    // prologue: x f32[k]; optional RMSNorm (norm_w != nullptr) then activation quantization into smem
    const float * x;
    const float * norm_w;
    float eps;
    double inv_k;              // 1.0 / k if k is a power of two, else 0 (rms_scale)
    int act_q8_0;              // 0: Q8_K activations, 1: Q8_0 activations
    int tiles_unit;            // identical for every segment of a launch (same K, same block width)
    int warps;                 // W = warps of the CTA that work on this mat-vec (all of them in k_matvec)
    int group;                 // G = warps sharing one 32-row unit (a divisor of W)
    int kpw;                   // = tiles_unit / group   (host-computed: integer divisions are ~20 instructions each on
    int groups_per_cta;        // = warps / group         the device, and these kernels are instruction-fetch bound)
    uint32_t grp_magic;        // warp / group == (warp * grp_magic) >> 16 for warp < 32
    int chain_mode;            // = chain_mode_of(group)
    uint32_t act_bytes;        // = act_smem_bytes(k, act_q8_0)
    uint32_t exch_words;       // floats of the exchange buffers in front of the hand-off region (0 unless CHAIN_EXCHANGE)
    uint32_t chain_bytes;      // = chain_smem_bytes(warps, group, nv)
    int stages;                // tiles of shared-memory ring per warp (>= 2)
    int stage_bytes;           // ring slot size = largest tile of the launch, 128-byte multiple
    int prefill;               // tiles per warp requested before griddepcontrol.wait (the rest follow the x loads)
    int nv;                    // chain values per tile (largest chain_values_of over the segments)
    int epi;                   // EPI_*: ONE kernel serves every mat-vec of the layer, so its code stays in the instruction cache
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
    PfRange pf[PF_RANGES];     // L2 look-ahead for the kernels that follow (bytes == 0: unused)
    const float * warm_x;      // B200_WARM: constant vector of k floats for the dry run of the prologue (nullptr: no dry run)
    unsigned long long * trace; // nullptr unless b200_trace_token is running
};

static constexpr int MV_MAX_WARPS = 16;                     // one persistent CTA per SM, 8..16 warps (host picks)
static constexpr int HANDOFF_WORDS = 12 * 32;               // final chain values of one unit: 12 fp32 chains x 32 rows
// values a tile publishes for the chain phase: s[8] (+ p[4] | sum p) as floats, d (+ dmin)
__host__ __device__ __forceinline__ int chain_values_of(int type) { return type == 12 ? 14 : type == 13 ? 11 : 9; }
// shared memory of the chain exchange when G > 1: per warp a double-buffered slot of `nv` x 32 floats, per group the
// final values
// K-split modes: 0 = none (G == 1, chains in registers); 1 = hand-off (small G: the chain state walks from warp to warp
// through a 1.5 KB buffer per group — the serial part is 12 FMAs per tile and the ring keeps its depth); 2 = exchange
// (large G: integers published per round, chains advanced in parallel over (row, chain))
enum { CHAIN_REGS = 0, CHAIN_HANDOFF = 1, CHAIN_EXCHANGE = 2 };
__host__ __device__ __forceinline__ int chain_mode_of(int G) { return G == 1 ? CHAIN_REGS : G <= 4 ? CHAIN_HANDOFF : CHAIN_EXCHANGE; }
__host__ __device__ __forceinline__ size_t chain_smem_bytes(int W, int G, int nv) {
    const int mode = chain_mode_of(G);
    if (mode == CHAIN_REGS) return 0;
    const size_t fin = (size_t) (W / G) * HANDOFF_WORDS * 4;
    return mode == CHAIN_HANDOFF ? fin : (size_t) W * 2 * nv * 128 + fin;
}

// ------------------------------------------------------------------------------------------------------------
// PDL (programmatic dependent launch): every kernel of the forward pass is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization. pdl_launch_dependents() lets the NEXT kernel's CTAs be
// scheduled as soon as SM resources free up (they run their data-independent prologue: barrier init, weight
// prefetch); pdl_wait() blocks until the PREVIOUS kernel has completed and its memory is visible. Rule kept by
// every kernel: no global read of anything a kernel writes, and no global write at all, before pdl_wait().
// ------------------------------------------------------------------------------------------------------------
// optional phase trace (b200_trace_token): thread 0 of every CTA stamps %globaltimer into tr[cta][phase]
static constexpr int TRACE_PHASES = 12;
__device__ __noinline__ void trace_stamp(unsigned long long * tr, int phase) {
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tr[(size_t) (blockIdx.y * gridDim.x + blockIdx.x) * TRACE_PHASES + phase] = t;
    }
}
// TR is a template parameter of every forward kernel: the production instantiation carries no trace code at all (these
// kernels run with a cold instruction cache, so code bytes are time — DESIGN.md "instruction footprint")
template <bool TR>
__device__ __forceinline__ void trace_mark(unsigned long long * tr, int phase) {
    if (TR) { if (tr != nullptr && threadIdx.x == 0) trace_stamp(tr, phase); }
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_v4(const void * p) {
    uint4 r;   // streaming 16-byte load: read-only path, no L1 allocation (weights are touched once per token)
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_u32(const void * p) { return __ldg((const uint32_t *) p); }
__device__ __forceinline__ uint32_t ldg_u16(const void * p) { return __ldg((const uint16_t *) p); }
__device__ __forceinline__ float h16_to_f32(uint32_t bits) { return __half2float(__ushort_as_half((unsigned short) bits)); }

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
__device__ __forceinline__ uint32_t word_of(const uint4 & v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ int      word_of(const int4 & v, int i)  { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// ggml_v_expf, AVX-512 variant (cpp/ggml/src/ggml.c:2447-2472), one lane. All steps are element-wise IEEE ops.
__device__ __noinline__ float scalef_slow(float j, int n) { return ldexpf(j, n); }   // sub-normal results: rare, out of line
__device__ __forceinline__ float v_expf(float x) {
    const float r = 0x1.8p23f;
    const float z = __fmaf_rn(x, 0x1.715476p+0f, r);
    const float n = __fsub_rn(z, r);
    const float b = __fmaf_rn(-n, 0x1.7f7d1cp-20f, __fmaf_rn(-n, 0x1.62e4p-1f, x));
    const bool big = fabsf(n) > 192.f;
    const float u = __fmul_rn(b, b);
    const float j = __fmaf_rn(__fmaf_rn(__fmaf_rn(0x1.0e4020p-7f, b, 0x1.573e2ep-5f), u,
                                         __fmaf_rn(0x1.555e66p-3f, b, 0x1.fffdb6p-2f)),
                              u, __fmaf_rn(0x1.ffffecp-1f, b, 1.0f));
    if (big) return n <= 0.f ? 0.f : INFINITY;
    // _mm512_scalef_ps: exact scaling by 2^n. j is in [0.70, 1.42], so for |n| <= 125 the result is a normal number
    // and the scaling is an addition to the exponent field; the (very rare) sub-normal range goes through ldexpf.
    const int ni = (int) n;
    if (ni >= -125 && ni <= 125) return __int_as_float(__float_as_int(j) + (ni << 23));
    return scalef_slow(j, ni);
}
// _mm512_reduce_add_ps of 16 values held by one thread (same tree as reduce_add16_shfl)
__device__ __forceinline__ float reduce_add16_regs(const float (&a)[16]) {
    float t[8], u[4];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = __fadd_rn(a[8 + i], a[i]);
#pragma unroll
    for (int i = 0; i < 4; i++) u[i] = __fadd_rn(t[4 + i], t[i]);
    return __fadd_rn(__fadd_rn(u[0], u[2]), __fadd_rn(u[1], u[3]));
}
// ggml_v_silu (cpp/ggml/src/ggml.c:2475-2482)
__device__ __forceinline__ float silu_exact(float x) {
    return __fdiv_rn(x, __fadd_rn(1.0f, v_expf(__fsub_rn(0.0f, x))));
}
// _mm512_reduce_add_ps over 16 consecutive lanes of a half-warp (lane & 15 = element), GCC's expansion:
// (a[8+i]+a[i]) -> (t[4+i]+t[i]) -> (u0+u2, u1+u3) -> sum. Result valid in the lane with (lane & 15) == 0.
__device__ __forceinline__ float reduce_add16_shfl(float v) {
    v = __fadd_rn(__shfl_down_sync(0xffffffffu, v, 8, 16), v);
    v = __fadd_rn(__shfl_down_sync(0xffffffffu, v, 4, 16), v);
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 2, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 1, 16));
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// activation quantization (block-cooperative, result in shared memory in natural order: all lanes of a warp work
// on the same block b of their own rows, so every activation read is a warp-wide broadcast)
//   Q8_K: q[k] int8 | dx[nb] f32 | bp[nb][8] int = sums of the 8 sub-blocks of 32 | as[nb][64] int = -32 * (sum of
//         the 4 quants of each 32-bit word): the "- 32" of Q6_K folded into the accumulator input of dp4a
//   Q8_0: q[k] int8 | dx[k/32] f32 (already rounded through fp16)
// ------------------------------------------------------------------------------------------------------------
struct ActSmem {
    int8_t * q;
    float  * dx;
    int    * bp;
    int    * as;
};
__host__ __device__ __forceinline__ size_t act_smem_bytes(int k, int act_q8_0) {
    size_t n = (size_t) k;
    n += (size_t) (act_q8_0 ? k / 32 : k / 256) * 4;
    n = (n + 15) / 16 * 16;
    if (!act_q8_0) n += (size_t) (k / 256) * 32 + (size_t) k;
    return (n + 15) / 16 * 16;
}
__device__ __forceinline__ ActSmem act_smem_carve(uint8_t * base, int k, int act_q8_0) {
    ActSmem a;
    a.q = (int8_t *) base;
    a.dx = (float *) (base + k);
    size_t off = (size_t) k + (size_t) (act_q8_0 ? k / 32 : k / 256) * 4;
    off = (off + 15) / 16 * 16;
    a.bp = (int *) (base + off);
    a.as = (int *) (base + off + (size_t) (k / 256) * 32);
    return a;
}

// One warp quantizes 256 consecutive values (8 per lane) to Q8_K — quantize_row_q8_K_ref
// (cpp/ggml/src/ggml-quants.c:3593-3630): `max` is the FIRST element of largest magnitude, iscale = -127/max,
// q = min(127, round_half_even(iscale*x)), d = 1/iscale.
// (__noinline__ + by-value operands: ONE copy of this code in the kernel — these short kernels run with a cold
// instruction cache, so code bytes are time; see DESIGN.md "instruction footprint")
// QB blocks (b, b + bstride, ...; the first `nvalid` of them exist, the others carry zeros and are not stored) go through
// the steps TOGETHER: the arg-max reductions, the two divisions and the rounding of one block are a dependent chain
// of several hundred cycles, and a warp that owns four blocks of a 14336-long vector (ffn_down's input) would
// otherwise walk four such chains back to back.
// Only QB = 1 is instantiated: running a warp's four blocks of a 14336-long vector through the steps together
// (QB = 4) shortens ffn_down's prologue by 1.3 us but the extra 600 instructions cost more than that (r01n/r01q A/B).
template <int QB>
__device__ __noinline__ void q8k_blocks_warp(float4 a0, float4 b0_, float4 a1, float4 b1_, float4 a2, float4 b2_, float4 a3, float4 b3_,
                                             int lane, int b, int bstride, int nvalid, ActSmem A) {
    const float4 va[4] = {a0, a1, a2, a3}, vb[4] = {b0_, b1_, b2_, b3_};
    float v[QB][8];
#pragma unroll
    for (int u = 0; u < QB; u++) {
        v[u][0] = va[u].x; v[u][1] = va[u].y; v[u][2] = va[u].z; v[u][3] = va[u].w;
        v[u][4] = vb[u].x; v[u][5] = vb[u].y; v[u][6] = vb[u].z; v[u][7] = vb[u].w;
    }
    float amax[QB], mval[QB];
    int   midx[QB];
#pragma unroll
    for (int u = 0; u < QB; u++) {
        amax[u] = 0.f; mval[u] = 0.f; midx[u] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float ax = fabsf(v[u][i]);
            if (ax > amax[u]) { amax[u] = ax; mval[u] = v[u][i]; midx[u] = lane * 8 + i; }
        }
    }
    // first element of largest magnitude over the warp: two hardware warp reductions (redux.sync) and one shuffle.
    // |x| >= 0, so the float bit patterns order like unsigned integers.
    unsigned amax_all[QB], first[QB];
#pragma unroll
    for (int u = 0; u < QB; u++) amax_all[u] = __reduce_max_sync(0xffffffffu, __float_as_uint(amax[u]));
#pragma unroll
    for (int u = 0; u < QB; u++) first[u] = __reduce_min_sync(0xffffffffu, __float_as_uint(amax[u]) == amax_all[u] ? (unsigned) midx[u] : 0xffffffffu);
#pragma unroll
    for (int u = 0; u < QB; u++) mval[u] = __shfl_sync(0xffffffffu, mval[u], (int) (first[u] >> 3));
    float iscale[QB], d[QB];
#pragma unroll
    for (int u = 0; u < QB; u++) {
        // an all-zero block keeps q = 0, d = 0 (cpp/ggml/src/ggml-quants.c:3607-3612): iscale 0 rounds every value to 0
        const float is = __fdiv_rn(-127.f, mval[u]);
        iscale[u] = amax_all[u] != 0u ? is : 0.f;
    }
#pragma unroll
    for (int u = 0; u < QB; u++) {
        const float dd = __fdiv_rn(1.f, iscale[u]);
        d[u] = amax_all[u] != 0u ? dd : 0.f;
    }
#pragma unroll
    for (int u = 0; u < QB; u++) {
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int qi = __float2int_rn(__fmul_rn(iscale[u], v[u][i]));
            qi = min(127, qi);
            if (i < 4) w0 |= ((uint32_t) (qi & 0xff)) << (8 * i);
            else       w1 |= ((uint32_t) (qi & 0xff)) << (8 * (i - 4));
        }
        const int a0s = __dp4a((int) w0, 0x01010101, 0), a1s = __dp4a((int) w1, 0x01010101, 0);   // word sums
        int s32 = a0s + a1s;
        s32 += __shfl_xor_sync(0xffffffffu, s32, 1);
        s32 += __shfl_xor_sync(0xffffffffu, s32, 2);             // sum of sub-block lane/4 (32 values)
        if (u < nvalid) {
            const int bb = b + u * bstride;
            *reinterpret_cast<uint2 *>(A.q + (size_t) bb * 256 + lane * 8) = make_uint2(w0, w1);
            *reinterpret_cast<int2 *>(A.as + (size_t) bb * 64 + lane * 2) = make_int2(-32 * a0s, -32 * a1s);
            if ((lane & 3) == 0) A.bp[(size_t) bb * 8 + (lane >> 2)] = s32;
            if (lane == 0) A.dx[bb] = d[u];
        }
    }
}
// One warp quantizes 256 consecutive values = 8 Q8_0 blocks of 32 — AVX path of quantize_row_q8_0
// (cpp/ggml/src/ggml-quants.c:936-1000): d = amax/127 kept as fp16, id = 127/amax, q = round_half_even(x*id).
__device__ __noinline__ void q80_blocks_warp(float4 va, float4 vb, int lane, int b256, ActSmem A) {
    const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
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
    *reinterpret_cast<uint2 *>(A.q + (size_t) b256 * 256 + lane * 8) = make_uint2(w0, w1);
    if ((lane & 3) == 0) A.dx[b256 * 8 + (lane >> 2)] = __half2float(__float2half_rn(d));
}

__device__ __forceinline__ void load8(const float * p, float (&v)[8]) {
    const float4 a0 = *reinterpret_cast<const float4 *>(p), a1 = *reinterpret_cast<const float4 *>(p + 4);
    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
}
__device__ __forceinline__ void ldg8(const float * p, float (&v)[8]) {
    const float4 a0 = __ldg(reinterpret_cast<const float4 *>(p)), a1 = __ldg(reinterpret_cast<const float4 *>(p + 4));
    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
}

// Block-cooperative prologue: optional RMSNorm(+weight) then activation quantization into shared memory.
// RMSNorm = ggml_compute_forward_rms_norm_f32 (double sum of float x*x, scale = 1/sqrtf(mean+eps), y = x*scale)
// followed by the separate ggml_mul with the norm weight (cpp/src/llama.cpp:7928-7958).
// Warp w owns blocks w, w+W, ... ; PRO_U of them are handled per pass with all their loads issued up front. With a
// norm the whole vector must fit ONE pass (k/256 <= PRO_U * W; the host checks), so x is read once and a single
// load latency sits in front of the mat-vec; `pre_w` then holds the warp's norm weights, fetched by the caller
// before griddepcontrol.wait (they are constants). `after_loads` runs once, right after the first pass' x loads are
// in flight (the caller tops up its weight ring there: the x loads must not queue behind those bulk copies).
static constexpr int PRO_U = 4;
// inv_k = 1.0 / k when k is a power of two (the multiplication is then exactly the division and the ~100-instruction
// double-precision division routine is never fetched), else 0
__device__ __forceinline__ float rms_scale(double tot, int k, double inv_k, float eps) {
    const float mean = inv_k != 0.0 ? (float) (tot * inv_k) : (float) (tot / (double) k);
    return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
}
// barrier of the first `nwarp` warps of the CTA (a mat-vec phase of the persistent kernel may run on fewer warps than the
// CTA has; barrier 15 is reserved for it, 1..8 are the K-split groups')
__device__ __forceinline__ void sync_warps(int nwarp) { asm volatile("bar.sync 15, %0;" :: "r"(nwarp * 32) : "memory"); }
// WHOLE_CTA: every warp of the CTA takes part (the stand-alone kernels): the barrier is a plain __syncthreads()
template <bool TR, bool WHOLE_CTA, typename F>
__device__ __forceinline__ void prologue_quantize(const float * __restrict__ x, bool norm, float eps, int k, double inv_k, int act_q8_0,
                                                  const ActSmem & A, double * red, const float (&pre_w)[PRO_U][8], F after_loads,
                                                  int nwarp_arg, unsigned long long * tr = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (the stand-alone kernels derive the count from blockDim as the measured round-1 code did: ptxas unrolls the cross-warp
    //  reduction below differently when the count is a run-time argument)
    const int nwarp = WHOLE_CTA ? (int) (blockDim.x >> 5) : nwarp_arg;
    const int n256 = k / 256;
    bool first = true;
    for (int b0 = warp; b0 < n256 || first; b0 += PRO_U * nwarp) {
        float v[PRO_U][8];
#pragma unroll
        for (int u = 0; u < PRO_U; u++) {
            const int b = b0 + u * nwarp;
            if (b < n256) load8(x + b * 256 + lane * 8, v[u]);
        }
        if (first) after_loads();
        float scale = 1.f;
        if (norm && first) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < PRO_U; u++) {
                const int b = b0 + u * nwarp;
                if (b < n256) {
#pragma unroll
                    for (int i = 0; i < 8; i++) s += (double) __fmul_rn(v[u][i], v[u][i]);
                }
            }
            s = warp_sum_d(s);
            if (lane == 0) red[warp] = s;
            trace_mark<TR>(tr, 5);
            if (WHOLE_CTA) __syncthreads(); else sync_warps(nwarp);
            double tot = 0.0;
            for (int w = 0; w < nwarp; w++) tot += red[w];
            scale = rms_scale(tot, k, inv_k, eps);
            trace_mark<TR>(tr, 6);
        }
        first = false;
        if (norm) {
#pragma unroll
            for (int u = 0; u < PRO_U; u++) {
                if (b0 + u * nwarp < n256) {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[u][i] = __fmul_rn(__fmul_rn(v[u][i], scale), pre_w[u][i]);
                }
            }
        }
        if (act_q8_0) {
#pragma unroll
            for (int u = 0; u < PRO_U; u++) {
                const int b = b0 + u * nwarp;
                if (b < n256) q80_blocks_warp(make_float4(v[u][0], v[u][1], v[u][2], v[u][3]), make_float4(v[u][4], v[u][5], v[u][6], v[u][7]), lane, b, A);
            }
        } else {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < PRO_U; u++)
                if (b0 + u * nwarp < n256)
                    q8k_blocks_warp<1>(make_float4(v[u][0], v[u][1], v[u][2], v[u][3]), make_float4(v[u][4], v[u][5], v[u][6], v[u][7]),
                                       z4, z4, z4, z4, z4, z4, lane, b0 + u * nwarp, nwarp, 1, A);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// cp.async helpers (attention kernels)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc) {
    const unsigned sa = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// dp4a with an UNSIGNED first operand (bytes 0..255) and a signed second one
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// byte i (compile-time) of a word, zero-extended: one PRMT
template <int I> __device__ __forceinline__ int byte_of(uint32_t w) { return (int) __byte_perm(w, 0u, 0x4440u | (unsigned) I); }

// ------------------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy wrappers (PTX; sm_90+ instructions, compiled for sm_100a)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
#ifndef B200_HANG_DEBUG
#define B200_HANG_DEBUG 0    // 1: mbarrier waits report and trap after ~2 s instead of spinning for ever (debug builds)
#endif
#if B200_HANG_DEBUG
__device__ int g_dbg_phase[256];     // phase each CTA is in (token kernel), for the timeout report
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if B200_HANG_DEBUG
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned spins = 0;
#endif
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#if B200_HANG_DEBUG
        if (!ok && (++spins & 255u) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) {
                if ((threadIdx.x & 31) == 0)
                    printf("booster_b200: mbarrier timeout: CTA %d warp %d bar 0x%x parity %u phase %d\n", (int) blockIdx.x, (int) threadIdx.x >> 5, bar, parity,
                           g_dbg_phase[blockIdx.x & 255]);
                __trap();
            }
        }
#endif
    } while (!ok);
}
// global -> shared bulk copy (TMA engine, no registers, no per-lane addressing); completes `bytes` on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// the same copy with an L2 evict-first hint: a tile is read exactly once per token, so after the copy its lines are the
// preferred victims and the look-ahead data that has NOT been consumed yet stays resident
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// L2 look-ahead (PfRange): lane 0 of warp `wid` (of `n_w` issuing warps in the whole grid) prefetches chunks wid,
// wid + n_w, ... of every range — the ranges arrive front first, in the order their consumer walks them
__device__ __forceinline__ void l2_prefetch_bulk(const void * p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void issue_l2_lookahead(const PfRange (&pf)[PF_RANGES], int wid, int n_w, int pos) {
#pragma unroll 1
    for (int r = 0; r < PF_RANGES; r++) {
        uint32_t bytes = pf[r].bytes;
        if (bytes == 0) continue;
        if (pf[r].per_pos) bytes = min(bytes, (uint32_t) (pos + 1) * pf[r].per_pos);
        bytes &= ~15u;
#pragma unroll 1
        for (uint32_t off = (uint32_t) wid * PF_CHUNK; off < bytes; off += (uint32_t) n_w * PF_CHUNK)
            l2_prefetch_bulk(pf[r].p + off, min(PF_CHUNK, bytes - off));
    }
}

// ------------------------------------------------------------------------------------------------------------
// One lane = one ROW. A tile is block b of 32 rows. Every warp streams ITS tiles through a private ring in shared
// memory: lane 0 issues one TMA bulk copy per tile (`stages - 1` tiles ahead), completion lands on an mbarrier, and
// each lane then reads its own row's fields with conflict-free 16-byte LDS. Per tile a lane computes the block's
// exact integers — s[m] = the m-th int32 lane of the reference's AVX2 `sumi`, p[l] = the l-th lane of its mins
// product — and advances its own fp32 chains (12 for Q4_K, 9 for Q5_K, 8 for Q6_K / Q8_0):
//   acc[m] = fma(d_b, (float) s[m], acc[m])                      (ggml-quants.c:6972, 7556, 8216, 5376)
//   Q4_K: acc[8+l] = fma(dmin_b, (float) p[l], acc[8+l])  (:6934)      Q5_K: acc[8] += dmin_b * (float) sum p  (:7516)
// Tile layouts (one contiguous chunk each; `c` = 16-byte chunk index, `l` = lane):
//   Q4_K 4608 B: qs uint4[8][32] | sd uint4[32] = {scales[12], d, dmin} of the lane's block
//   Q5_K 5632 B: the same | qh uint4[2][32]
//   Q6_K 6720 B: ql uint4[8][32] | scales int8[16] as uint4[32] | qh uint4[4][32] | d u16[32]
//   Q8_0 1088 B: qs uint4[2][32] | d u16[32]                        (blocks of 32 weights)
// ------------------------------------------------------------------------------------------------------------
struct BlockInts { int s[8]; int p[4]; float d, dmin; };

__device__ __forceinline__ uint4    lds_u4(const uint8_t * p)  { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint32_t lds_u32(const uint8_t * p) { return *reinterpret_cast<const uint32_t *>(p); }
// 16 activation bytes of the current block, same address in every lane (broadcast)
__device__ __forceinline__ int4 act16(const int8_t * ab, int slot) { return *reinterpret_cast<const int4 *>(ab + slot * 16); }

template <int TYPE> struct TypeTag { static constexpr int value = TYPE; };
template <int TYPE> __device__ __forceinline__ constexpr int n_chains() { return TYPE == T_Q4_K ? 12 : TYPE == T_Q5_K ? 9 : 8; }

// `sl` = the tile in shared memory + lane*16. The loops over the four 64-weight groups are deliberately NOT unrolled
// (instruction footprint): their accumulators are indexed by (h, wi) only.
template <bool Q5>
__device__ __forceinline__ void ints_q45k(const uint8_t * sl, const int8_t * ab, const int * bp, float yd, BlockInts & o) {
    const uint4 sd = lds_u4(sl + 4096);
    const uint32_t s0 = sd.x, s1 = sd.y, s2 = sd.z, dmw = sd.w;
    // the 8 scale bytes and 8 min bytes (get_scale_min_k4 packing, cpp/ggml/src/ggml-quants.c:1891-1898)
    const uint32_t sc_a = s0 & 0x3f3f3f3fu, m_a = s1 & 0x3f3f3f3fu;
    const uint32_t sc_b = (s2 & 0x0f0f0f0fu) | ((s0 >> 2) & 0x30303030u);
    const uint32_t m_b  = ((s2 >> 4) & 0x0f0f0f0fu) | ((s1 >> 2) & 0x30303030u);
    uint4 qh[2];
    if (Q5) { qh[0] = lds_u4(sl + 4608); qh[1] = lds_u4(sl + 4608 + 512); }
    // s = s_lo + (s_hi16 >> 4): the high nibbles are multiplied IN PLACE (mask 0xf0 = 16 x value, unsigned dp4a);
    // their scaled sum is an exact multiple of 16, so the shift is exact
    int s_lo[8], s_hi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { s_lo[i] = 0; s_hi[i] = 0; }
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        const uint32_t scw = (j & 2) ? sc_b : sc_a;
        const int sh = (j & 1) * 16;
        const int sc_lo = (int) ((scw >> sh) & 0xffu), sc_hi = (int) ((scw >> (sh + 8)) & 0xffu);
        const uint8_t * q = sl + j * 1024;
        const int8_t * aj = ab + j * 64;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 w = lds_u4(q + h * 512);
            const int4 alo = *reinterpret_cast<const int4 *>(aj + h * 16), ahi = *reinterpret_cast<const int4 *>(aj + 32 + h * 16);
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                const uint32_t W = word_of(w, wi);
                if (!Q5) {
                    s_lo[4 * h + wi] += sc_lo * __dp4a((int) (W & 0x0f0f0f0fu), word_of(alo, wi), 0);
                    s_hi[4 * h + wi] += sc_hi * dp4a_us(W & 0xf0f0f0f0u, word_of(ahi, wi), 0);
                } else {
                    // qh bit 2j -> +16 on the low-nibble weight, bit 2j+1 -> +16 on the high-nibble one
                    const uint32_t H = word_of(qh[h], wi) >> (2 * j);
                    const uint32_t lo = (W & 0x0f0f0f0fu) | ((H & 0x01010101u) << 4);
                    const uint32_t hi = ((W >> 4) & 0x0f0f0f0fu) | (((H >> 1) & 0x01010101u) << 4);
                    s_lo[4 * h + wi] += sc_lo * __dp4a((int) lo, word_of(alo, wi), 0);
                    s_hi[4 * h + wi] += 16 * sc_hi * __dp4a((int) hi, word_of(ahi, wi), 0);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) o.s[i] = s_lo[i] + (s_hi[i] >> 4);
    // mins: lane l of the reference's _mm_madd_epi16(mins, q8s) = m[2l]*bsum[2l] + m[2l+1]*bsum[2l+1]
    const int4 bp0 = *reinterpret_cast<const int4 *>(bp), bp1 = *reinterpret_cast<const int4 *>(bp + 4);
    o.p[0] = byte_of<0>(m_a) * bp0.x + byte_of<1>(m_a) * bp0.y;
    o.p[1] = byte_of<2>(m_a) * bp0.z + byte_of<3>(m_a) * bp0.w;
    o.p[2] = byte_of<0>(m_b) * bp1.x + byte_of<1>(m_b) * bp1.y;
    o.p[3] = byte_of<2>(m_b) * bp1.z + byte_of<3>(m_b) * bp1.w;
    const __half2 dmh = *reinterpret_cast<const __half2 *>(&dmw);
    o.d    = __fmul_rn(yd, __low2float(dmh));            // y[i].d * fp16(x[i].d)
    o.dmin = __fmul_rn(-yd, __high2float(dmh));          // -y[i].d * fp16(x[i].dmin)
}

// Q6_K: the weight is q - 32 with q in 0..63; sum (q - 32) a = dp4a_u8(q, a) - 32 * sum(a), and the second term
// (per 32-bit word of the activation block) comes pre-computed from the prologue as the accumulator input of dp4a.
// The loop over (half n, 16-byte column mq) is not unrolled; the four groups g inside it are.
__device__ __forceinline__ void ints_q6k(const uint8_t * sl, const uint8_t * tile, int lane, const int8_t * ab, const int * asb,
                                         float yd, BlockInts & o) {
    const uint4 scv = lds_u4(sl + 4096);
#pragma unroll
    for (int i = 0; i < 8; i++) o.s[i] = 0;
    // layout of a super-block: dequantize_row_q6_K (cpp/ggml/src/ggml-quants.c:2970-3000); half n, group g of 32
    // weights: ql byte 64n + 32(g&1) + l, nibble g>>1; qh byte 32n + l, bits 2g..2g+1; scale 8n + 2g + l/16
#pragma unroll 1
    for (int n = 0; n < 2; n++) {
        // scales 8n .. 8n+7 = two words; scale of group g, column mq = byte 2g + mq of the pair
        const uint32_t scw0 = n ? scv.z : scv.x, scw1 = n ? scv.w : scv.y;
#pragma unroll 1
        for (int mq = 0; mq < 2; mq++) {
            const uint4 qh = lds_u4(sl + 4608 + (2 * n + mq) * 512);
            const uint32_t sw0 = scw0 >> (8 * mq), sw1 = scw1 >> (8 * mq);
            int part[4] = {0, 0, 0, 0};
#pragma unroll
            for (int gl = 0; gl < 2; gl++) {                      // one ql chunk serves groups gl (low nibble) and gl+2 (high)
                const uint4 ql = lds_u4(sl + (4 * n + 2 * gl + mq) * 512);
#pragma unroll
                for (int gh = 0; gh < 2; gh++) {
                    const int g = gl + 2 * gh;
                    const int si = 8 * n + 2 * g + mq;
                    const int sc = (int) (int8_t) (((g >> 1) ? sw1 : sw0) >> (16 * (g & 1)));
                    const int4 a = act16(ab, si);
                    const int4 c32 = *reinterpret_cast<const int4 *>(asb + 4 * si);
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        const uint32_t QL = word_of(ql, wi), QH = word_of(qh, wi);
                        const uint32_t lo = gh ? ((QL >> 4) & 0x0f0f0f0fu) : (QL & 0x0f0f0f0fu);
                        const uint32_t h2 = g == 0 ? (QH << 4) : g == 1 ? (QH << 2) : g == 2 ? QH : (QH >> 2);
                        const uint32_t q  = lo | (h2 & 0x30303030u);
                        part[wi] += sc * dp4a_us(q, word_of(a, wi), word_of(c32, wi));
                    }
                }
            }
            // o.s[4 mq + wi] += part[wi] without a dynamically indexed register array
#pragma unroll
            for (int wi = 0; wi < 4; wi++) { if (mq == 0) o.s[wi] += part[wi]; else o.s[4 + wi] += part[wi]; }
        }
    }
    const float dw = __half2float(*reinterpret_cast<const __half *>(tile + 6656 + lane * 2));
    o.d = __fmul_rn(yd, dw);                              // y[i].d * fp16(x[i].d)
    o.dmin = 0.f;
}

__device__ __forceinline__ void ints_q80(const uint8_t * sl, const uint8_t * tile, int lane, const int8_t * ab, float yd, BlockInts & o) {
    const uint4 w0 = lds_u4(sl), w1 = lds_u4(sl + 512);
    const int4 a0 = act16(ab, 0), a1 = act16(ab, 1);
#pragma unroll
    for (int wi = 0; wi < 4; wi++) {
        o.s[wi]     = __dp4a((int) word_of(w0, wi), word_of(a0, wi), 0);
        o.s[4 + wi] = __dp4a((int) word_of(w1, wi), word_of(a1, wi), 0);
    }
    const float dw = __half2float(*reinterpret_cast<const __half *>(tile + 1024 + lane * 2));
    o.d = __fmul_rn(dw, yd);                              // fp16(x.d) * fp16(y.d)
    o.dmin = 0.f;
}

template <int TYPE>
__device__ __forceinline__ void tile_ints(const uint8_t * tile, int lane, int t, const ActSmem & A, BlockInts & bi) {
    const uint8_t * sl = tile + lane * 16;
    if (TYPE == T_Q4_K)      ints_q45k<false>(sl, A.q + (size_t) t * 256, A.bp + (size_t) t * 8, A.dx[t], bi);
    else if (TYPE == T_Q5_K) ints_q45k<true>(sl, A.q + (size_t) t * 256, A.bp + (size_t) t * 8, A.dx[t], bi);
    else if (TYPE == T_Q6_K) ints_q6k(sl, tile, lane, A.q + (size_t) t * 256, A.as + (size_t) t * 64, A.dx[t], bi);
    else                     ints_q80(sl, tile, lane, A.q + (size_t) t * 32, A.dx[t], bi);
}

// hsum_float_8 (cpp/ggml/src/ggml-quants.c:47-53) + the type's tail, from the row's chain values
template <int TYPE>
__device__ __forceinline__ float finish_row(const float * c) {
    const float r0 = __fadd_rn(c[4], c[0]), r1 = __fadd_rn(c[5], c[1]), r2 = __fadd_rn(c[6], c[2]), r3 = __fadd_rn(c[7], c[3]);
    const float h = __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
    if (TYPE == T_Q4_K) return __fadd_rn(h, __fadd_rn(__fadd_rn(c[8], c[10]), __fadd_rn(c[9], c[11])));   // + acc_m
    if (TYPE == T_Q5_K) return __fadd_rn(h, c[8]);                                                       // + summs
    return h;
}

// Chain phase of one round (G > 1): this lane advances its CPW chain slots over the round's G tiles in order. Slot i
// owns chain c = w + i*G of the lane's row: value at voff[i], multiplier (d for c < 8, dmin above) at moff[i] of a
// tile's published record. Slots beyond the type's chain count point at valid words and accumulate garbage that is
// never read, so the loop has no branches; q5tail marks the slot holding Q5_K's summs chain (add of a product, not
// an fma: ggml-quants.c:7516).
template <int CPW, bool Q5>
__device__ __forceinline__ void chain_round(const float * rb, int G, int tile_stride, const int (&voff)[6], const int (&moff)[6],
                                            int q5slot, float (&acc)[12]) {
    // (unrolled by 4 when a lane has one or two chains: the LDS of four tiles are in flight together and only the FMAs
    //  are serial)
    constexpr int UNR = CPW <= 2 ? B200_CHAIN_UNROLL : 1;
#pragma unroll UNR
    for (int ww = 0; ww < G; ww++) {
        const float * tb = rb + (size_t) ww * tile_stride;
#pragma unroll
        for (int i = 0; i < CPW; i++) {
            const float v = tb[voff[i]], m = tb[moff[i]];
            if (Q5) {
                const float f = __fmaf_rn(m, v, acc[i]), g = __fadd_rn(acc[i], __fmul_rn(m, v));
                acc[i] = i == q5slot ? g : f;
            } else {
                acc[i] = __fmaf_rn(m, v, acc[i]);
            }
        }
    }
}

// a work unit resolved against the launch's segments: type, first tile in HBM, first output row
struct UnitDesc { int type; int row0; uint32_t bytes; const uint8_t * tiles; };
__device__ __forceinline__ UnitDesc describe_unit(const MatvecArgs & a, int unit) {
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

// ------------------------------------------------------------------------------------------------------------
// The fused quantized mat-vec: [RMSNorm] + activation quant (prologue) -> exact W.x -> epilogue.
// A work unit is 32 rows x all TU blocks of K. G warps (a divisor of TU chosen by the host so that matrices with few
// rows still occupy every SM) share a unit: warp w of the group owns tiles t = w, w+G, ... of every unit of its
// group. G == 1: the row's fp32 chains live in the lane's registers. G > 1: the unit is processed in rounds of G
// tiles; every warp computes the INTEGERS of its tile (order-free, exact) and publishes them as floats, one named
// barrier, then the fp32 chains — which must run in block order — are advanced over the round's G tiles by the lanes
// of the whole group in parallel: lane (row) of warp w owns chains c = w, w+G, ... of its row (32 x 12 independent
// chains per unit). Nothing is serialised across warps any more; a chain step costs one LDS + one FMA.
// The code of a unit is specialised on the unit's block type.
// ------------------------------------------------------------------------------------------------------------
// the row's epilogue (one call per unit; out of line: four block types share one copy)
__device__ __noinline__ void matvec_epilogue(const MatvecArgs & a, float val, int row, int lane, float pre0, float pre1, int cell) {
    const float oth = __shfl_xor_sync(0xffffffffu, val, 1);    // partner row (2i <-> 2i+1)
    if (a.epi == EPI_STORE) {
        a.out[row] = val;
    } else if (a.epi == EPI_RESID) {
        a.out[row] = __fadd_rn(val, pre0);                     // ggml_add(cur, inpSA / ffn_inp): llama.cpp:8865, 8901
    } else if (a.epi == EPI_SILU) {
        // rows (2r, 2r+1) = (gate r, up r): silu(gate) * up, cpp/src/llama.cpp:7960-8085
        if ((lane & 1) == 0) a.out[row >> 1] = __fmul_rn(silu_exact(val), oth);
    } else {  // EPI_QKV
        const float v0 = (lane & 1) ? oth : val, v1 = (lane & 1) ? val : oth;   // (x0, x1) of this row's RoPE pair
        if (row < a.n_q + a.n_k) {
            // RoPE NORM mode on the pair (x0, x1): cpp/ggml/src/ggml.c:14121-14135
            const float y = (lane & 1) ? __fadd_rn(__fmul_rn(v0, pre1), __fmul_rn(v1, pre0))
                                       : __fsub_rn(__fmul_rn(v0, pre0), __fmul_rn(v1, pre1));
            if (row < a.n_q) a.q_out[row] = y;
            else a.k_cache[(size_t) cell * a.kv_dim + (row - a.n_q)] = __float2half_rn(y);   // K post-RoPE as f16: llama.cpp:7849-7853
        } else {
            a.v_cache[(size_t) cell * a.kv_dim + (row - a.n_q - a.n_k)] = __float2half_rn(val);
        }
    }
}

// per-thread state of one mat-vec between its two halves: mv_begin (everything that does not depend on the input vector:
// shared-memory carve-up, barrier init, work split, the first weight tiles requested, norm weights) and mv_run (prologue on
// x, the tile loop, epilogues). k_matvec runs them around griddepcontrol.wait; the persistent per-token kernel runs
// mv_begin of the NEXT phase before it waits at the grid barrier, so HBM streams through the barrier.
struct MvState {
    uint8_t * ring; uint8_t * act_base;
    ActSmem A;
    float * cbuf; float * fin;
    uint32_t full0, edge_in, edge_out, ring_u32;
    int grp, w, bar_id, bar_threads;
    int group_global, n_groups, my_units, n_items;
    int pi, pj, pk, ps;
    UnitDesc pd;
    uint64_t pol;
    float ww[PRO_U][8];        // the warp's norm weights (constants: requested before the input exists)
};

// producer side: item pi = (unit pj, tile w + G*pk) goes to ring slot pi % S
__device__ __forceinline__ void mv_issue_next(const MatvecArgs & a, MvState & s, int lane) {
    if (s.pi >= s.n_items) return;
    if (lane == 0) {
        mbar_expect_tx(s.full0 + 8 * s.ps, s.pd.bytes);
        bulk_g2s_hint(s.ring_u32 + (uint32_t) s.ps * a.stage_bytes, s.pd.tiles + (size_t) (s.w + a.group * s.pk) * s.pd.bytes, s.pd.bytes,
                      s.full0 + 8 * s.ps, s.pol);
    }
    s.pi++; s.ps = s.ps + 1 == a.stages ? 0 : s.ps + 1;
    if (++s.pk == a.kpw) {
        s.pk = 0; s.pj++;
        if (s.pi < s.n_items) s.pd = describe_unit(a, s.group_global + s.pj * s.n_groups);
    }
}

// end of a mat-vec inside the persistent kernel: the warp's mbarriers are invalidated before their shared memory is
// re-used by the next phase (initialising a location that still holds a valid mbarrier object is undefined — measured: a
// later phase with the same layout then waits for ever on its first tile). Call after a barrier of the mat-vec's warps.
__device__ __forceinline__ void mv_end(const MatvecArgs & a, const MvState & s) {
    if ((threadIdx.x & 31) == 0) {
#pragma unroll 1
        for (int st = 0; st < a.stages; st++) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(s.full0 + 8 * st) : "memory");
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(s.edge_in) : "memory");
    }
}

// cta / n_cta: this CTA's index and the number of CTAs that share the mat-vec; n_prefill: tiles per warp requested now
template <bool TR>
__device__ __forceinline__ void mv_begin(const MatvecArgs & a, uint8_t * smem_raw, int cta, int n_cta, int n_prefill, MvState & s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = a.warps;
    const int G = a.group, S = a.stages;
    // shared memory: ring (128-byte aligned slots) | activations | hand-off buffers | mbarriers
    s.ring = smem_raw + (size_t) warp * S * a.stage_bytes;
    s.act_base = smem_raw + (size_t) W * S * a.stage_bytes;
    const size_t act_bytes = a.act_bytes;
    s.A = act_smem_carve(s.act_base, a.k, a.act_q8_0);
    s.grp = (int) (((uint32_t) warp * a.grp_magic) >> 16); s.w = warp - s.grp * G;   // group inside the CTA, warp inside the group
    // chain region: exchange: cbuf[parity][warp of the CTA][value][lane] | fin[group][chain][lane]; hand-off: fin only
    s.cbuf = reinterpret_cast<float *>(s.act_base + act_bytes);
    s.fin  = s.cbuf + a.exch_words + (size_t) s.grp * HANDOFF_WORDS;
    uint64_t * bars = reinterpret_cast<uint64_t *>(s.act_base + act_bytes + a.chain_bytes);
    s.full0    = smem_u32(bars + warp * S);                            // my ring slots' "tile landed" barriers
    s.edge_in  = smem_u32(bars + W * S + warp);                        // hand-off: "chain state for me is published"
    s.edge_out = smem_u32(bars + W * S + s.grp * G + (s.w + 1 == G ? 0 : s.w + 1));
    s.ring_u32 = smem_u32(s.ring);
    s.bar_id = 1 + s.grp; s.bar_threads = G * 32;                      // the group's named barrier
    if (lane == 0) {                                           // each warp: its own ring barriers and its edge barrier
#pragma unroll 1
        for (int st = 0; st < S; st++) mbar_init(s.full0 + 8 * st, 1);
        mbar_init(s.edge_in, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();                                              // (the edge barriers of OTHER warps are first used after the
                                                               //  prologue's barrier)
    // group-major mapping: unit u -> CTA u % n_cta, group (u / n_cta) % groups_per_cta, so that mat-vecs with few
    // units spread over every SM
    s.group_global = s.grp * n_cta + cta;
    s.n_groups = n_cta * a.groups_per_cta;
    s.my_units = s.group_global < a.n_units ? (a.n_units - s.group_global + s.n_groups - 1) / s.n_groups : 0;
    s.n_items = s.my_units * a.kpw;
    s.pi = 0; s.pj = 0; s.pk = 0; s.ps = 0;
    s.pd = describe_unit(a, s.my_units > 0 ? s.group_global : 0);
    // weight tiles are read exactly once per token: the copies carry the L2 evict-first hint, so streaming 4.6 GB of
    // them per token does not push the activations, the K/V rows and the kernels' own code out of L2
    s.pol = l2_policy_evict_first();
    // weights do not depend on x: (part of) the ring is filled before the input exists
    const int prefill = min(n_prefill, S - 1);
#pragma unroll 1
    for (int st = 0; st < prefill; st++) mv_issue_next(a, s, lane);
    // norm weights are constants too: the warp's blocks (the host guarantees k/256 <= PRO_U * W when there is a norm)
#pragma unroll
    for (int u = 0; u < PRO_U; u++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s.ww[u][i] = 0.f;
    }
    if (a.norm_w != nullptr) {
#pragma unroll
        for (int u = 0; u < PRO_U; u++) {
            const int b = warp + u * W;
            if (b < a.k / 256) ldg8(a.norm_w + b * 256 + lane * 8, s.ww[u]);
        }
    }
}

template <bool TR>
__device__ __forceinline__ void mv_run(const MatvecArgs & a, double * red_smem, int n_prefilled, MvState & s) {
    const int EPI = a.epi;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = a.warps;
    const int G = a.group, TU = a.tiles_unit, S = a.stages, KPW = a.kpw;
    const int NV = a.nv, mode = a.chain_mode;
    const int grp = s.grp, w = s.w;
    uint8_t * const ring = s.ring;
    const ActSmem A = s.A;
    float * const cbuf = s.cbuf; float * const fin = s.fin;
    const uint32_t full0 = s.full0, edge_in = s.edge_in, edge_out = s.edge_out;
    const int bar_id = s.bar_id, bar_threads = s.bar_threads;
    const int group_global = s.group_global, n_groups = s.n_groups, my_units = s.my_units;
    const bool norm = a.norm_w != nullptr;
    const int pos = EPI == EPI_QKV ? a.st->pos : 0;            // in flight during the prologue
    const int cell = EPI == EPI_QKV ? a.st->cell : 0;
    prologue_quantize<TR, false>(a.x, norm, a.eps, a.k, a.inv_k, a.act_q8_0, A, red_smem, s.ww,
                      [&]() {
#pragma unroll 1
                          for (int st = n_prefilled; st < S - 1; st++) mv_issue_next(a, s, lane);
                      }, W, a.trace);
    sync_warps(W);                                             // activations + every warp's barrier inits are visible
    trace_mark<TR>(a.trace, 3);

    // ---- consumer side
    int cs = 0, cpar = 0, rnd = 0, n_in = 0;
    for (int j = 0; j < my_units; j++) {
        const UnitDesc cd = describe_unit(a, group_global + j * n_groups);
        auto unit_body = [&](auto tag) {
            constexpr int TYPE = decltype(tag)::value;
            constexpr int NCH = n_chains<TYPE>();
            constexpr int CPW = 6;                             // chain slots per lane when G > 1: ceil(12 / 2)
            float acc[12];                                     // G == 1: the row's chains; G > 1: [0, CPW) = chains w, w+G, ...
#pragma unroll
            for (int c = 0; c < 12; c++) acc[c] = 0.f;
            // G > 1: slot i of this lane = chain w + i*G (value and multiplier offsets inside a tile's record)
            int voff[6], moff[6], q5slot = -1;
            const int cpw = G >= 12 ? 1 : G >= 6 ? 2 : G >= 4 ? 3 : 6;
#pragma unroll
            for (int i = 0; i < CPW; i++) {
                const int c = w + i * G;
                voff[i] = (c < NCH ? c : 0) * 32;
                moff[i] = (c < 8 || c >= NCH ? NCH : NCH + 1) * 32;
                if (TYPE == T_Q5_K && c == 8) q5slot = i;
            }
            for (int k = 0; k < KPW; k++) {
                const int t = w + G * k;
                mv_issue_next(a, s, lane);                     // keeps S-1 tiles in flight (slot of the previous item is free)
                // epilogue operands of a unit that completes with this tile: fetched now, used after the chain
                float pre0 = 0.f, pre1 = 0.f;
                if (t == TU - 1) {
                    const int prow = cd.row0 + lane;
                    if (EPI == EPI_RESID) pre0 = a.resid[prow];
                    if (EPI == EPI_QKV && prow < a.n_q + a.n_k) {
                        const float2 cs2 = a.rope[(size_t) pos * (a.head_dim / 2) + (((prow & ~1) & (a.head_dim - 1)) >> 1)];
                        pre0 = cs2.x; pre1 = cs2.y;
                    }
                }
                mbar_wait(full0 + 8 * cs, (uint32_t) cpar);
                if (j == 0 && k == 0) trace_mark<TR>(a.trace, 7);  // first tile landed
                BlockInts bi;
                tile_ints<TYPE>(ring + (size_t) cs * a.stage_bytes, lane, t, A, bi);
                __syncwarp();                                  // every lane is done reading the slot before it is refilled
                if (j == 0 && k == 0) trace_mark<TR>(a.trace, 8);  // first tile's integers done
                cs = cs + 1 == S ? 0 : cs + 1; if (cs == 0) cpar ^= 1;
                float val;
                if (mode != CHAIN_EXCHANGE) {
                    if (mode == CHAIN_HANDOFF) {
                        // the previous step of the group's tile sequence is done and its state published
                        if (j > 0 || t > 0) { mbar_wait(edge_in, (uint32_t) (n_in & 1)); n_in++; }
                        if (t > 0) {
#pragma unroll
                            for (int c = 0; c < NCH; c++) acc[c] = fin[c * 32 + lane];
                        } else {
#pragma unroll
                            for (int c = 0; c < NCH; c++) acc[c] = 0.f;
                        }
                    }
                    // ---- chain step in registers, strictly in block order
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[c] = __fmaf_rn(bi.d, (float) bi.s[c], acc[c]);
                    if (TYPE == T_Q4_K) {
#pragma unroll
                        for (int l = 0; l < 4; l++) acc[8 + l] = __fmaf_rn(bi.dmin, (float) bi.p[l], acc[8 + l]);
                    } else if (TYPE == T_Q5_K) {
                        acc[8] = __fadd_rn(acc[8], __fmul_rn(bi.dmin, (float) (bi.p[0] + bi.p[1] + bi.p[2] + bi.p[3])));
                    }
                    if (mode == CHAIN_HANDOFF && !(j == my_units - 1 && t == TU - 1)) {
                        if (t != TU - 1) {
#pragma unroll
                            for (int c = 0; c < NCH; c++) fin[c * 32 + lane] = acc[c];
                        }
                        mbar_arrive(edge_out);                 // release: my lane's stores above are visible to the waiter
                    }
                    if (t != TU - 1) continue;
                    val = finish_row<TYPE>(acc);
                } else {
                    // ---- publish this tile's integers (as exact floats) and scales
                    float * cb = cbuf + ((size_t) ((rnd & 1) * W + warp) * NV) * 32 + lane;
#pragma unroll
                    for (int c = 0; c < 8; c++) cb[c * 32] = (float) bi.s[c];
                    if (TYPE == T_Q4_K) {
#pragma unroll
                        for (int l = 0; l < 4; l++) cb[(8 + l) * 32] = (float) bi.p[l];
                    } else if (TYPE == T_Q5_K) {
                        cb[8 * 32] = (float) (bi.p[0] + bi.p[1] + bi.p[2] + bi.p[3]);
                    }
                    cb[NCH * 32] = bi.d;
                    if (TYPE == T_Q4_K || TYPE == T_Q5_K) cb[(NCH + 1) * 32] = bi.dmin;
                    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_threads) : "memory");
                    // ---- chain phase of the round: tiles t0 .. t0+G-1 in order, this lane's chains c = w + i*G
                    const float * rb = cbuf + ((size_t) ((rnd & 1) * W + grp * G) * NV) * 32 + lane;
                    rnd++;
                    if (j == 0 && k == 0) trace_mark<TR>(a.trace, 9);   // first round: every warp's integers published
                    switch (cpw) {
                        case 1:  chain_round<1, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                        case 2:  chain_round<2, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                        case 3:  chain_round<3, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                        default: chain_round<6, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                    }
                    if (j == 0 && k == 0) trace_mark<TR>(a.trace, 11);  // first round's chains advanced
                    if (k != KPW - 1) continue;
                    // ---- unit complete: gather the row's chains
#pragma unroll
                    for (int i = 0; i < CPW; i++) {
                        const int c = w + i * G;
                        if (c < NCH) fin[c * 32 + lane] = acc[i];
                    }
                    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_threads) : "memory");
                    if (w != G - 1) continue;                  // the warp that fetched the epilogue operands finishes the rows
                    float cv[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; c++) cv[c] = fin[c * 32 + lane];
                    val = finish_row<TYPE>(cv);
                }

                matvec_epilogue(a, val, cd.row0 + lane, lane, pre0, pre1, cell);
            }
        };
        switch (cd.type) {
            case T_Q4_K: unit_body(TypeTag<T_Q4_K>{}); break;
            case T_Q5_K: unit_body(TypeTag<T_Q5_K>{}); break;
            case T_Q6_K: unit_body(TypeTag<T_Q6_K>{}); break;
            default:     unit_body(TypeTag<T_Q8_0>{}); break;
        }
    }
    trace_mark<TR>(a.trace, 10);                                   // warp 0 out of work
}

// The stand-alone kernel keeps its own copy of the body (the one the round-1 measurements were made with): expressed through
// mv_begin / mv_run it compiles to the same register count but measures 2.4 % slower per token on the same box
// (548 -> 535 tok/s, gpurun_out/r2k) — ptxas schedules the monolithic body better than the state-struct version.
template <bool TR>
__global__ void __launch_bounds__(MV_MAX_WARPS * 32, 1) k_matvec(const __grid_constant__ MatvecArgs a) {
    const int EPI = a.epi;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ double red_smem[MV_MAX_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int G = a.group, TU = a.tiles_unit, S = a.stages, KPW = a.kpw;
    // shared memory: ring (128-byte aligned slots) | activations | hand-off buffers | mbarriers
    uint8_t * ring = smem_raw + (size_t) warp * S * a.stage_bytes;
    uint8_t * act_base = smem_raw + (size_t) W * S * a.stage_bytes;
    const size_t act_bytes = a.act_bytes;
    const ActSmem A = act_smem_carve(act_base, a.k, a.act_q8_0);
    const int grp = (int) (((uint32_t) warp * a.grp_magic) >> 16), w = warp - grp * G;   // group inside the CTA, warp inside the group
    const int NV = a.nv;
    const int mode = a.chain_mode;
    // chain region: exchange: cbuf[parity][warp of the CTA][value][lane] | fin[group][chain][lane]; hand-off: fin only
    float * cbuf = reinterpret_cast<float *>(act_base + act_bytes);
    float * fin  = cbuf + a.exch_words + (size_t) grp * HANDOFF_WORDS;
    uint64_t * bars = reinterpret_cast<uint64_t *>(act_base + act_bytes + a.chain_bytes);
    const uint32_t full0   = smem_u32(bars + warp * S);                            // my ring slots' "tile landed" barriers
    const uint32_t edge_in  = smem_u32(bars + W * S + warp);                        // hand-off: "chain state for me is published"
    const uint32_t edge_out = smem_u32(bars + W * S + grp * G + (w + 1 == G ? 0 : w + 1));
    const uint32_t ring_u32 = smem_u32(ring);
    const int bar_id = 1 + grp, bar_threads = G * 32;                              // the group's named barrier

    trace_mark<TR>(a.trace, 0);
    pdl_launch_dependents();                                   // the next kernel may start its own weight prefetch
    if (lane == 0) {                                           // each warp: its own ring barriers and its edge barrier
#pragma unroll 1
        for (int s = 0; s < S; s++) mbar_init(full0 + 8 * s, 1);
        mbar_init(edge_in, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();                                              // (the edge barriers of OTHER warps are first used after the
                                                               //  prologue's __syncthreads)

    // group-major mapping: unit u -> CTA u % grid, group (u / grid) % groups_per_cta, so that launches with few
    // units spread over every SM
    const int groups_per_cta = a.groups_per_cta;
    const int group_global = grp * gridDim.x + blockIdx.x;
    const int n_groups = gridDim.x * groups_per_cta;
    const int my_units = group_global < a.n_units ? (a.n_units - group_global + n_groups - 1) / n_groups : 0;
    const int n_items = my_units * KPW;

    // ---- producer side: item pi = (unit pj, tile w + G*pk) goes to ring slot pi % S
    int pi = 0, pj = 0, pk = 0, ps = 0;
    UnitDesc pd = describe_unit(a, my_units > 0 ? group_global : 0);
    // weight tiles are read exactly once per token: the copies carry the L2 evict-first hint, so streaming 4.6 GB of
    // them per token does not push the activations, the K/V rows and the kernels' own code out of L2
    const uint64_t pol = l2_policy_evict_first();
    auto issue_next = [&]() {
        if (pi >= n_items) return;
        if (lane == 0) {
            mbar_expect_tx(full0 + 8 * ps, pd.bytes);
            bulk_g2s_hint(ring_u32 + (uint32_t) ps * a.stage_bytes, pd.tiles + (size_t) (w + G * pk) * pd.bytes, pd.bytes, full0 + 8 * ps, pol);
        }
        pi++; ps = ps + 1 == S ? 0 : ps + 1;
        if (++pk == KPW) {
            pk = 0; pj++;
            if (pi < n_items) pd = describe_unit(a, group_global + pj * n_groups);
        }
    };
    // weights do not depend on x: part of the ring is filled before the wait, the rest once the x loads are in flight
    const int prefill = min(a.prefill, S - 1);
#pragma unroll 1
    for (int s = 0; s < prefill; s++) issue_next();
    // norm weights are constants too: the warp's blocks (the host guarantees k/256 <= PRO_U * W when there is a norm)
    const bool norm = a.norm_w != nullptr;
    float ww[PRO_U][8] = {};
    if (norm) {
#pragma unroll
        for (int u = 0; u < PRO_U; u++) {
            const int b = warp + u * W;
            if (b < a.k / 256) ldg8(a.norm_w + b * 256 + lane * 8, ww[u]);
        }
    }

    // L2 look-ahead for the kernels that follow (after this CTA's own first tiles are requested). DecodeState is
    // written by the previous TOKEN's last kernel, so pos may be read before the wait.
#if B200_LOOKAHEAD
    if (lane == 0 && a.pf[0].bytes) issue_l2_lookahead(a.pf, warp * gridDim.x + blockIdx.x, W * gridDim.x, a.st ? a.st->pos : 0);
#endif

    trace_mark<TR>(a.trace, 1);
    pdl_wait();                                                // x (and everything else the previous kernels wrote) is visible
    trace_mark<TR>(a.trace, 2);
    // (pos, ...) and (cell, ...) as two 16-byte loads of the DecodeState, in flight during the prologue
    int pos = 0, cell = 0;
    if (EPI == EPI_QKV) {
        const int4 s0 = *reinterpret_cast<const int4 *>(a.st), s1 = *(reinterpret_cast<const int4 *>(a.st) + 1);
        pos = s0.y; cell = s1.x;
    }
    prologue_quantize<TR, true>(a.x, norm, a.eps, a.k, a.inv_k, a.act_q8_0, A, red_smem, ww,
                      [&]() {
#pragma unroll 1
                          for (int s = prefill; s < S - 1; s++) issue_next();
                      }, W, a.trace);
    __syncthreads();                                           // activations + every warp's barrier inits are visible
    trace_mark<TR>(a.trace, 3);

    // ---- consumer side
    int cs = 0, cpar = 0, rnd = 0, n_in = 0;
    for (int j = 0; j < my_units; j++) {
        const UnitDesc cd = describe_unit(a, group_global + j * n_groups);
        auto unit_body = [&](auto tag) {
            constexpr int TYPE = decltype(tag)::value;
            constexpr int NCH = n_chains<TYPE>();
            constexpr int CPW = 6;                             // chain slots per lane when G > 1: ceil(12 / 2)
            float acc[12];                                     // G == 1: the row's chains; G > 1: [0, CPW) = chains w, w+G, ...
#pragma unroll
            for (int c = 0; c < 12; c++) acc[c] = 0.f;
            // G > 1: slot i of this lane = chain w + i*G (value and multiplier offsets inside a tile's record)
            int voff[6], moff[6], q5slot = -1;
            const int cpw = G >= 12 ? 1 : G >= 6 ? 2 : G >= 4 ? 3 : 6;
#pragma unroll
            for (int i = 0; i < CPW; i++) {
                const int c = w + i * G;
                voff[i] = (c < NCH ? c : 0) * 32;
                moff[i] = (c < 8 || c >= NCH ? NCH : NCH + 1) * 32;
                if (TYPE == T_Q5_K && c == 8) q5slot = i;
            }
            for (int k = 0; k < KPW; k++) {
                const int t = w + G * k;
                issue_next();                                  // keeps S-1 tiles in flight (slot of the previous item is free)
                // epilogue operands of a unit that completes with this tile: fetched now, used after the chain
                float pre0 = 0.f, pre1 = 0.f;
                if (t == TU - 1) {
                    const int prow = cd.row0 + lane;
                    if (EPI == EPI_RESID) pre0 = a.resid[prow];
                    if (EPI == EPI_QKV && prow < a.n_q + a.n_k) {
                        const float2 cs2 = a.rope[(size_t) pos * (a.head_dim / 2) + (((prow & ~1) & (a.head_dim - 1)) >> 1)];
                        pre0 = cs2.x; pre1 = cs2.y;
                    }
                }
                mbar_wait(full0 + 8 * cs, (uint32_t) cpar);
                if (j == 0 && k == 0) trace_mark<TR>(a.trace, 7);  // first tile landed
                BlockInts bi;
                tile_ints<TYPE>(ring + (size_t) cs * a.stage_bytes, lane, t, A, bi);
                __syncwarp();                                  // every lane is done reading the slot before it is refilled
                if (j == 0 && k == 0) trace_mark<TR>(a.trace, 8);  // first tile's integers done
                cs = cs + 1 == S ? 0 : cs + 1; if (cs == 0) cpar ^= 1;
                float val;
                if (mode != CHAIN_EXCHANGE) {
                    if (mode == CHAIN_HANDOFF) {
                        // the previous step of the group's tile sequence is done and its state published
                        if (j > 0 || t > 0) { mbar_wait(edge_in, (uint32_t) (n_in & 1)); n_in++; }
                        if (t > 0) {
#pragma unroll
                            for (int c = 0; c < NCH; c++) acc[c] = fin[c * 32 + lane];
                        } else {
#pragma unroll
                            for (int c = 0; c < NCH; c++) acc[c] = 0.f;
                        }
                    }
                    // ---- chain step in registers, strictly in block order
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[c] = __fmaf_rn(bi.d, (float) bi.s[c], acc[c]);
                    if (TYPE == T_Q4_K) {
#pragma unroll
                        for (int l = 0; l < 4; l++) acc[8 + l] = __fmaf_rn(bi.dmin, (float) bi.p[l], acc[8 + l]);
                    } else if (TYPE == T_Q5_K) {
                        acc[8] = __fadd_rn(acc[8], __fmul_rn(bi.dmin, (float) (bi.p[0] + bi.p[1] + bi.p[2] + bi.p[3])));
                    }
                    if (mode == CHAIN_HANDOFF && !(j == my_units - 1 && t == TU - 1)) {
                        if (t != TU - 1) {
#pragma unroll
                            for (int c = 0; c < NCH; c++) fin[c * 32 + lane] = acc[c];
                        }
                        mbar_arrive(edge_out);                 // release: my lane's stores above are visible to the waiter
                    }
                    if (t != TU - 1) continue;
                    val = finish_row<TYPE>(acc);
                } else {
                    // ---- publish this tile's integers (as exact floats) and scales
                    float * cb = cbuf + ((size_t) ((rnd & 1) * W + warp) * NV) * 32 + lane;
#pragma unroll
                    for (int c = 0; c < 8; c++) cb[c * 32] = (float) bi.s[c];
                    if (TYPE == T_Q4_K) {
#pragma unroll
                        for (int l = 0; l < 4; l++) cb[(8 + l) * 32] = (float) bi.p[l];
                    } else if (TYPE == T_Q5_K) {
                        cb[8 * 32] = (float) (bi.p[0] + bi.p[1] + bi.p[2] + bi.p[3]);
                    }
                    cb[NCH * 32] = bi.d;
                    if (TYPE == T_Q4_K || TYPE == T_Q5_K) cb[(NCH + 1) * 32] = bi.dmin;
                    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_threads) : "memory");
                    // ---- chain phase of the round: tiles t0 .. t0+G-1 in order, this lane's chains c = w + i*G
                    const float * rb = cbuf + ((size_t) ((rnd & 1) * W + grp * G) * NV) * 32 + lane;
                    rnd++;
                    if (j == 0 && k == 0) trace_mark<TR>(a.trace, 9);   // first round: every warp's integers published
                    switch (cpw) {
                        case 1:  chain_round<1, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                        case 2:  chain_round<2, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                        case 3:  chain_round<3, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                        default: chain_round<6, TYPE == T_Q5_K>(rb, G, NV * 32, voff, moff, q5slot, acc); break;
                    }
                    if (j == 0 && k == 0) trace_mark<TR>(a.trace, 11);  // first round's chains advanced
                    if (k != KPW - 1) continue;
                    // ---- unit complete: gather the row's chains
#pragma unroll
                    for (int i = 0; i < CPW; i++) {
                        const int c = w + i * G;
                        if (c < NCH) fin[c * 32 + lane] = acc[i];
                    }
                    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_threads) : "memory");
                    if (w != G - 1) continue;                  // the warp that fetched the epilogue operands finishes the rows
                    float cv[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; c++) cv[c] = fin[c * 32 + lane];
                    val = finish_row<TYPE>(cv);
                }

                matvec_epilogue(a, val, cd.row0 + lane, lane, pre0, pre1, cell);
            }
        };
        switch (cd.type) {
            case T_Q4_K: unit_body(TypeTag<T_Q4_K>{}); break;
            case T_Q5_K: unit_body(TypeTag<T_Q5_K>{}); break;
            case T_Q6_K: unit_body(TypeTag<T_Q6_K>{}); break;
            default:     unit_body(TypeTag<T_Q8_0>{}); break;
        }
    }
    trace_mark<TR>(a.trace, 10);                                   // warp 0 out of work
    if (TR) { if (a.trace != nullptr) { __syncthreads(); trace_mark<TR>(a.trace, 4); } }
}

// ------------------------------------------------------------------------------------------------------------
// stand-alone activation quantization (operator-level tests; same device functions as the prologue).
// out = ggml block structs: block_q8_K {float d; int8 qs[256]; int16 bsums[16]} (292 B), block_q8_0 {half d;
// int8 qs[32]} (34 B)  (cpp/ggml/src/ggml-common.h:311-315, 186-190). bsums are re-derived from the quants.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_quantize_export(const float * __restrict__ x, int k, int act_q8_0, uint8_t * __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ double red_smem[MV_MAX_WARPS];
    const ActSmem A = act_smem_carve(smem_raw, k, act_q8_0);
    const float no_w[PRO_U][8] = {};
    prologue_quantize<false, true>(x, false, 0.f, k, 0.0, act_q8_0, A, red_smem, no_w, []() {}, (int) (blockDim.x >> 5));
    __syncthreads();
    if (!act_q8_0) {
        const int nb = k / 256;
        for (int b = 0; b < nb; b++) {
            uint8_t * o = out + (size_t) b * 292;
            if (threadIdx.x == 0) *reinterpret_cast<float *>(o) = A.dx[b];
            for (int e = threadIdx.x; e < 256; e += blockDim.x) o[4 + e] = (uint8_t) A.q[(size_t) b * 256 + e];
            for (int i = threadIdx.x; i < 16; i += blockDim.x) {
                int s = 0;
                for (int e = 16 * i; e < 16 * i + 16; e++) s += A.q[(size_t) b * 256 + e];
                o[260 + 2 * i] = (uint8_t) (s & 0xff); o[261 + 2 * i] = (uint8_t) ((s >> 8) & 0xff);
            }
        }
    } else {
        const int nb32 = k / 32;
        for (int b = threadIdx.x; b < nb32; b += blockDim.x) {
            uint8_t * o = out + (size_t) b * 34;
            *reinterpret_cast<__half *>(o) = __float2half_rn(A.dx[b]);     // dx is an exact f16 value; 34*b is even
            for (int e = 0; e < 32; e++) o[2 + e] = (uint8_t) A.q[(size_t) b * 32 + e];
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
// decode attention, default (non-flash) route (cpp/src/llama.cpp:8248-8297), exact order, three launches:
//   k_attn_scores : S[h][t] = (K[t] . q[h]) * scale for t <= pos, -inf for the padded tail (n_kv is padded to 32,
//                   cpp/src/llama.cpp:14698; the mask adds -inf there). 4 lanes per key row; lane c4 owns the
//                   tinyBLAS chains 4c4..4c4+3 (element 16s+c of step s) and the _mm512_reduce_add_ps tree is
//                   completed with two shuffles. round_q: ggml_vec_dot_f16 order (4 accumulators x 16 lanes).
//   k_attn_softmax: per head: max, p = ggml_v_expf(s - max), per-16 partial sums (_mm512_reduce_add_ps),
//                   double sum, p *= (float)(1/sum)   (cpp/ggml/src/ggml.c:13682-13778, 2619-2640)
//   k_attn_pv     : out[h][d] = sum_t V[t][d] p[t] as tinyBLAS computes it: chain c = t mod 16 over t, then the
//                   reduce tree. One thread per (kv head, d, c), all GQA heads of the group share each V load.
// ------------------------------------------------------------------------------------------------------------
static constexpr int ATT_THREADS = 256;
static constexpr int ATT_TILE    = 64;       // key positions per scores CTA
static constexpr int ATT_MAX_GQA = 8;
static constexpr int PV_DIMS     = 16;       // dims per P.V CTA (x 16 chains = 256 threads; 16 dims = one 32-byte sector)
static constexpr int PV_CHUNK    = 1024;     // positions of p staged in shared memory at a time

struct AttnArgs {
    const float * q;          // [n_head][hd] post-RoPE
    const __half * k_cache;   // [n_ctx][kv_dim]
    const __half * v_cache;
    float * S;                // [n_head][s_stride] scores, then probabilities
    int s_stride;             // >= n_ctx padded to 32
    float * out;              // [n_head*hd]  (kqv_merged_cont)
    int n_head, n_head_kv, head_dim, kv_dim;
    float scale;
    const DecodeState * st;
    int n_kv_override;        // >0: use instead of st->pos+1 (operator-level test)
    int round_q_override;     // operator-level test of the batch>1 arithmetic
    int p_chunk;              // positions of p staged in shared memory by k_attn_pv (multiple of PV_BATCH)
    PfRange pf[PF_RANGES];    // L2 look-ahead issued by k_attn_scores
    PfRange pf2[PF_RANGES];   // ... and by k_attn_softmax_pv
    const int32_t * cell_pos; // [n_ctx] position held by each KV cell (-1: empty); read only when st->managed
    // prompt batches: blockIdx.z = token of the batch; token z uses q + z*zq, S + z*zs, out + z*zq and attends
    // n_kv_override + z cells (its own position included). All zero for a single token.
    int zq, zs;
    unsigned long long * trace;
};
__device__ __forceinline__ int attn_n_kv(const AttnArgs & a) { return a.n_kv_override > 0 ? a.n_kv_override + (int) blockIdx.z : a.st->n_kv; }
// the cell this token's K / V rows were just written to (by the QKV kernel; every other cell is older)
__device__ __forceinline__ int attn_cur_cell(const AttnArgs & a, int n_kv) { return a.n_kv_override > 0 ? n_kv - 1 : a.st->cell; }
// the KQ mask (cpp/src/llama.cpp:14132-14200): cell t is attended iff it holds a position of the sequence that is <= the
// token's. Without a context shift cell t holds position t and the test is t < n_kv.
__device__ __forceinline__ bool attn_cell_visible(const DecodeState * st, const int32_t * cell_pos, int t, int n_kv) {
    if (t >= n_kv) return false;
    if (st == nullptr || !st->managed) return true;
    const int p = cell_pos[t];
    return p >= 0 && p <= st->pos;
}

template <int GQA, bool TR>
__global__ void __launch_bounds__(ATT_THREADS) k_attn_scores(const AttnArgs a) {
    constexpr int HD = 128;
    __shared__ __align__(16) float qs[GQA][HD];
    const int g = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x;
    trace_mark<TR>(a.trace, 0);
    // (cell, n_kv, managed) in ONE 16-byte load: DecodeState is written by the previous TOKEN's last kernel
    int n_kv, cur, managed = 0;
    if (a.n_kv_override > 0) { n_kv = a.n_kv_override + (int) blockIdx.z; cur = n_kv - 1; }
    else { const int4 s1 = *(reinterpret_cast<const int4 *>(a.st) + 1); cur = s1.x; n_kv = s1.y; managed = s1.z; }
    const int n_pad = (n_kv + 31) / 32 * 32;
    const int t = tile * ATT_TILE + (tid >> 2), c4 = tid & 3;
    // K rows of EARLIER positions were written by earlier tokens: their loads go out before griddepcontrol.wait
    // (8 x 8-byte loads per lane in flight); only the row of the current position has to wait for the QKV kernel
    uint2 kv[8];
#pragma unroll
    for (int s = 0; s < 8; s++) kv[s] = make_uint2(0u, 0u);
    const uint2 * kr = reinterpret_cast<const uint2 *>(a.k_cache + (size_t) t * a.kv_dim + g * HD + 4 * c4);
    if (t < n_kv && t != cur) {
#pragma unroll
        for (int s = 0; s < 8; s++) kv[s] = __ldg(kr + s * 4);            // 4 halfs at element 16s + 4c4
    }
#if B200_LOOKAHEAD
    if ((tid & 31) == 0 && a.pf[0].bytes)
        issue_l2_lookahead(a.pf, (tid >> 5) * (gridDim.x * gridDim.y) + tile * gridDim.x + g, (ATT_THREADS / 32) * gridDim.x * gridDim.y, n_kv - 1);
#endif
    pdl_wait();                                               // q and this token's K row come from the QKV kernel
    pdl_launch_dependents();                                  // AFTER the wait: the next kernel may touch K/V/q before ITS wait
    trace_mark<TR>(a.trace, 1);
    if (tile * ATT_TILE >= n_pad) return;
    int round_q = a.round_q_override, pos = 0;
    if (a.st) { const int4 s0 = *reinterpret_cast<const int4 *>(a.st); pos = s0.y; round_q = s0.z; }
    if (t == cur) {
#pragma unroll
        for (int s = 0; s < 8; s++) kv[s] = kr[s * 4];                    // plain loads: written by the previous kernel
    }
    // the KQ mask (cpp/src/llama.cpp:14132-14200): without a context shift cell t holds position t and the test is t < n_kv
    bool visible = t < n_kv;
    if (managed && visible) { const int p = a.cell_pos[t]; visible = p >= 0 && p <= pos; }
    for (int i = tid; i < GQA * HD; i += ATT_THREADS) {
        float v = a.q[(size_t) blockIdx.z * a.zq + (size_t) (g * GQA) * HD + i];
        if (round_q) v = __half2float(__float2half_rn(v));    // src1 converted to the vec_dot_type F16 (ggml.c:12345-12371)
        (&qs[0][0])[i] = v;
    }
    float kf[8][4];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[s].x));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&kv[s].y));
        kf[s][0] = f0.x; kf[s][1] = f0.y; kf[s][2] = f1.x; kf[s][3] = f1.y;
    }
    __syncthreads();
    if (t < n_pad) {                                          // n_pad % 32 == 0: whole warps take this branch together
#pragma unroll 1
        for (int h = 0; h < GQA; h++) {                       // rolled: instruction footprint
            float ch[4];
            if (!round_q) {
                // tinyBLAS<16>: lane c: acc = fma(k[16s+c], q[16s+c], acc), s = 0..7
#pragma unroll
                for (int e = 0; e < 4; e++) ch[e] = 0.f;
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    const float4 qv = *reinterpret_cast<const float4 *>(&qs[h][16 * s + 4 * c4]);
                    ch[0] = __fmaf_rn(kf[s][0], qv.x, ch[0]); ch[1] = __fmaf_rn(kf[s][1], qv.y, ch[1]);
                    ch[2] = __fmaf_rn(kf[s][2], qv.z, ch[2]); ch[3] = __fmaf_rn(kf[s][3], qv.w, ch[3]);
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
            if (c4 == 0) a.S[(size_t) blockIdx.z * a.zs + (size_t) (g * GQA + h) * a.s_stride + t] = visible ? __fmul_rn(res, a.scale) : -INFINITY;
        }
    }
    trace_mark<TR>(a.trace, 2);
}

__global__ void __launch_bounds__(256) k_attn_softmax(const AttnArgs a) {
    __shared__ float redf[8];
    __shared__ double redd[8];
    const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_pad = (attn_n_kv(a) + 31) / 32 * 32;
    float * S = a.S + (size_t) h * a.s_stride;
    float mx = -INFINITY;
    for (int i = tid; i < n_pad; i += 256) mx = fmaxf(mx, S[i]);
    mx = warp_max(mx);
    if (lane == 0) redf[warp] = mx;
    __syncthreads();
    mx = redf[0];
#pragma unroll
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, redf[w]);
    // n_pad is a multiple of 32 and the stride is 256, so every warp iteration is either fully in or fully out of
    // range and half-warps coincide with the reference's 16-wide vectors. The per-vector sums (float,
    // _mm512_reduce_add_ps order) are accumulated in double; double addition of <= 512 such floats is exact up to
    // one final rounding far below float resolution, so the tree order below equals the reference's serial order
    // after the cast to float.
    double part = 0.0;
    for (int i = tid; i < n_pad; i += 256) {
        const float p = v_expf(__fsub_rn(S[i], mx));
        S[i] = p;
        const float gs = reduce_add16_shfl(p);
        if ((lane & 15) == 0) part += (double) gs;
    }
    part = warp_sum_d(part);
    if (lane == 0) redd[warp] = part;
    __syncthreads();
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) sum += redd[w];
    const float inv = (float) (1.0 / sum);
    for (int i = tid; i < n_pad; i += 256) S[i] = __fmul_rn(S[i], inv);
}

static constexpr int PV_BATCH = 256;   // granularity of the p/V chunk

template <int GQA>
__global__ void __launch_bounds__(PV_DIMS * 16) k_attn_pv(const AttnArgs a) {
    constexpr int HD = 128;
    constexpr int NT = PV_DIMS * 16;
    extern __shared__ __align__(16) uint8_t pv_dyn[];              // [GQA][p_chunk] f32 | [p_chunk][16] f16
    __shared__ float red[GQA][16][PV_DIMS + 1];
    const int PCH = a.p_chunk;
    float * ps = reinterpret_cast<float *>(pv_dyn);
    __half (*vs)[PV_DIMS] = reinterpret_cast<__half (*)[PV_DIMS]>(pv_dyn + (size_t) GQA * PCH * 4);
    // thread = (chain c, dim dl); a CTA owns 16 dims (one 32-byte sector per V row) of one KV head
    const int g = blockIdx.x, c = threadIdx.x / PV_DIMS, dl = threadIdx.x % PV_DIMS;
    const int n_kv = attn_n_kv(a);
    const int n_pad = (n_kv + 31) / 32 * 32;
    const __half * vbase = a.v_cache + g * HD + blockIdx.y * PV_DIMS;
    float acc[GQA];
#pragma unroll
    for (int h = 0; h < GQA; h++) acc[h] = 0.f;
    for (int t0 = 0; t0 < n_pad; t0 += PCH) {
        const int len = min(PCH, n_pad - t0), rows = min(len, n_kv - t0);     // rows beyond n_kv are never read (p = 0)
        if (t0) __syncthreads();
        // every byte of the chunk in flight at once: V rows (2 x 16 B each) and the GQA probability rows
        for (int i = threadIdx.x; i < rows * 2; i += NT)
            cp_async16(&vs[i >> 1][(i & 1) * 8], vbase + (size_t) (t0 + (i >> 1)) * a.kv_dim + (i & 1) * 8);
        for (int i = threadIdx.x; i < GQA * (len / 4); i += NT) {
            const int h = i / (len / 4), j = i - h * (len / 4);
            cp_async16(ps + (size_t) h * PCH + 4 * j, a.S + (size_t) (g * GQA + h) * a.s_stride + t0 + 4 * j);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        const int steps = len / 16;
#pragma unroll 8
        for (int s = 0; s < steps; s++) {
            const int tt = 16 * s + c;
            // slots at or beyond n_kv have p == 0 exactly; their V bytes were not loaded and must not be multiplied
            const float v = t0 + tt < n_kv ? __half2float(vs[tt][dl]) : 0.f;
#pragma unroll
            for (int h = 0; h < GQA; h++) acc[h] = __fmaf_rn(v, ps[(size_t) h * PCH + tt], acc[h]);
        }
    }
#pragma unroll
    for (int h = 0; h < GQA; h++) red[h][c][dl] = acc[h];
    __syncthreads();
    for (int i = threadIdx.x; i < GQA * PV_DIMS; i += NT) {
        const int h = i / PV_DIMS, dd = i % PV_DIMS;
        // _mm512_reduce_add_ps over the 16 chains
        float t3[8], t6[4];
#pragma unroll
        for (int j = 0; j < 8; j++) t3[j] = __fadd_rn(red[h][8 + j][dd], red[h][j][dd]);
#pragma unroll
        for (int j = 0; j < 4; j++) t6[j] = __fadd_rn(t3[4 + j], t3[j]);
        a.out[(size_t) (g * GQA + h) * HD + blockIdx.y * PV_DIMS + dd] =
            __fadd_rn(__fadd_rn(t6[0], t6[2]), __fadd_rn(t6[1], t6[3]));     // kqv_merged_cont layout: [n_head*hd]
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_attn_softmax_pv: soft_max_ext + P.V of one layer in ONE launch, no inter-CTA communication. A CTA owns
// (KV head g, a slice of PVS_DIMS output dims) and ALL positions (the 16 tinyBLAS chains of an output element run
// over t in order, so positions cannot be split). Every CTA of a KV head normalises the GQA score rows itself — the
// exp work is repeated HD/PVS_DIMS times across CTAs, which is far cheaper than a grid-wide exchange of max and sum —
// with exactly k_attn_softmax's arithmetic, on a shared-memory copy; then thread (h, c, dl) runs chain c of output
// dim dl of head h. The CTA's V slice is independent of the scores and is put in flight BEFORE griddepcontrol.wait
// (k_attn_scores waits for the QKV kernel before it lets this kernel launch, so K/V/q are already visible).
//   grid (n_head_kv, HD / PVS_DIMS), block GQA * pvs_th(GQA) threads (16 chains x dim groups per head); shared: ps [GQA][n_pad] f32 | vs [v_chunk][8] f16
// ------------------------------------------------------------------------------------------------------------
static constexpr int PVS_DIMS = 8;            // dims per CTA: 8 halfs = one 16-byte cp.async per position
// dims per thread: 1 (128 threads per head) up to GQA 4, 2 (64 threads per head) for GQA 8 — at most 512 threads
__host__ __device__ constexpr int pvs_dpt(int gqa) { return gqa >= 8 ? 2 : 1; }
__host__ __device__ constexpr int pvs_th(int gqa) { return 16 * (PVS_DIMS / pvs_dpt(gqa)); }

template <int GQA, bool TR>
__global__ void __launch_bounds__(GQA * pvs_th(GQA)) k_attn_softmax_pv(const AttnArgs a) {
    constexpr int HD = 128;
    constexpr int DPT = pvs_dpt(GQA);
    constexpr int TH = pvs_th(GQA);                            // threads per head: 16 chains x (8 / DPT) dim groups
    constexpr int NT = GQA * TH;
    constexpr int NW = TH / 32;
    extern __shared__ __align__(16) uint8_t sp_dyn[];
    __shared__ float  redf[GQA][NW];
    __shared__ double redd[GQA][NW];
    __shared__ float  red[GQA][16][PVS_DIMS + 1];
    const int g = blockIdx.x, slice = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int h = tid / TH, ht = tid % TH;                     // head of the group, thread within the head
    const int c = ht / (PVS_DIMS / DPT), dp = ht % (PVS_DIMS / DPT);   // chain, dim group
    const int w = ht >> 5;                                     // warp within the head

    trace_mark<TR>(a.trace, 0);
    pdl_launch_dependents();
    const int n_kv = attn_n_kv(a);                             // DecodeState is written by the previous TOKEN's last kernel
    const int n_pad = (n_kv + 31) / 32 * 32;
    const int VCH = a.p_chunk;                                 // positions of V staged at a time (multiple of 32)
    float * ps = reinterpret_cast<float *>(sp_dyn);            // [GQA][n_pad]
    __half (*vs)[PVS_DIMS] = reinterpret_cast<__half (*)[PVS_DIMS]>(sp_dyn + (size_t) GQA * n_pad * 4);
    const __half * vbase = a.v_cache + g * HD + slice * PVS_DIMS;
    // V rows of one chunk; rows in [n_kv, n_pad) are zero-filled: p == 0 there and the product must be 0, never NaN
    auto stage_v = [&](int t0) {
        const int len = min(VCH, n_pad - t0);
        for (int i = tid; i < len; i += NT) {
            if (t0 + i < n_kv) cp_async16(&vs[i][0], vbase + (size_t) (t0 + i) * a.kv_dim);
            else *reinterpret_cast<uint4 *>(&vs[i][0]) = make_uint4(0u, 0u, 0u, 0u);
        }
        cp_async_commit();
    };
    stage_v(0);
#if B200_LOOKAHEAD
    if (lane == 0 && a.pf2[0].bytes)
        issue_l2_lookahead(a.pf2, (tid >> 5) * (gridDim.x * gridDim.y) + slice * gridDim.x + g, (NT / 32) * gridDim.x * gridDim.y, n_kv - 1);
#endif
    pdl_wait();                                               // the raw scores are complete
    trace_mark<TR>(a.trace, 1);
    float * row = ps + (size_t) h * n_pad;
    {
        const float * Sg = a.S + (size_t) blockIdx.z * a.zs + (size_t) (g * GQA + h) * a.s_stride;
        for (int i = ht; i < n_pad / 4; i += TH) cp_async16(row + 4 * i, Sg + 4 * i);
        cp_async_commit();
        cp_async_wait<0>();                                    // (also completes this thread's V copies)
    }
    asm volatile("bar.sync %0, %1;" :: "r"(1 + h), "r"(TH) : "memory");   // the head's row is in shared memory
    trace_mark<TR>(a.trace, 2);
    // soft_max_ext of the head's row (cpp/ggml/src/ggml.c:13682-13778): a thread owns whole 16-element vectors of the
    // reference's loop, so the _mm512_reduce_add_ps tree is register arithmetic; the per-vector float sums are
    // accumulated in double (order-insensitive here: see k_attn_softmax)
    const int n16 = n_pad / 16;
    float mx = -INFINITY;
    // a thread visits the four float4 of its vector starting at (ht >> 1) & 3: the 8 lanes of a quarter-warp then hit 8
    // different 16-byte bank groups. The reduce tree pairs float4 f with f+2 and adds the two halves — both
    // commutative — so the rotation leaves every bit of the result unchanged.
    const int rot = (ht >> 1) & 3;
    for (int gi = ht; gi < n16; gi += TH) {
        const float4 * r4 = reinterpret_cast<const float4 *>(row + 16 * gi);
#pragma unroll
        for (int q = 0; q < 4; q++) { const float4 v = r4[(q + rot) & 3]; mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w)); }
    }
    mx = warp_max(mx);
    if (lane == 0) redf[h][w] = mx;
    asm volatile("bar.sync %0, %1;" :: "r"(1 + h), "r"(TH) : "memory");
    mx = redf[h][0];
#pragma unroll
    for (int j = 1; j < NW; j++) mx = fmaxf(mx, redf[h][j]);
    double part = 0.0;
    for (int gi = ht; gi < n16; gi += TH) {
        float4 * r4 = reinterpret_cast<float4 *>(row + 16 * gi);
        // _mm512_reduce_add_ps of the vector's 16 exponentials: (a[8+i] + a[i]) pairs float4 f with f+2, so the vector is
        // walked as two such pairs in a ROLLED loop (8 inlined ggml_v_expf instead of 16: instruction footprint)
        float4 t[2];
#pragma unroll 1
        for (int pq = 0; pq < 2; pq++) {
            const int f0 = (pq + rot) & 3, f1 = f0 ^ 2;
            float4 lo = r4[f0], hi = r4[f1];
            lo.x = v_expf(__fsub_rn(lo.x, mx)); lo.y = v_expf(__fsub_rn(lo.y, mx)); lo.z = v_expf(__fsub_rn(lo.z, mx)); lo.w = v_expf(__fsub_rn(lo.w, mx));
            hi.x = v_expf(__fsub_rn(hi.x, mx)); hi.y = v_expf(__fsub_rn(hi.y, mx)); hi.z = v_expf(__fsub_rn(hi.z, mx)); hi.w = v_expf(__fsub_rn(hi.w, mx));
            r4[f0] = lo; r4[f1] = hi;
            // fp32 addition is commutative: which of the pair is the "upper" float4 does not change the sum
            const float4 tt = make_float4(__fadd_rn(hi.x, lo.x), __fadd_rn(hi.y, lo.y), __fadd_rn(hi.z, lo.z), __fadd_rn(hi.w, lo.w));
            // the pair {f, f+2} with f even is t[0..3] of the tree, the odd one is t[4..7]
            if (f0 & 1) t[1] = tt; else t[0] = tt;
        }
        const float u0 = __fadd_rn(t[1].x, t[0].x), u1 = __fadd_rn(t[1].y, t[0].y), u2 = __fadd_rn(t[1].z, t[0].z), u3 = __fadd_rn(t[1].w, t[0].w);
        part += (double) __fadd_rn(__fadd_rn(u0, u2), __fadd_rn(u1, u3));
    }
    part = warp_sum_d(part);
    if (lane == 0) redd[h][w] = part;
    asm volatile("bar.sync %0, %1;" :: "r"(1 + h), "r"(TH) : "memory");
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < NW; j++) sum += redd[h][j];
    const float inv = (float) (1.0 / sum);
    for (int gi = ht; gi < n16; gi += TH) {                    // own vectors only
        float4 * r4 = reinterpret_cast<float4 *>(row + 16 * gi);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float4 v = r4[(q + rot) & 3];
            v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv); v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
            r4[(q + rot) & 3] = v;
        }
    }
    __syncthreads();                                           // all rows normalised, every thread's V copies landed
    trace_mark<TR>(a.trace, 3);

    // P.V: chain c of (head h, dims 2dp, 2dp+1): acc = fma(V[t][d], p[t], acc) over t = c, c+16, ...
    float acc0 = 0.f, acc1 = 0.f;
    for (int t0 = 0; t0 < n_pad; t0 += VCH) {
        const int len = min(VCH, n_pad - t0);
        if (t0) {
            __syncthreads();
            stage_v(t0);
            cp_async_wait<0>();
            __syncthreads();
        }
        const int steps = len / 16;
        const float * pr = row + t0 + c;
        if (DPT == 2) {
            const __half2 * vr = reinterpret_cast<const __half2 *>(&vs[c][2 * dp]);
#pragma unroll 8
            for (int s = 0; s < steps; s++) {
                const float2 v = __half22float2(vr[(size_t) s * 16 * (PVS_DIMS / 2)]);
                const float p = pr[16 * s];
                acc0 = __fmaf_rn(v.x, p, acc0);
                acc1 = __fmaf_rn(v.y, p, acc1);
            }
        } else {
            const __half * vr = &vs[c][dp];
#pragma unroll 8
            for (int s = 0; s < steps; s++)
                acc0 = __fmaf_rn(__half2float(vr[(size_t) s * 16 * PVS_DIMS]), pr[16 * s], acc0);
        }
    }
    if (DPT == 2) { red[h][c][2 * dp] = acc0; red[h][c][2 * dp + 1] = acc1; }
    else          red[h][c][dp] = acc0;
    __syncthreads();
    trace_mark<TR>(a.trace, 4);
    if (tid < GQA * PVS_DIMS) {
        const int hh = tid / PVS_DIMS, dd = tid % PVS_DIMS;
        float t3[8], t6[4];                                   // _mm512_reduce_add_ps over the 16 chains
#pragma unroll
        for (int j = 0; j < 8; j++) t3[j] = __fadd_rn(red[hh][8 + j][dd], red[hh][j][dd]);
#pragma unroll
        for (int j = 0; j < 4; j++) t6[j] = __fadd_rn(t3[4 + j], t3[j]);
        a.out[(size_t) blockIdx.z * a.zq + (size_t) (g * GQA + hh) * HD + slice * PVS_DIMS + dd] =
            __fadd_rn(__fadd_rn(t6[0], t6[2]), __fadd_rn(t6[1], t6[3]));     // kqv_merged_cont layout: [n_head*hd]
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
        st->token = idx; st->pos += 1; st->step += 1; st->cell += 1; st->n_kv += 1;   // (device loop: only without a context shift)
    } else {
        out_tokens[0] = idx;
    }
}

__global__ void k_set_state(DecodeState * st, const DecodeState v) { *st = v; }

// advance (pos, step) on stages that do not sample (pipeline stages other than the last)
__global__ void k_advance(DecodeState * st) {
    if (threadIdx.x == 0) { st->pos += 1; st->step += 1; st->cell += 1; st->n_kv += 1; }
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
