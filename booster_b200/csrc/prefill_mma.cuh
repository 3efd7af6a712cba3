// prefill_mma.cuh — prompt batches of K-quant matrices (Q4_K / Q5_K / Q6_K x Q8_K) on the tensor cores, BIT-EXACT with the
// reference's CPU arithmetic (the counterpart of ggml-cuda's mul_mat_q, cpp/ggml/src/ggml-cuda/mmq.cuh).
//
// What has to be reproduced (cpp/ggml/src/ggml-quants.c:6914-6977, 7487-7560, 8145-8220): per output element (row, token) and
// per 256-weight super-block b, the AVX2 code forms EIGHT 32-bit integers — lane m sums scale[g] * w * a over bytes 4m..4m+3 of
// every 32-byte group g — and runs acc[m] = fma(d_b, (float) isum[m], acc[m]) block after block (plus 4 / 1 chains of
// mins x bsums). The fp32 chains are ordered; the integers are not. So:
//
//   * lane m of a super-block is a K = 32 contraction (8 groups x 4 bytes) with the 6-bit / 8-bit sub-block scale folded INTO
//     the weight: w * scale <= 15*63 = 945 (Q4_K), 31*63 = 1953 (Q5_K) — both exact in fp16 (integers <= 2048) — and
//     (q-32) * scale <= 32*128 = 4096 for Q6_K, which is the sum of two fp16-exact integers: the even part ((q & ~1) - 32) * s
//     and the odd bit (q & 1) * s. Activations are int8, exact in fp16. Every product is an integer below 2^19 and every
//     32-term sum is below 2^24, so an fp16 MMA with fp32 accumulation returns EXACTLY (float) isum[m]: no rounding anywhere.
//     (tests/test_ops_gpu.py::test_mul_mat_batch_extremes drives the worst-case magnitudes through the hardware.)
//   * the mins chains are contractions too: bsum (|.| <= 4064) = lo (0..63) + hi (multiple of 64), both fp16-exact, K = 16
//     = 8 groups x (lo, hi); for Q4_K the four lanes l of the reference's _mm_madd_epi16 are four B columns per token.
//   * the chain step acc = fma(d, C, acc) then runs on the MMA's C fragment in registers, super-block after super-block.
//
// mma.sync.m16n8k16 (HMMA) is the default, not tcgen05: the accumulator of every (super-block, lane) must come back to the
// register file for its own fp32 chain step — 8 + 4 FFMA per 256 MACs and output — so the kernel is bound by that CUDA-core
// epilogue, by the 12 live fp32 chains per output element (a 64 x 32 tile fills the register file) and by fragment traffic
// through shared memory, not by tensor throughput (ncu: HMMA pipe 22-26 % busy). The tcgen05 / TMEM variant is
// prefill_umma.cuh: bit-exact too, measured slower (its header says why).
//
// Shape: CTA = 64 rows (two 32-row units of the tiled weight layout) x 32 tokens, 16 warps. Per super-block (K step):
//   TMA bulk copies: the two raw weight tiles + the chunk's activation record block (fp16, MMA order, written by
//     k_quant_batch) into a ring of stages (mbarrier full flags)
//   expand: 512 threads turn the raw tiles into the fp16 A operand (scale folded in) in shared memory, XOR-swizzled for
//     conflict-free ldmatrix; the operand is double-buffered: step b+1 is expanded while step b runs (one barrier per step)
//   mma: warp (m = w & 7, rh = w >> 3) owns lane-slice m of 32 rows x 32 tokens: 16 HMMA + 32 chain FFMA per super-block;
//     the mins tiles are spread over the 16 warps
// After the last super-block the 12 chains of every output meet in shared memory, finish_row() + the layer epilogues of the
// per-token kernel follow (RoPE + KV rows, + residual, SiLU * up).
#pragma once
#include "prefill.cuh"

namespace b200 {

static constexpr int MB_NT = 32;               // tokens per CTA
static constexpr int MB_ROWS = 64;             // rows per CTA (two work units)
static constexpr int MB_WARPS = 16;
static constexpr int MB_MAX_STAGES = 6;
// activation record block of one (32-token chunk, super-block): one contiguous bulk copy
//   main  [8 lanes m][32 tokens][32 fp16]: kk = 4 g + i <- a[32 g + 4 m + i]; 16-byte chunk c of a row sits at c ^ ((t >> 1) & 3)
//   yd    f32[32]                          Q8_K scale of the token's block
//   m4    [32 tokens x 4 lanes l][16 fp16]: kk = 2 g + part (lo, hi of bsum[g]), non-zero for g in {2l, 2l+1}; chunk ^ ((n >> 2) & 1)
//   m5    [32 tokens][16 fp16]: the same, dense (Q5_K sums all eight groups)
static constexpr int MB_REC_MAIN = 8 * MB_NT * 64;
static constexpr int MB_OFF_YD = MB_REC_MAIN;
static constexpr int MB_OFF_M4 = MB_OFF_YD + MB_NT * 4;
static constexpr int MB_OFF_M5 = MB_OFF_M4 + MB_NT * 128;
static constexpr int MB_REC_BYTES = MB_OFF_M5 + MB_NT * 32;
__host__ __device__ __forceinline__ int mb_rec_copy_bytes(bool q4, bool q5) { return q5 ? MB_REC_BYTES : q4 ? MB_OFF_M5 : MB_OFF_M4; }
// A operand in shared memory: part 0 [8 m][64 rows][32 fp16] (| part 1 when the launch has a Q6_K segment) | mins [64][16 fp16] |
// f32 d[64] | f32 dmin[64]
static constexpr int MB_A_PART = 8 * MB_ROWS * 64;
static constexpr int MB_A_MINS = MB_ROWS * 32;
static constexpr int MB_A_SCAL = MB_ROWS * 8;
__host__ __device__ __forceinline__ int mb_a_bytes(bool q6) { return (q6 ? 2 : 1) * MB_A_PART + MB_A_MINS + MB_A_SCAL; }
static constexpr int MB_CH_STRIDE = MB_ROWS + 4;                     // floats per (chain, token) row of the final exchange: +4 keeps the
                                                                     // fragment-order stores (8 rows x 4 token pairs per instruction) conflict-free
static constexpr int MB_CHAIN_BYTES = 12 * MB_NT * MB_CH_STRIDE * 4; // final chain exchange (re-uses the operand / stage memory)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
// C += A(16x16, row) * B(16x8, col), fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D = A * B (no accumulator input: the zero registers are shared, not re-materialised per MMA)
__device__ __forceinline__ void mma_f16_z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 bits_h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t h2_ints(int lo, int hi) { return h2_bits(__halves2half2(__int2half_rn(lo), __int2half_rn(hi))); }
// half2(v, v) for an integer 0 <= v < 1024 on the ALU / FMA pipes (0x6400 | v is 1024 + v in fp16), not the conversion unit
__device__ __forceinline__ __half2 h2_small(int v) {
    const uint32_t b = 0x64006400u | (uint32_t) v | ((uint32_t) v << 16);
    return __hsub2(bits_h2(b), bits_h2(0x64006400u));
}
__device__ __forceinline__ int sbyte_of(uint32_t w, int i) { return (int) (int8_t) (w >> (8 * i)); }

// k_quant_batch, MMA layout: the quantized image of token t (ActSmem, natural order) -> its rows of the chunk's record blocks
__device__ __forceinline__ void mb_write_records(const ActSmem & A, int n256, uint8_t * rec, int t, int lane, int warp, int W) {
    const int chunk = t / MB_NT, j = t % MB_NT;
    for (int b = warp; b < n256; b += W) {
        uint8_t * r = rec + ((size_t) chunk * n256 + b) * MB_REC_BYTES;
        const int8_t * qb = A.q + (size_t) b * 256;
        {   // lane = (m, c): 16-byte chunk c of lane-slice m = groups 2c, 2c+1, bytes 4m..4m+3
            const int m = lane >> 2, c = lane & 3;
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(qb + 64 * c + 4 * m);
            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(qb + 64 * c + 32 + 4 * m);
            uint4 o;
            o.x = h2_ints(sbyte_of(w0, 0), sbyte_of(w0, 1)); o.y = h2_ints(sbyte_of(w0, 2), sbyte_of(w0, 3));
            o.z = h2_ints(sbyte_of(w1, 0), sbyte_of(w1, 1)); o.w = h2_ints(sbyte_of(w1, 2), sbyte_of(w1, 3));
            *reinterpret_cast<uint4 *>(r + m * (MB_NT * 64) + j * 64 + ((c ^ ((j >> 1) & 3)) << 4)) = o;
        }
        if (lane == 0) *reinterpret_cast<float *>(r + MB_OFF_YD + j * 4) = A.dx[b];
        if (lane < 10) {
            const int * bp = A.bp + (size_t) b * 8;
            const int ch = lane & 1;
            uint4 z = make_uint4(0u, 0u, 0u, 0u);
            if (lane < 8) {
                const int l = lane >> 1;
                if (ch == (l >> 1)) {
                    const int s0 = bp[2 * l], s1 = bp[2 * l + 1];
                    const int l0 = s0 & 63, l1 = s1 & 63;
                    const uint32_t p0 = h2_ints(l0, s0 - l0), p1 = h2_ints(l1, s1 - l1);
                    if (l & 1) { z.z = p0; z.w = p1; } else { z.x = p0; z.y = p1; }
                }
                *reinterpret_cast<uint4 *>(r + MB_OFF_M4 + (4 * j + l) * 32 + ((ch ^ (j & 1)) << 4)) = z;
            } else {
                int s[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { s[i] = bp[4 * ch + i]; lo[i] = s[i] & 63; }
                z.x = h2_ints(lo[0], s[0] - lo[0]); z.y = h2_ints(lo[1], s[1] - lo[1]);
                z.z = h2_ints(lo[2], s[2] - lo[2]); z.w = h2_ints(lo[3], s[3] - lo[3]);
                *reinterpret_cast<uint4 *>(r + MB_OFF_M5 + j * 32 + ((ch ^ ((j >> 2) & 1)) << 4)) = z;
            }
        }
    }
}

// raw tiles of the CTA's two units -> fp16 A operand. warp = (unit u, 16-byte chunk c of the quants); lane = row of the unit
template <int TYPE>
__device__ __forceinline__ void mb_expand(const uint8_t * raw, uint32_t raw_stride, uint8_t * As, uint32_t mins_off, int warp, int lane) {
    const int u = warp >> 3, c = warp & 7, r = u * 32 + lane;
    const uint8_t * tile = raw + (size_t) u * raw_stride, * sl = tile + lane * 16;
    const uint32_t sw = (uint32_t) ((r >> 1) & 3);
    float * scal = reinterpret_cast<float *>(As + mins_off + MB_A_MINS);
    if (TYPE == T_Q4_K || TYPE == T_Q5_K) {
        const int j = c >> 1, h = c & 1;                       // 64-weight group j, byte half h: lanes m = 4h .. 4h+3
        const uint4 W4 = lds_u4(sl + c * 512);
        const uint4 sd = lds_u4(sl + 4096);
        const uint32_t sc_a = sd.x & 0x3f3f3f3fu, sc_b = (sd.z & 0x0f0f0f0fu) | ((sd.x >> 2) & 0x30303030u);
        const uint32_t scw = (j & 2) ? sc_b : sc_a;
        const int sh = (j & 1) * 16;
        const int sc_lo = (int) ((scw >> sh) & 0xffu), sc_hi = (int) ((scw >> (sh + 8)) & 0xffu);
        // (1024 + v) * s - 1024 s = v * s: one HFMA2 per pair, exact (v s <= 1953 is an fp16 integer, 1024 s <= 64512 too)
        const __half2 s_lo = h2_small(sc_lo), s_hi = h2_small(sc_hi);
        const __half2 m1024 = bits_h2(0xe400e400u);                     // -1024: -1024 s is an fp16 integer (|.| <= 64512)
        const __half2 o_lo = __hmul2(s_lo, m1024), o_hi = __hmul2(s_hi, m1024);
        uint4 H4 = make_uint4(0u, 0u, 0u, 0u);
        if (TYPE == T_Q5_K) H4 = lds_u4(sl + 4608 + h * 512);
        uint8_t * dst = As + r * 64 + ((((uint32_t) j) ^ sw) << 4);
#pragma unroll
        for (int wi = 0; wi < 4; wi++) {
            const uint32_t Wd = word_of(W4, wi);
            const uint32_t t01 = __byte_perm(Wd, 0u, 0x4140u), t23 = __byte_perm(Wd, 0u, 0x4342u);
            uint32_t l01 = t01 & 0x000f000fu, l23 = t23 & 0x000f000fu, h01 = (t01 >> 4) & 0x000f000fu, h23 = (t23 >> 4) & 0x000f000fu;
            if (TYPE == T_Q5_K) {   // qh bit 2j -> +16 on the low-nibble weight, bit 2j+1 -> +16 on the high-nibble one
                const uint32_t Hs = word_of(H4, wi) >> (2 * j);
                const uint32_t u01 = __byte_perm(Hs, 0u, 0x4140u), u23 = __byte_perm(Hs, 0u, 0x4342u);
                l01 |= (u01 & 0x00010001u) << 4; l23 |= (u23 & 0x00010001u) << 4;
                h01 |= (u01 & 0x00020002u) << 3; h23 |= (u23 & 0x00020002u) << 3;
            }
            uint4 o;
            o.x = h2_bits(__hfma2(bits_h2(l01 | 0x64006400u), s_lo, o_lo));
            o.y = h2_bits(__hfma2(bits_h2(l23 | 0x64006400u), s_lo, o_lo));
            o.z = h2_bits(__hfma2(bits_h2(h01 | 0x64006400u), s_hi, o_hi));
            o.w = h2_bits(__hfma2(bits_h2(h23 | 0x64006400u), s_hi, o_hi));
            *reinterpret_cast<uint4 *>(dst + (4 * h + wi) * (MB_ROWS * 64)) = o;
        }
        if (c < 2) {   // mins operand: kk = 2g + part <- m[g] for both parts; chunk c = mins 4c..4c+3
            const uint32_t m_a = sd.y & 0x3f3f3f3fu, m_b = ((sd.z >> 4) & 0x0f0f0f0fu) | ((sd.y >> 2) & 0x30303030u);
            const uint32_t mw = c ? m_b : m_a;
            uint4 o;
            o.x = h2_bits(h2_small((int) (mw & 0xffu)));
            o.y = h2_bits(h2_small((int) ((mw >> 8) & 0xffu)));
            o.z = h2_bits(h2_small((int) ((mw >> 16) & 0xffu)));
            o.w = h2_bits(h2_small((int) (mw >> 24)));
            *reinterpret_cast<uint4 *>(As + mins_off + r * 32 + ((((uint32_t) c) ^ ((r >> 2) & 1)) << 4)) = o;
            if (c == 0) {
                const __half2 dmh = bits_h2(sd.w);
                scal[r] = __low2float(dmh); scal[MB_ROWS + r] = __high2float(dmh);
            }
        }
    } else {   // Q6_K: chunk c of ql = (half n, group pair gl, 16-byte column mq) -> groups g = gl (low nibble), gl + 2 (high)
        const int n = c >> 2, gl = (c >> 1) & 1, mq = c & 1;
        const uint4 ql = lds_u4(sl + c * 512), qh = lds_u4(sl + 4608 + (2 * n + mq) * 512), scv = lds_u4(sl + 4096);
        const __half2 c1056 = __half2half2(__int2half_rn(1056)), c1024 = __half2half2(__int2half_rn(1024));
#pragma unroll
        for (int gh = 0; gh < 2; gh++) {
            const int g = gl + 2 * gh, G = 4 * n + g, si = 8 * n + 2 * g + mq;
            const uint32_t scw = (si >> 2) == 0 ? scv.x : (si >> 2) == 1 ? scv.y : (si >> 2) == 2 ? scv.z : scv.w;
            const int sc = (int) (int8_t) (scw >> (8 * (si & 3)));
            const __half2 s2 = __half2half2(__int2half_rn(sc));
            uint8_t * dst = As + r * 64 + ((((uint32_t) (G >> 1)) ^ sw) << 4) + (G & 1) * 8;
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                const uint32_t QL = word_of(ql, wi), QH = word_of(qh, wi);
                const uint32_t lo = gh ? ((QL >> 4) & 0x0f0f0f0fu) : (QL & 0x0f0f0f0fu);
                const uint32_t q = lo | (((QH >> (2 * g)) & 0x03030303u) << 4);            // 0..63 per byte
                const uint32_t t01 = __byte_perm(q, 0u, 0x4140u), t23 = __byte_perm(q, 0u, 0x4342u);
                // q - 32 = ((q & ~1) - 32) + (q & 1): both terms times the scale are fp16 integers
                uint2 e, o;
                e.x = h2_bits(__hmul2(__hsub2(bits_h2((t01 & 0x003e003eu) | 0x64006400u), c1056), s2));
                e.y = h2_bits(__hmul2(__hsub2(bits_h2((t23 & 0x003e003eu) | 0x64006400u), c1056), s2));
                o.x = h2_bits(__hmul2(__hsub2(bits_h2((t01 & 0x00010001u) | 0x64006400u), c1024), s2));
                o.y = h2_bits(__hmul2(__hsub2(bits_h2((t23 & 0x00010001u) | 0x64006400u), c1024), s2));
                uint8_t * d = dst + (4 * mq + wi) * (MB_ROWS * 64);
                *reinterpret_cast<uint2 *>(d) = o;
                *reinterpret_cast<uint2 *>(d + MB_A_PART) = e;
            }
        }
        if (c == 0) scal[r] = __half2float(*reinterpret_cast<const __half *>(tile + 6656 + lane * 2));
    }
}

// lane-slice m = warp & 7 of rows 32 rh .. 32 rh + 31 x 32 tokens: exact (float) isum in C, then the chain step
template <int TYPE>
__device__ __forceinline__ void mb_mma_main(const uint8_t * As, const uint8_t * rec, const float * scal, int warp, int lane,
                                            float (&acc)[2][4][4]) {
    constexpr int PARTS = TYPE == T_Q6_K ? 2 : 1;
    const int m = warp & 7, rh = warp >> 3;
    const uint32_t Bs = smem_u32(rec) + m * (MB_NT * 64);
    uint32_t bf[4][4];
    {
        const int tl = lane & 7, kc = lane >> 3;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) ldsm_x4(bf[nt], Bs + (nt * 8 + tl) * 64 + ((kc ^ ((tl >> 1) & 3)) << 4));
    }
    float2 yd[4];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) yd[nt] = *reinterpret_cast<const float2 *>(rec + MB_OFF_YD + (nt * 8 + 2 * (lane & 3)) * 4);
    const uint32_t Am = smem_u32(As) + m * (MB_ROWS * 64);
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const int rl = rh * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kc = lane >> 4;
        const uint32_t arow = Am + rl * 64, sw = (uint32_t) ((rl >> 1) & 3);
        uint32_t af[PARTS][2][4];
#pragma unroll
        for (int p = 0; p < PARTS; p++) {
#pragma unroll
            for (int ks = 0; ks < 2; ks++) ldsm_x4(af[p][ks], arow + p * MB_A_PART + ((((uint32_t) (2 * ks + kc)) ^ sw) << 4));
        }
        const int r0 = rh * 32 + mt * 16 + (lane >> 2);
        const float dw0 = scal[r0], dw1 = scal[r0 + 8];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            float c[4];
            mma_f16_z(c, af[0][0], bf[nt][0], bf[nt][1]);
            mma_f16(c, af[0][1], bf[nt][2], bf[nt][3]);
            if (PARTS == 2) {
                mma_f16(c, af[PARTS - 1][0], bf[nt][0], bf[nt][1]);
                mma_f16(c, af[PARTS - 1][1], bf[nt][2], bf[nt][3]);
            }
            // d = y[i].d * fp16(x[i].d) (ggml-quants.c:6922), acc[m] = fma(d, (float) isum[m], acc[m]) (:6974)
            acc[mt][nt][0] = __fmaf_rn(__fmul_rn(yd[nt].x, dw0), c[0], acc[mt][nt][0]);
            acc[mt][nt][1] = __fmaf_rn(__fmul_rn(yd[nt].y, dw0), c[1], acc[mt][nt][1]);
            acc[mt][nt][2] = __fmaf_rn(__fmul_rn(yd[nt].x, dw1), c[2], acc[mt][nt][2]);
            acc[mt][nt][3] = __fmaf_rn(__fmul_rn(yd[nt].y, dw1), c[3], acc[mt][nt][3]);
        }
    }
}

// the mins operand of rows 16 mt .. 16 mt + 15 (ldmatrix.x4: a0..a3 of one m16k16 tile)
__device__ __forceinline__ void mb_mins_afrag(const uint8_t * As, uint32_t mins_off, int mt, int lane, uint32_t (&af)[4]) {
    const int rl = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kc = lane >> 4;
    ldsm_x4(af, smem_u32(As) + mins_off + rl * 32 + ((((uint32_t) kc) ^ ((rl >> 2) & 1)) << 4));
}
// Q4_K: acc_m lane l = fma(dmin, (float) (m[2l] bsum[2l] + m[2l+1] bsum[2l+1]), .) (ggml-quants.c:6931-6934); columns = (token, l).
// warp = (16-row tile mt = w & 3, 8 tokens tq = w >> 2)
__device__ __forceinline__ void mb_mma_mins4(const uint8_t * As, uint32_t mins_off, const uint8_t * rec, const float * scal, int warp, int lane,
                                             float (&accm)[4][4]) {
    const int mt = warp & 3, tq = warp >> 2;
    uint32_t af[4];
    mb_mins_afrag(As, mins_off, mt, lane, af);
    const int r0 = mt * 16 + (lane >> 2);
    const float dm0 = scal[MB_ROWS + r0], dm1 = scal[MB_ROWS + r0 + 8];
    const float * ydp = reinterpret_cast<const float *>(rec + MB_OFF_YD);
#pragma unroll
    for (int p = 0; p < 2; p++) {
        uint32_t bf[4];
        {
            const int mi = lane >> 3, n = 32 * tq + 8 * (2 * p + (mi >> 1)) + (lane & 7), kc = mi & 1;
            ldsm_x4(bf, smem_u32(rec) + MB_OFF_M4 + n * 32 + ((((uint32_t) kc) ^ ((n >> 2) & 1)) << 4));
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int nt = 2 * p + q;
            float c[4];
            mma_f16_z(c, af, bf[2 * q], bf[2 * q + 1]);
            const float nyd = -ydp[8 * tq + 2 * nt + ((lane & 3) >> 1)];
            const float d0 = __fmul_rn(nyd, dm0), d1 = __fmul_rn(nyd, dm1);     // dmin = -y[i].d * fp16(x[i].dmin)
            accm[nt][0] = __fmaf_rn(d0, c[0], accm[nt][0]);
            accm[nt][1] = __fmaf_rn(d0, c[1], accm[nt][1]);
            accm[nt][2] = __fmaf_rn(d1, c[2], accm[nt][2]);
            accm[nt][3] = __fmaf_rn(d1, c[3], accm[nt][3]);
        }
    }
}
// Q5_K: summs += dmin * (float) sum_g m[g] bsum[g] (ggml-quants.c:7516); columns = tokens. warp = (mt = w & 3, 8 tokens nt = w >> 2)
__device__ __forceinline__ void mb_mma_mins5(const uint8_t * As, uint32_t mins_off, const uint8_t * rec, const float * scal, int warp, int lane,
                                             float (&acc5)[4]) {
    const int mt = warp & 3, nt = warp >> 2;
    uint32_t af[4], bf[2];
    mb_mins_afrag(As, mins_off, mt, lane, af);
    {
        const int t = 8 * nt + (lane & 7), kc = (lane >> 3) & 1;
        ldsm_x2(bf, smem_u32(rec) + MB_OFF_M5 + t * 32 + ((((uint32_t) kc) ^ ((t >> 2) & 1)) << 4));
    }
    float c[4];
    mma_f16_z(c, af, bf[0], bf[1]);
    const int r0 = mt * 16 + (lane >> 2);
    const float dm0 = scal[MB_ROWS + r0], dm1 = scal[MB_ROWS + r0 + 8];
    const float2 y = *reinterpret_cast<const float2 *>(rec + MB_OFF_YD + (8 * nt + 2 * (lane & 3)) * 4);
    acc5[0] = __fadd_rn(acc5[0], __fmul_rn(__fmul_rn(-y.x, dm0), c[0]));
    acc5[1] = __fadd_rn(acc5[1], __fmul_rn(__fmul_rn(-y.y, dm0), c[1]));
    acc5[2] = __fadd_rn(acc5[2], __fmul_rn(__fmul_rn(-y.x, dm1), c[2]));
    acc5[3] = __fadd_rn(acc5[3], __fmul_rn(__fmul_rn(-y.y, dm1), c[3]));
}

__global__ void __launch_bounds__(MB_WARPS * 32, 1) k_mma_batch(const __grid_constant__ MatmulBatchArgs a) {
    extern __shared__ __align__(128) uint8_t mb_smem[];
    __shared__ __align__(8) uint64_t bars[MB_MAX_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, unit0 = 2 * blockIdx.y;
    const UnitDesc ud = pb_describe_unit(a, unit0), ud1 = pb_describe_unit(a, unit0 + 1);   // same segment (host: even unit counts)
    // the A operand is double-buffered: step b+1 is expanded while the tensor cores work on step b (one barrier per K step)
    uint8_t * stages = mb_smem + 2 * a.mb_a_bytes;
    const uint32_t mins_off = a.mb_a_bytes - (MB_A_MINS + MB_A_SCAL);
    const int n_steps = a.tiles_unit, n_stages = a.mb_stages;
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint8_t * rec_chunk = a.rec + (size_t) chunk * n_steps * MB_REC_BYTES;
    auto issue = [&](int step, int s) {   // thread 0: the two raw tiles and the chunk's record block of K step `step`
        const uint32_t dst = smem_u32(stages + (size_t) s * a.mb_stage_bytes), bar = bar0 + 8 * s;
        mbar_expect_tx(bar, 2 * ud.bytes + a.mb_rec_copy);
        bulk_g2s(dst, ud.tiles + (size_t) step * ud.bytes, ud.bytes, bar);
        bulk_g2s(dst + a.mb_raw_stride, ud1.tiles + (size_t) step * ud.bytes, ud.bytes, bar);
        bulk_g2s(dst + 2 * a.mb_raw_stride, rec_chunk + (size_t) step * MB_REC_BYTES, a.mb_rec_copy, bar);
    };
    if (tid == 0) { for (int s = 0; s < n_stages && s < n_steps; s++) issue(s, s); }

    auto body = [&](auto tag) {
        constexpr int TYPE = decltype(tag)::value;
        float acc[2][4][4];
        float accm[4][4];
#pragma unroll
        for (int i = 0; i < 2; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int e = 0; e < 4; e++) { acc[i][j][e] = 0.f; accm[j][e] = 0.f; }
            }
        }
        mbar_wait(bar0, 0u);
        mb_expand<TYPE>(stages, a.mb_raw_stride, mb_smem, mins_off, warp, lane);
        __syncthreads();
        int s = 0, s1 = n_stages > 1 ? 1 : 0;                    // stage of step b, of step b + 1
        uint32_t par1 = n_stages > 1 ? 0u : 1u;                  // full-flag parity of step b + 1's use of its stage
        for (int step = 0; step < n_steps; step++) {
            const uint8_t * As = mb_smem + (size_t) (step & 1) * a.mb_a_bytes;
            const float * scal = reinterpret_cast<const float *>(As + mins_off + MB_A_MINS);
            const uint8_t * rec = stages + (size_t) s * a.mb_stage_bytes + 2 * a.mb_raw_stride;
            mb_mma_main<TYPE>(As, rec, scal, warp, lane, acc);
            if (TYPE == T_Q4_K) mb_mma_mins4(As, mins_off, rec, scal, warp, lane, accm);
            if (TYPE == T_Q5_K) mb_mma_mins5(As, mins_off, rec, scal, warp, lane, accm[0]);
            if (step + 1 < n_steps) {
                mbar_wait(bar0 + 8 * s1, par1);
                mb_expand<TYPE>(stages + (size_t) s1 * a.mb_stage_bytes, a.mb_raw_stride, mb_smem + (size_t) ((step + 1) & 1) * a.mb_a_bytes, mins_off, warp, lane);
            }
            __syncthreads();                                   // A[step & 1] and stage s are free, A[(step + 1) & 1] is complete
            if (tid == 0 && step + n_stages < n_steps) issue(step + n_stages, s);
            s = s1;
            if (++s1 == n_stages) { s1 = 0; par1 ^= 1u; }
        }
        // the 12 chains of every output meet in shared memory: CH[c][token][row]
        float * CH = reinterpret_cast<float *>(mb_smem);
        {
            const int m = warp & 7, rh = warp >> 3;
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int row = rh * 32 + mt * 16 + (lane >> 2) + 8 * (e >> 1), t = nt * 8 + 2 * (lane & 3) + (e & 1);
                        CH[(m * MB_NT + t) * MB_CH_STRIDE + row] = acc[mt][nt][e];
                    }
                }
            }
            if (TYPE == T_Q4_K) {
                const int mt = warp & 3, tq = warp >> 2;
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int row = mt * 16 + (lane >> 2) + 8 * (e >> 1), t = 8 * tq + 2 * nt + ((lane & 3) >> 1), l = 2 * (lane & 1) + (e & 1);
                        CH[((8 + l) * MB_NT + t) * MB_CH_STRIDE + row] = accm[nt][e];
                    }
                }
            }
            if (TYPE == T_Q5_K) {
                const int mt = warp & 3, nt = warp >> 2;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int row = mt * 16 + (lane >> 2) + 8 * (e >> 1), t = 8 * nt + 2 * (lane & 3) + (e & 1);
                    CH[(8 * MB_NT + t) * MB_CH_STRIDE + row] = accm[0][e];
                }
            }
        }
        __syncthreads();
        {
            const int r = tid & (MB_ROWS - 1), tb = tid >> 6;
#pragma unroll 1
            for (int jj = 0; jj < 4; jj++) {
                const int t = tb + 8 * jj;
                float c[12];
#pragma unroll
                for (int i = 0; i < n_chains<TYPE>(); i++) c[i] = CH[(i * MB_NT + t) * MB_CH_STRIDE + r];
                const float val = finish_row<TYPE>(c);
                pb_epilogue(a, val, ud.row0 + r, lane, chunk * MB_NT + t);
            }
        }
    };
    switch (ud.type) {
        case T_Q4_K: body(TypeTag<T_Q4_K>{}); break;
        case T_Q5_K: body(TypeTag<T_Q5_K>{}); break;
        default:     body(TypeTag<T_Q6_K>{}); break;
    }
}


// ------------------------------------------------------------------------------------------------------------
// Q8_0 x Q8_0 prompt batches on the tensor cores. The reference (tinyBLAS_Q0_AVX, cpp/ggml/src/llamafile/sgemm.cpp) keeps 8
// fp32 lanes per output: lane m sums the FOUR products of bytes 4m..4m+3 of a 32-weight block (dpbusd) and runs
// acc[m] = fma(d_w d_a, (float) isum[m], acc[m]) block after block. A 4-term contraction is exactly a dp4a — but it is also
// an m16n8k16 MMA with a block-diagonal B: K = 16 = four lanes x four bytes, the eight B columns = (2 tokens x 4 lanes), column
// (t, q) carrying token t's bytes only in rows 4q..4q+3. One HMMA then returns 128 lane sums as exact floats (|isum| <= 4 *
// 127 * 127 < 2^24, operands are int8 values in fp16), at a quarter of the tensor density but OFF the CUDA cores, which keep
// only the ordered chain step (8 FFMA per output and block instead of 8 dp4a + 8 FADD + 8 FFMA).
//   CTA 64 rows x 32 tokens, 16 warps; per K step of 256 weights: TMA bulk copies of the 2 x 8 raw tiles + the chunk's fp16
//   activation rows; 512 threads expand int8 -> fp16 A rows (double-buffered); warp (16-row tile mt = w & 3, 8 tokens
//   tq = w >> 2): per block 2 ldmatrix.x4 (A), 8 predicated 4-byte B loads, 8 HMMA, 32 chain FFMA.
// ------------------------------------------------------------------------------------------------------------
// activation rows of a (32-token chunk, 256-weight step): token row = [block pair p (4)][element pair s (8)][w (4)] half2, w = 2 (block & 1) +
// 16-group g: the four B words a thread needs for two blocks are one 16-byte load
static constexpr int Q80_ROW = 512;
static constexpr int Q80_OFF_DX = MB_NT * Q80_ROW;                     // f32 [32 tokens][8 blocks]
static constexpr int Q80_REC_BYTES = Q80_OFF_DX + MB_NT * 8 * 4;       // 17408
static constexpr int Q80_A_BYTES = 8 * 2 * MB_ROWS * 32 + 8 * MB_ROWS * 4;   // fp16 rows [block][16-group][row][16] | f32 d [row][block]

// k_quant_batch_mma, layout 3: token t's Q8_0 image (natural order) -> fp16 words in fragment order + block scales
__device__ __forceinline__ void q80_write_records(const ActSmem & A, int n256, uint8_t * rec, int t, int lane, int warp, int W) {
    const int chunk = t / MB_NT, j = t % MB_NT;
    for (int b = warp; b < n256; b += W) {
        uint8_t * r = rec + ((size_t) chunk * n256 + b) * Q80_REC_BYTES;
        const int8_t * qb = A.q + (size_t) b * 256;
        const int p = lane >> 3, sl = lane & 7;
        uint32_t o[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const int8_t * src = qb + (2 * p + (w >> 1)) * 32 + (w & 1) * 16 + 2 * sl;
            o[w] = h2_ints((int) src[0], (int) src[1]);
        }
        *reinterpret_cast<uint4 *>(r + j * Q80_ROW + lane * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        if (lane < 8) *reinterpret_cast<float *>(r + Q80_OFF_DX + (j * 8 + lane) * 4) = A.dx[b * 8 + lane];
    }
}

__global__ void __launch_bounds__(MB_WARPS * 32, 1) k_mma_batch_q80(const __grid_constant__ MatmulBatchArgs a) {
    extern __shared__ __align__(128) uint8_t q8_smem[];
    __shared__ __align__(8) uint64_t bars[MB_MAX_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, unit0 = 2 * blockIdx.y;
    const UnitDesc ud = pb_describe_unit(a, unit0), ud1 = pb_describe_unit(a, unit0 + 1);
    uint8_t * stages = q8_smem + 2 * Q80_A_BYTES;
    const int n_steps = a.tiles_unit / 8, n_stages = a.mb_stages;      // a K step = eight 32-weight tiles per unit
    const uint32_t raw_bytes = 8 * ud.bytes;
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint8_t * rec_chunk = a.rec + (size_t) chunk * n_steps * Q80_REC_BYTES;
    auto issue = [&](int step, int s) {
        const uint32_t dst = smem_u32(stages + (size_t) s * a.mb_stage_bytes), bar = bar0 + 8 * s;
        mbar_expect_tx(bar, 2 * raw_bytes + Q80_REC_BYTES);
        bulk_g2s(dst, ud.tiles + (size_t) step * raw_bytes, raw_bytes, bar);
        bulk_g2s(dst + a.mb_raw_stride, ud1.tiles + (size_t) step * raw_bytes, raw_bytes, bar);
        bulk_g2s(dst + 2 * a.mb_raw_stride, rec_chunk + (size_t) step * Q80_REC_BYTES, Q80_REC_BYTES, bar);
    };
    if (tid == 0) { for (int s = 0; s < n_stages && s < n_steps; s++) issue(s, s); }

    // int8 -> fp16 rows of the A operand: warp = (unit u, tile tt), lane = row; both 16-byte halves of the block
    auto expand = [&](const uint8_t * stage, uint8_t * As) {
        const int u = warp >> 3, tt = warp & 7, r = u * 32 + lane;
        const uint8_t * tile = stage + (size_t) u * a.mb_raw_stride + (size_t) tt * ud.bytes;
        const __half2 c1152 = __half2half2(__int2half_rn(1152));
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 w = lds_u4(tile + h * 512 + lane * 16);
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t x = word_of(w, i) ^ 0x80808080u;        // b + 128 in 0..255; (0x6400 | v) = 1024 + v
                o[2 * i]     = h2_bits(__hsub2(bits_h2(__byte_perm(x, 0x64646464u, 0x4140u)), c1152));
                o[2 * i + 1] = h2_bits(__hsub2(bits_h2(__byte_perm(x, 0x64646464u, 0x4342u)), c1152));
            }
            uint8_t * dst = As + (size_t) (tt * 2 + h) * (MB_ROWS * 32) + r * 32;
            const uint32_t sw = (uint32_t) ((r >> 2) & 1);
            *reinterpret_cast<uint4 *>(dst + ((0u ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4 *>(dst + ((1u ^ sw) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        float * dwb = reinterpret_cast<float *>(As + 8 * 2 * MB_ROWS * 32);
        dwb[r * 8 + tt] = __half2float(*reinterpret_cast<const __half *>(tile + 1024 + lane * 2));
    };

    const int mt = warp & 3, tq = warp >> 2;
    float acc[4][2][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][g][e] = 0.f;
        }
    }
    mbar_wait(bar0, 0u);
    expand(stages, q8_smem);
    __syncthreads();
    int s = 0, s1 = n_stages > 1 ? 1 : 0;
    uint32_t par1 = n_stages > 1 ? 0u : 1u;
    // fragment constants: A rows for ldmatrix; the B element this thread supplies: rows k = 2 (lane & 3) [+8], column (token
    // lane >> 4, lane-in-group q = (lane >> 2) & 3); it is non-zero only when k's group of four equals q
    const int rl = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kc = lane >> 4;
    const uint32_t a_off = (uint32_t) (rl * 32) + ((((uint32_t) kc) ^ ((rl >> 2) & 1)) << 4);
    const int q = (lane >> 2) & 3, kq = (lane & 3) >> 1;
    const bool b_on = (q & 1) == kq;                       // k / 4 == q  (b0 holds k < 8: q in {0, 1}; b1 holds k >= 8: q in {2, 3})
    const uint32_t m0 = (b_on && q < 2) ? 0xffffffffu : 0u, m1 = (b_on && q >= 2) ? 0xffffffffu : 0u;
    const uint32_t b_slot = (uint32_t) ((lane & 3) + (q >= 2 ? 4 : 0)) * 16;      // element pair 2 (lane & 3) [+8] of the 16-group
    const int r0 = mt * 16 + (lane >> 2);
    for (int step = 0; step < n_steps; step++) {
        const uint8_t * As = q8_smem + (size_t) (step & 1) * Q80_A_BYTES;
        const float * dwb = reinterpret_cast<const float *>(As + 8 * 2 * MB_ROWS * 32);
        const uint8_t * rec = stages + (size_t) s * a.mb_stage_bytes + 2 * a.mb_raw_stride;
        const uint8_t * dxr = rec + Q80_OFF_DX;
#pragma unroll 1
        for (int p = 0; p < 4; p++) {                          // two blocks per iteration: every operand is a vector load
            uint32_t af[2][2][4];
#pragma unroll
            for (int bb = 0; bb < 2; bb++) {
#pragma unroll
                for (int g = 0; g < 2; g++) ldsm_x4(af[bb][g], smem_u32(As) + (uint32_t) ((2 * p + bb) * 2 + g) * (MB_ROWS * 32) + a_off);
            }
            const float2 dwa = *reinterpret_cast<const float2 *>(dwb + r0 * 8 + 2 * p), dwc = *reinterpret_cast<const float2 *>(dwb + (r0 + 8) * 8 + 2 * p);
            uint4 bv[4];
            float2 dxv[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int tk = 8 * tq + 2 * nt;
                bv[nt] = *reinterpret_cast<const uint4 *>(rec + (size_t) (tk + (lane >> 4)) * Q80_ROW + p * 128 + b_slot);
                dxv[nt] = *reinterpret_cast<const float2 *>(dxr + ((tk + ((lane & 3) >> 1)) * 8 + 2 * p) * 4);
            }
#pragma unroll
            for (int bb = 0; bb < 2; bb++) {
                float c[4][2][4];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        const uint32_t v = word_of(bv[nt], 2 * bb + g);
                        mma_f16_z(c[nt][g], af[bb][g], v & m0, v & m1);
                    }
                }
                const float dw0 = bb ? dwa.y : dwa.x, dw1 = bb ? dwc.y : dwc.x;
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const float dx = bb ? dxv[nt].y : dxv[nt].x;
                    const float d0 = __fmul_rn(dw0, dx), d1 = __fmul_rn(dw1, dx);            // fp16(x.d) * fp16(y.d)
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        acc[nt][g][0] = __fmaf_rn(d0, c[nt][g][0], acc[nt][g][0]);
                        acc[nt][g][1] = __fmaf_rn(d0, c[nt][g][1], acc[nt][g][1]);
                        acc[nt][g][2] = __fmaf_rn(d1, c[nt][g][2], acc[nt][g][2]);
                        acc[nt][g][3] = __fmaf_rn(d1, c[nt][g][3], acc[nt][g][3]);
                    }
                }
            }
        }
        if (step + 1 < n_steps) {
            mbar_wait(bar0 + 8 * s1, par1);
            expand(stages + (size_t) s1 * a.mb_stage_bytes, q8_smem + (size_t) ((step + 1) & 1) * Q80_A_BYTES);
        }
        __syncthreads();
        if (tid == 0 && step + n_stages < n_steps) issue(step + n_stages, s);
        s = s1;
        if (++s1 == n_stages) { s1 = 0; par1 ^= 1u; }
    }
    // the 8 lanes of every output meet in shared memory: CH[m][token][row]
    float * CH = reinterpret_cast<float *>(q8_smem);
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int row = r0 + 8 * (e >> 1), t = 8 * tq + 2 * nt + ((lane & 3) >> 1), m = 4 * g + 2 * (lane & 1) + (e & 1);
                CH[(m * MB_NT + t) * MB_CH_STRIDE + row] = acc[nt][g][e];
            }
        }
    }
    __syncthreads();
    {
        const int r = tid & (MB_ROWS - 1), tb = tid >> 6;
#pragma unroll 1
        for (int jj = 0; jj < 4; jj++) {
            const int t = tb + 8 * jj;
            float c[12];
#pragma unroll
            for (int i = 0; i < 8; i++) c[i] = CH[(i * MB_NT + t) * MB_CH_STRIDE + r];
            const float val = finish_row<T_Q8_0>(c);
            pb_epilogue(a, val, ud.row0 + r, lane, chunk * MB_NT + t);
        }
    }
}

}  // namespace b200
