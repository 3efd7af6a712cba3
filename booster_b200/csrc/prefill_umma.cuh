// prefill_umma.cuh — prompt batches of K-quant matrices on the 5th-generation tensor cores: tcgen05.mma (kind::f16) with the
// accumulators in tensor memory. Same exact-integer formulation as prefill_mma.cuh (read its header first): per super-block b and
// AVX2 lane-slice m the contraction  isum[m] = sum_g scale[g] * sum_i w[g][4m+i] * a[g][4m+i]  is a K = 32 fp16 MMA whose operands
// (weight x sub-block scale, int8 activation) and whose every partial sum are integers below 2^24, so D = (float) isum[m] exactly.
//
// Why this kernel exists next to the mma.sync one: tcgen05 moves the contraction off the SM's issue slots and its accumulators
// out of the register file; what remains is CUDA-core work — expanding the 4.5-6.5 bit weights to fp16 operands and the ordered
// fp32 chain step acc[m] = fma(d_b, D_m, acc[m]) on every drained accumulator.
// MEASURED (B200, 8B Q4_K_M, 512-token batch; profiles/r02h_prefill_umma_ncu_summary.txt): bit-exact, but SLOWER than k_mma_batch
// (136 vs 92 ms per batch at the time of the A/B). The 12 fp32 chains per output element must stay in registers, which caps the
// tile at 128 rows x 16 tokens: every expanded A tile is used for 16 tokens only (32 in k_mma_batch), SS-mode MMAs re-read the
// 64 KB A tile from shared memory for 0.5 MMAC, and a third of the samples wait on the MMA-completion mbarrier because the
// single A buffer serialises expansion and MMA. The HMMA kernel is not tensor-bound either (HMMA pipe 22-26 % busy), so the
// 16x higher tcgen05 rate buys nothing here. Kept selectable (b200_set_prefill_mma(2)) as the measured tcgen05 variant.
//
// CTA = 128 rows (four 32-row units) x 16 tokens, 16 warps, one CTA per SM (TMEM: all 512 columns).
//   TMA bulk copies   raw weight tiles of the four units + the chunk's activation record block (fp16, canonical no-swizzle
//                     K-major core-matrix order, written by k_quant_batch_mma) -> ring of stages, mbarrier full flags
//   expand            512 threads: raw tiles -> fp16 A operand in shared memory, core matrices [row/8][k-chunk][row%8][8 halfs]
//                     (SBO = 512 B between 8-row groups, LBO = 128 B between the two k-chunks of an MMA), generic->async proxy fence
//   issue (1 thread)  per lane-slice m: D[buf][16 m .. 16 m + 15] = A_m (128 x 32) * B_m (32 x 16): two K = 16 tcgen05.mma (four for
//                     Q6_K: even part + odd bit of the weight), mins: one N = 64 (Q4_K: columns (token, l)) or N = 16 (Q5_K) MMA;
//                     tcgen05.commit -> mbarrier
//   drain + chains    warp w reads TMEM lanes 32 (w & 3) .. +31 (one output row per thread), tokens 4 (w >> 2) .. +3: tcgen05.ld
//                     32x32b, then the chain FFMAs; the 12 chains of an output element never leave the thread's registers
// Software pipeline (one __syncthreads per K step): iteration b waits for MMA(b), expands step b+1 into the (single) A buffer,
// issues MMA(b+1) into the other accumulator buffer, then drains accumulator buffer b & 1 while the tensor core works.
#pragma once
#include "prefill_mma.cuh"

namespace b200 {

static constexpr int UM_NT = 16;               // tokens per CTA
static constexpr int UM_ROWS = 128;            // rows per CTA (four work units)
static constexpr int UM_WARPS = 16;
static constexpr int UM_MAX_STAGES = 6;
// activation record block of one (16-token chunk, super-block)
//   main  [8 lanes m][token/8][k-chunk j (4)][token%8][8 fp16]: chunk j = groups 2j, 2j+1 of lane-slice m, bytes in the order
//         (0, 2, 1, 3) (the order the expansion extracts nibble pairs in)
//   yd    f32[16]
//   m4    [(4 token + l)/8 (8)][k-chunk (2)][.%8][8 fp16]   kk = 2 g + part, non-zero for g in {2l, 2l+1}   (Q4_K)
//   m5    [token/8 (2)][k-chunk (2)][token%8][8 fp16]       the same, dense                                 (Q5_K)
static constexpr int UM_REC_MAIN = 8 * UM_NT * 64;             // 8192
static constexpr int UM_OFF_YD = UM_REC_MAIN;
static constexpr int UM_OFF_M4 = UM_OFF_YD + UM_NT * 4;        // 8256
static constexpr int UM_OFF_M5 = UM_OFF_M4 + UM_NT * 4 * 32;   // 10304
static constexpr int UM_REC_BYTES = UM_OFF_M5 + UM_NT * 32;    // 10816
__host__ __device__ __forceinline__ int um_rec_copy_bytes(bool q4, bool q5) { return q5 ? UM_REC_BYTES : q4 ? UM_OFF_M5 : UM_OFF_M4; }
// A operand: part 0 [8 m][16 row groups][4 k-chunks][8 rows][16 B] (| part 1 for Q6_K) | mins [16][2][8][16 B] | f32 (d, dmin)[2 buffers][2][128]
static constexpr int UM_A_SLICE = UM_ROWS * 64;                // 8192
static constexpr int UM_A_PART = 8 * UM_A_SLICE;               // 65536
static constexpr int UM_A_MINS = UM_ROWS * 32;                 // 4096
static constexpr int UM_A_SCAL = 2 * 2 * UM_ROWS * 4;          // 2048
__host__ __device__ __forceinline__ int um_a_bytes(bool q6) { return (q6 ? 2 : 1) * UM_A_PART + UM_A_MINS + UM_A_SCAL; }
static constexpr uint32_t UM_D_BUF = 256;      // TMEM columns per accumulator buffer: 8 x 16 main | 64 mins

// ---- tcgen05 / TMEM wrappers (PTX ISA 8.6+, sm_100a) ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc_512(uint32_t smem_dst) {   // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {    // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(512u) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1
__device__ __forceinline__ uint64_t um_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((saddr >> 4) & 0x3fffu) | ((uint64_t) ((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t) ((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t um_idesc(int n) { return (1u << 4) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (UM_ROWS >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// mbarrier wait that gives up (trap) instead of hanging the GPU if a completion never arrives
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity, int what) {
    uint32_t ok;
    unsigned spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) {
            if ((threadIdx.x & 31) == 0) printf("booster_b200: k_umma_batch mbarrier timeout (%d) CTA (%d,%d) warp %d\n", what, (int) blockIdx.x, (int) blockIdx.y, (int) threadIdx.x >> 5);
            __trap();
        }
    } while (!ok);
}

// k_quant_batch_mma, tcgen05 layout: token t's quantized image -> its rows of the chunk's record blocks
__device__ __forceinline__ void um_write_records(const ActSmem & A, int n256, uint8_t * rec, int t, int lane, int warp, int W) {
    const int chunk = t / UM_NT, j = t % UM_NT;
    for (int b = warp; b < n256; b += W) {
        uint8_t * r = rec + ((size_t) chunk * n256 + b) * UM_REC_BYTES;
        const int8_t * qb = A.q + (size_t) b * 256;
        {   // lane = (m, c): k-chunk c of lane-slice m = groups 2c, 2c+1, bytes 4m..4m+3 in the order (0, 2, 1, 3)
            const int m = lane >> 2, c = lane & 3;
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(qb + 64 * c + 4 * m);
            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(qb + 64 * c + 32 + 4 * m);
            uint4 o;
            o.x = h2_ints(sbyte_of(w0, 0), sbyte_of(w0, 2)); o.y = h2_ints(sbyte_of(w0, 1), sbyte_of(w0, 3));
            o.z = h2_ints(sbyte_of(w1, 0), sbyte_of(w1, 2)); o.w = h2_ints(sbyte_of(w1, 1), sbyte_of(w1, 3));
            *reinterpret_cast<uint4 *>(r + m * (UM_NT * 64) + (j >> 3) * 512 + c * 128 + (j & 7) * 16) = o;
        }
        if (lane == 0) *reinterpret_cast<float *>(r + UM_OFF_YD + j * 4) = A.dx[b];
        if (lane < 10) {
            const int * bp = A.bp + (size_t) b * 8;
            const int ch = lane & 1;
            uint4 z = make_uint4(0u, 0u, 0u, 0u);
            if (lane < 8) {
                const int l = lane >> 1, n = 4 * j + l;
                if (ch == (l >> 1)) {
                    const int s0 = bp[2 * l], s1 = bp[2 * l + 1];
                    const int l0 = s0 & 63, l1 = s1 & 63;
                    const uint32_t p0 = h2_ints(l0, s0 - l0), p1 = h2_ints(l1, s1 - l1);
                    if (l & 1) { z.z = p0; z.w = p1; } else { z.x = p0; z.y = p1; }
                }
                *reinterpret_cast<uint4 *>(r + UM_OFF_M4 + (n >> 3) * 256 + ch * 128 + (n & 7) * 16) = z;
            } else {
                int s[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { s[i] = bp[4 * ch + i]; lo[i] = s[i] & 63; }
                z.x = h2_ints(lo[0], s[0] - lo[0]); z.y = h2_ints(lo[1], s[1] - lo[1]);
                z.z = h2_ints(lo[2], s[2] - lo[2]); z.w = h2_ints(lo[3], s[3] - lo[3]);
                *reinterpret_cast<uint4 *>(r + UM_OFF_M5 + (j >> 3) * 256 + ch * 128 + (j & 7) * 16) = z;
            }
        }
    }
}

// k_quant_batch with the tensor-core record layouts (one CTA per token: the decode path's prologue, then the image is written out
// for k_mma_batch (layout 1) or k_umma_batch (layout 2))
__global__ void __launch_bounds__(512) k_quant_batch_mma(const QuantBatchArgs a) {
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
    if (a.layout == 3) q80_write_records(A, a.k / 256, a.rec, t, lane, warp, W);
    else if (a.layout == 2) um_write_records(A, a.k / 256, a.rec, t, lane, warp, W);
    else mb_write_records(A, a.k / 256, a.rec, t, lane, warp, W);
}

// raw tile of unit u, 16-byte quant chunk c -> core-matrix rows of the A operand. lane = row of the unit.
template <int TYPE>
__device__ __forceinline__ void um_expand_task(const uint8_t * tile, uint8_t * As, uint32_t mins_off, float * scal, int u, int c, int lane) {
    const int r = u * 32 + lane;
    const uint8_t * sl = tile + lane * 16;
    uint8_t * arow = As + (r >> 3) * 512 + (r & 7) * 16;      // + slice * 8192 + k-chunk * 128
    if (TYPE == T_Q4_K || TYPE == T_Q5_K) {
        const int j = c >> 1, h = c & 1;
        const uint4 W4 = lds_u4(sl + c * 512);
        const uint4 sd = lds_u4(sl + 4096);
        const uint32_t sc_a = sd.x & 0x3f3f3f3fu, sc_b = (sd.z & 0x0f0f0f0fu) | ((sd.x >> 2) & 0x30303030u);
        const uint32_t scw = (j & 2) ? sc_b : sc_a;
        const int sh = (j & 1) * 16;
        const int sc_lo = (int) ((scw >> sh) & 0xffu), sc_hi = (int) ((scw >> (sh + 8)) & 0xffu);
        const __half2 s_lo = __half2half2(__int2half_rn(sc_lo)), s_hi = __half2half2(__int2half_rn(sc_hi));
        const __half2 o_lo = __half2half2(__int2half_rn(-1024 * sc_lo)), o_hi = __half2half2(__int2half_rn(-1024 * sc_hi));
        uint4 H4 = make_uint4(0u, 0u, 0u, 0u);
        if (TYPE == T_Q5_K) H4 = lds_u4(sl + 4608 + h * 512);
#pragma unroll
        for (int wi = 0; wi < 4; wi++) {
            const uint32_t Wd = word_of(W4, wi);
            // nibble pairs straight out of the word: bytes (0, 2) and (1, 3) — the record blocks use the same order
            uint32_t l02 = Wd & 0x000f000fu, h02 = (Wd >> 4) & 0x000f000fu, l13 = (Wd >> 8) & 0x000f000fu, h13 = (Wd >> 12) & 0x000f000fu;
            if (TYPE == T_Q5_K) {   // qh bit 2j -> +16 on the low-nibble weight, bit 2j+1 -> +16 on the high-nibble one
                const uint32_t Hs = word_of(H4, wi) >> (2 * j);
                l02 |= (Hs & 0x00010001u) << 4;        h02 |= (Hs & 0x00020002u) << 3;
                l13 |= ((Hs >> 8) & 0x00010001u) << 4; h13 |= ((Hs >> 8) & 0x00020002u) << 3;
            }
            uint4 o;
            o.x = h2_bits(__hfma2(bits_h2(l02 | 0x64006400u), s_lo, o_lo));
            o.y = h2_bits(__hfma2(bits_h2(l13 | 0x64006400u), s_lo, o_lo));
            o.z = h2_bits(__hfma2(bits_h2(h02 | 0x64006400u), s_hi, o_hi));
            o.w = h2_bits(__hfma2(bits_h2(h13 | 0x64006400u), s_hi, o_hi));
            *reinterpret_cast<uint4 *>(arow + (4 * h + wi) * UM_A_SLICE + j * 128) = o;
        }
        if (c < 2) {   // mins operand: kk = 2g + part <- m[g]; k-chunk c = mins 4c..4c+3
            const uint32_t m_a = sd.y & 0x3f3f3f3fu, m_b = ((sd.z >> 4) & 0x0f0f0f0fu) | ((sd.y >> 2) & 0x30303030u);
            const uint32_t mw = c ? m_b : m_a;
            uint4 o;
            o.x = h2_ints((int) (mw & 0xffu), (int) (mw & 0xffu));
            o.y = h2_ints((int) ((mw >> 8) & 0xffu), (int) ((mw >> 8) & 0xffu));
            o.z = h2_ints((int) ((mw >> 16) & 0xffu), (int) ((mw >> 16) & 0xffu));
            o.w = h2_ints((int) (mw >> 24), (int) (mw >> 24));
            *reinterpret_cast<uint4 *>(As + mins_off + (r >> 3) * 256 + c * 128 + (r & 7) * 16) = o;
            if (c == 0) {
                const __half2 dmh = bits_h2(sd.w);
                scal[r] = __low2float(dmh); scal[UM_ROWS + r] = __high2float(dmh);
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
            uint8_t * dst = arow + (G >> 1) * 128 + (G & 1) * 8;
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                const uint32_t QL = word_of(ql, wi), QH = word_of(qh, wi);
                const uint32_t lo = gh ? ((QL >> 4) & 0x0f0f0f0fu) : (QL & 0x0f0f0f0fu);
                const uint32_t q = lo | (((QH >> (2 * g)) & 0x03030303u) << 4);            // 0..63 per byte
                const uint32_t t02 = q & 0x00ff00ffu, t13 = (q >> 8) & 0x00ff00ffu;        // byte pairs (0, 2), (1, 3)
                // q - 32 = ((q & ~1) - 32) + (q & 1): both terms times the scale are fp16 integers
                uint2 e, o;
                e.x = h2_bits(__hmul2(__hsub2(bits_h2((t02 & 0x003e003eu) | 0x64006400u), c1056), s2));
                e.y = h2_bits(__hmul2(__hsub2(bits_h2((t13 & 0x003e003eu) | 0x64006400u), c1056), s2));
                o.x = h2_bits(__hmul2(__hsub2(bits_h2((t02 & 0x00010001u) | 0x64006400u), c1024), s2));
                o.y = h2_bits(__hmul2(__hsub2(bits_h2((t13 & 0x00010001u) | 0x64006400u), c1024), s2));
                uint8_t * d = dst + (4 * mq + wi) * UM_A_SLICE;
                *reinterpret_cast<uint2 *>(d) = o;
                *reinterpret_cast<uint2 *>(d + UM_A_PART) = e;
            }
        }
        if (c == 0) scal[r] = __half2float(*reinterpret_cast<const __half *>(tile + 6656 + lane * 2));
    }
}

__global__ void __launch_bounds__(UM_WARPS * 32, 1) k_umma_batch(const __grid_constant__ MatmulBatchArgs a) {
    extern __shared__ __align__(128) uint8_t um_smem[];
    __shared__ __align__(8) uint64_t bars[UM_MAX_STAGES + 2];      // full[stage] | mma done[2]
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, unit0 = 4 * blockIdx.y;
    UnitDesc ud[4];
#pragma unroll
    for (int u = 0; u < 4; u++) ud[u] = pb_describe_unit(a, unit0 + u);     // same segment (host: unit counts are multiples of 4)
    uint8_t * As = um_smem;
    uint8_t * stages = um_smem + a.mb_a_bytes;
    const uint32_t mins_off = a.mb_a_bytes - (UM_A_MINS + UM_A_SCAL);
    float * scal0 = reinterpret_cast<float *>(As + mins_off + UM_A_MINS);
    const int n_steps = a.tiles_unit, n_stages = a.mb_stages;
    const uint32_t bar0 = smem_u32(&bars[0]), mbar0 = smem_u32(&bars[UM_MAX_STAGES]);
    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) mbar_init(bar0 + 8 * s, 1);
        mbar_init(mbar0, 1); mbar_init(mbar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) tmem_alloc_512(smem_u32(&tmem_base_s));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint8_t * rec_chunk = a.rec + (size_t) chunk * n_steps * UM_REC_BYTES;
    const uint32_t tile_bytes = ud[0].bytes;
    auto issue_load = [&](int step, int s) {   // thread 0: four raw tiles + the chunk's record block of K step `step`
        const uint32_t dst = smem_u32(stages + (size_t) s * a.mb_stage_bytes), bar = bar0 + 8 * s;
        mbar_expect_tx(bar, 4 * tile_bytes + a.mb_rec_copy);
#pragma unroll
        for (int u = 0; u < 4; u++) bulk_g2s(dst + u * a.mb_raw_stride, ud[u].tiles + (size_t) step * tile_bytes, tile_bytes, bar);
        bulk_g2s(dst + 4 * a.mb_raw_stride, rec_chunk + (size_t) step * UM_REC_BYTES, a.mb_rec_copy, bar);
    };
    if (tid == 0) { for (int s = 0; s < n_stages && s < n_steps; s++) issue_load(s, s); }

    auto body = [&](auto tag) {
        constexpr int TYPE = decltype(tag)::value;
        constexpr int PARTS = TYPE == T_Q6_K ? 2 : 1;
        const int q = warp & 3, cg = warp >> 2;                 // TMEM lane quarter (output rows 32q..32q+31), token group
        const int row = 32 * q + lane;
        float acc[4][8], accm[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int m = 0; m < 8; m++) acc[i][m] = 0.f;
#pragma unroll
            for (int l = 0; l < 4; l++) accm[i][l] = 0.f;
        }
        auto expand = [&](int step, int s) {                    // every thread: two (unit, chunk) tasks
            const uint8_t * stage = stages + (size_t) s * a.mb_stage_bytes;
            float * scal = scal0 + (step & 1) * (2 * UM_ROWS);
            const int u = warp >> 2, c0 = (warp & 3) * 2;
            um_expand_task<TYPE>(stage + (size_t) u * a.mb_raw_stride, As, mins_off, scal, u, c0, lane);
            um_expand_task<TYPE>(stage + (size_t) u * a.mb_raw_stride, As, mins_off, scal, u, c0 + 1, lane);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the tensor core's reads
        };
        auto issue_mma = [&](int step, int s) {                 // thread 0
            const uint32_t a0 = smem_u32(As), r0 = smem_u32(stages + (size_t) s * a.mb_stage_bytes + 4 * a.mb_raw_stride);
            const uint32_t d0 = tmem + (uint32_t) (step & 1) * UM_D_BUF;
#pragma unroll 1
            for (int m = 0; m < 8; m++) {
#pragma unroll
                for (int p = 0; p < PARTS; p++) {
#pragma unroll
                    for (int ks = 0; ks < 2; ks++)
                        umma_f16(d0 + 16 * m, um_desc(a0 + p * UM_A_PART + m * UM_A_SLICE + ks * 256, 128, 512),
                                 um_desc(r0 + m * (UM_NT * 64) + ks * 256, 128, 512), um_idesc(16), (p | ks) ? 1u : 0u);
                }
            }
            if (TYPE == T_Q4_K) umma_f16(d0 + 128, um_desc(a0 + mins_off, 128, 256), um_desc(r0 + UM_OFF_M4, 128, 256), um_idesc(64), 0u);
            if (TYPE == T_Q5_K) umma_f16(d0 + 128, um_desc(a0 + mins_off, 128, 256), um_desc(r0 + UM_OFF_M5, 128, 256), um_idesc(16), 0u);
            umma_commit(mbar0 + 8 * (step & 1));
        };
        // prologue: step 0 expanded and its MMAs issued
        mbar_wait_bounded(bar0, 0u, 0);
        expand(0, 0);
        tc_fence_before();
        __syncthreads();
        if (tid == 0) { tc_fence_after(); issue_mma(0, 0); }
        int s = 0, s1 = n_stages > 1 ? 1 : 0;                   // stage of step b, of step b + 1
        uint32_t par1 = n_stages > 1 ? 0u : 1u;                 // full-flag parity of step b + 1's use of its stage
        for (int b = 0; b < n_steps; b++) {
            mbar_wait_bounded(mbar0 + 8 * (b & 1), (uint32_t) ((b >> 1) & 1), 1);      // MMA(b) complete: A, B(b) are free
            tc_fence_after();
            // what the chain step of b needs from shared memory, before anything of it is re-used
            const uint8_t * rec = stages + (size_t) s * a.mb_stage_bytes + 4 * a.mb_raw_stride;
            const float4 yd = *reinterpret_cast<const float4 *>(rec + UM_OFF_YD + cg * 16);
            const float * scal = scal0 + (b & 1) * (2 * UM_ROWS);
            const float dw = scal[row], dmw = TYPE == T_Q6_K ? 0.f : scal[UM_ROWS + row];
            if (b + 1 < n_steps) {
                mbar_wait_bounded(bar0 + 8 * s1, par1, 2);
                expand(b + 1, s1);
            }
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                if (b + 1 < n_steps) issue_mma(b + 1, s1);
                if (b + n_stages < n_steps) issue_load(b + n_stages, s);
            }
            // drain accumulator buffer b & 1 (the tensor core is working on the other one)
            const uint32_t dT = tmem + ((uint32_t) (32 * q) << 16) + (uint32_t) (b & 1) * UM_D_BUF;
            const float ydv[4] = {yd.x, yd.y, yd.z, yd.w};
            float d[4];
#pragma unroll
            for (int i = 0; i < 4; i++) d[i] = __fmul_rn(ydv[i], dw);        // d = y[i].d * fp16(x[i].d)
            {
                float v[8][4];
#pragma unroll
                for (int m = 0; m < 8; m++) tmem_ld4(dT + 16 * m + 4 * cg, v[m]);
                tmem_ld_wait();
#pragma unroll
                for (int m = 0; m < 8; m++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) acc[i][m] = __fmaf_rn(d[i], v[m][i], acc[i][m]);
                }
            }
            if (TYPE == T_Q4_K) {
                float v[4][4];                                               // columns 4 t + l of the token group
#pragma unroll
                for (int i = 0; i < 4; i++) tmem_ld4(dT + 128 + 16 * cg + 4 * i, v[i]);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float dm = __fmul_rn(-ydv[i], dmw);                // dmin = -y[i].d * fp16(x[i].dmin)
#pragma unroll
                    for (int l = 0; l < 4; l++) accm[i][l] = __fmaf_rn(dm, v[i][l], accm[i][l]);
                }
            }
            if (TYPE == T_Q5_K) {
                float v[4];
                tmem_ld4(dT + 128 + 4 * cg, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) accm[i][0] = __fadd_rn(accm[i][0], __fmul_rn(__fmul_rn(-ydv[i], dmw), v[i]));
            }
            s = s1;
            if (++s1 == n_stages) { s1 = 0; par1 ^= 1u; }
        }
        // every chain of an output element is in this thread: finish the row, then the layer epilogue
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float c[12];
#pragma unroll
            for (int m = 0; m < 8; m++) c[m] = acc[i][m];
#pragma unroll
            for (int l = 0; l < 4; l++) c[8 + l] = accm[i][l];
            const float val = finish_row<TYPE>(c);
            pb_epilogue(a, val, ud[0].row0 + row, lane, chunk * UM_NT + 4 * cg + i);
        }
    };
    switch (ud[0].type) {
        case T_Q4_K: body(TypeTag<T_Q4_K>{}); break;
        case T_Q5_K: body(TypeTag<T_Q5_K>{}); break;
        default:     body(TypeTag<T_Q6_K>{}); break;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc_512(tmem); }
}

}  // namespace b200
