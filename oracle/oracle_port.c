/* oracle/oracle_port.c — TEST INFRASTRUCTURE, NOT PRODUCT ("port" oracle).
 *
 * A plain-C restatement of the reference's CPU arithmetic for the quantized LLaMA decode path, written
 * from the reference sources under /root/reference/cpp (every function cites the file:line it follows).
 * It exists so that parity can be checked where /root/reference is absent (the GPU box) and so that the
 * algorithm is stated once in readable scalar code. It is PINNED against the real reference: tests/ compare
 * it with oracle/_ref (the unmodified reference compiled by oracle/Makefile) and with the golden vectors under
 * tests/golden/ that oracle/_ref generated (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library. The product (booster_b200/) never does.
 *
 * EVERY floating-point operation is restated in the reference's exact order for its AVX-512 ("native",
 * -march=native on Sapphire Rapids, what Booster's own build flags produce) build: the K-quant dot products follow
 * the AVX2 code path (8 fp32 lanes, FMA chains over super-blocks, fixed horizontal add), K.q and V.p follow
 * tinyBLAS<16> (16-lane FMA chains + _mm512_reduce_add_ps), batch>1 K.q follows ggml_vec_dot_f16, silu and softmax
 * use the ggml_v_expf polynomial, row sums are accumulated in double. Result: whole-model logits are
 * BIT-IDENTICAL to the reference's (tests/test_oracle_pinned.py checks array_equal against golden vectors and
 * against oracle/_ref live). Compile with -ffp-contract=off: contraction is written explicitly with fmaf where the
 * reference uses FMA.
 */
#define _GNU_SOURCE   /* M_PI under -std=c11 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QK_K 256

/* ---- fp16 <-> fp32, IEEE round-to-nearest-even (what F16C _cvtss_sh / _cvtsh_ss do; ggml-impl.h) ---------- */
static float h2f(uint16_t h) {
    const uint32_t s = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1f, m = h & 0x3ffu, u;
    if (e == 0) {
        if (m == 0) u = s;
        else { e = 1; while (!(m & 0x400u)) { m <<= 1; e--; } m &= 0x3ffu; u = s | ((e + 112) << 23) | (m << 13); }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}
static uint16_t f2h(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    const uint32_t s = (u >> 16) & 0x8000u;
    const int32_t e = (int32_t)((u >> 23) & 0xff) - 127 + 15;
    uint32_t m = u & 0x7fffffu;
    if (((u >> 23) & 0xff) == 0xff) return (uint16_t)(s | 0x7c00u | (m ? 0x200u : 0));
    if (e >= 31) return (uint16_t)(s | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t) s;
        m |= 0x800000u;
        const int shift = 14 - e;
        uint32_t r = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1))) r++;
        return (uint16_t)(s | r);
    }
    uint32_t r = ((uint32_t) e << 10) | (m >> 13);
    const uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) r++;
    return (uint16_t)(s | r);
}
static uint16_t rd16(const uint8_t * p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* round-half-even of a float to int: nearest_int(), cpp/ggml/src/ggml-quants.c:1632-1637 (magic-number add) */
static int nearest_int(float fval) {
    float val = fval + 12582912.f;
    int i; memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

/* ---- activation quantization ----------------------------------------------------------------------------- */
/* quantize_row_q8_K_ref, cpp/ggml/src/ggml-quants.c:3593-3630. out: block_q8_K {float d; int8 qs[256];
 * int16 bsums[16]} = 292 bytes per block (cpp/ggml/src/ggml-common.h:311-315). */
void port_quantize_row_q8_K(const float * x, uint8_t * out, int64_t k) {
    for (int64_t b = 0; b < k / QK_K; b++, x += QK_K, out += 292) {
        float max = 0, amax = 0;
        for (int j = 0; j < QK_K; j++) { const float ax = fabsf(x[j]); if (ax > amax) { amax = ax; max = x[j]; } }
        int8_t * qs = (int8_t *)(out + 4);
        if (!amax) { memset(out, 0, 292); continue; }
        const float iscale = -127.f / max;
        for (int j = 0; j < QK_K; j++) { int v = nearest_int(iscale * x[j]); qs[j] = (int8_t)(v < 127 ? v : 127); }
        for (int j = 0; j < 16; j++) {
            int sum = 0;
            for (int i = 0; i < 16; i++) sum += qs[16 * j + i];
            const int16_t s16 = (int16_t) sum; memcpy(out + 260 + 2 * j, &s16, 2);
        }
        const float d = 1 / iscale; memcpy(out, &d, 4);
    }
}
/* quantize_row_q8_0, AVX path, cpp/ggml/src/ggml-quants.c:936-1000: d = amax/127 -> fp16, id = 127/amax,
 * q = round-half-even(x*id). out: block_q8_0 {half d; int8 qs[32]} = 34 bytes (ggml-common.h:186-190). */
void port_quantize_row_q8_0(const float * x, uint8_t * out, int64_t k) {
    for (int64_t b = 0; b < k / 32; b++, x += 32, out += 34) {
        float amax = 0;
        for (int j = 0; j < 32; j++) { const float ax = fabsf(x[j]); if (ax > amax) amax = ax; }
        const float d = amax / 127.f;
        const uint16_t dh = f2h(d); out[0] = (uint8_t)(dh & 0xff); out[1] = (uint8_t)(dh >> 8);
        const float id = amax != 0.0f ? 127.f / amax : 0.0f;
        for (int j = 0; j < 32; j++) out[2 + j] = (uint8_t)(int8_t) lrintf(x[j] * id);   /* default rounding mode = RNE */
    }
}

/* ---- block formats (cpp/ggml/src/ggml-common.h:186-316) and scale unpacking ------------------------------- */
/* get_scale_min_k4, cpp/ggml/src/ggml-quants.c:1891-1898 */
static void scale_min_k4(int j, const uint8_t * q, int * sc, int * m) {
    if (j < 4) { *sc = q[j] & 63; *m = q[j + 4] & 63; }
    else { *sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); *m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

/* dequantize_row_q4_K :2548, q5_K :2756, q6_K :2970, q8_0 :1609 (cpp/ggml/src/ggml-quants.c);
 * this is also the embedding get_rows path (cpp/ggml/src/ggml.c:13186). */
void port_dequantize_row(int type, const uint8_t * x, float * y, int64_t k) {
    if (type == 0) { memcpy(y, x, (size_t) k * 4); return; }
    if (type == 1) { for (int64_t i = 0; i < k; i++) y[i] = h2f(rd16(x + 2 * i)); return; }
    if (type == 8) {
        for (int64_t b = 0; b < k / 32; b++, x += 34) {
            const float d = h2f(rd16(x));
            for (int j = 0; j < 32; j++) *y++ = (int8_t) x[2 + j] * d;
        }
    } else if (type == 12 || type == 13) {
        const int bb = type == 12 ? 144 : 176;
        for (int64_t b = 0; b < k / QK_K; b++, x += bb) {
            const float d = h2f(rd16(x)), min = h2f(rd16(x + 2));
            const uint8_t * scales = x + 4;
            const uint8_t * qh = x + 16;                       /* q5_K only */
            const uint8_t * q = type == 12 ? x + 16 : x + 48;
            int is = 0; uint8_t u1 = 1, u2 = 2;
            for (int j = 0; j < QK_K; j += 64) {
                int sc, m;
                scale_min_k4(is + 0, scales, &sc, &m); const float d1 = d * sc, m1 = min * m;
                scale_min_k4(is + 1, scales, &sc, &m); const float d2 = d * sc, m2 = min * m;
                for (int l = 0; l < 32; l++) *y++ = d1 * ((q[l] & 0xF) + (type == 13 && (qh[l] & u1) ? 16 : 0)) - m1;
                for (int l = 0; l < 32; l++) *y++ = d2 * ((q[l] >> 4) + (type == 13 && (qh[l] & u2) ? 16 : 0)) - m2;
                q += 32; is += 2; u1 <<= 2; u2 <<= 2;
            }
        }
    } else if (type == 14) {
        for (int64_t b = 0; b < k / QK_K; b++, x += 210) {
            const float d = h2f(rd16(x + 208));
            const uint8_t * ql = x, * qh = x + 128; const int8_t * sc = (const int8_t *)(x + 192);
            for (int n = 0; n < QK_K; n += 128) {
                for (int l = 0; l < 32; l++) {
                    const int is = l / 16;
                    const int8_t q1 = (int8_t)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    const int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    const int8_t q3 = (int8_t)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    const int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    y[l] = d * sc[is] * q1; y[l + 32] = d * sc[is + 2] * q2;
                    y[l + 64] = d * sc[is + 4] * q3; y[l + 96] = d * sc[is + 6] * q4;
                }
                y += 128; ql += 64; qh += 32; sc += 8;
            }
        }
    }
}

/* ---- quantized dot products, in the reference's AVX2 lane order ------------------------------------------- */
/* hsum_float_8, cpp/ggml/src/ggml-quants.c:47-53 */
static float hsum8(const float * a) {
    const float r0 = a[4] + a[0], r1 = a[5] + a[1], r2 = a[6] + a[2], r3 = a[7] + a[3];
    const float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}
/* the 8 scale bytes and 8 min bytes of a Q4_K/Q5_K super-block (utmp shuffle, ggml-quants.c:6925-6930) */
static void k4_all(const uint8_t * scales, int * sc8, int * m8) {
    for (int j = 0; j < 8; j++) scale_min_k4(j, scales, &sc8[j], &m8[j]);
}

/* ggml_vec_dot_q4_K_q8_K (AVX2 :6914-6977) and ggml_vec_dot_q5_K_q8_K (AVX2 :7487-7560):
 * lane m of the 8-wide int32 accumulator collects bytes 4m..4m+3 of every 32-byte group. */
static float vec_dot_q45_K(int n, const uint8_t * x, const uint8_t * y, int q5) {
    const int nb = n / QK_K, bb = q5 ? 176 : 144;
    float acc[8] = {0}, acc_m[4] = {0}, summs = 0.f;
    for (int i = 0; i < nb; i++, x += bb, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int16_t bsums[16]; memcpy(bsums, y + 260, 32);
        const float d = yd * h2f(rd16(x)), dmin = -yd * h2f(rd16(x + 2));
        int sc8[8], m8[8]; k4_all(x + 4, sc8, m8);
        const uint8_t * qh = x + 16;
        const uint8_t * q4 = q5 ? x + 48 : x + 16;
        /* mins: prod lane l = m[2l]*(bsums pair 2l) + m[2l+1]*(bsums pair 2l+1) */
        int prod[4];
        for (int l = 0; l < 4; l++)
            prod[l] = m8[2 * l] * (bsums[4 * l] + bsums[4 * l + 1]) + m8[2 * l + 1] * (bsums[4 * l + 2] + bsums[4 * l + 3]);
        if (!q5) for (int l = 0; l < 4; l++) acc_m[l] = fmaf(dmin, (float) prod[l], acc_m[l]);
        else     summs += dmin * (float)(prod[0] + prod[1] + prod[2] + prod[3]);
        int sumi[8] = {0};
        for (int j = 0; j < 4; j++) {
            for (int m = 0; m < 8; m++) {
                int lo = 0, hi = 0;
                for (int t = 0; t < 4; t++) {
                    const int l = 4 * m + t;
                    int ql = q4[32 * j + l] & 0xF, qhh = q4[32 * j + l] >> 4;
                    if (q5) { ql += ((qh[l] >> (2 * j)) & 1) << 4; qhh += ((qh[l] >> (2 * j + 1)) & 1) << 4; }
                    lo += ql * q8[64 * j + l];
                    hi += qhh * q8[64 * j + 32 + l];
                }
                sumi[m] += sc8[2 * j] * lo + sc8[2 * j + 1] * hi;
            }
        }
        for (int m = 0; m < 8; m++) acc[m] = fmaf(d, (float) sumi[m], acc[m]);
    }
    if (!q5) {
        const float a0 = acc_m[0] + acc_m[2], a1 = acc_m[1] + acc_m[3];
        return hsum8(acc) + (a0 + a1);
    }
    return hsum8(acc) + summs;
}
/* ggml_vec_dot_q6_K_q8_K, AVX2 :8145-8220: lanes 0-3 use scale 2k, lanes 4-7 scale 2k+1 of each 32-group */
static float vec_dot_q6_K(int n, const uint8_t * x, const uint8_t * y) {
    const int nb = n / QK_K;
    float acc[8] = {0};
    for (int i = 0; i < nb; i++, x += 210, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        const float d = yd * h2f(rd16(x + 208));
        const uint8_t * ql = x, * qh = x + 128; const int8_t * sc = (const int8_t *)(x + 192);
        int sumi[8] = {0};
        for (int j = 0; j < 2; j++) {
            for (int g = 0; g < 4; g++) {
                for (int m = 0; m < 8; m++) {
                    int s = 0;
                    for (int t = 0; t < 4; t++) {
                        const int l = 4 * m + t;
                        const uint8_t qlb = ql[64 * j + 32 * (g & 1) + l];
                        const int lo = (g >> 1) ? (qlb >> 4) : (qlb & 0xF);
                        const int q = (lo | (((qh[32 * j + l] >> (2 * g)) & 3) << 4)) - 32;
                        s += q * q8[128 * j + 32 * g + l];
                    }
                    sumi[m] += sc[8 * j + 2 * g + (m >> 2)] * s;
                }
            }
        }
        for (int m = 0; m < 8; m++) acc[m] = fmaf(d, (float) sumi[m], acc[m]);
    }
    return hsum8(acc);
}
/* ggml_vec_dot_q8_0_q8_0, AVX2 :5361-5382 (and tinyBLAS_Q0_AVX, cpp/ggml/src/llamafile/sgemm.cpp, which the
 * model path takes for Q8_0 x Q8_0: same 8-lane fma(d_w*d_x, sum4, acc) structure) */
static float vec_dot_q8_0(int n, const uint8_t * x, const uint8_t * y) {
    float acc[8] = {0};
    for (int i = 0; i < n / 32; i++, x += 34, y += 34) {
        const float d = h2f(rd16(x)) * h2f(rd16(y));
        for (int m = 0; m < 8; m++) {
            int s = 0;
            for (int t = 0; t < 4; t++) s += (int8_t) x[2 + 4 * m + t] * (int8_t) y[2 + 4 * m + t];
            acc[m] = fmaf(d, (float) s, acc[m]);
        }
    }
    return hsum8(acc);
}

static int64_t row_bytes(int type, int64_t k) {
    switch (type) { case 0: return k * 4; case 1: return k * 2; case 8: return k / 32 * 34; case 12: return k / 256 * 144;
                    case 13: return k / 256 * 176; case 14: return k / 256 * 210; default: return 0; }
}

/* ggml_compute_forward_mul_mat for one activation row (cpp/ggml/src/ggml.c:12277-12490): quantize x to the
 * weight type's vec_dot_type (Q8_K for K-quants, Q8_0 for Q8_0; type_traits cpp/ggml/src/ggml.c:769-855), then one
 * vec_dot per weight row. y[n_rows]. */
void port_mul_mat_vec(int type, const uint8_t * w, int64_t n_rows, int64_t k, const float * x, float * y) {
    const int64_t rb = row_bytes(type, k);
    uint8_t * xq = (uint8_t *) malloc((size_t)(type == 8 ? k / 32 * 34 : k / 256 * 292));
    if (type == 8) port_quantize_row_q8_0(x, xq, k); else port_quantize_row_q8_K(x, xq, k);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; r++) {
        const uint8_t * wr = w + r * rb;
        y[r] = type == 8 ? vec_dot_q8_0((int) k, wr, xq) : type == 14 ? vec_dot_q6_K((int) k, wr, xq)
                                                                       : vec_dot_q45_K((int) k, wr, xq, type == 13);
    }
    free(xq);
}

/* ---- float ops ------------------------------------------------------------------------------------------------ */
/* ggml_compute_forward_rms_norm_f32 (cpp/ggml/src/ggml.c:11850-11896) then ggml_mul by the weight
 * (llm_build_norm, cpp/src/llama.cpp:7928-7958) */
void port_rms_norm(const float * x, const float * w, int64_t k, float eps, float * y) {
    double sum = 0.0;
    for (int64_t i = 0; i < k; i++) sum += (double)(x[i] * x[i]);
    const float mean = (float)(sum / (double) k);
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = 0; i < k; i++) { const float v = x[i] * scale; y[i] = w ? v * w[i] : v; }
}

/* ggml_rope_yarn_corr_dim(s) (cpp/ggml/src/ggml.c:14011-14041) */
static float yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float) M_PI)) / (2 * logf(base));
}
/* ggml_rope_cache_init + rope_yarn + NORM-mode rotation (cpp/ggml/src/ggml.c:13987-14031, 14121-14135): theta starts
 * at pos and is multiplied by theta_scale per pair; with ext_factor != 0 the angle is the YaRN mix of the interpolated
 * and the extrapolated one and the magnitude gets the 0.1*ln(1/freq_scale) correction; cos/sin carry attn_factor.
 * beta_fast = 32, beta_slow = 1 are the context defaults (cpp/src/llama.cpp:16441-16442). */
void port_rope_ext(float * x, int n_heads, int head_dim, int pos, float freq_base, float freq_scale, const float * ff,
                   float ext_factor, float attn_factor, int n_ctx_orig) {
    const float theta_scale = powf(freq_base, -2.0f / head_dim);
    float corr[2];
    {
        const float start = floorf(yarn_corr_dim(head_dim, n_ctx_orig, 32.0f, freq_base));
        const float end   = ceilf(yarn_corr_dim(head_dim, n_ctx_orig, 1.0f, freq_base));
        corr[0] = start > 0 ? start : 0;
        corr[1] = end < head_dim - 1 ? end : head_dim - 1;
    }
    for (int h = 0; h < n_heads; h++) {
        float theta = (float) pos;
        float * p = x + (int64_t) h * head_dim;
        for (int i0 = 0; i0 < head_dim; i0 += 2) {
            const float f = ff ? ff[i0 / 2] : 1.0f;
            const float theta_extrap = theta / f;
            const float theta_interp = freq_scale * theta_extrap;
            float th = theta_interp, mscale = attn_factor;
            if (ext_factor != 0.0f) {
                const float hl = corr[1] - corr[0];
                const float y = (i0 / 2 - corr[0]) / (0.001f > hl ? 0.001f : hl);
                const float y01 = y < 0 ? 0 : (y > 1 ? 1 : y);          /* MIN(1, MAX(0, y)) */
                const float ramp_mix = (1 - y01) * ext_factor;
                th = theta_interp * (1 - ramp_mix) + theta_extrap * ramp_mix;
                mscale *= 1.0f + 0.1f * logf(1.0f / freq_scale);
            }
            const float c = cosf(th) * mscale, s = sinf(th) * mscale;
            const float x0 = p[i0], x1 = p[i0 + 1];
            p[i0] = x0 * c - x1 * s;
            p[i0 + 1] = x0 * s + x1 * c;
            theta *= theta_scale;
        }
    }
}
void port_rope(float * x, int n_heads, int head_dim, int pos, float freq_base, float freq_scale, const float * ff) {
    port_rope_ext(x, n_heads, head_dim, pos, freq_base, freq_scale, ff, 0.0f, 1.0f, 4096);
}

/* ---- the reference's SIMD exp/silu, lane-exact -------------------------------------------------------------
 * ggml_v_expf, AVX-512 variant (cpp/ggml/src/ggml.c:2447-2472): every step is an element-wise IEEE operation
 * (fmadd / fnmadd / mul / scalef), restated here with fmaf and ldexpf. The AVX2 variant of the reference
 * (:2487-) scales differently in its last step and is NOT bit-identical; this port follows the "native"
 * (AVX-512) build, which is what Booster's -march=native produces on the hosts in question. */
static float v_expf(float x) {
    const float r = 0x1.8p23f;
    const float z = fmaf(x, 0x1.715476p+0f, r);
    const float n = z - r;
    const float b = fmaf(-n, 0x1.7f7d1cp-20f, fmaf(-n, 0x1.62e4p-1f, x));
    const int   big = fabsf(n) > 192.f;
    const float u = b * b;
    const float j = fmaf(fmaf(fmaf(0x1.0e4020p-7f, b, 0x1.573e2ep-5f), u, fmaf(0x1.555e66p-3f, b, 0x1.fffdb6p-2f)), u,
                         fmaf(0x1.ffffecp-1f, b, 1.0f));
    if (big) return n <= 0.f ? 0.f : INFINITY;       /* also catches x = -inf (n = -inf) */
    return ldexpf(j, (int) n);                        /* _mm512_scalef_ps(j, n): exact scaling by 2^n */
}
/* ggml_v_silu (cpp/ggml/src/ggml.c:2475-2482): x / (1 + exp(0 - x)) */
static float silu(float x) { return x / (1.0f + v_expf(0.0f - x)); }

/* _mm512_reduce_add_ps as GCC expands it (avx512fintrin.h __MM512_REDUCE_OP): upper+lower 256, upper+lower 128,
 * then (0+2), (1+3), and the final pair */
static float reduce_add16(const float * a) {
    float t3[8], t6[4];
    for (int i = 0; i < 8; i++) t3[i] = a[8 + i] + a[i];
    for (int i = 0; i < 4; i++) t6[i] = t3[4 + i] + t3[i];
    const float t80 = t6[0] + t6[2], t81 = t6[1] + t6[3];
    return t80 + t81;
}

/* ---- the model ------------------------------------------------------------------------------------------------ */
typedef struct { int32_t type; int32_t pad; const uint8_t * data; int64_t rows, k; } port_mat;
typedef struct {
    port_mat wq, wk, wv, wo, gate, up, down;
    const float * attn_norm; const float * ffn_norm;
} port_layer;
typedef struct {
    int32_t n_layer, n_embd, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx;
    float rms_eps, rope_freq_base, rope_freq_scale;
    float yarn_ext_factor, yarn_attn_factor;   /* cparams.yarn_* (cpp/src/llama.cpp:16686-16690); 0 and 1 without YaRN */
    int32_t n_ctx_orig;
    const float * rope_freq_factors;
    port_mat tok_embd, output;
    const float * output_norm;
    const port_layer * layers;
    uint16_t * k_cache;     /* [n_layer][n_ctx][kv_dim] f16 bits, K post-RoPE (llm_build_kv_store, llama.cpp:7849-7853) */
    uint16_t * v_cache;     /* same layout (the reference's non-FA cache is transposed; layout is not arithmetic) */
    float * tap_l_out;      /* optional [n_layer][n_embd]: l_out of the LAST token of the call */
    float * tap_q;          /* optional [n_layer][n_head*head_dim]: Qcur (post-RoPE) of the last token */
    float * tap_kqv;        /* optional [n_layer][n_head*head_dim]: kqv_merged_cont of the last token */
} port_model;

/* default (non-flash) attention for one query token at position pos (llm_build_kqv, cpp/src/llama.cpp:8248-8297),
 * in the exact operation order of the reference's AVX-512 build:
 *   kq   batch 1 : tinyBLAS<16> F16xF32 (cpp/ggml/src/llamafile/sgemm.cpp:408-430): 16 lanes, lane c chains
 *                  fma(k[16s+c], q[16s+c], acc) over s, then _mm512_reduce_add_ps
 *        batch>1 : q rounded to f16, ggml_vec_dot_f16 (cpp/ggml/src/ggml.c:2038-2075): 4 accumulators x 16 lanes
 *                  over steps of 64, reduced (0+2),(1+3),(0+1), then _mm512_reduce_add_ps
 *   soft_max_ext : x*scale (+mask), max, ggml_v_expf(x-max) per 16 lanes, double sum of the per-vector
 *                  _mm512_reduce_add_ps, then p *= (float)(1/sum)   (cpp/ggml/src/ggml.c:13682-13778, 2619-2640)
 *   kqv          : tinyBLAS<16> over the (32-padded) kv length: lane c chains fma(v[16s+c][d], p[16s+c], acc)
 * n_kv is padded to 32 by the reference (kv_self.n, cpp/src/llama.cpp:14698); padded slots have p = 0. */
static void attention_impl(const float * q, const uint16_t * kc, const uint16_t * vc, int n_kv, int n_head, int n_head_kv,
                           int hd, float scale, int round_q, const int32_t * cell_pos, int pos, float * out);
void port_attention(const float * q, const uint16_t * kc, const uint16_t * vc, int n_kv, int n_head, int n_head_kv,
                    int hd, float scale, int round_q, float * out) {
    attention_impl(q, kc, vc, n_kv, n_head, n_head_kv, hd, scale, round_q, NULL, 0, out);
}
/* cell_pos != NULL: the KV cells are managed like struct llama_kv_cache after llama_kv_cache_seq_rm / seq_add / seq_div — cell t
 * holds position cell_pos[t] (-1: empty) and the query at position pos sees the cells with 0 <= cell_pos[t] <= pos; every other
 * cell gets -INFINITY from KQ_mask (llama_set_inputs, cpp/src/llama.cpp:14132-14200) and so probability 0, which the P.V chains
 * still step through (fma(v, 0, acc)), in CELL order. */
static void attention_impl(const float * q, const uint16_t * kc, const uint16_t * vc, int n_kv, int n_head, int n_head_kv,
                           int hd, float scale, int round_q, const int32_t * cell_pos, int pos, float * out) {
    const int kvd = n_head_kv * hd, gqa = n_head / n_head_kv;
    const int n_pad = (n_kv + 31) / 32 * 32;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < n_head; h++) {
        const int g = h / gqa;
        float * p = (float *) malloc(sizeof(float) * (size_t) n_pad);
        float qh[512];
        for (int d = 0; d < hd; d++) qh[d] = round_q ? h2f(f2h(q[h * hd + d])) : q[h * hd + d];
        float max = -INFINITY;
        for (int t = 0; t < n_pad; t++) {
            if (t >= n_kv) { p[t] = -INFINITY; continue; }
            if (cell_pos && (cell_pos[t] < 0 || cell_pos[t] > pos)) { p[t] = -INFINITY; continue; }
            const uint16_t * kr = kc + (int64_t) t * kvd + g * hd;
            float s;
            if (!round_q) {
                float acc[16] = {0};
                for (int l = 0; l < hd; l += 16)
                    for (int c = 0; c < 16; c++) acc[c] = fmaf(h2f(kr[l + c]), qh[l + c], acc[c]);
                s = reduce_add16(acc);
            } else {
                float acc[4][16] = {{0}};
                for (int i = 0; i < hd; i += 64)
                    for (int j = 0; j < 4; j++)
                        for (int c = 0; c < 16; c++) acc[j][c] = fmaf(h2f(kr[i + 16 * j + c]), qh[i + 16 * j + c], acc[j][c]);
                float r[16];
                for (int c = 0; c < 16; c++) r[c] = (acc[0][c] + acc[2][c]) + (acc[1][c] + acc[3][c]);
                s = reduce_add16(r);
            }
            p[t] = s * scale;
            if (p[t] > max) max = p[t];
        }
        double sum = 0.0;
        for (int t0 = 0; t0 < n_pad; t0 += 16) {
            float v[16];
            for (int c = 0; c < 16; c++) { v[c] = v_expf(p[t0 + c] - max); p[t0 + c] = v[c]; }
            sum += (double) reduce_add16(v);
        }
        const float inv = (float) (1.0 / sum);
        for (int t = 0; t < n_pad; t++) p[t] *= inv;
        for (int d = 0; d < hd; d++) {
            float acc[16] = {0};
            for (int t0 = 0; t0 < n_kv; t0 += 16)
                for (int c = 0; c < 16 && t0 + c < n_kv; c++)
                    acc[c] = fmaf(h2f(vc[(int64_t) (t0 + c) * kvd + g * hd + d]), p[t0 + c], acc[c]);
            out[h * hd + d] = reduce_add16(acc);
        }
        free(p);
    }
}

static void attention(const port_model * M, int il, const float * q, int pos, int round_q, float * out) {
    const int kvd = M->n_head_kv * M->head_dim;
    port_attention(q, M->k_cache + (int64_t) il * M->n_ctx * kvd, M->v_cache + (int64_t) il * M->n_ctx * kvd, pos + 1,
                   M->n_head, M->n_head_kv, M->head_dim, 1.0f / sqrtf((float) M->head_dim), round_q, out);
}

/* llama_decode(ctx, llama_batch_get_one(tokens, n, pos0, 0)) + llama_get_logits for LLM_ARCH_LLAMA
 * (llama_decode_internal cpp/src/llama.cpp:14537-14840; graph build_llama :8781-8925). Tokens are processed one
 * after the other — per-token arithmetic of a batch is independent in the reference except for the f16 rounding
 * of q when n > 1. logits[n_vocab] = last token's row. Returns 0, or 1 if positions exceed n_ctx. */
static int decode_impl(const port_model * M, const int32_t * tokens, int n, int pos0, const int32_t * cell_of,
                       const int32_t * cell_pos, int n_kv_cells, float * logits);
int port_decode(const port_model * M, const int32_t * tokens, int n, int pos0, float * logits) {
    return decode_impl(M, tokens, n, pos0, NULL, NULL, 0, logits);
}
/* the same with managed KV cells (after a context shift / Self-Extend, cpp/bridge.cpp:487-524): token t of the batch is written
 * to cell cell_of[t] (llama_kv_cache_find_slot, cpp/src/llama.cpp:3028-3125), cell_pos[n_ctx] holds every cell's position AFTER
 * the batch was placed, n_kv_cells = llama_kv_cache_cell_max (:3397-3407; padded to 32 inside the attention like kv_self.n). */
int port_decode_cells(const port_model * M, const int32_t * tokens, int n, int pos0, const int32_t * cell_of,
                      const int32_t * cell_pos, int n_kv_cells, float * logits) {
    return decode_impl(M, tokens, n, pos0, cell_of, cell_pos, n_kv_cells, logits);
}
/* llama_kv_cache_update -> build_k_shift (cpp/src/llama.cpp:15245-15277, 8482-8510): EVERY cell's cached K row (post-RoPE, f16)
 * is rotated in place by that cell's accumulated position delta — ggml_compute_forward_rope_f16 (cpp/ggml/src/ggml.c:14169-14291):
 * f16 -> f32, the un-fused rotation, f32 -> f16; delta 0 is cos = attn_factor-scaled 1, sin = 0. */
void port_k_shift(const port_model * M, const int32_t * delta) {
    const int hd = M->head_dim, KVD = M->n_head_kv * hd;
    float * row = malloc(sizeof(float) * KVD);
    for (int il = 0; il < M->n_layer; il++) {
        for (int i = 0; i < M->n_ctx; i++) {
            uint16_t * kc = M->k_cache + ((int64_t) il * M->n_ctx + i) * KVD;
            for (int j = 0; j < KVD; j++) row[j] = h2f(kc[j]);
            port_rope_ext(row, M->n_head_kv, hd, delta[i], M->rope_freq_base, M->rope_freq_scale, M->rope_freq_factors,
                          M->yarn_ext_factor, M->yarn_attn_factor, M->n_ctx_orig);
            for (int j = 0; j < KVD; j++) kc[j] = f2h(row[j]);
        }
    }
    free(row);
}
static int decode_impl(const port_model * M, const int32_t * tokens, int n, int pos0, const int32_t * cell_of,
                       const int32_t * cell_pos, int n_kv_cells, float * logits) {
    const int E = M->n_embd, hd = M->head_dim, QD = M->n_head * hd, KVD = M->n_head_kv * hd, FF = M->n_ff;
    if (pos0 < 0 || (!cell_of && pos0 + n > M->n_ctx)) return 1;
    float * x = malloc(sizeof(float) * E), * nx = malloc(sizeof(float) * E), * q = malloc(sizeof(float) * QD);
    float * kk = malloc(sizeof(float) * KVD), * vv = malloc(sizeof(float) * KVD), * att = malloc(sizeof(float) * QD);
    float * tmp = malloc(sizeof(float) * E), * g = malloc(sizeof(float) * FF), * u = malloc(sizeof(float) * FF);
    const int round_q = n > 1;
    for (int t = 0; t < n; t++) {
        const int pos = pos0 + t, last = t == n - 1;
        /* inp_embd = get_rows(tok_embd, token): cpp/src/llama.cpp:7802-7828 */
        port_dequantize_row(M->tok_embd.type, M->tok_embd.data + (int64_t) tokens[t] * row_bytes(M->tok_embd.type, E), x, E);
        for (int il = 0; il < M->n_layer; il++) {
            const port_layer * L = &M->layers[il];
            port_rms_norm(x, L->attn_norm, E, M->rms_eps, nx);
            port_mul_mat_vec(L->wq.type, L->wq.data, QD, E, nx, q);
            port_mul_mat_vec(L->wk.type, L->wk.data, KVD, E, nx, kk);
            port_mul_mat_vec(L->wv.type, L->wv.data, KVD, E, nx, vv);
            port_rope_ext(q, M->n_head, hd, pos, M->rope_freq_base, M->rope_freq_scale, M->rope_freq_factors,
                          M->yarn_ext_factor, M->yarn_attn_factor, M->n_ctx_orig);
            port_rope_ext(kk, M->n_head_kv, hd, pos, M->rope_freq_base, M->rope_freq_scale, M->rope_freq_factors,
                          M->yarn_ext_factor, M->yarn_attn_factor, M->n_ctx_orig);
            const int cell = cell_of ? cell_of[t] : pos;
            uint16_t * kc = M->k_cache + ((int64_t) il * M->n_ctx + cell) * KVD;
            uint16_t * vc = M->v_cache + ((int64_t) il * M->n_ctx + cell) * KVD;
            for (int i = 0; i < KVD; i++) { kc[i] = f2h(kk[i]); vc[i] = f2h(vv[i]); }
            if (cell_of)
                attention_impl(q, M->k_cache + (int64_t) il * M->n_ctx * KVD, M->v_cache + (int64_t) il * M->n_ctx * KVD, n_kv_cells,
                               M->n_head, M->n_head_kv, hd, 1.0f / sqrtf((float) hd), round_q, cell_pos, pos, att);
            else
                attention(M, il, q, pos, round_q, att);
            if (last && M->tap_q)   memcpy(M->tap_q + (int64_t) il * QD, q, sizeof(float) * QD);
            if (last && M->tap_kqv) memcpy(M->tap_kqv + (int64_t) il * QD, att, sizeof(float) * QD);
            port_mul_mat_vec(L->wo.type, L->wo.data, E, QD, att, tmp);
            for (int i = 0; i < E; i++) x[i] = tmp[i] + x[i];                 /* ffn_inp = cur + inpSA, :8865 */
            port_rms_norm(x, L->ffn_norm, E, M->rms_eps, nx);
            port_mul_mat_vec(L->up.type, L->up.data, FF, E, nx, u);          /* llm_build_ffn :7960-8085 */
            port_mul_mat_vec(L->gate.type, L->gate.data, FF, E, nx, g);
            for (int i = 0; i < FF; i++) g[i] = silu(g[i]) * u[i];
            port_mul_mat_vec(L->down.type, L->down.data, E, FF, g, tmp);
            for (int i = 0; i < E; i++) x[i] = tmp[i] + x[i];                 /* l_out = cur + ffn_inp, :8901 */
            if (last && M->tap_l_out) memcpy(M->tap_l_out + (int64_t) il * E, x, sizeof(float) * E);
        }
        if (last && logits) {
            port_rms_norm(x, M->output_norm, E, M->rms_eps, nx);
            port_mul_mat_vec(M->output.type, M->output.data, M->n_vocab, E, nx, logits);
        }
    }
    free(x); free(nx); free(q); free(kk); free(vv); free(att); free(tmp); free(g); free(u);
    return 0;
}
