// engine.cu — the B200 decode engine behind include/booster_b200.h.
//
// A fixed-function LLaMA-architecture decoder (not a graph interpreter): the per-token work of
// llama_decode_internal (cpp/src/llama.cpp:14537-14840) + build_llama (:8781-8925) is a fixed sequence of
// 6 fused kernels per layer, captured once into a CUDA graph and replayed per token; the per-token scalars
// (token, pos) live in device memory (DecodeState) so the graph never needs re-instantiation.
//
//   per layer:  k_matvec<QKV>   [RMSNorm(attn_norm)+Q8_K quant] -> wq|wk|wv -> RoPE -> q (f32), K/V (f16 cache)
//               k_attn_scores      raw K.q scores of the f16 cache (GQA group per CTA, 64 positions each)
//               k_attn_softmax_pv  soft-max of the score rows + P.V -> kqv_merged_cont
//               k_matvec<RESID> [quant] -> wo -> + residual                          (ffn_inp)
//               k_matvec<SILU>  [RMSNorm(ffn_norm)+quant] -> gate|up -> silu(g)*u    (ffn_gate_par)
//               k_matvec<RESID> [quant] -> down -> + residual                        (l_out)
//   head:       k_matvec<STORE> [RMSNorm(output_norm)+quant] -> output -> logits
//
// There is no CPU fallback: every entry point fails with an error when no CUDA device is present.
#include "../../include/booster_b200.h"
#include "gguf.hpp"
#include "tokenizer.hpp"
#include "kernels.cuh"
#ifndef B200_TOKEN_KERNEL
#define B200_TOKEN_KERNEL 1      // 0: build without the persistent per-token kernel (A/B: code size of the module)
#endif
#include "token_kernel.cuh"
#include "prefill.cuh"
#include "prefill_mma.cuh"
#include "prefill_umma.cuh"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <map>
#include <mutex>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

using namespace b200;

// ------------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int set_err(const std::string & s) { g_err = s; return 1; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)

extern "C" const char * b200_last_error(void) { return g_err.c_str(); }
extern "C" const char * b200_version(void) { return "booster_b200 0.1 (sm_100a)"; }
extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
static void require_gpu() {
    if (b200_device_count() <= 0)
        throw std::runtime_error("no CUDA device visible: booster_b200 has no CPU fallback (the CUDA path is the product)");
}

// ------------------------------------------------------------------------------------------------------------
// re-tiling kernels: ggml block layout -> tiles (one thread per block; one-time at load). `vstride`/`voff` place
// a tensor's row r at virtual row r*vstride + voff of the destination (gate/up interleave: vstride 2, voff 0/1).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy16(uint8_t * dst, const uint8_t * src) {
    for (int j = 0; j < 16; j++) dst[j] = src[j];
}
__global__ void k_retile(int type, const uint8_t * __restrict__ src, int64_t n_rows, int nb, int vstride, int voff,
                         uint8_t * tiles) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;   // source block index = row*nb + bi
    if (i >= n_rows * nb) return;
    const int64_t row = i / nb, bi = i % nb;
    const int64_t vrow = row * vstride + voff;                            // destination (virtual) row
    const int64_t T = (vrow >> 5) * nb + bi; const int lane = (int) (vrow & 31);   // tile = block bi of 32 rows, one row per lane
    uint8_t * p0 = tiles + (size_t) T * tile_bytes_of(type);    // the tile is one contiguous chunk; field offsets below
    uint8_t * p1 = p0 + 4096, * p2 = p0 + 4608;
    uint8_t * p3 = p0 + (type == T_Q6_K ? 6656 : 1024);
    if (type == T_Q4_K) {
        const uint8_t * b = src + i * 144;   // {half d, half dmin, u8 scales[12], u8 qs[128]}  ggml-common.h:267-277
        for (int c = 0; c < 8; c++) copy16(p0 + ((size_t) c * 32 + lane) * 16, b + 16 + 16 * c);
        for (int j = 0; j < 12; j++) p1[(size_t) lane * 16 + j] = b[4 + j];        // sd: scales[12] | d | dmin
        for (int j = 0; j < 4; j++) p1[(size_t) lane * 16 + 12 + j] = b[j];
    } else if (type == T_Q5_K) {
        const uint8_t * b = src + i * 176;   // {half d, half dmin, u8 scales[12], u8 qh[32], u8 qs[128]}  :284-295
        for (int c = 0; c < 8; c++) copy16(p0 + ((size_t) c * 32 + lane) * 16, b + 48 + 16 * c);
        for (int c = 0; c < 2; c++) copy16(p2 + ((size_t) c * 32 + lane) * 16, b + 16 + 16 * c);
        for (int j = 0; j < 12; j++) p1[(size_t) lane * 16 + j] = b[4 + j];
        for (int j = 0; j < 4; j++) p1[(size_t) lane * 16 + 12 + j] = b[j];
    } else if (type == T_Q6_K) {
        const uint8_t * b = src + i * 210;   // {u8 ql[128], u8 qh[64], i8 scales[16], half d}  :302-307
        for (int c = 0; c < 8; c++) copy16(p0 + ((size_t) c * 32 + lane) * 16, b + 16 * c);
        for (int c = 0; c < 4; c++) copy16(p2 + ((size_t) c * 32 + lane) * 16, b + 128 + 16 * c);
        for (int j = 0; j < 16; j++) p1[(size_t) lane * 16 + j] = b[192 + j];
        p3[lane * 2] = b[208]; p3[lane * 2 + 1] = b[209];
    } else {                                 // T_Q8_0: {half d, i8 qs[32]}  :186-190
        const uint8_t * b = src + i * 34;
        for (int c = 0; c < 2; c++) copy16(p0 + ((size_t) c * 32 + lane) * 16, b + 2 + 16 * c);
        p3[lane * 2] = b[0]; p3[lane * 2 + 1] = b[1];
    }
}

// ------------------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------------------
struct DevMat {
    TMat m;
    void * alloc = nullptr;
    size_t bytes = 0;       // algorithmic bytes (== ggml tensor bytes)
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct HostTensor { int type; const void * data; int64_t rows; int64_t k; };

// upload n_src tensors of identical type/shape (1, or 2 for the gate/up interleave) as ONE tiled virtual matrix
static DevMat upload_tiled(const HostTensor * src, int n_src, cudaStream_t st) {
    const int type = src[0].type;
    const int64_t k = src[0].k, rows1 = src[0].rows;
    for (int i = 1; i < n_src; i++)
        if (src[i].type != type || src[i].k != k || src[i].rows != rows1) throw std::runtime_error("interleaved tensors must share type and shape");
    if (k % 256 != 0) throw std::runtime_error("matrix inner dimension must be a multiple of 256 (got " + std::to_string(k) + ")");
    int blk_bytes = 0, wpb = 256;
    switch (type) {
        case T_Q4_K: blk_bytes = 144; break;
        case T_Q5_K: blk_bytes = 176; break;
        case T_Q6_K: blk_bytes = 210; break;
        case T_Q8_0: blk_bytes = 34; wpb = 32; break;
        default: throw std::runtime_error("unsupported matrix type " + std::to_string(type) + " (supported: Q4_K, Q5_K, Q6_K, Q8_0)");
    }
    DevMat d;
    const int nb = (int) (k / wpb);
    const int64_t vrows = rows1 * n_src;
    if (vrows % 32 != 0) throw std::runtime_error("row count " + std::to_string(vrows) + " is not a multiple of the work-unit height 32");
    d.m.type = type; d.m.n_rows = (int) vrows; d.m.nb = nb; d.m.rows_unit = 32;
    d.m.tiles_unit = nb; d.m.n_units = (int) (vrows / 32); d.m.tile_bytes = tile_bytes_of(type);
    const size_t n_tiles = (size_t) vrows * nb / 32;
    const size_t raw1 = (size_t) blk_bytes * nb * rows1;
    d.bytes = raw1 * n_src;
    uint8_t * base = nullptr;
    CU(cudaMalloc(&base, align_up((size_t) tile_bytes_of(type) * n_tiles, 256)));
    d.alloc = base;
    uint8_t * tmp = nullptr;
    try {
        CU(cudaMalloc(&tmp, raw1));
        for (int i = 0; i < n_src; i++) {
            CU(cudaMemcpyAsync(tmp, src[i].data, raw1, cudaMemcpyHostToDevice, st));
            const int64_t n_blocks = rows1 * nb;
            k_retile<<<(unsigned) ((n_blocks + 127) / 128), 128, 0, st>>>(type, tmp, rows1, nb, n_src, i, base);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(st));
        }
    } catch (...) { cudaFree(tmp); cudaFree(base); throw; }
    CU(cudaFree(tmp));
    d.m.p0 = base;
    return d;
}
// n tensors with the same inner dimension stacked row-wise in ONE allocation; adjacent tensors of the same block type
// form one segment (wq|wk always, wq|wk|wv when attn_v is not bumped to another type), so the kernel's unit lookup is
// arithmetic on at most two segments
struct DevStack {
    void * alloc = nullptr;
    size_t bytes = 0;          // algorithmic bytes
    TMat seg[3];
    int n_seg = 0;
    int n_units = 0;
};
static DevStack upload_stack(const HostTensor * src, int n_src, cudaStream_t st) {
    DevStack d;
    const int64_t k = src[0].k;
    size_t total = 0;
    for (int i = 0; i < n_src; i++) {
        if (src[i].k != k) throw std::runtime_error("stacked tensors must share the inner dimension");
        if (src[i].rows % 32 != 0) throw std::runtime_error("row count is not a multiple of the work-unit height 32");
        const int wpb = src[i].type == T_Q8_0 ? 32 : 256;
        if (k % 256 != 0) throw std::runtime_error("matrix inner dimension must be a multiple of 256");
        total += (size_t) tile_bytes_of(src[i].type) * (size_t) (src[i].rows / 32) * (size_t) (k / wpb);
        if (i > 0 && ((src[i].type == T_Q8_0) != (src[0].type == T_Q8_0))) throw std::runtime_error("mixing Q8_0 and K-quant matrices inside one launch is not supported");
    }
    uint8_t * base = nullptr;
    CU(cudaMalloc(&base, align_up(total, 256)));
    d.alloc = base;
    size_t off = 0;
    uint8_t * tmp = nullptr;
    try {
    for (int i = 0; i < n_src; i++) {
        const int type = src[i].type;
        int blk_bytes = 0, wpb = 256;
        switch (type) {
            case T_Q4_K: blk_bytes = 144; break;
            case T_Q5_K: blk_bytes = 176; break;
            case T_Q6_K: blk_bytes = 210; break;
            case T_Q8_0: blk_bytes = 34; wpb = 32; break;
            default: throw std::runtime_error("unsupported matrix type " + std::to_string(type) + " (supported: Q4_K, Q5_K, Q6_K, Q8_0)");
        }
        const int nb = (int) (k / wpb);
        const size_t raw = (size_t) blk_bytes * nb * src[i].rows;
        CU(cudaMalloc(&tmp, raw));
        CU(cudaMemcpyAsync(tmp, src[i].data, raw, cudaMemcpyHostToDevice, st));
        const int64_t n_blocks = src[i].rows * nb;
        k_retile<<<(unsigned) ((n_blocks + 127) / 128), 128, 0, st>>>(type, tmp, src[i].rows, nb, 1, 0, base + off);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));
        CU(cudaFree(tmp));
        tmp = nullptr;
        d.bytes += raw;
        if (d.n_seg > 0 && d.seg[d.n_seg - 1].type == type) {
            TMat & m = d.seg[d.n_seg - 1];
            m.n_rows += (int) src[i].rows; m.n_units += (int) (src[i].rows / 32);
        } else {
            TMat & m = d.seg[d.n_seg++];
            m.type = type; m.n_rows = (int) src[i].rows; m.nb = nb; m.rows_unit = 32; m.tiles_unit = nb;
            m.n_units = (int) (src[i].rows / 32); m.tile_bytes = tile_bytes_of(type); m.p0 = base + off;
        }
        d.n_units += (int) (src[i].rows / 32);
        off += (size_t) tile_bytes_of(type) * (size_t) (src[i].rows / 32) * nb;
    }
    } catch (...) { cudaFree(tmp); cudaFree(base); throw; }
    return d;
}

static DevMat upload_matrix(int type, const void * host, int64_t n_rows, int64_t k, cudaStream_t st) {
    const HostTensor h{type, host, n_rows, k};
    return upload_tiled(&h, 1, st);
}

struct LayerW {
    DevStack qkv;                           // attn_q | attn_k | attn_v stacked row-wise
    DevMat wo, gateup, down;                // gateup = ffn_gate/ffn_up interleaved row by row
    float * attn_norm = nullptr;
    float * ffn_norm = nullptr;
};

struct b200_model {
    int device = 0;
    int n_vocab = 0, n_embd = 0, n_layer = 0, n_head = 0, n_head_kv = 0, n_ff = 0, head_dim = 0, n_ctx_train = 0;
    int layer_begin = 0, layer_end = 0;
    int ftype = -1;
    float rms_eps = 1e-5f;
    float rope_freq_base = 10000.f, rope_freq_scale = 1.f;
    int   rope_dim = 0;
    int   n_ctx_orig = 0;
    float yarn_ext_factor = 0.f, yarn_attn_factor = 1.f, yarn_beta_fast = 32.f, yarn_beta_slow = 1.f;
    std::vector<float> rope_freq_factors;   // rope_freqs.weight (llama 3.1), empty if absent
    std::vector<LayerW> layers;             // indexed by il - layer_begin
    // first stage
    int      embd_type = 0;
    uint8_t * embd_rows = nullptr;          // token_embd in ORIGINAL ggml layout (one row read per token)
    size_t   embd_row_bytes = 0;
    // last stage
    float *  output_norm = nullptr;
    DevMat   output;
    int64_t  weight_bytes = 0;
    // tokenizer metadata kept for the bridge (host side)
    std::string tok_model;
    std::vector<std::string> tok_tokens;
    int32_t tok_eos = -1, tok_bos = -1;
    bool has_embd() const { return layer_begin == 0; }
    bool has_head() const { return layer_end == n_layer; }
};

static float * upload_f32(const gguf_tensor * t, int64_t n, cudaStream_t st) {
    if (!t) throw std::runtime_error("missing norm tensor");
    if (t->type != 0) throw std::runtime_error("norm tensor " + t->name + " must be F32");
    if ((int64_t) t->ne[0] != n) throw std::runtime_error("norm tensor " + t->name + " has wrong length");
    float * d = nullptr;
    CU(cudaMalloc(&d, (size_t) n * 4));
    CU(cudaMemcpyAsync(d, t->data, (size_t) n * 4, cudaMemcpyHostToDevice, st));
    return d;
}

static DevMat upload_named(const gguf_file & g, const std::string & name, int64_t rows, int64_t k, cudaStream_t st) {
    const gguf_tensor * t = g.find(name);
    if (!t) throw std::runtime_error("missing tensor " + name);
    if ((int64_t) t->ne[0] != k || (int64_t) t->ne[1] != rows)
        throw std::runtime_error("tensor " + name + " has shape [" + std::to_string(t->ne[0]) + "," + std::to_string(t->ne[1]) +
                                 "], expected [" + std::to_string(k) + "," + std::to_string(rows) + "]");
    return upload_matrix((int) t->type, t->data, rows, k, st);
}

extern "C" void b200_model_free(b200_model * m);
extern "C" b200_model * b200_model_load(const char * path, int device, int layer_begin, int layer_end) {
    std::unique_ptr<b200_model> m;
    cudaStream_t st = nullptr;
    try {
        require_gpu();
        if (!path) throw std::runtime_error("null model path");
        gguf_file g;
        const std::string err = g.open(path);
        if (!err.empty()) throw std::runtime_error(err);
        const std::string arch = g.get_s("general.architecture", "");
        if (arch != "llama") throw std::runtime_error("unsupported architecture '" + arch + "' (this path covers LLM_ARCH_LLAMA only)");
        m = std::make_unique<b200_model>();
        m->device = device;
        CU(cudaSetDevice(device));
        // hparams: cpp/src/llama.cpp:4571-4700
        m->n_ctx_train = (int) g.get_u("llama.context_length", 0);
        m->n_embd      = (int) g.get_u("llama.embedding_length", 0);
        m->n_layer     = (int) g.get_u("llama.block_count", 0);
        m->n_ff        = (int) g.get_u("llama.feed_forward_length", 0);
        m->n_head      = (int) g.get_u("llama.attention.head_count", 0);
        m->n_head_kv   = (int) g.get_u("llama.attention.head_count_kv", (uint64_t) m->n_head);
        m->rms_eps     = (float) g.get_f("llama.attention.layer_norm_rms_epsilon", 1e-5);
        m->ftype       = (int) g.get_u("general.file_type", (uint64_t) -1);
        if (m->n_embd <= 0 || m->n_layer <= 0 || m->n_head <= 0 || m->n_head_kv <= 0 || m->n_ff <= 0) throw std::runtime_error("missing llama.* hyper-parameters");
        m->head_dim    = (int) g.get_u("llama.attention.key_length", (uint64_t) (m->n_embd / m->n_head));
        m->rope_dim    = (int) g.get_u("llama.rope.dimension_count", (uint64_t) m->head_dim);
        if (m->rope_dim != m->head_dim) throw std::runtime_error("n_rot != head_dim is not supported on this path");
        if (m->head_dim % 8 != 0 || m->head_dim > 256) throw std::runtime_error("head_dim must be a multiple of 8 and <= 256");
        if (m->n_head % m->n_head_kv != 0 || m->n_head / m->n_head_kv > ATT_MAX_GQA) throw std::runtime_error("unsupported GQA ratio");
        m->rope_freq_base = (float) g.get_f("llama.rope.freq_base", 10000.0);
        // rope scaling: cpp/src/llama.cpp:4630-4650, 16655-16690
        const std::string rs = g.get_s("llama.rope.scaling.type", "linear");
        float ropescale = (float) g.get_f("llama.rope.scaling.factor", 0.0);
        if (ropescale == 0.f) ropescale = (float) g.get_f("llama.rope.scale_linear", 0.0);
        m->rope_freq_scale = ropescale == 0.f ? 1.f : 1.f / ropescale;
        if (rs == "none") m->rope_freq_scale = 1.f;
        m->n_ctx_orig = (int) g.get_u("llama.rope.scaling.original_context_length", (uint64_t) m->n_ctx_train);
        m->yarn_ext_factor = rs == "yarn" ? 1.f : 0.f;
        m->yarn_attn_factor = (float) g.get_f("llama.rope.scaling.attn_factor", 1.0);
        const gguf_tensor * te = g.find("token_embd.weight");
        if (!te) throw std::runtime_error("missing token_embd.weight");
        if ((int64_t) te->ne[0] != m->n_embd) throw std::runtime_error("token_embd.weight rows must have n_embd elements");
        if (te->ne[1] == 0 || te->ne[1] > (uint64_t) INT32_MAX) throw std::runtime_error("bad vocabulary size");
        m->n_vocab = (int) te->ne[1];
        if (layer_end < 0 || layer_end > m->n_layer) layer_end = m->n_layer;
        if (layer_begin < 0 || layer_begin >= layer_end) throw std::runtime_error("bad layer range");
        m->layer_begin = layer_begin; m->layer_end = layer_end;

        m->tok_model = g.get_s("tokenizer.ggml.model", "no_vocab");
        auto itk = g.kv.find("tokenizer.ggml.tokens");
        if (itk != g.kv.end()) m->tok_tokens = itk->second.arr_s;
        m->tok_eos = (int32_t) g.get_u("tokenizer.ggml.eos_token_id", (uint64_t) -1);
        m->tok_bos = (int32_t) g.get_u("tokenizer.ggml.bos_token_id", (uint64_t) -1);

        CU(cudaStreamCreate(&st));
        const int E = m->n_embd, HD = m->head_dim, KV = m->n_head_kv * HD, Q = m->n_head * HD, FF = m->n_ff;
        if (const gguf_tensor * rf = g.find("rope_freqs.weight")) {
            if (rf->type != 0) throw std::runtime_error("rope_freqs.weight must be F32");
            if ((int64_t) rf->ne[0] < m->head_dim / 2) throw std::runtime_error("rope_freqs.weight is shorter than head_dim / 2");
            m->rope_freq_factors.assign((const float *) rf->data, (const float *) rf->data + rf->ne[0]);
        }
        int64_t wb = 0;
        for (int il = layer_begin; il < layer_end; il++) {
            const std::string p = "blk." + std::to_string(il) + ".";
            m->layers.emplace_back();          // registered first: a failure below frees what this layer already uploaded
            LayerW & L = m->layers.back();
            L.attn_norm = upload_f32(g.find(p + "attn_norm.weight"), E, st);
            L.ffn_norm  = upload_f32(g.find(p + "ffn_norm.weight"), E, st);
            {
                const char * nm[3] = { "attn_q.weight", "attn_k.weight", "attn_v.weight" };
                const int64_t rows[3] = { Q, KV, KV };
                HostTensor hs[3];
                for (int i = 0; i < 3; i++) {
                    const gguf_tensor * t = g.find(p + nm[i]);
                    if (!t) throw std::runtime_error("missing tensor " + p + nm[i]);
                    if ((int64_t) t->ne[0] != E || (int64_t) t->ne[1] != rows[i]) throw std::runtime_error("tensor " + p + nm[i] + " has an unexpected shape");
                    hs[i] = HostTensor{(int) t->type, t->data, rows[i], E};
                }
                L.qkv = upload_stack(hs, 3, st);
            }
            L.wo   = upload_named(g, p + "attn_output.weight", E, Q, st);
            {
                const gguf_tensor * tg = g.find(p + "ffn_gate.weight"), * tu = g.find(p + "ffn_up.weight");
                if (!tg || !tu) throw std::runtime_error("missing ffn_gate/ffn_up in layer " + std::to_string(il));
                if ((int64_t) tg->ne[0] != E || (int64_t) tg->ne[1] != FF || (int64_t) tu->ne[0] != E || (int64_t) tu->ne[1] != FF)
                    throw std::runtime_error("ffn_gate/ffn_up have unexpected shapes");
                if (tg->type != tu->type) throw std::runtime_error("ffn_gate and ffn_up must share a block type");
                const HostTensor hs[2] = { {(int) tg->type, tg->data, FF, E}, {(int) tu->type, tu->data, FF, E} };
                L.gateup = upload_tiled(hs, 2, st);
            }
            L.down = upload_named(g, p + "ffn_down.weight", E, FF, st);
            const bool q80 = L.qkv.seg[0].type == T_Q8_0;
            for (const DevMat * d : { &L.wo, &L.gateup, &L.down })
                if ((d->m.type == T_Q8_0) != q80) throw std::runtime_error("mixing Q8_0 and K-quant matrices inside one layer is not supported");
            wb += (int64_t) (L.qkv.bytes + L.wo.bytes + L.gateup.bytes + L.down.bytes) + 2 * (int64_t) E * 4;
        }
        if (m->has_embd()) {
            m->embd_type = (int) te->type;
            m->embd_row_bytes = (size_t) ggml_row_bytes(te->type, te->ne[0]);
            if (m->embd_row_bytes == 0) throw std::runtime_error("unsupported token_embd type " + std::to_string(te->type));
            CU(cudaMalloc(&m->embd_rows, te->nbytes));
            CU(cudaMemcpyAsync(m->embd_rows, te->data, te->nbytes, cudaMemcpyHostToDevice, st));
        }
        if (m->has_head()) {
            m->output_norm = upload_f32(g.find("output_norm.weight"), E, st);
            // tied embeddings: output falls back to token_embd (cpp/src/llama.cpp:6077-6084)
            const std::string on = g.find("output.weight") ? "output.weight" : "token_embd.weight";
            m->output = upload_named(g, on, m->n_vocab, E, st);
            wb += (int64_t) m->output.bytes + (int64_t) E * 4;
        }
        CU(cudaStreamSynchronize(st));
        CU(cudaStreamDestroy(st));
        st = nullptr;
        m->weight_bytes = wb;
        return m.release();
    } catch (const std::exception & e) {
        set_err(e.what());
        // a failed load (out of memory, bad tensor) must not keep the layers it already uploaded
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        if (m) b200_model_free(m.release());
        cudaGetLastError();
        return nullptr;
    }
}

extern "C" void b200_model_free(b200_model * m) {
    if (!m) return;
    cudaSetDevice(m->device);
    for (auto & L : m->layers) {
        cudaFree(L.qkv.alloc);
        for (DevMat * d : { &L.wo, &L.gateup, &L.down }) cudaFree(d->alloc);
        cudaFree(L.attn_norm); cudaFree(L.ffn_norm);
    }
    cudaFree(m->embd_rows); cudaFree(m->output_norm); cudaFree(m->output.alloc);
    delete m;
}

extern "C" int b200_model_info(const b200_model * m, int32_t info[B200_INFO_COUNT]) {
    if (!m) return set_err("null model");
    std::memset(info, 0, sizeof(int32_t) * B200_INFO_COUNT);
    info[B200_INFO_N_VOCAB] = m->n_vocab; info[B200_INFO_N_EMBD] = m->n_embd; info[B200_INFO_N_LAYER] = m->n_layer;
    info[B200_INFO_N_HEAD] = m->n_head; info[B200_INFO_N_HEAD_KV] = m->n_head_kv; info[B200_INFO_N_FF] = m->n_ff;
    info[B200_INFO_HEAD_DIM] = m->head_dim; info[B200_INFO_N_CTX_TRAIN] = m->n_ctx_train;
    info[B200_INFO_LAYER_BEGIN] = m->layer_begin; info[B200_INFO_LAYER_END] = m->layer_end; info[B200_INFO_FTYPE] = m->ftype;
    return 0;
}
extern "C" int64_t b200_model_weight_bytes(const b200_model * m) { return m ? m->weight_bytes : 0; }

// ------------------------------------------------------------------------------------------------------------
// RoPE table on the host — restates ggml_rope_cache_init + rope_yarn + corr dims
// (cpp/ggml/src/ggml.c:13987-14041): theta starts at pos and is multiplied by theta_scale each pair.
// Built with the host libm so it is bit-identical to the CPU reference on the same machine.
// ------------------------------------------------------------------------------------------------------------
static float yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float) M_PI)) / (2 * logf(base));
}
// (cos, sin) of the hd/2 pairs for position p — ggml_rope_cache_init; p may be negative (the K-shift rotates by a delta)
static void build_rope_row(const b200_model & m, int p, float2 * row);
static void build_rope_table(const b200_model & m, int n_ctx, std::vector<float2> & tab) {
    const int hd = m.head_dim;
    tab.resize((size_t) n_ctx * (hd / 2));
    for (int p = 0; p < n_ctx; p++) build_rope_row(m, p, tab.data() + (size_t) p * (hd / 2));
}
static void build_rope_row(const b200_model & m, int p, float2 * row) {
    const int hd = m.head_dim;
    const float theta_scale = powf(m.rope_freq_base, -2.0f / hd);
    float corr[2];
    {
        const float start = floorf(yarn_corr_dim(hd, m.n_ctx_orig, m.yarn_beta_fast, m.rope_freq_base));
        const float end   = ceilf(yarn_corr_dim(hd, m.n_ctx_orig, m.yarn_beta_slow, m.rope_freq_base));
        corr[0] = std::max(0.f, start);
        corr[1] = std::min((float) (hd - 1), end);
    }
    {
        float theta = (float) p;
        for (int i0 = 0; i0 < hd; i0 += 2) {
            const float ff = m.rope_freq_factors.empty() ? 1.0f : m.rope_freq_factors[(size_t) (i0 / 2)];
            const float theta_extrap = theta / ff;
            float theta_interp = m.rope_freq_scale * theta_extrap;
            float th = theta_interp;
            float mscale = m.yarn_attn_factor;
            if (m.yarn_ext_factor != 0.0f) {
                const float y = (i0 / 2 - corr[0]) / std::max(0.001f, corr[1] - corr[0]);
                const float ramp_mix = (1 - std::min(1.f, std::max(0.f, y))) * m.yarn_ext_factor;
                th = theta_interp * (1 - ramp_mix) + theta_extrap * ramp_mix;
                mscale *= 1.0f + 0.1f * logf(1.0f / m.rope_freq_scale);
            }
            row[i0 / 2] = make_float2(cosf(th) * mscale, sinf(th) * mscale);
            theta *= theta_scale;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// NCCL, loaded lazily (single-GPU use never touches it)
// ------------------------------------------------------------------------------------------------------------
struct NcclId { char b[128]; };   // ncclUniqueId is passed BY VALUE (nccl.h: struct { char internal[128]; })
struct NcclApi {
    void * h = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char * (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static void nccl_load() {
    if (g_nccl.h) return;
    const char * names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char * n : names) { g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
    if (!g_nccl.h) throw std::runtime_error("cannot dlopen libnccl.so.2 (needed for multi-GPU pipeline)");
    auto sym = [&](const char * s) { void * p = dlsym(g_nccl.h, s); if (!p) throw std::runtime_error(std::string("missing NCCL symbol ") + s); return p; };
    *(void **) &g_nccl.GetUniqueId    = sym("ncclGetUniqueId");
    *(void **) &g_nccl.CommInitRank   = sym("ncclCommInitRank");
    *(void **) &g_nccl.Send           = sym("ncclSend");
    *(void **) &g_nccl.Recv           = sym("ncclRecv");
    *(void **) &g_nccl.GroupStart     = sym("ncclGroupStart");
    *(void **) &g_nccl.GroupEnd       = sym("ncclGroupEnd");
    *(void **) &g_nccl.CommDestroy    = sym("ncclCommDestroy");
    *(void **) &g_nccl.GetErrorString = sym("ncclGetErrorString");
}
#define NC(call) do { int r_ = (call); if (r_ != 0) throw std::runtime_error(std::string(#call) + ": " + g_nccl.GetErrorString(r_)); } while (0)
static constexpr int NCCL_FLOAT32 = 7, NCCL_INT32 = 2;   // ncclDataType_t (nccl.h)

// ------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------
struct TapStore { std::map<std::string, std::vector<float>> v; };

struct b200_ctx {
    b200_model * m = nullptr;
    int device = 0;
    int n_ctx = 0;
    cudaStream_t st = nullptr;
    // activations
    float * x = nullptr;        // residual stream [n_embd]
    float * q = nullptr;        // [n_head*hd]
    float * att = nullptr;      // kqv_merged_cont [n_head*hd]
    float * ffh = nullptr;      // silu(gate)*up [n_ff]
    float * warm_x = nullptr;   // constant vector [max(n_ff, n_embd)] read by the prologue's dry run (B200_WARM)
    float * logits = nullptr;   // [n_vocab]
    float * S = nullptr;        // attention scores / probabilities [n_head][n_ctx]
    unsigned int * tickets = nullptr;
    unsigned long long * amax_key = nullptr;
    std::vector<__half *> kc, vc;   // per local layer
    float2 * rope = nullptr;
    DecodeState * d_state = nullptr;
    DecodeState * h_state = nullptr;   // pinned
    float * h_logits = nullptr;        // pinned
    int32_t * d_out_tokens = nullptr;
    int out_tokens_cap = 0;
    cudaGraphExec_t g_logits = nullptr;   // H2D state -> forward -> D2H logits
    cudaGraphExec_t g_greedy = nullptr;   // forward -> argmax -> advance
    cudaGraphExec_t g_pipe = nullptr;     // pipeline stage step (recv -> forward -> send)
    cudaGraphExec_t g_step = nullptr;     // H2D state -> forward -> arg-max -> D2H token (b200_step_greedy)
    int64_t n_step = 0;
    int32_t * h_tok = nullptr;            // pinned: the token b200_step_greedy hands back
    int64_t n_logits = 0, n_greedy = 0, n_pipe = 0;   // kernels per replay of g_logits / g_greedy / g_pipe
    int pipe_calls = 0;                   // b200_pipeline_generate_greedy calls so far (the first one runs un-graphed: NCCL connects lazily)
    bool pipe_graph_failed = false;
    int sm_count = 148;
    bool taps = false;
    TapStore tapstore;
    // KV cells (struct llama_kv_cache, cpp/src/llama.cpp:2495-2539). Until the first llama_kv_cache_seq_rm / seq_add a sequence
    // only grows and cell == position (n_hi positions written so far); afterwards the cells are managed like the reference's:
    // find_slot from `head`, positions and pending K-shift deltas per cell, attention masked by cell_pos on the device.
    struct KvCells {
        bool managed = false;
        int n_hi = 0;                      // identity mode: highest position written + 1
        std::vector<int32_t> pos, delta;   // managed mode, per cell
        int head = 0, used = 0;
        bool has_shift = false;
    } cells;
    int32_t * d_cell_pos = nullptr;        // [n_ctx] device copy of cells.pos
    // prompt batches (prefill.cuh): activations of up to pb_cap tokens, allocated at the first batch
    struct PrefillBufs {
        int cap = 0;
        int att_z = 64;          // tokens per attention launch group (bounded by the score buffer)
        float * X = nullptr, * Q = nullptr, * ATT = nullptr, * FFH = nullptr, * S = nullptr;
        uint8_t * rec = nullptr;
        int32_t * tokens = nullptr;
    } pb;
    // persistent per-token kernel (token_kernel.cuh): phase list in device memory, grid-barrier counter, transposed scores
    Phase * d_plan = nullptr;
    int n_phases = 0;
    size_t token_smem = 0;
    unsigned long long * d_bar = nullptr;
    float * S_T = nullptr;
    int token_state = 0;               // 0: not built yet, 1: usable, -1: not usable for this model / context
    unsigned long long * d_ttrace = nullptr;   // b200_trace_phases
    bool ttracing = false;
    // pipeline
    // direct NVLink hand-off between the stage processes (b200_p2p_*): this rank's inbox (device memory, exported through CUDA
    // IPC), the next rank's inbox and rank 0's inbox mapped into this process, and the per-direction message counters
    struct P2PInbox * inbox = nullptr;          // mine
    struct P2PInbox * next_inbox = nullptr;     // rank + 1's (x hand-off)
    struct P2PInbox * first_inbox = nullptr;    // rank 0's (token hand-back; only on the last rank)
    unsigned long long * p2p_counts = nullptr;  // device: [0] x sent, [1] x received, [2] tokens sent, [3] tokens received
    bool p2p = false;
    void * comm = nullptr;
    int rank = 0, world = 1;
    cudaEvent_t ev_done = nullptr;     // in-process stage chain hand-off: this stage's l_out is complete
    cudaEvent_t ev_taken = nullptr;    // ... and the NEXT stage has copied it out of c->x (recorded on the consumer's stream)
    bool taken_pending = false;        // ev_taken was recorded since this stage last waited on it
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // device time of the last generate/decode call
    float last_device_ms = 0.f;
    // phase trace (b200_trace_token): [launch][TRACE_CTAS][TRACE_PHASES] globaltimer stamps
    unsigned long long * d_trace = nullptr;
    bool tracing = false;
    int trace_seq = 0;
    std::vector<std::array<int, 3>> trace_meta;   // per launch: (kind, ctas, 0)
    // per-launch profiling (bench.py's live roofline): event pairs around every launch of one token
    bool prof = false;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_ev;   // (kind, (start, stop))
    // counters
    int64_t launches = 0;
    double t_prompt_us = 0, t_gen_us = 0;
    int64_t n_prompt = 0, n_gen = 0;
};

extern "C" int b200_n_ctx(const b200_ctx * c) { return c ? c->n_ctx : 0; }
extern "C" int64_t b200_kernel_launches(const b200_ctx * c) { return c ? c->launches : 0; }

enum { KIND_EMBED = 0, KIND_QKV, KIND_ATTN, KIND_WO, KIND_GATEUP, KIND_DOWN, KIND_HEAD, KIND_ATTN_PV, KIND_COUNT };
// (thread_local: up to 8 pods run doInference concurrently, each on its own OS thread — SURVEY §8b threading)
static thread_local int g_kind = KIND_EMBED;   // set by enqueue_forward before each launch group
static thread_local int g_only_kind = -1;       // b200_profile_kind: enqueue_forward launches only this kind (-1: everything)
// function attributes (dynamic shared-memory opt-in) are raised lazily, per device, from whichever pod's thread gets
// there first
static std::mutex g_attr_mu;
static bool want(int kind) { return g_only_kind < 0 || g_only_kind == kind; }
// raise a kernel's dynamic shared-memory limit at most once per size and device; `have` (the per-device high-water mark) is
// read and written under g_attr_mu only — pods on different OS threads reach the same kernels concurrently
template <typename K>
static void raise_smem(K kern, size_t smem, size_t (&have)[64], int device) {
    std::lock_guard<std::mutex> lk(g_attr_mu);
    if (smem <= have[device & 63]) return;
    cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e_ != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e_));
    have[device & 63] = smem;
}
struct ProfScope {
    b200_ctx * c; cudaEvent_t a = nullptr, b = nullptr;
    explicit ProfScope(b200_ctx * c_);
    ~ProfScope();
};

// Every kernel of the forward pass goes through here: cudaLaunchKernelEx with the programmatic-stream-serialization
// attribute (PDL), so that inside the captured graph consecutive kernels are joined by programmatic edges — the next
// kernel's CTAs start (barrier init, weight prefetch) while the previous one drains, and block in griddepcontrol.wait
// until its results are visible. BOOSTER_B200_NO_PDL=1 falls back to plain stream order (A/B measurements).
static bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char * e = getenv("BOOSTER_B200_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
template <typename Arg>
static void launch_fwd(void (*kern)(Arg), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Arg & arg, bool pdl = true) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kern, arg));
}

static constexpr int TRACE_CTAS = 512, TRACE_MAX_LAUNCHES = 1024;
static unsigned long long * trace_slot(b200_ctx * c, int ctas) {
    if (!c->tracing) return nullptr;
    if (ctas > TRACE_CTAS || c->trace_seq >= TRACE_MAX_LAUNCHES) throw std::runtime_error("trace buffer too small");
    c->trace_meta.push_back({g_kind, ctas, 0});
    return c->d_trace + (size_t) (c->trace_seq++) * TRACE_CTAS * TRACE_PHASES;
}

// launch shape of one mat-vec: W warps per CTA (one CTA per SM), S ring stages per warp, G warps per 32-row unit. Pick the
// combination with the most concurrently active warps (every unit resident in as few waves as possible), then the
// deepest ring that still fits the shared memory. Pure host arithmetic (b200_op_launch_shape exposes it to the tests).
static constexpr size_t MV_SMEM_BUDGET = 227 * 1024 - 512;    // dynamic shared memory of k_matvec
static void pick_launch_shape(int n_units, int tiles_unit, int k, bool norm, int act_q8_0, int stage_bytes, int nv, int sm_count,
                              int & bestW, int & bestG, int & bestS, size_t budget = MV_SMEM_BUDGET) {
    const size_t act_bytes = act_smem_bytes(k, act_q8_0);
    bestW = 0; bestG = 1; bestS = 0;
    double best = -1;
    for (int W = MV_MAX_WARPS; W >= 4; W -= 2) {
        for (int G = W; G >= 1; G--) {
            if (W % G) continue;
            if (G > 1 && (long long) n_units * G > (long long) sm_count * W) continue;   // sharing only within one wave
            if (tiles_unit % G) continue;                                               // warp w owns tiles w, w+G, ... of every unit
            if (norm && k / 256 > PRO_U * W) continue;                                  // the normed vector is quantized in one pass
            const size_t fixed = act_bytes + chain_smem_bytes(W, G, nv) + (size_t) W * 4 * 8 + (size_t) W * 8;
            if (fixed + (size_t) W * 2 * stage_bytes > budget) continue;
            const int S = (int) std::min<size_t>(4, (budget - fixed) / ((size_t) W * stage_bytes));
            const long long warps = (long long) n_units * G, slots = (long long) sm_count * W;
            const long long waves = (warps + slots - 1) / slots;
            const double active = (double) warps / (double) waves + 0.01 * S;   // average concurrently running warps
            if (active > best) { best = active; bestW = W; bestG = G; bestS = S; }
            break;   // largest feasible G for this W
        }
    }
}
// (W, G, S) the engine would launch a mat-vec with: `types` = block types of the launch's segments, n_units = rows / 32
extern "C" int b200_op_launch_shape(const int32_t * types, int n_types, int64_t n_units, int64_t k, int norm, int sm_count, int32_t wgs[3]) {
    if (!types || n_types <= 0 || !wgs || k <= 0 || k % 256) return set_err("bad arguments");
    int sb = 0, nv = 0;
    for (int i = 0; i < n_types; i++) { sb = std::max(sb, tile_bytes_of(types[i])); nv = std::max(nv, chain_values_of(types[i])); }
    const int q80 = types[0] == T_Q8_0;
    int W, G, S;
    pick_launch_shape((int) n_units, (int) (k / (q80 ? 32 : 256)), (int) k, norm != 0, q80, (sb + 127) / 128 * 128, nv, sm_count, W, G, S);
    wgs[0] = W; wgs[1] = G; wgs[2] = S;
    return W == 0 ? set_err("activation vector too long for the shared-memory budget") : 0;
}

// derived fields of a mat-vec (launch shape, shared-memory layout); returns the dynamic shared memory it needs
static size_t shape_matvec(MatvecArgs & a, int epi, int sm_count, size_t budget) {
    a.epi = epi;
    a.tiles_unit = a.seg[0].tiles_unit;
    for (int i = 0; i < a.n_seg; i++)
        if (a.seg[i].tiles_unit != a.tiles_unit) throw std::runtime_error("segments of one launch must share the unit shape");
    int sb = 0;
    for (int i = 0; i < a.n_seg; i++) sb = std::max(sb, tile_bytes_of(a.seg[i].type));
    a.stage_bytes = (sb + 127) / 128 * 128;
    int nv = 0;
    for (int i = 0; i < a.n_seg; i++) nv = std::max(nv, chain_values_of(a.seg[i].type));
    a.nv = nv;
    const size_t act_bytes = act_smem_bytes(a.k, a.act_q8_0);
    int bestW = 0, bestG = 1, bestS = 0;
    pick_launch_shape(a.n_units, a.tiles_unit, a.k, a.norm_w != nullptr, a.act_q8_0, a.stage_bytes, nv, sm_count, bestW, bestG, bestS, budget);
    if (bestW == 0) throw std::runtime_error("activation vector too long for the shared-memory budget");
    a.group = bestG; a.stages = bestS; a.warps = bestW;
    a.inv_k = (a.k & (a.k - 1)) == 0 ? 1.0 / (double) a.k : 0.0;
    a.chain_mode = chain_mode_of(bestG); a.act_bytes = (uint32_t) act_bytes; a.chain_bytes = (uint32_t) chain_smem_bytes(bestW, bestG, nv);
    a.exch_words = a.chain_mode == CHAIN_EXCHANGE ? (uint32_t) bestW * 2 * (uint32_t) nv * 32 : 0;
    a.kpw = a.tiles_unit / bestG; a.groups_per_cta = bestW / bestG; a.grp_magic = (uint32_t) (65536 / bestG + 1);
    const size_t smem = (size_t) bestW * a.stages * a.stage_bytes + act_bytes + chain_smem_bytes(bestW, a.group, nv) + (size_t) bestW * a.stages * 8 + (size_t) bestW * 8;
    if (smem > budget + 256) throw std::runtime_error("activation vector too long for the shared-memory budget");
    return smem;
}

static void launch_matvec(b200_ctx * c, const MatvecArgs & a_in, int epi) {
    ProfScope ps(c);
    MatvecArgs a = a_in;
    const size_t smem = shape_matvec(a, epi, c->sm_count, MV_SMEM_BUDGET);
    static int prefill_env = -1;
    if (prefill_env < 0) { const char * e = getenv("BOOSTER_B200_PREFILL"); prefill_env = e ? atoi(e) : 0; }
    a.prefill = prefill_env;
    const int W = a.warps;
    static size_t attr_smem[64] = {0};   // per device (function attributes are per device)
    {
        std::lock_guard<std::mutex> attr_lock(g_attr_mu);
        if (smem > attr_smem[c->device & 63]) {
            CU(cudaFuncSetAttribute(k_matvec<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            CU(cudaFuncSetAttribute(k_matvec<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            attr_smem[c->device & 63] = smem;
        }
    }
    const int grid = std::max(1, std::min(a.n_units, c->sm_count));
    a.trace = trace_slot(c, grid);
    launch_fwd(a.trace ? k_matvec<true> : k_matvec<false>, dim3((unsigned) grid), dim3((unsigned) (W * 32)), smem, c->st, a);
    c->launches++;
}

// two launches per layer: raw scores (k_attn_scores, every SM busy) then softmax + P.V (k_attn_softmax_pv); returns
// false when the GQA score rows of the context do not fit the shared memory (the three-kernel route below takes over)
template <int GQA>
// nz > 1: a prompt batch, token z of it in blockIdx.z (plain stream order then: the batch's own K / V rows come from the
// kernel just before, and the scores kernel reads older rows ahead of its dependency wait)
static bool launch_attention_2k(b200_ctx * c, const AttnArgs & a_in, int n_ctx_pad, int nz = 1) {
    AttnArgs a = a_in;
    const size_t row_bytes = (size_t) GQA * n_ctx_pad * 4;
    const size_t budget = 200 * 1024;
    if (row_bytes + (size_t) PV_BATCH * 16 > budget) return false;
    int vch = (int) ((budget - row_bytes) / 16) / PV_BATCH * PV_BATCH;
    vch = std::min(vch, (n_ctx_pad + PV_BATCH - 1) / PV_BATCH * PV_BATCH);
    a.p_chunk = vch;
    size_t smem = row_bytes + (size_t) vch * 16;
    // at most one CTA per SM: the block scheduler then spreads the (n_head_kv x 16) CTAs over distinct SMs even when
    // they are launched early (PDL) next to the tail of the previous kernel
    if ((int) a.n_head_kv * (128 / PVS_DIMS) <= c->sm_count) smem = std::max(smem, (size_t) 116 * 1024);
    static size_t attr[64] = {0};
    const int dv = c->device & 63;
    std::unique_lock<std::mutex> attr_lock(g_attr_mu);
    if (smem > attr[dv]) {
        CU(cudaFuncSetAttribute(k_attn_softmax_pv<GQA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        CU(cudaFuncSetAttribute(k_attn_softmax_pv<GQA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr[dv] = smem;
    }
    attr_lock.unlock();
    {
        g_kind = KIND_ATTN; ProfScope ps(c);
        const dim3 gs((unsigned) a.n_head_kv, (unsigned) ((n_ctx_pad + ATT_TILE - 1) / ATT_TILE), (unsigned) nz);
        a.trace = nz == 1 ? trace_slot(c, (int) (gs.x * gs.y)) : nullptr;
        launch_fwd(a.trace ? k_attn_scores<GQA, true> : k_attn_scores<GQA, false>, gs, dim3(ATT_THREADS), 0, c->st, a, nz == 1);
    }
    {
        g_kind = KIND_ATTN_PV; ProfScope ps(c);
        const dim3 gp((unsigned) a.n_head_kv, (unsigned) (128 / PVS_DIMS), (unsigned) nz);
        a.trace = nz == 1 ? trace_slot(c, (int) (gp.x * gp.y)) : nullptr;
        launch_fwd(a.trace ? k_attn_softmax_pv<GQA, true> : k_attn_softmax_pv<GQA, false>, gp, dim3((unsigned) (GQA * pvs_th(GQA))), smem, c->st, a, nz == 1);
    }
    c->launches += 2;
    return true;
}

// attention route: 0 = automatic (two launches when the GQA score rows fit one CTA's shared memory, else the
// long-context three-kernel route), 1 = always the long-context route (tests keep both covered at every size)
static int g_attn_route = 0;
extern "C" void b200_set_attention_route(int route) { g_attn_route = route; }
static bool attn_2k_enabled() {
    static int v = -1;
    if (v < 0) { const char * e = getenv("BOOSTER_B200_ATTN_SPLIT"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1 && g_attn_route != 1;
}

template <int GQA>
static void launch_attention_t(b200_ctx * c, const AttnArgs & a_in, int n_ctx_pad) {
    if (attn_2k_enabled() && launch_attention_2k<GQA>(c, a_in, n_ctx_pad)) return;
    // long-context route (the GQA score rows do not fit one CTA's shared memory): scores, softmax, chunked P.V
    AttnArgs a = a_in;
    int pch = (200 * 1024) / (GQA * 4 + 32) / PV_BATCH * PV_BATCH;
    pch = std::min(pch, (n_ctx_pad + PV_BATCH - 1) / PV_BATCH * PV_BATCH);
    a.p_chunk = pch;
    const size_t pv_smem = (size_t) pch * (GQA * 4 + 32);
    static size_t attr_pv[64] = {0};   // per device (function attributes are per device)
    const int dv = c->device & 63;
    {
        std::lock_guard<std::mutex> attr_lock(g_attr_mu);
        if (pv_smem > attr_pv[dv]) { CU(cudaFuncSetAttribute(k_attn_pv<GQA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pv_smem)); attr_pv[dv] = pv_smem; }
    }
    const dim3 gs((unsigned) a.n_head_kv, (unsigned) ((n_ctx_pad + ATT_TILE - 1) / ATT_TILE));
    {
        g_kind = KIND_ATTN; ProfScope ps(c);
        launch_fwd(k_attn_scores<GQA, false>, gs, dim3(ATT_THREADS), 0, c->st, a);
        k_attn_softmax<<<a.n_head, 256, 0, c->st>>>(a);
    }
    {
        g_kind = KIND_ATTN_PV; ProfScope ps(c);
        const dim3 gp((unsigned) a.n_head_kv, (unsigned) (128 / PV_DIMS));
        k_attn_pv<GQA><<<gp, PV_DIMS * 16, pv_smem, c->st>>>(a);
    }
    c->launches += 3;
}
static void launch_attention(b200_ctx * c, const AttnArgs & a, int n_ctx_pad) {
    if (a.head_dim != 128) throw std::runtime_error("attention kernels are specialised for head_dim 128");
    switch (a.n_head / a.n_head_kv) {
        case 1: launch_attention_t<1>(c, a, n_ctx_pad); break;
        case 2: launch_attention_t<2>(c, a, n_ctx_pad); break;
        case 4: launch_attention_t<4>(c, a, n_ctx_pad); break;
        case 8: launch_attention_t<8>(c, a, n_ctx_pad); break;
        default: throw std::runtime_error("GQA ratio must be 1, 2, 4 or 8");
    }
}

ProfScope::ProfScope(b200_ctx * c_) : c(c_) {
    if (!c->prof) return;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, c->st));
}
ProfScope::~ProfScope() {
    if (!c->prof) return;
    cudaEventRecord(b, c->st);
    c->prof_ev.push_back({g_kind, {a, b}});
}

static void tap(b200_ctx * c, const std::string & name, int il, const float * dptr, size_t n) {
    if (!c->taps) return;
    CU(cudaStreamSynchronize(c->st));
    std::vector<float> h(n);
    CU(cudaMemcpy(h.data(), dptr, n * 4, cudaMemcpyDeviceToHost));
    c->tapstore.v[name + "-" + std::to_string(il)] = std::move(h);
}

// L2 look-ahead plan (kernels.cuh PfRange): which later kernel's bytes each kernel of a layer prefetches into L2.
//   QKV      -> this layer's K and V rows 0..pos, wo                      (consumed by scores, softmax+P.V, wo)
//   scores   -> gate|up, first part        softmax+P.V -> gate|up, middle        wo -> gate|up, rest
//   gate|up  -> down                       down        -> the next layer's QKV stack (or the head's first tiles)
// The short kernels between two big mat-vecs are latency-bound and leave HBM idle; with the plan HBM streams the whole
// time and the big mat-vecs find (most of) their tiles in L2. BOOSTER_B200_PF="kvwo,gu_scores,gu_pv,gu_wo,down,next"
// (fractions of the consumer's bytes, 0 = off) tunes it; BOOSTER_B200_PF=0 disables the look-ahead.
struct PfPlan { float kvwo = 1.f, gu_scores = 0.f, gu_pv = 0.f, gu_wo = 0.f, down = 0.f, next = 1.f; };
static const PfPlan & pf_plan() {
    static PfPlan p; static bool init = false;
    if (!init) {
        init = true;
        if (const char * e = getenv("BOOSTER_B200_PF")) {
            float v[6] = {0, 0, 0, 0, 0, 0};
            const int n = sscanf(e, "%f,%f,%f,%f,%f,%f", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
            if (n == 1 && v[0] == 0.f) { p = PfPlan{0, 0, 0, 0, 0, 0}; }
            else if (n == 6) { p = PfPlan{v[0], v[1], v[2], v[3], v[4], v[5]}; }
        }
    }
    return p;
}
static size_t tiled_bytes(const TMat & m) { return (size_t) m.n_units * m.tiles_unit * m.tile_bytes; }
static size_t tiled_bytes(const DevStack & s) { size_t n = 0; for (int i = 0; i < s.n_seg; i++) n += tiled_bytes(s.seg[i]); return n; }
// bytes [from, to) (fractions) of a tiled matrix as a look-ahead range, cut at tile-size-independent 16-byte multiples
static PfRange pf_slice(const uint8_t * base, size_t total, double from, double to) {
    PfRange r{nullptr, 0, 0};
    from = std::min(1.0, std::max(0.0, from)); to = std::min(1.0, std::max(from, to));
    const size_t b0 = (size_t) (total * from) / 8192 * 8192, b1 = to >= 1.0 ? total : (size_t) (total * to) / 8192 * 8192;
    if (b1 <= b0) return r;
    r.p = base + b0; r.bytes = (uint32_t) std::min<size_t>(b1 - b0, 0xfffffff0u);
    return r;
}
static void pf_push(PfRange (&pf)[PF_RANGES], const PfRange & r) {
    if (!r.bytes) return;
    for (int i = 0; i < PF_RANGES; i++) if (!pf[i].bytes) { pf[i] = r; return; }
}

// arguments of the mat-vecs of one layer / of the head (shared by the per-kernel path and the persistent kernel's plan)
static MatvecArgs args_qkv(b200_ctx * c, int li) {
    b200_model & m = *c->m; LayerW & L = m.layers[(size_t) li];
    const int E = m.n_embd, HD = m.head_dim, KVD = m.n_head_kv * HD, QD = m.n_head * HD;
    MatvecArgs a{};
    for (int i = 0; i < L.qkv.n_seg; i++) a.seg[i] = L.qkv.seg[i];
    a.n_seg = L.qkv.n_seg; a.n_units = L.qkv.n_units; a.k = E;
    a.x = c->x; a.norm_w = L.attn_norm; a.eps = m.rms_eps; a.act_q8_0 = L.qkv.seg[0].type == T_Q8_0;
    a.q_out = c->q; a.k_cache = c->kc[(size_t) li]; a.v_cache = c->vc[(size_t) li];
    a.n_q = QD; a.n_k = KVD; a.head_dim = HD; a.kv_dim = KVD; a.rope = c->rope; a.st = c->d_state;
    return a;
}
static MatvecArgs args_wo(b200_ctx * c, int li) {
    b200_model & m = *c->m; LayerW & L = m.layers[(size_t) li];
    MatvecArgs a{};
    a.seg[0] = L.wo.m; a.n_seg = 1; a.n_units = L.wo.m.n_units; a.k = m.n_head * m.head_dim;
    a.x = c->att; a.norm_w = nullptr; a.act_q8_0 = L.qkv.seg[0].type == T_Q8_0;
    a.out = c->x; a.resid = c->x; a.st = c->d_state;
    return a;
}
static MatvecArgs args_gateup(b200_ctx * c, int li) {
    b200_model & m = *c->m; LayerW & L = m.layers[(size_t) li];
    MatvecArgs a{};
    a.seg[0] = L.gateup.m; a.n_seg = 1; a.n_units = L.gateup.m.n_units; a.k = m.n_embd;
    a.x = c->x; a.norm_w = L.ffn_norm; a.eps = m.rms_eps; a.act_q8_0 = L.qkv.seg[0].type == T_Q8_0;
    a.out = c->ffh; a.st = c->d_state;
    return a;
}
static MatvecArgs args_down(b200_ctx * c, int li) {
    b200_model & m = *c->m; LayerW & L = m.layers[(size_t) li];
    MatvecArgs a{};
    a.seg[0] = L.down.m; a.n_seg = 1; a.n_units = L.down.m.n_units; a.k = m.n_ff;
    a.x = c->ffh; a.norm_w = nullptr; a.act_q8_0 = L.qkv.seg[0].type == T_Q8_0;
    a.out = c->x; a.resid = c->x; a.st = c->d_state;
    return a;
}
static MatvecArgs args_head(b200_ctx * c) {
    b200_model & m = *c->m;
    MatvecArgs a{};
    a.seg[0] = m.output.m; a.n_seg = 1; a.n_units = m.output.m.n_units; a.k = m.n_embd;
    a.x = c->x; a.norm_w = m.output_norm; a.eps = m.rms_eps; a.act_q8_0 = m.output.m.type == T_Q8_0;
    a.out = c->logits;
    return a;
}

// ------------------------------------------------------------------------------------------------------------
// persistent per-token kernel (token_kernel.cuh): phase list of this stage, built once per context
// ------------------------------------------------------------------------------------------------------------
#if B200_TOKEN_KERNEL
// shared memory the kernel's own __shared__ variables take (phase descriptors, attention scratch): the dynamic budget is
// what is left of the 227 KB a CTA may have
template <int GQA>
static size_t token_kernel_static_smem() {
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, k_token<GQA, true>));
    return fa.sharedSizeBytes;
}
static size_t token_kernel_budget(int gqa) {
    size_t st = 0;
    switch (gqa) {
        case 1: st = token_kernel_static_smem<1>(); break;
        case 2: st = token_kernel_static_smem<2>(); break;
        case 4: st = token_kernel_static_smem<4>(); break;
        default: st = token_kernel_static_smem<8>(); break;
    }
    return 227 * 1024 - (st + 127) / 128 * 128 - 256;
}
template <int GQA>
static void token_kernel_prepare(b200_ctx * c, size_t smem) {
    static size_t attr[64] = {0};      // per device; the attribute only ever grows (other contexts keep launching with theirs)
    const int dv = c->device & 63;
    if (smem > attr[dv]) {
        for (auto kern : { k_token<GQA, false>, k_token<GQA, true> })
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr[dv] = smem;
    }
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_token<GQA, false>, TK_THREADS, smem));
    if (nb < 1) throw std::runtime_error("the per-token kernel does not fit one SM");
}
// Which path decodes a token: 0 = one kernel per operator joined by programmatic dependent launch (default: measured
// faster, 546 vs 371 tok/s on the 8B Q4_K_M config — the grid barrier costs 3 us against ~2 us for a PDL kernel boundary
// and does not overlap the way the next kernel's early CTAs do), 1 = the persistent per-token kernel.
// BOOSTER_B200_TOKEN_KERNEL=1 or b200_set_token_kernel(1) selects it for contexts created afterwards.
static int g_token_kernel = -1;
extern "C" void b200_set_token_kernel(int on) { g_token_kernel = on ? 1 : 0; }
static bool token_kernel_enabled() {
    if (g_token_kernel < 0) { const char * e = getenv("BOOSTER_B200_TOKEN_KERNEL"); g_token_kernel = (e && e[0] == '1') ? 1 : 0; }
    return g_token_kernel == 1;
}
// returns false (and remembers it) when this model / context cannot use the persistent kernel; the per-kernel path runs then
static bool token_kernel_build(b200_ctx * c) {
    if (c->token_state != 0) return c->token_state > 0;
    c->token_state = -1;
    b200_model & m = *c->m;
    const int HD = m.head_dim, KVD = m.n_head_kv * HD, gqa = m.n_head / m.n_head_kv;
    if (HD != 128 || (gqa != 1 && gqa != 2 && gqa != 4 && gqa != 8) || m.layers.empty()) return false;
    try {
        const size_t TK_SMEM_BUDGET = token_kernel_budget(gqa);
        const int rs = (c->n_ctx / 16 + 3) / 4 * 4;
        const size_t ps_bytes = (size_t) gqa * 16 * rs * 4;
        if (ps_bytes + 64 * 16 > TK_SMEM_BUDGET) return false;            // the GQA score rows of the context must fit one CTA
        int v_chunk = (int) std::min<size_t>((size_t) (c->n_ctx + 63) / 64 * 64, (TK_SMEM_BUDGET - ps_bytes) / 16 / 64 * 64);
        size_t smem = ps_bytes + (size_t) v_chunk * 16;
        std::vector<Phase> plan;
        auto push_mv = [&](MatvecArgs a, int epi) {
            Phase P{};
            P.kind = PH_MATVEC;
            smem = std::max(smem, shape_matvec(a, epi, c->sm_count, TK_SMEM_BUDGET));
            // tiles per warp requested before the grid barrier (the whole ring but one slot); BOOSTER_B200_TK_PREFILL overrides (A/B)
            static int pf_env = -2;
            if (pf_env == -2) { const char * e = getenv("BOOSTER_B200_TK_PREFILL"); pf_env = e ? atoi(e) : -1; }
            a.prefill = pf_env >= 0 ? std::min(pf_env, a.stages - 1) : a.stages - 1;
            P.mv = a;
            plan.push_back(P);
        };
        for (int li = 0; li < (int) m.layers.size(); li++) {
            push_mv(args_qkv(c, li), EPI_QKV);
            Phase P{};
            P.at.q = c->q; P.at.k_cache = c->kc[(size_t) li]; P.at.v_cache = c->vc[(size_t) li];
            P.at.rs = rs; P.at.out = c->att; P.at.n_head_kv = m.n_head_kv; P.at.kv_dim = KVD;
            P.at.scale = 1.0f / sqrtf((float) HD); P.at.st = c->d_state; P.at.v_chunk = v_chunk; P.at.cell_pos = c->d_cell_pos;
            P.kind = PH_SCORES; plan.push_back(P);
            P.kind = PH_SOFTMAX_PV; plan.push_back(P);
            push_mv(args_wo(c, li), EPI_RESID);
            push_mv(args_gateup(c, li), EPI_SILU);
            push_mv(args_down(c, li), EPI_RESID);
        }
        if (m.has_head()) push_mv(args_head(c), EPI_STORE);
        CU(cudaMalloc(&c->S_T, (size_t) m.n_head * 16 * rs * 4));
        for (auto & P : plan) P.at.S = c->S_T;
        CU(cudaMalloc(&c->d_plan, plan.size() * sizeof(Phase)));
        CU(cudaMemcpy(c->d_plan, plan.data(), plan.size() * sizeof(Phase), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_bar, 8));
        CU(cudaMemset(c->d_bar, 0, 8));
        c->n_phases = (int) plan.size();
        c->token_smem = smem;
        {
            std::lock_guard<std::mutex> attr_lock(g_attr_mu);
            switch (gqa) {
                case 1: token_kernel_prepare<1>(c, smem); break;
                case 2: token_kernel_prepare<2>(c, smem); break;
                case 4: token_kernel_prepare<4>(c, smem); break;
                default: token_kernel_prepare<8>(c, smem); break;
            }
        }
        c->token_state = 1;
        return true;
    } catch (const std::exception & e) {
        fprintf(stderr, "booster_b200: per-token kernel not usable (%s); using one kernel per operator\n", e.what());
        cudaGetLastError();
        return false;
    }
}
template <int GQA>
static void token_kernel_launch(b200_ctx * c) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned) c->sm_count); cfg.blockDim = dim3(TK_THREADS); cfg.dynamicSmemBytes = c->token_smem; cfg.stream = c->st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;                 // all CTAs co-resident, or the launch fails: the grid barrier needs it
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    unsigned long long * tr = c->ttracing ? c->d_ttrace : nullptr;
    CU(cudaLaunchKernelEx(&cfg, tr ? k_token<GQA, true> : k_token<GQA, false>, (const Phase *) c->d_plan, c->n_phases, c->d_bar, tr));
    c->launches++;
}

#else
extern "C" void b200_set_token_kernel(int) {}
static bool token_kernel_enabled() { return false; }
static bool token_kernel_build(b200_ctx *) { return false; }
template <int GQA> static void token_kernel_launch(b200_ctx *) { throw std::runtime_error("built without the per-token kernel"); }
#endif
// enqueue the layers of this stage (+ embedding on the first stage, + head on the last) for ONE token whose
// scalars are in c->d_state
static void enqueue_forward(b200_ctx * c) {
    b200_model & m = *c->m;
    const int E = m.n_embd, HD = m.head_dim, KVD = m.n_head_kv * HD, QD = m.n_head * HD, FF = m.n_ff;
    if (m.has_embd() && want(KIND_EMBED)) {
        const int thr = 256;
        g_kind = KIND_EMBED;
        ProfScope ps(c);
        k_embed<<<(E + thr - 1) / thr, thr, 0, c->st>>>(m.embd_type, m.embd_rows, m.embd_row_bytes, E, c->d_state, 0, c->x);
        c->launches++;
    }
    // the persistent kernel runs everything after the embedding row; taps, per-launch profiling and the per-kernel phase
    // trace need the one-kernel-per-operator path
    if (c->token_state > 0 && !c->taps && !c->prof && !c->tracing && g_only_kind < 0) {
        switch (m.n_head / m.n_head_kv) {
            case 1: token_kernel_launch<1>(c); break;
            case 2: token_kernel_launch<2>(c); break;
            case 4: token_kernel_launch<4>(c); break;
            default: token_kernel_launch<8>(c); break;
        }
        return;
    }
    const PfPlan & pp = pf_plan();
    for (int li = 0; li < (int) m.layers.size(); li++) {
        LayerW & L = m.layers[(size_t) li];
        const int il = m.layer_begin + li;
        const size_t gu_bytes = tiled_bytes(L.gateup.m);
        (void) pp; (void) gu_bytes;
        if (want(KIND_QKV)) {   // QKV
            g_kind = KIND_QKV;
            MatvecArgs a = args_qkv(c, li);
#if B200_LOOKAHEAD
            if (pp.kvwo > 0.f) {
                const uint32_t row = (uint32_t) KVD * 2, all = (uint32_t) std::min<size_t>((size_t) c->n_ctx * row, 0xfffffff0u);
                pf_push(a.pf, PfRange{(const uint8_t *) c->kc[(size_t) li], all, row});
                pf_push(a.pf, PfRange{(const uint8_t *) c->vc[(size_t) li], all, row});
                pf_push(a.pf, pf_slice(L.wo.m.p0, tiled_bytes(L.wo.m), 0.0, pp.kvwo));
            }
#endif
            launch_matvec(c, a, EPI_QKV);
            tap(c, "Qcur", il, c->q, (size_t) QD);
        }
        if (want(KIND_ATTN) || want(KIND_ATTN_PV)) {   // attention
            g_kind = KIND_ATTN;
            AttnArgs a{};
            a.q = c->q; a.k_cache = c->kc[(size_t) li]; a.v_cache = c->vc[(size_t) li];
            a.S = c->S; a.s_stride = c->n_ctx; a.out = c->att;
            a.n_head = m.n_head; a.n_head_kv = m.n_head_kv; a.head_dim = HD; a.kv_dim = KVD;
            a.scale = 1.0f / sqrtf((float) HD);
            a.st = c->d_state; a.n_kv_override = 0; a.round_q_override = 0; a.cell_pos = c->d_cell_pos;
#if B200_LOOKAHEAD
            pf_push(a.pf,  pf_slice(L.gateup.m.p0, gu_bytes, 0.0, pp.gu_scores));
            pf_push(a.pf2, pf_slice(L.gateup.m.p0, gu_bytes, pp.gu_scores, pp.gu_scores + pp.gu_pv));
#endif
            launch_attention(c, a, c->n_ctx);
            tap(c, "kqv_merged_cont", il, c->att, (size_t) QD);
        }
        if (want(KIND_WO)) {   // wo + residual
            g_kind = KIND_WO;
            MatvecArgs a = args_wo(c, li);
#if B200_LOOKAHEAD
            pf_push(a.pf, pf_slice(L.gateup.m.p0, gu_bytes, pp.gu_scores + pp.gu_pv, pp.gu_scores + pp.gu_pv + pp.gu_wo));
#endif
            launch_matvec(c, a, EPI_RESID);
            tap(c, "ffn_inp", il, c->x, (size_t) E);
        }
        if (want(KIND_GATEUP)) {   // gate/up (interleaved virtual matrix) + SiLU*mul
            g_kind = KIND_GATEUP;
            MatvecArgs a = args_gateup(c, li);
#if B200_LOOKAHEAD
            pf_push(a.pf, pf_slice(L.down.m.p0, tiled_bytes(L.down.m), 0.0, pp.down));
#endif
            launch_matvec(c, a, EPI_SILU);
            tap(c, "ffn_gate_par", il, c->ffh, (size_t) FF);
        }
        if (want(KIND_DOWN)) {   // down + residual
            g_kind = KIND_DOWN;
            MatvecArgs a = args_down(c, li);
#if B200_LOOKAHEAD
            if (li + 1 < (int) m.layers.size()) {
                const DevStack & nq = m.layers[(size_t) li + 1].qkv;
                pf_push(a.pf, pf_slice(nq.seg[0].p0, tiled_bytes(nq), 0.0, pp.next));
            } else if (m.has_head()) {
                // the head is 431 MB: only its first tiles (what the layer stack's look-ahead depth allows)
                const size_t hb = tiled_bytes(m.output.m);
                pf_push(a.pf, pf_slice(m.output.m.p0, hb, 0.0, pp.next * std::min(1.0, 32.0e6 / (double) hb)));
            }
#endif
            launch_matvec(c, a, EPI_RESID);
            tap(c, "l_out", il, c->x, (size_t) E);
        }
    }
    if (m.has_head() && want(KIND_HEAD)) {
        g_kind = KIND_HEAD;
        MatvecArgs a = args_head(c);
        launch_matvec(c, a, EPI_STORE);
        tap(c, "result_output", -1, c->logits, (size_t) m.n_vocab);
    }
}

static void enqueue_argmax(b200_ctx * c, int advance) {
    k_argmax_partial<<<c->sm_count, 256, 0, c->st>>>(c->logits, c->m->n_vocab, c->amax_key);
    k_argmax_finish<<<1, 32, 0, c->st>>>(c->amax_key, c->d_state, c->d_out_tokens, advance);
    c->launches += 2;
}

static cudaGraphExec_t capture(b200_ctx * c, const std::function<void()> & body, int64_t * n_kernels) {
    cudaGraph_t g;
    const int64_t l0 = c->launches;
    CU(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
    body();
    CU(cudaStreamEndCapture(c->st, &g));
    *n_kernels = c->launches - l0;   // kernels one replay of this graph launches
    c->launches = l0;                // capture is not execution
    cudaGraphExec_t ge;
    CU(cudaGraphInstantiate(&ge, g, 0));
    CU(cudaGraphDestroy(g));
    return ge;
}

extern "C" void b200_ctx_free(b200_ctx * c);
extern "C" b200_ctx * b200_ctx_new(b200_model * m, int n_ctx) {
    std::unique_ptr<b200_ctx> c;
    try {
        require_gpu();
        if (!m) throw std::runtime_error("null model");
        c = std::make_unique<b200_ctx>();
        c->m = m;
        c->device = m->device;
        CU(cudaSetDevice(m->device));
        if (n_ctx <= 0) n_ctx = m->n_ctx_train;
        if (n_ctx <= 0) throw std::runtime_error("n_ctx must be positive (the model names no llama.context_length either)");
        c->n_ctx = (n_ctx + 31) / 32 * 32;                       // GGML_PAD(n_ctx, 32): cpp/src/llama.cpp:16655
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, m->device));
        c->sm_count = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        const int E = m->n_embd, HD = m->head_dim, KVD = m->n_head_kv * HD, QD = m->n_head * HD;
        CU(cudaMalloc(&c->x, (size_t) E * 4));
        CU(cudaMalloc(&c->q, (size_t) QD * 4));
        CU(cudaMalloc(&c->att, (size_t) QD * 4));
        CU(cudaMalloc(&c->ffh, (size_t) m->n_ff * 4));
        {
            std::vector<float> w((size_t) std::max(m->n_ff, m->n_embd));
            for (size_t i = 0; i < w.size(); i++) w[i] = 0.25f + 0.001f * (float) (i % 97);
            CU(cudaMalloc(&c->warm_x, w.size() * 4));
            CU(cudaMemcpy(c->warm_x, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&c->logits, (size_t) m->n_vocab * 4));
        CU(cudaMalloc(&c->S, (size_t) m->n_head * c->n_ctx * 4));
        CU(cudaMalloc(&c->tickets, (size_t) m->n_head_kv * 4));
        CU(cudaMemsetAsync(c->tickets, 0, (size_t) m->n_head_kv * 4, c->st));
        CU(cudaMalloc(&c->d_cell_pos, (size_t) c->n_ctx * 4));
        CU(cudaMemsetAsync(c->d_cell_pos, 0xff, (size_t) c->n_ctx * 4, c->st));
        CU(cudaMalloc(&c->amax_key, 8));
        CU(cudaMemsetAsync(c->amax_key, 0, 8, c->st));
        if (HD != 128) throw std::runtime_error("attention kernels are specialised for head_dim 128");
        for (size_t i = 0; i < m->layers.size(); i++) {
            __half * k = nullptr, * v = nullptr;
            CU(cudaMalloc(&k, (size_t) c->n_ctx * KVD * 2));
            CU(cudaMalloc(&v, (size_t) c->n_ctx * KVD * 2));
            // on the context's own (non-blocking) stream: the legacy default stream does not order against it
            CU(cudaMemsetAsync(k, 0, (size_t) c->n_ctx * KVD * 2, c->st));
            CU(cudaMemsetAsync(v, 0, (size_t) c->n_ctx * KVD * 2, c->st));
            c->kc.push_back(k); c->vc.push_back(v);
        }
        std::vector<float2> tab;
        build_rope_table(*m, c->n_ctx, tab);
        CU(cudaMalloc(&c->rope, tab.size() * sizeof(float2)));
        CU(cudaMemcpyAsync(c->rope, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice, c->st));
        CU(cudaMalloc(&c->d_state, sizeof(DecodeState)));
        CU(cudaMemsetAsync(c->d_state, 0, sizeof(DecodeState), c->st));
        CU(cudaMallocHost(&c->h_state, sizeof(DecodeState)));
        CU(cudaMallocHost(&c->h_logits, (size_t) m->n_vocab * 4));
        CU(cudaMallocHost(&c->h_tok, sizeof(int32_t)));
        c->out_tokens_cap = c->n_ctx + 8;
        CU(cudaMalloc(&c->d_out_tokens, (size_t) c->out_tokens_cap * 4));
        CU(cudaStreamSynchronize(c->st));
        CU(cudaDeviceSynchronize());
        // the persistent kernel's phase list is built now: it allocates, and the first decode may already be a capture
        if (token_kernel_enabled()) token_kernel_build(c.get());
        return c.release();
    } catch (const std::exception & e) {
        set_err(e.what());
        if (c) b200_ctx_free(c.release());     // whatever was allocated so far (cudaFree(nullptr) is a no-op)
        cudaGetLastError();
        return nullptr;
    }
}

extern "C" void b200_ctx_free(b200_ctx * c) {
    if (!c) return;
    cudaSetDevice(c->m->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->g_logits) cudaGraphExecDestroy(c->g_logits);
    if (c->g_greedy) cudaGraphExecDestroy(c->g_greedy);
    if (c->g_pipe)   cudaGraphExecDestroy(c->g_pipe);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->next_inbox) cudaIpcCloseMemHandle(c->next_inbox);
    if (c->first_inbox) cudaIpcCloseMemHandle(c->first_inbox);
    cudaFree(c->inbox); cudaFree(c->p2p_counts);
    for (auto p : c->kc) cudaFree(p);
    for (auto p : c->vc) cudaFree(p);
    cudaFree(c->x); cudaFree(c->q); cudaFree(c->att); cudaFree(c->ffh); cudaFree(c->warm_x); cudaFree(c->logits);
    cudaFree(c->d_cell_pos);
    cudaFree(c->pb.X); cudaFree(c->pb.Q); cudaFree(c->pb.ATT); cudaFree(c->pb.FFH); cudaFree(c->pb.S); cudaFree(c->pb.rec); cudaFree(c->pb.tokens);
    cudaFree(c->d_trace); cudaFree(c->d_plan); cudaFree(c->d_bar); cudaFree(c->S_T); cudaFree(c->d_ttrace);
    cudaFree(c->S); cudaFree(c->tickets); cudaFree(c->amax_key); cudaFree(c->rope); cudaFree(c->d_state); cudaFree(c->d_out_tokens);
    cudaFreeHost(c->h_state); cudaFreeHost(c->h_logits); cudaFreeHost(c->h_tok);
    if (c->g_step) cudaGraphExecDestroy(c->g_step);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    if (c->ev_taken) cudaEventDestroy(c->ev_taken);
    if (c->ev_t0) { cudaEventDestroy(c->ev_t0); cudaEventDestroy(c->ev_t1); }
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

extern "C" void b200_kv_clear(b200_ctx * c) {
    if (!c) return;
    // llama_kv_cache_clear (cpp/src/llama.cpp:3135-3152): every cell empty, head 0. Attention only reads cells that hold a
    // position, so the rows themselves need no clearing.
    cudaSetDevice(c->m->device);
    cudaStreamSynchronize(c->st);
    c->cells = b200_ctx::KvCells();
}

// ------------------------------------------------------------------------------------------------------------
// KV cells: find_slot / seq_rm / seq_add / K-shift (cpp/src/llama.cpp:3028-3125, 3154-3206, 3268-3314, 8482-8510, 15245-15277)
// ------------------------------------------------------------------------------------------------------------
// K-shift: the cached (post-RoPE) K rows of cells whose position changed are rotated by the accumulated delta, in place, as
// ggml_rope_ext_inplace on the f16 cache does (ggml_compute_forward_rope_f16, cpp/ggml/src/ggml.c:14169-14291): f16 -> f32,
// x0*cos - x1*sin / x0*sin + x1*cos (un-fused), f32 -> f16. row_of[cell] selects the (cos, sin) row of the cell's delta
// (-1: untouched — delta 0 is the identity unless the model scales cos/sin by an attention factor, then row 0 holds it).
__global__ void k_kshift(__half * k_cache, int n_cells, int kv_dim, int head_dim, const int32_t * __restrict__ row_of,
                         const float2 * __restrict__ rows) {
    const int half_dim = head_dim / 2, pairs = kv_dim / 2;
    const int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t) n_cells * pairs) return;
    const int cell = (int) (idx / pairs), pr = (int) (idx % pairs);
    const int r = row_of[cell];
    if (r < 0) return;
    const float2 cs = rows[(size_t) r * half_dim + (pr % half_dim)];
    __half * p = k_cache + (size_t) cell * kv_dim + 2 * pr;
    const float x0 = __half2float(p[0]), x1 = __half2float(p[1]);
    p[0] = __float2half_rn(__fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y)));
    p[1] = __float2half_rn(__fadd_rn(__fmul_rn(x0, cs.y), __fmul_rn(x1, cs.x)));
}

static void kv_enter_managed(b200_ctx * c) {
    auto & k = c->cells;
    if (k.managed) return;
    k.managed = true;
    k.pos.assign((size_t) c->n_ctx, -1);
    k.delta.assign((size_t) c->n_ctx, 0);
    for (int i = 0; i < k.n_hi && i < c->n_ctx; i++) k.pos[(size_t) i] = i;
    k.used = std::min(k.n_hi, c->n_ctx);
    k.head = k.used >= c->n_ctx ? 0 : k.used;       // cpp/src/llama.cpp:14821-14826: head += n_tokens, wrapped
    k.has_shift = false;
    CU(cudaMemcpyAsync(c->d_cell_pos, k.pos.data(), (size_t) c->n_ctx * 4, cudaMemcpyHostToDevice, c->st));
    CU(cudaStreamSynchronize(c->st));
}
static int kv_cell_max(const b200_ctx * c) {            // llama_kv_cache_cell_max (cpp/src/llama.cpp:3397-3407)
    for (int i = c->n_ctx; i > 0; i--) if (c->cells.pos[(size_t) i - 1] >= 0) return i;
    return 0;
}
// llama_kv_cache_update_internal: apply the pending K-shift to every local layer, then clear the deltas
static void kv_apply_shift(b200_ctx * c) {
    auto & k = c->cells;
    if (!k.managed || !k.has_shift) return;
    b200_model & m = *c->m;
    const int hd = m.head_dim;
    // the reference rotates EVERY cell by its delta; delta 0 is the identity only when cos(0) * mscale == 1
    float2 probe[128];
    std::vector<float2> row0((size_t) hd / 2);
    build_rope_row(m, 0, row0.data());
    (void) probe;
    const bool identity0 = row0[0].x == 1.0f && row0[0].y == 0.0f;
    std::vector<int32_t> deltas;                        // distinct deltas -> table rows
    std::vector<int32_t> row_of((size_t) c->n_ctx, -1);
    for (int i = 0; i < c->n_ctx; i++) {
        const int d = k.delta[(size_t) i];
        if (d == 0 && identity0) continue;
        size_t r = 0;
        while (r < deltas.size() && deltas[r] != d) r++;
        if (r == deltas.size()) deltas.push_back(d);
        row_of[(size_t) i] = (int32_t) r;
    }
    if (!deltas.empty()) {
        std::vector<float2> rows(deltas.size() * (size_t) (hd / 2));
        for (size_t r = 0; r < deltas.size(); r++) build_rope_row(m, deltas[r], rows.data() + r * (size_t) (hd / 2));
        int32_t * d_row_of = nullptr; float2 * d_rows = nullptr;
        CU(cudaMalloc(&d_row_of, row_of.size() * 4));
        CU(cudaMalloc(&d_rows, rows.size() * sizeof(float2)));
        CU(cudaMemcpyAsync(d_row_of, row_of.data(), row_of.size() * 4, cudaMemcpyHostToDevice, c->st));
        CU(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(float2), cudaMemcpyHostToDevice, c->st));
        const int kvd = m.n_head_kv * hd;
        const int64_t n = (int64_t) c->n_ctx * (kvd / 2);
        for (size_t li = 0; li < c->kc.size(); li++) {
            k_kshift<<<(unsigned) ((n + 255) / 256), 256, 0, c->st>>>(c->kc[li], c->n_ctx, kvd, hd, d_row_of, d_rows);
            c->launches++;
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->st));
        cudaFree(d_row_of); cudaFree(d_rows);
    }
    k.has_shift = false;
    std::fill(k.delta.begin(), k.delta.end(), 0);
}
// llama_kv_cache_find_slot for n tokens of sequence 0 at positions pos0..; returns the first cell (-1: no room), and the
// number of cells to attend. Updates the device copy of the cells' positions.
static int kv_find_slot(b200_ctx * c, int pos0, int n, int * n_kv) {
    auto & k = c->cells;
    const int size = c->n_ctx;
    if (n > size) return -1;
    if (k.head > k.used + 2 * n) k.head = 0;            // cpp/src/llama.cpp:14684-14688
    int n_tested = 0;
    while (true) {
        if (k.head + n > size) { n_tested += size - k.head; k.head = 0; continue; }
        bool found = true;
        for (int i = 0; i < n; i++) {
            if (k.pos[(size_t) (k.head + i)] >= 0) { found = false; k.head += i + 1; n_tested += i + 1; break; }
        }
        if (found) break;
        if (n_tested >= size) return -1;
    }
    for (int i = 0; i < n; i++) k.pos[(size_t) (k.head + i)] = pos0 + i;
    k.used += n;
    const int first = k.head;
    CU(cudaMemcpyAsync(c->d_cell_pos + first, k.pos.data() + first, (size_t) n * 4, cudaMemcpyHostToDevice, c->st));
    *n_kv = kv_cell_max(c);
    k.head += n;                                         // cpp/src/llama.cpp:14821-14826 (after the graph ran)
    if (k.head >= size) k.head = 0;
    return first;
}
// the per-token scalars of token `pos` (cell and attention length included) for a context whose cache was never shifted
static DecodeState identity_state(b200_ctx * c, int32_t token, int pos, int round_q) {
    if (c->cells.managed) throw std::runtime_error("this entry point does not support a context-shifted KV cache; use b200_decode / b200_step_greedy");
    DecodeState hs{};
    hs.token = token; hs.pos = pos; hs.round_q = round_q; hs.step = 0; hs.cell = pos; hs.n_kv = pos + 1; hs.managed = 0;
    return hs;
}
// state of token i of a batch of n at positions pos0.. : cells from find_slot once the cache is managed
struct BatchPlace { int first = -1, n_kv = 0; };
static BatchPlace place_batch(b200_ctx * c, int pos0, int n) {
    BatchPlace b;
    if (!c->cells.managed) return b;
    kv_apply_shift(c);
    b.first = kv_find_slot(c, pos0, n, &b.n_kv);
    if (b.first < 0) throw std::runtime_error("no free KV cells for the batch");   // llama_decode returns 1: cpp/src/llama.cpp:14690
    return b;
}
static DecodeState token_state(b200_ctx * c, const BatchPlace & b, int i, int32_t token, int pos, int round_q) {
    if (!c->cells.managed) return identity_state(c, token, pos, round_q);
    DecodeState hs{};
    hs.token = token; hs.pos = pos; hs.round_q = round_q; hs.step = 0; hs.cell = b.first + i; hs.n_kv = b.n_kv; hs.managed = 1;
    return hs;
}
static void note_positions(b200_ctx * c, int pos_end) { if (!c->cells.managed) c->cells.n_hi = std::max(c->cells.n_hi, pos_end); }

// llama_kv_cache_seq_rm(ctx, 0, p0, p1) (cpp/bridge.cpp:500; cpp/src/llama.cpp:3154-3206)
extern "C" int b200_kv_seq_rm(b200_ctx * c, int p0, int p1) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        CU(cudaSetDevice(c->m->device));
        kv_enter_managed(c);
        auto & k = c->cells;
        if (p0 < 0) p0 = 0;
        if (p1 < 0) p1 = INT32_MAX;
        int new_head = c->n_ctx;
        for (int i = 0; i < c->n_ctx; i++) {
            if (k.pos[(size_t) i] >= p0 && k.pos[(size_t) i] < p1) {
                k.used--;
                k.pos[(size_t) i] = -1;
                if (new_head == c->n_ctx) new_head = i;
            }
        }
        if (new_head != c->n_ctx && new_head < k.head) k.head = new_head;
        CU(cudaMemcpyAsync(c->d_cell_pos, k.pos.data(), (size_t) c->n_ctx * 4, cudaMemcpyHostToDevice, c->st));
        CU(cudaStreamSynchronize(c->st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
// llama_kv_cache_seq_add(ctx, 0, p0, p1, delta) (cpp/bridge.cpp:501; cpp/src/llama.cpp:3268-3314): positions move now, the K
// rows are re-rotated by the accumulated delta at the next decode (llama_kv_cache_update)
extern "C" int b200_kv_seq_add(b200_ctx * c, int p0, int p1, int delta) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        CU(cudaSetDevice(c->m->device));
        kv_enter_managed(c);
        auto & k = c->cells;
        if (p0 < 0) p0 = 0;
        if (p1 < 0) p1 = INT32_MAX;
        if (p0 == p1) return 0;
        int new_head = c->n_ctx;
        for (int i = 0; i < c->n_ctx; i++) {
            if (k.pos[(size_t) i] >= p0 && k.pos[(size_t) i] < p1) {
                k.has_shift = true;
                k.pos[(size_t) i] += delta;
                k.delta[(size_t) i] += delta;
                if (k.pos[(size_t) i] < 0) {
                    k.used--;
                    k.pos[(size_t) i] = -1;
                    if (new_head == c->n_ctx) new_head = i;
                }
            }
        }
        k.head = new_head != c->n_ctx ? new_head : 0;
        CU(cudaMemcpyAsync(c->d_cell_pos, k.pos.data(), (size_t) c->n_ctx * 4, cudaMemcpyHostToDevice, c->st));
        CU(cudaStreamSynchronize(c->st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

// llama_kv_cache_seq_div(ctx, 0, p0, p1, d) (cpp/bridge.cpp:518, Self-Extend; cpp/src/llama.cpp:3316-3349): the positions of the
// cells in [p0, p1) are divided by d (integer division); the K rows are re-rotated by the accumulated delta at the next decode
// like after a seq_add. Several cells may then hold the same position; attention only compares positions.
extern "C" int b200_kv_seq_div(b200_ctx * c, int p0, int p1, int d) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        if (d <= 0) throw std::runtime_error("b200_kv_seq_div: the divisor must be positive");
        CU(cudaSetDevice(c->m->device));
        kv_enter_managed(c);
        auto & k = c->cells;
        if (p0 < 0) p0 = 0;
        if (p1 < 0) p1 = INT32_MAX;
        if (p0 == p1) return 0;
        for (int i = 0; i < c->n_ctx; i++) {
            if (k.pos[(size_t) i] >= p0 && k.pos[(size_t) i] < p1) {
                k.has_shift = true;
                const int32_t p_old = k.pos[(size_t) i];
                k.pos[(size_t) i] /= d;
                k.delta[(size_t) i] += k.pos[(size_t) i] - p_old;
            }
        }
        CU(cudaMemcpyAsync(c->d_cell_pos, k.pos.data(), (size_t) c->n_ctx * 4, cudaMemcpyHostToDevice, c->st));
        CU(cudaStreamSynchronize(c->st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

// Rows [pos0, pos0 + n) of one local layer's K (post-RoPE) and V cache, f16 bits [n][n_head_kv * head_dim] — the view
// llama_kv_cache exposes through k_l / v_l (cpp/src/llama.cpp:2495-2539; V returned position-major). Tests use the pair
// to start both sides of a parity check from the same cache at n_kv in the thousands, and to check the K-shift.
extern "C" int b200_kv_write(b200_ctx * c, int layer, int pos0, int n, const uint16_t * k_rows, const uint16_t * v_rows) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        const int li = layer - c->m->layer_begin;
        if (li < 0 || li >= (int) c->kc.size()) throw std::runtime_error("layer is not on this stage");
        if (pos0 < 0 || n < 0 || pos0 + n > c->n_ctx) throw std::runtime_error("positions exceed n_ctx");
        const size_t row = (size_t) c->m->n_head_kv * c->m->head_dim * 2;
        CU(cudaSetDevice(c->m->device));
        if (c->cells.managed) throw std::runtime_error("b200_kv_write addresses rows by position: not available after a context shift");
        note_positions(c, pos0 + n);
        if (k_rows) CU(cudaMemcpyAsync((uint8_t *) c->kc[(size_t) li] + pos0 * row, k_rows, n * row, cudaMemcpyHostToDevice, c->st));
        if (v_rows) CU(cudaMemcpyAsync((uint8_t *) c->vc[(size_t) li] + pos0 * row, v_rows, n * row, cudaMemcpyHostToDevice, c->st));
        CU(cudaStreamSynchronize(c->st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
extern "C" int b200_kv_read(b200_ctx * c, int layer, int pos0, int n, uint16_t * k_rows, uint16_t * v_rows) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        const int li = layer - c->m->layer_begin;
        if (li < 0 || li >= (int) c->kc.size()) throw std::runtime_error("layer is not on this stage");
        if (pos0 < 0 || n < 0 || pos0 + n > c->n_ctx) throw std::runtime_error("positions exceed n_ctx");
        const size_t row = (size_t) c->m->n_head_kv * c->m->head_dim * 2;
        CU(cudaSetDevice(c->m->device));
        if (k_rows) CU(cudaMemcpyAsync(k_rows, (const uint8_t *) c->kc[(size_t) li] + pos0 * row, n * row, cudaMemcpyDeviceToHost, c->st));
        if (v_rows) CU(cudaMemcpyAsync(v_rows, (const uint8_t *) c->vc[(size_t) li] + pos0 * row, n * row, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" void b200_set_taps(b200_ctx * c, int enable) { if (c) { c->taps = enable != 0; c->tapstore.v.clear(); } }
extern "C" int64_t b200_get_tap(b200_ctx * c, const char * name, int layer, float * out, int64_t cap) {
    if (!c || !name) return 0;
    auto it = c->tapstore.v.find(std::string(name) + "-" + std::to_string(layer));
    if (it == c->tapstore.v.end()) return 0;
    const int64_t n = (int64_t) it->second.size();
    if (out) std::memcpy(out, it->second.data(), (size_t) std::min(n, cap) * 4);
    return n;
}
extern "C" void b200_timings(b200_ctx * c, double * tp, int64_t * np, double * tg, int64_t * ng) {
    if (!c || !tp || !np || !tg || !ng) return;
    *tp = c->t_prompt_us; *np = c->n_prompt; *tg = c->t_gen_us; *ng = c->n_gen;
}
extern "C" void b200_reset_timings(b200_ctx * c) { if (!c) return; c->t_prompt_us = c->t_gen_us = 0; c->n_prompt = c->n_gen = 0; }

static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------------------
// prompt batches (prefill.cuh): llama_decode with n_tokens > 1 as batched kernels — every weight tile is fetched once per 64
// tokens. Same arithmetic as the token-by-token path (same device functions), so the logits are bit-identical to it.
// ------------------------------------------------------------------------------------------------------------
static constexpr int PB_MAX_T = 512;               // tokens per pass (the reference's n_ubatch: cpp/common/common.h:81)
static constexpr size_t PB_SCORE_BYTES = (size_t) 256 << 20;   // score buffer budget: tokens per attention launch group = what fits (>= 64)
static int g_prefill_batch = -1;       // 1: prompt batches run the batched kernels (default), 0: token by token (A/B, tests)
extern "C" void b200_set_prefill_batch(int on) { g_prefill_batch = on ? 1 : 0; }
static bool prefill_batch_enabled() {
    if (g_prefill_batch < 0) { const char * e = getenv("BOOSTER_B200_PREFILL_BATCH"); g_prefill_batch = (e && e[0] == '0') ? 0 : 1; }
    return g_prefill_batch == 1;
}
static int g_prefill_attn_batch = 1;   // 1: k_attn_softmax_rows + k_attn_pv_batch, 0: the per-token attention kernels over blockIdx.z
extern "C" void b200_set_prefill_attn_batch(int on) { g_prefill_attn_batch = on ? 1 : 0; }
static bool prefill_batch_usable(const b200_ctx * c, int n) {
    const b200_model & m = *c->m;
    if (!prefill_batch_enabled() || n < 8 || c->taps || c->cells.managed || m.head_dim != 128) return false;   // (stages too: b200_stage_forward_batch)
    const int gqa = m.n_head / m.n_head_kv;
    if (gqa != 1 && gqa != 2 && gqa != 4 && gqa != 8) return false;
    if (g_prefill_attn_batch == 0 && (size_t) gqa * c->n_ctx * 4 + (size_t) PV_BATCH * 16 > 200 * 1024) return false;   // the two-launch attention must fit
    if (m.n_ff % 256 || m.n_embd % 256 || std::max(m.n_ff, m.n_embd) / 256 > 4 * 16 * 4) return false;
    return true;
}
static void prefill_alloc(b200_ctx * c) {
    if (c->pb.cap) return;
    const b200_model & m = *c->m;
    const int T = PB_MAX_T, QD = m.n_head * m.head_dim;
    const size_t kmax = (size_t) std::max(std::max(m.n_ff, m.n_embd), QD);
    CU(cudaMalloc(&c->pb.X, (size_t) T * m.n_embd * 4));
    CU(cudaMalloc(&c->pb.Q, (size_t) T * QD * 4));
    CU(cudaMalloc(&c->pb.ATT, (size_t) T * QD * 4));
    CU(cudaMalloc(&c->pb.FFH, (size_t) T * m.n_ff * 4));
    {
        const size_t per_token = (size_t) m.n_head * c->n_ctx * 4;
        c->pb.att_z = (int) std::max<size_t>(64, std::min<size_t>((size_t) T, PB_SCORE_BYTES / per_token / 32 * 32));
        if (g_prefill_attn_batch == 0) c->pb.att_z = 64;   // (the per-token kernels index blockIdx.z the same way)
    }
    CU(cudaMalloc(&c->pb.S, (size_t) c->pb.att_z * m.n_head * c->n_ctx * 4));
    const size_t rec_bytes = std::max((size_t) (T / PB_CHUNK) * (kmax / 256) * PB_CHUNK * pb_record_bytes(0, 1),
                                      std::max((size_t) (T / MB_NT) * (kmax / 256) * MB_REC_BYTES, (size_t) (T / UM_NT) * (kmax / 256) * UM_REC_BYTES));
    CU(cudaMalloc(&c->pb.rec, rec_bytes));
    CU(cudaMemset(c->pb.rec, 0, rec_bytes));       // records of tokens beyond a partial last chunk are read (and discarded)
    CU(cudaMalloc(&c->pb.tokens, (size_t) T * 4));
    c->pb.cap = T;
}
static bool has_q6k(const TMat * seg, int n_seg) { for (int i = 0; i < n_seg; i++) if (seg[i].type == T_Q6_K) return true; return false; }
// K-quant launches run on the tensor cores (prefill_mma.cuh) when every segment has an even number of 32-row units
static int g_prefill_mma = -1;         // 1 (default, the fastest measured): mma.sync (k_mma_batch), 2: tcgen05 (k_umma_batch), 0: the dp4a kernel
extern "C" void b200_set_prefill_mma(int mode) { g_prefill_mma = mode < 0 ? 0 : mode > 2 ? 2 : mode; }
// record layout / kernel of a K-quant launch: 2 = k_umma_batch (segments of a multiple of 4 units), 1 = k_mma_batch (even), 0 = dp4a
static int pb_mma_mode(const MatvecArgs & mv) {
    if (g_prefill_mma < 0) { const char * e = getenv("BOOSTER_B200_PREFILL_MMA"); g_prefill_mma = e ? std::max(0, std::min(2, atoi(e))) : 1; }
    if (g_prefill_mma == 0) return 0;
    if (mv.act_q8_0) {   // Q8_0 x Q8_0: block-diagonal HMMA (k_mma_batch_q80)
        for (int i = 0; i < mv.n_seg; i++) if (mv.seg[i].type != T_Q8_0 || (mv.seg[i].n_units & 1)) return 0;
        return 3;
    }
    int mode = g_prefill_mma;
    for (int i = 0; i < mv.n_seg; i++) {
        const int t = mv.seg[i].type;
        if (t != T_Q4_K && t != T_Q5_K && t != T_Q6_K) return 0;
        if (mv.seg[i].n_units & 1) return 0;
        if (mv.seg[i].n_units & 3) mode = 1;
    }
    return mode;
}
static bool pb_use_mma(const MatvecArgs & mv) { return pb_mma_mode(mv) != 0; }
static void pb_quant(b200_ctx * c, const float * X, int k, int T, const float * norm_w, int act_q8_0, int with_as, int mma = 0) {
    QuantBatchArgs a{};
    a.X = X; a.k = k; a.T = T; a.norm_w = norm_w; a.eps = norm_w ? c->m->rms_eps : 0.f; a.inv_k = (k & (k - 1)) == 0 ? 1.0 / (double) k : 0.0;
    a.act_q8_0 = act_q8_0; a.with_as = with_as; a.rec = c->pb.rec; a.layout = mma;
    const size_t smem = act_smem_bytes(k, act_q8_0);
    static size_t attr[64] = {0};
    if (smem > 48 * 1024) raise_smem(k_quant_batch, smem, attr, c->device);
    if (norm_w && k / 256 > PRO_U * 16) throw std::runtime_error("normed vector too long for the batched quantizer");
    if (mma) {
        static size_t attr_m[64] = {0};
        if (smem > 48 * 1024) raise_smem(k_quant_batch_mma, smem, attr_m, c->device);
        k_quant_batch_mma<<<(unsigned) T, 512, smem, c->st>>>(a);
    } else {
        k_quant_batch<<<(unsigned) T, 512, smem, c->st>>>(a);
    }
    c->launches++;
}
static void pb_matmul(b200_ctx * c, const MatvecArgs & mv, int epi, int T, int pos0, float * out, int out_stride, const float * resid, int with_as) {
    MatmulBatchArgs a{};
    for (int i = 0; i < mv.n_seg; i++) a.seg[i] = mv.seg[i];
    a.n_seg = mv.n_seg; a.n_units = mv.n_units; a.k = mv.k; a.tiles_unit = mv.seg[0].tiles_unit;
    a.act_q8_0 = mv.act_q8_0; a.with_as = with_as; a.rec = c->pb.rec; a.T = T; a.epi = epi;
    a.out = out; a.out_stride = out_stride; a.resid = resid; a.resid_stride = out_stride;
    a.q_out = c->pb.Q; a.q_stride = mv.n_q; a.k_cache = mv.k_cache; a.v_cache = mv.v_cache;
    a.n_q = mv.n_q; a.n_k = mv.n_k; a.head_dim = mv.head_dim; a.kv_dim = mv.kv_dim; a.rope = mv.rope; a.pos0 = pos0;
    int sb = 0;
    for (int i = 0; i < mv.n_seg; i++) sb = std::max(sb, tile_bytes_of(mv.seg[i].type));
    const int mma_mode = pb_mma_mode(mv);
    bool q4 = false, q5 = false, q6 = false;
    for (int i = 0; i < mv.n_seg; i++) { q4 |= mv.seg[i].type == T_Q4_K; q5 |= mv.seg[i].type == T_Q5_K; q6 |= mv.seg[i].type == T_Q6_K; }
    if (mma_mode == 3) {
        a.mb_a_bytes = (uint32_t) Q80_A_BYTES;
        a.mb_raw_stride = (uint32_t) ((8 * sb + 127) / 128 * 128);
        a.mb_rec_copy = (uint32_t) Q80_REC_BYTES;
        a.mb_stage_bytes = 2 * a.mb_raw_stride + (a.mb_rec_copy + 127) / 128 * 128;
        a.mb_stages = (int) std::min<size_t>(MB_MAX_STAGES, ((size_t) 226 * 1024 - 2 * a.mb_a_bytes) / a.mb_stage_bytes);
        if (a.mb_stages < 2) throw std::runtime_error("k_mma_batch_q80: shared memory layout does not fit");
        const size_t smem_q = (size_t) 2 * a.mb_a_bytes + (size_t) a.mb_stages * a.mb_stage_bytes;
        static size_t attr_q[64] = {0};
        raise_smem(k_mma_batch_q80, smem_q, attr_q, c->device);
        const dim3 grid_q((unsigned) ((T + MB_NT - 1) / MB_NT), (unsigned) (a.n_units / 2));
        k_mma_batch_q80<<<grid_q, MB_WARPS * 32, smem_q, c->st>>>(a);
        c->launches++;
        return;
    }
    if (mma_mode == 2) {
        a.mb_a_bytes = (uint32_t) um_a_bytes(q6);
        a.mb_raw_stride = (uint32_t) ((sb + 127) / 128 * 128);
        a.mb_rec_copy = (uint32_t) um_rec_copy_bytes(q4, q5);
        a.mb_stage_bytes = 4 * a.mb_raw_stride + (a.mb_rec_copy + 127) / 128 * 128;
        const size_t budget = 226 * 1024;
        a.mb_stages = (int) std::min<size_t>(UM_MAX_STAGES, (budget - a.mb_a_bytes) / a.mb_stage_bytes);
        if (a.mb_stages < 2) throw std::runtime_error("k_umma_batch: shared memory layout does not fit");
        const size_t smem_u = std::max((size_t) a.mb_a_bytes + (size_t) a.mb_stages * a.mb_stage_bytes, (size_t) 120 * 1024);   // one CTA per SM (TMEM)
        static size_t attr_u[64] = {0};
        raise_smem(k_umma_batch, smem_u, attr_u, c->device);
        const dim3 grid_u((unsigned) ((T + UM_NT - 1) / UM_NT), (unsigned) (a.n_units / 4));
        k_umma_batch<<<grid_u, UM_WARPS * 32, smem_u, c->st>>>(a);
        c->launches++;
        return;
    }
    if (mma_mode == 1) {
        a.mb_a_bytes = (uint32_t) mb_a_bytes(q6);             // one of the two A buffers
        a.mb_raw_stride = (uint32_t) ((sb + 127) / 128 * 128);
        a.mb_rec_copy = (uint32_t) mb_rec_copy_bytes(q4, q5);
        a.mb_stage_bytes = 2 * a.mb_raw_stride + (a.mb_rec_copy + 127) / 128 * 128;
        const size_t budget = 226 * 1024;
        a.mb_stages = (int) std::min<size_t>(MB_MAX_STAGES, (budget - 2 * a.mb_a_bytes) / a.mb_stage_bytes);
        if (a.mb_stages < 2) throw std::runtime_error("k_mma_batch: shared memory layout does not fit");
        const size_t smem_m = std::max((size_t) 2 * a.mb_a_bytes + (size_t) a.mb_stages * a.mb_stage_bytes, (size_t) MB_CHAIN_BYTES);
        static size_t attr_m[64] = {0};
        raise_smem(k_mma_batch, smem_m, attr_m, c->device);
        const dim3 grid_m((unsigned) ((T + MB_NT - 1) / MB_NT), (unsigned) (a.n_units / 2));
        k_mma_batch<<<grid_m, MB_WARPS * 32, smem_m, c->st>>>(a);
        c->launches++;
        return;
    }
    const size_t smem = PB_STAGES * pb_stage_bytes(sb, a.act_q8_0, with_as);
    static size_t attr[64] = {0};
    raise_smem(k_matmul_batch, smem, attr, c->device);
    const dim3 grid((unsigned) ((T + PB_CHUNK - 1) / PB_CHUNK), (unsigned) a.n_units);
    k_matmul_batch<<<grid, PB_WARPS * 32, smem, c->st>>>(a, sb);
    c->launches++;
}
template <int GQA>
static void pb_attention(b200_ctx * c, int li, int T, int pos0) {
    const b200_model & m = *c->m;
    const int HD = m.head_dim, KVD = m.n_head_kv * HD, QD = m.n_head * HD;
    for (int z0 = 0; z0 < T; z0 += c->pb.att_z) {
        const int nz = std::min(c->pb.att_z, T - z0);
        AttnArgs a{};
        a.q = c->pb.Q + (size_t) z0 * QD; a.k_cache = c->kc[(size_t) li]; a.v_cache = c->vc[(size_t) li];
        a.S = c->pb.S; a.s_stride = c->n_ctx; a.out = c->pb.ATT + (size_t) z0 * QD;
        a.n_head = m.n_head; a.n_head_kv = m.n_head_kv; a.head_dim = HD; a.kv_dim = KVD;
        a.scale = 1.0f / sqrtf((float) HD);
        a.st = nullptr; a.n_kv_override = pos0 + z0 + 1; a.round_q_override = 1; a.cell_pos = nullptr;   // batch > 1: q rounded to f16
        a.zq = QD; a.zs = m.n_head * c->n_ctx;
        const int n_pad_max = std::min((pos0 + z0 + nz + 31) / 32 * 32, c->n_ctx);
        if (g_prefill_attn_batch == 0) {   // A/B: the per-token kernels with the token in blockIdx.z
            if (!launch_attention_2k<GQA>(c, a, n_pad_max, nz)) throw std::runtime_error("batched attention does not fit");
            continue;
        }
        const dim3 gs((unsigned) a.n_head_kv, (unsigned) ((n_pad_max + ATT_TILE - 1) / ATT_TILE), (unsigned) ((nz + SCB_TQ - 1) / SCB_TQ));
        k_attn_scores_batch<GQA><<<gs, ATT_THREADS, 0, c->st>>>(a, nz);
        k_attn_softmax_rows<<<(unsigned) ((nz * a.n_head + 7) / 8), 256, 0, c->st>>>(a, nz);
        static size_t attr_pvb[64] = {0};
        raise_smem(k_attn_pv_batch<GQA>, pvb_smem_bytes(), attr_pvb, c->device);
        constexpr int TQ = PVB_ROWS / GQA;
        const dim3 gp((unsigned) a.n_head_kv, (unsigned) (128 / PVB_DIMS), (unsigned) ((nz + TQ - 1) / TQ));
        k_attn_pv_batch<GQA><<<gp, 256, pvb_smem_bytes(), c->st>>>(a, nz);
        c->launches += 3;
    }
}
// the context's layers (a stage's share of them) over the residual streams c->pb.X[T][n_embd] of T tokens at positions p0..
static void prefill_layers(b200_ctx * c, int T, int p0) {
    b200_model & m = *c->m;
    const int E = m.n_embd, FF = m.n_ff, QD = m.n_head * m.head_dim;
    for (int li = 0; li < (int) m.layers.size(); li++) {
        LayerW & L = m.layers[(size_t) li];
        const int q80 = L.qkv.seg[0].type == T_Q8_0;
        const MatvecArgs aq = args_qkv(c, li), ao = args_wo(c, li), ag = args_gateup(c, li), ad = args_down(c, li);
        int was = has_q6k(aq.seg, aq.n_seg);
        pb_quant(c, c->pb.X, E, T, L.attn_norm, q80, was, pb_mma_mode(aq));
        pb_matmul(c, aq, EPI_QKV, T, p0, nullptr, 0, nullptr, was);
        switch (m.n_head / m.n_head_kv) {
            case 1: pb_attention<1>(c, li, T, p0); break;
            case 2: pb_attention<2>(c, li, T, p0); break;
            case 4: pb_attention<4>(c, li, T, p0); break;
            default: pb_attention<8>(c, li, T, p0); break;
        }
        was = has_q6k(ao.seg, 1);
        pb_quant(c, c->pb.ATT, QD, T, nullptr, q80, was, pb_mma_mode(ao));
        pb_matmul(c, ao, EPI_RESID, T, p0, c->pb.X, E, c->pb.X, was);
        was = has_q6k(ag.seg, 1);
        pb_quant(c, c->pb.X, E, T, L.ffn_norm, q80, was, pb_mma_mode(ag));
        pb_matmul(c, ag, EPI_SILU, T, p0, c->pb.FFH, FF, nullptr, was);
        was = has_q6k(ad.seg, 1);
        pb_quant(c, c->pb.FFH, FF, T, nullptr, q80, was, pb_mma_mode(ad));
        pb_matmul(c, ad, EPI_RESID, T, p0, c->pb.X, E, c->pb.X, was);
    }
    CU(cudaGetLastError());
}
static void prefill_embed(b200_ctx * c, const int32_t * tokens, int T) {
    b200_model & m = *c->m;
    CU(cudaMemcpyAsync(c->pb.tokens, tokens, (size_t) T * 4, cudaMemcpyHostToDevice, c->st));
    k_embed_batch<<<dim3((unsigned) ((m.n_embd + 255) / 256), (unsigned) T), 256, 0, c->st>>>(m.embd_type, m.embd_rows, m.embd_row_bytes, m.n_embd, c->pb.tokens, c->pb.X);
    c->launches++;
}
// tokens[0..n) at positions pos0..: the whole prompt batch through every layer; leaves the LAST token's residual stream in c->x
static void prefill_batch(b200_ctx * c, const int32_t * tokens, int n, int pos0) {
    b200_model & m = *c->m;
    prefill_alloc(c);
    for (int t0 = 0; t0 < n; t0 += PB_MAX_T) {
        const int T = std::min(PB_MAX_T, n - t0);
        prefill_embed(c, tokens + t0, T);
        prefill_layers(c, T, pos0 + t0);
    }
    const int last = (n - 1) % PB_MAX_T;
    CU(cudaMemcpyAsync(c->x, c->pb.X + (size_t) last * m.n_embd, (size_t) m.n_embd * 4, cudaMemcpyDeviceToDevice, c->st));
}

// In-process layer split (bridge pods over several GPUs): one prompt chunk of n <= 512 tokens through THIS stage's layers with
// the batched kernels. The first stage embeds the tokens; every other stage takes the predecessor's residual streams
// [n][n_embd] with one peer copy (the batch counterpart of b200_stage_forward's hand-off, same events); the last stage
// leaves the logits of the chunk's last token. Returns 2 when the batched kernels cannot run this chunk (the caller then
// feeds the tokens one by one through b200_stage_forward).
extern "C" int b200_stage_batch_usable(b200_ctx * c, int n) { return c && n <= PB_MAX_T && prefill_batch_usable(c, n) ? 1 : 0; }
extern "C" int b200_stage_forward_batch(b200_ctx * c, const int32_t * tokens, int n, int pos0, b200_ctx * prev) {
    try {
        require_gpu();
        if (!c || !tokens || n <= 0) throw std::runtime_error("bad arguments");
        b200_model & m = *c->m;
        if (pos0 < 0 || pos0 + n > c->n_ctx) throw std::runtime_error("positions exceed n_ctx");
        if (m.has_embd() != (prev == nullptr)) throw std::runtime_error("stage chain mismatch: only the first stage has no predecessor");
        if (m.has_embd()) for (int i = 0; i < n; i++) if (tokens[i] < 0 || tokens[i] >= m.n_vocab) throw std::runtime_error("token id out of range");
        if (n > PB_MAX_T || !prefill_batch_usable(c, n) || (prev && !prev->pb.cap)) return 2;
        CU(cudaSetDevice(m.device));
        prefill_alloc(c);
        if (c->taken_pending) { CU(cudaStreamWaitEvent(c->st, c->ev_taken, 0)); c->taken_pending = false; }
        if (prev) {
            if (!prev->ev_done) { CU(cudaSetDevice(prev->m->device)); CU(cudaEventCreateWithFlags(&prev->ev_done, cudaEventDisableTiming)); CU(cudaSetDevice(m.device)); }
            if (!prev->ev_taken) CU(cudaEventCreateWithFlags(&prev->ev_taken, cudaEventDisableTiming));
            CU(cudaSetDevice(prev->m->device));
            CU(cudaEventRecord(prev->ev_done, prev->st));
            CU(cudaSetDevice(m.device));
            CU(cudaStreamWaitEvent(c->st, prev->ev_done, 0));
            CU(cudaMemcpyPeerAsync(c->pb.X, m.device, prev->pb.X, prev->m->device, (size_t) n * m.n_embd * 4, c->st));
            CU(cudaEventRecord(prev->ev_taken, c->st));
            prev->taken_pending = true;
        } else {
            prefill_embed(c, tokens, n);
        }
        note_positions(c, pos0 + n);
        prefill_layers(c, n, pos0);
        if (m.has_head()) {
            CU(cudaMemcpyAsync(c->x, c->pb.X + (size_t) (n - 1) * m.n_embd, (size_t) m.n_embd * 4, cudaMemcpyDeviceToDevice, c->st));
            g_only_kind = KIND_HEAD;
            try { enqueue_forward(c); } catch (...) { g_only_kind = -1; throw; }
            g_only_kind = -1;
        }
        CU(cudaGetLastError());
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" int b200_decode(b200_ctx * c, const int32_t * tokens, int n, int pos0, float * logits_out) {
    try {
        require_gpu();
        if (!c || !tokens || n <= 0) throw std::runtime_error("bad arguments");
        b200_model & m = *c->m;
        if (!m.has_embd() || !m.has_head()) throw std::runtime_error("b200_decode needs a single-stage model; use b200_pipeline_decode");
        if (pos0 < 0 || pos0 + n > c->n_ctx) throw std::runtime_error("positions exceed n_ctx");   // find_slot failure: llama.cpp:14690
        for (int i = 0; i < n; i++) if (tokens[i] < 0 || tokens[i] >= m.n_vocab) throw std::runtime_error("token id out of range");
        CU(cudaSetDevice(m.device));
        const double t0 = now_us();
        // reference semantics for batch > 1: q is rounded to f16 before K.q (cpp/ggml/src/ggml.c:12345-12371)
        const int round_q = n > 1 ? 1 : 0;
        if (prefill_batch_usable(c, n)) {   // (b200_decode is single-stage: checked above)
            // the prompt batch through the batched kernels, then the head on the last token's residual stream
            note_positions(c, pos0 + n);
            prefill_batch(c, tokens, n, pos0);
            g_only_kind = KIND_HEAD;
            try { enqueue_forward(c); } catch (...) { g_only_kind = -1; throw; }
            g_only_kind = -1;
            if (logits_out) CU(cudaMemcpyAsync(c->h_logits, c->logits, (size_t) m.n_vocab * 4, cudaMemcpyDeviceToHost, c->st));
            CU(cudaStreamSynchronize(c->st));
            if (logits_out) std::memcpy(logits_out, c->h_logits, (size_t) m.n_vocab * 4);
            c->t_prompt_us += now_us() - t0; c->n_prompt += n;
            return 0;
        }
        const BatchPlace place = place_batch(c, pos0, n);
        note_positions(c, pos0 + n);
        for (int i = 0; i < n; i++) {
            const bool last = i == n - 1;
            *c->h_state = token_state(c, place, i, tokens[i], pos0 + i, round_q);
            if (c->taps) {
                CU(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(DecodeState), cudaMemcpyHostToDevice, c->st));
                enqueue_forward(c);
                CU(cudaStreamSynchronize(c->st));
                if (last && logits_out) CU(cudaMemcpy(logits_out, c->logits, (size_t) m.n_vocab * 4, cudaMemcpyDeviceToHost));
                continue;
            }
            if (!c->g_logits) {
                c->g_logits = capture(c, [&]() {
                    CU(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(DecodeState), cudaMemcpyHostToDevice, c->st));
                    enqueue_forward(c);
                    CU(cudaMemcpyAsync(c->h_logits, c->logits, (size_t) m.n_vocab * 4, cudaMemcpyDeviceToHost, c->st));
                }, &c->n_logits);
            }
            CU(cudaGraphLaunch(c->g_logits, c->st));
            c->launches += c->n_logits;
            CU(cudaStreamSynchronize(c->st));   // h_state is reused for the next token
            if (last && logits_out) std::memcpy(logits_out, c->h_logits, (size_t) m.n_vocab * 4);
        }
        const double dt = now_us() - t0;
        if (n > 1) { c->t_prompt_us += dt; c->n_prompt += n; } else { c->t_gen_us += dt; c->n_gen += 1; }
        return 0;
    } catch (const std::exception & e) {
        return set_err(e.what());
    }
}

extern "C" int b200_generate_greedy(b200_ctx * c, int32_t first_token, int pos0, int n_steps, int32_t * out_tokens) {
    try {
        require_gpu();
        if (!c || n_steps <= 0) throw std::runtime_error("bad arguments");
        b200_model & m = *c->m;
        if (!m.has_embd() || !m.has_head()) throw std::runtime_error("single-stage model required; use b200_pipeline_generate_greedy");
        if (pos0 < 0 || pos0 + n_steps > c->n_ctx) throw std::runtime_error("positions exceed n_ctx");
        if (n_steps > c->out_tokens_cap) throw std::runtime_error("n_steps too large");
        if (first_token < 0 || first_token >= m.n_vocab) throw std::runtime_error("token id out of range");
        CU(cudaSetDevice(m.device));
        const double t0 = now_us();
        *c->h_state = identity_state(c, first_token, pos0, 0);
        note_positions(c, pos0 + n_steps);
        CU(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(DecodeState), cudaMemcpyHostToDevice, c->st));
        if (!c->g_greedy) {
            CU(cudaStreamSynchronize(c->st));
            c->g_greedy = capture(c, [&]() {
                enqueue_forward(c);
                enqueue_argmax(c, 1);
            }, &c->n_greedy);
        }
        if (!c->ev_t0) { CU(cudaEventCreate(&c->ev_t0)); CU(cudaEventCreate(&c->ev_t1)); }
        CU(cudaEventRecord(c->ev_t0, c->st));
        // BOOSTER_B200_NO_GRAPH=1 (A/B): the same kernels as plain stream launches — the token's scalars live in device
        // memory, so nothing but the launch mechanism changes
        static int no_graph = -1;
        if (no_graph < 0) { const char * e = getenv("BOOSTER_B200_NO_GRAPH"); no_graph = (e && e[0] == '1') ? 1 : 0; }
        if (no_graph) {
            for (int s = 0; s < n_steps; s++) { enqueue_forward(c); enqueue_argmax(c, 1); }
        } else {
            for (int s = 0; s < n_steps; s++) CU(cudaGraphLaunch(c->g_greedy, c->st));
            c->launches += c->n_greedy * (int64_t) n_steps;
        }
        CU(cudaEventRecord(c->ev_t1, c->st));
        if (out_tokens) CU(cudaMemcpyAsync(out_tokens, c->d_out_tokens, (size_t) n_steps * 4, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        CU(cudaEventElapsedTime(&c->last_device_ms, c->ev_t0, c->ev_t1));
        c->t_gen_us += now_us() - t0; c->n_gen += n_steps;
        return 0;
    } catch (const std::exception & e) {
        return set_err(e.what());
    }
}

// One generated token through a single-stage context, the way the bridge's generation loop needs it: decode `token`
// at `pos` (batch-1 arithmetic) and hand back the arg-max of its logits — llama_decode + sample_top_token
// (cpp/bridge.cpp:549-560, 962-981) as ONE CUDA-graph replay (state H2D -> forward -> arg-max -> 4-byte D2H) and one
// synchronisation, instead of ~195 plain launches per token.
extern "C" int b200_step_greedy(b200_ctx * c, int32_t token, int pos, int32_t * next_token) {
    try {
        require_gpu();
        if (!c || !next_token) throw std::runtime_error("bad arguments");
        b200_model & m = *c->m;
        if (!m.has_embd() || !m.has_head()) throw std::runtime_error("b200_step_greedy needs a single-stage context (use b200_stage_forward for layer splits)");
        if (pos < 0 || pos >= c->n_ctx) throw std::runtime_error("position exceeds n_ctx");
        if (token < 0 || token >= m.n_vocab) throw std::runtime_error("token id out of range");
        CU(cudaSetDevice(m.device));
        const BatchPlace place = place_batch(c, pos, 1);
        note_positions(c, pos + 1);
        *c->h_state = token_state(c, place, 0, token, pos, 0);
        if (c->taps) {                                         // taps need the un-graphed path
            CU(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(DecodeState), cudaMemcpyHostToDevice, c->st));
            enqueue_forward(c);
            enqueue_argmax(c, 0);
            CU(cudaMemcpyAsync(c->h_tok, c->d_out_tokens, 4, cudaMemcpyDeviceToHost, c->st));
        } else {
            if (!c->g_step) {
                c->g_step = capture(c, [&]() {
                    CU(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(DecodeState), cudaMemcpyHostToDevice, c->st));
                    enqueue_forward(c);
                    enqueue_argmax(c, 0);
                    CU(cudaMemcpyAsync(c->h_tok, c->d_out_tokens, 4, cudaMemcpyDeviceToHost, c->st));
                }, &c->n_step);
            }
            CU(cudaGraphLaunch(c->g_step, c->st));
            c->launches += c->n_step;
        }
        CU(cudaStreamSynchronize(c->st));
        *next_token = *c->h_tok;
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" float b200_last_device_ms(const b200_ctx * c) { return c ? c->last_device_ms : 0.f; }

// One token, un-graphed, with a CUDA-event pair around every launch on the engine's own stream:
// ms_by_kind[k] = summed device time of kind k, n_by_kind[k] = launches of that kind
// (kinds: 0 embed, 1 qkv, 2 attention(partial+combine), 3 wo, 4 gate/up, 5 down, 6 head).
extern "C" int b200_profile_token(b200_ctx * c, int32_t token, int pos, float ms_by_kind[8], int32_t n_by_kind[8]) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        b200_model & m = *c->m;
        if (pos < 0 || pos >= c->n_ctx || token < 0 || token >= m.n_vocab) throw std::runtime_error("bad token/pos");
        CU(cudaSetDevice(m.device));
        const DecodeState hs = identity_state(c, token, pos, 0);
        k_set_state<<<1, 1, 0, c->st>>>(c->d_state, hs);
        c->prof = true; c->prof_ev.clear();
        enqueue_forward(c);
        c->prof = false;
        CU(cudaStreamSynchronize(c->st));
        for (int k = 0; k < 8; k++) { ms_by_kind[k] = 0.f; n_by_kind[k] = 0; }
        for (auto & pe : c->prof_ev) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, pe.second.first, pe.second.second));
            ms_by_kind[pe.first] += ms; n_by_kind[pe.first] += 1;
            cudaEventDestroy(pe.second.first); cudaEventDestroy(pe.second.second);
        }
        c->prof_ev.clear();
        return 0;
    } catch (const std::exception & e) { if (c) c->prof = false; return set_err(e.what()); }
}

// Kernel-in-isolation timing: the launches of ONE kind (e.g. the gate|up mat-vec) of every layer of this stage, back to
// back on the engine's stream (each layer has its own weights, so nothing is re-read from L2), `reps` times, between ONE
// pair of CUDA events. ms_total / n_launches is the kernel's steady-state launch-to-launch time without the token's
// dependency chain around it (the next launch's weight prefetch overlaps the previous one's tail, as in the token).
extern "C" int b200_profile_kind(b200_ctx * c, int kind, int pos, int reps, float * ms_total, int32_t * n_launches) {
    try {
        require_gpu();
        if (!c || kind < 0 || kind >= KIND_COUNT || reps <= 0) throw std::runtime_error("bad arguments");
        b200_model & m = *c->m;
        if (pos < 0 || pos >= c->n_ctx) throw std::runtime_error("bad pos");
        CU(cudaSetDevice(m.device));
        const DecodeState hs = identity_state(c, 0, pos, 0);
        k_set_state<<<1, 1, 0, c->st>>>(c->d_state, hs);
        if (!c->ev_t0) { CU(cudaEventCreate(&c->ev_t0)); CU(cudaEventCreate(&c->ev_t1)); }
        const int64_t l0 = c->launches;
        g_only_kind = kind;
        try {
            enqueue_forward(c);                                  // warm-up round
            CU(cudaEventRecord(c->ev_t0, c->st));
            const int64_t l1 = c->launches;
            for (int r = 0; r < reps; r++) enqueue_forward(c);
            CU(cudaEventRecord(c->ev_t1, c->st));
            *n_launches = (int32_t) (c->launches - l1);
        } catch (...) { g_only_kind = -1; throw; }
        g_only_kind = -1;
        (void) l0;
        CU(cudaStreamSynchronize(c->st));
        CU(cudaEventElapsedTime(ms_total, c->ev_t0, c->ev_t1));
        return 0;
    } catch (const std::exception & e) { g_only_kind = -1; return set_err(e.what()); }
}

// One token through a freshly captured graph whose kernels stamp %globaltimer at their phase boundaries (thread 0 of
// every CTA). out = [n_launches][TRACE_CTAS=512][8] u64 (0 = not stamped), meta = [n_launches][2] (kind, ctas).
// Phases: k_matvec 0 start | 1 ring filled, before griddepcontrol.wait | 2 after it | 3 prologue done | 4 end;
//         k_attn_scores 0 start | 1 after wait | 2 scores written;
//         k_attn_softmax_pv 0 start | 1 after wait | 2 score rows staged | 3 normalised | 4 P.V done.
extern "C" int64_t b200_trace_token(b200_ctx * c, int32_t token, int pos, int reps, uint64_t * out, int64_t cap_words, int32_t * meta, int64_t cap_meta) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        b200_model & m = *c->m;
        if (pos < 0 || pos >= c->n_ctx || token < 0 || token >= m.n_vocab) throw std::runtime_error("bad token/pos");
        CU(cudaSetDevice(m.device));
        const size_t words = (size_t) TRACE_MAX_LAUNCHES * TRACE_CTAS * TRACE_PHASES;
        if (!c->d_trace) CU(cudaMalloc(&c->d_trace, words * 8));
        const DecodeState hs = identity_state(c, token, pos, 0);
        c->tracing = true; c->trace_seq = 0; c->trace_meta.clear();
        int64_t nk = 0;
        cudaGraphExec_t ge = nullptr;
        try {
            ge = capture(c, [&]() { enqueue_forward(c); enqueue_argmax(c, 0); }, &nk);
        } catch (...) { c->tracing = false; throw; }
        c->tracing = false;
        for (int r = 0; r < std::max(1, reps); r++) {
            k_set_state<<<1, 1, 0, c->st>>>(c->d_state, hs);
            CU(cudaMemsetAsync(c->d_trace, 0, words * 8, c->st));
            CU(cudaStreamSynchronize(c->st));
            CU(cudaGraphLaunch(ge, c->st));
            CU(cudaStreamSynchronize(c->st));
        }
        cudaGraphExecDestroy(ge);
        const int64_t nl = c->trace_seq;
        const int64_t need = nl * TRACE_CTAS * TRACE_PHASES;
        if (out && cap_words >= need) CU(cudaMemcpy(out, c->d_trace, (size_t) need * 8, cudaMemcpyDeviceToHost));
        if (meta) for (int64_t i = 0; i < nl && 2 * i + 1 < cap_meta; i++) { meta[2 * i] = c->trace_meta[(size_t) i][0]; meta[2 * i + 1] = c->trace_meta[(size_t) i][1]; }
        return nl;
    } catch (const std::exception & e) { if (c) c->tracing = false; set_err(e.what()); return -1; }
}

// One token through the persistent kernel's tracing instantiation: out = [n_phases][n_ctas][4] globaltimer stamps (ns):
// 0 the phase's dependent half starts | 1 it is done | 2 arrived at the grid barrier and the next phase's independent half
// issued | 3 barrier passed. kinds[n_phases] = 0 mat-vec, 1 attention scores, 2 soft-max + P.V. Returns n_phases (-1: error,
// 0: the persistent kernel is not in use for this context).
extern "C" int64_t b200_trace_phases(b200_ctx * c, int32_t token, int pos, int reps, uint64_t * out, int64_t cap_words, int32_t * kinds,
                                     int64_t cap_kinds, int32_t * n_ctas) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        b200_model & m = *c->m;
        if (pos < 0 || pos >= c->n_ctx || token < 0 || token >= m.n_vocab) throw std::runtime_error("bad token/pos");
        CU(cudaSetDevice(m.device));
        if (c->token_state <= 0) return 0;
        const size_t words = (size_t) c->n_phases * c->sm_count * TK_TRACE_SLOTS;
        if (!c->d_ttrace) CU(cudaMalloc(&c->d_ttrace, words * 8));
        const DecodeState hs = identity_state(c, token, pos, 0);
        for (int r = 0; r < std::max(1, reps); r++) {
            k_set_state<<<1, 1, 0, c->st>>>(c->d_state, hs);
            CU(cudaMemsetAsync(c->d_ttrace, 0, words * 8, c->st));
            c->ttracing = true;
            try { enqueue_forward(c); } catch (...) { c->ttracing = false; throw; }
            c->ttracing = false;
            CU(cudaStreamSynchronize(c->st));
        }
        if (out && cap_words >= (int64_t) words) CU(cudaMemcpy(out, c->d_ttrace, words * 8, cudaMemcpyDeviceToHost));
        if (kinds) {
            std::vector<Phase> plan((size_t) c->n_phases);
            CU(cudaMemcpy(plan.data(), c->d_plan, plan.size() * sizeof(Phase), cudaMemcpyDeviceToHost));
            for (int i = 0; i < c->n_phases && i < cap_kinds; i++) kinds[i] = plan[(size_t) i].kind == PH_MATVEC ? 10 + plan[(size_t) i].mv.epi : plan[(size_t) i].kind;
        }
        if (n_ctas) *n_ctas = c->sm_count;
        return c->n_phases;
    } catch (const std::exception & e) { set_err(e.what()); return -1; }
}

// ------------------------------------------------------------------------------------------------------------
// pipeline over NCCL: rank r holds one stage; one send/recv of f32[n_embd] per boundary per token
// ------------------------------------------------------------------------------------------------------------
extern "C" int b200_comm_unique_id(uint8_t id[128]) {
    try { if (!id) throw std::runtime_error("null id"); nccl_load(); NC(g_nccl.GetUniqueId(id)); return 0; } catch (const std::exception & e) { return set_err(e.what()); }
}
extern "C" int b200_comm_init(b200_ctx * c, int rank, int world, const uint8_t id[128]) {
    try {
        require_gpu();
        if (!c || !id || world < 1 || rank < 0 || rank >= world) throw std::runtime_error("bad arguments");
        nccl_load();
        CU(cudaSetDevice(c->m->device));
        NcclId nid; std::memcpy(nid.b, id, 128);
        NC(g_nccl.CommInitRank(&c->comm, world, nid, rank));
        c->rank = rank; c->world = world;
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

// ------------------------------------------------------------------------------------------------------------
// direct peer hand-off (replaces the ncclSend / ncclRecv pair of a stage boundary, which costs ~20 us of launch + proxy
// latency for a 16 KiB message that NVLink moves in well under a microsecond): the producer's last step is a small kernel
// that STORES the residual stream into the consumer's inbox over NVLink and then publishes a sequence number
// (st.release.sys); the consumer's first step spins on that number (ld.acquire.sys) and copies the vector out.
// The inboxes are device allocations exported through CUDA IPC (one process per GPU). Message i of a direction carries
// sequence number i: both ends count in device memory, so the kernels are captured once into the stage's CUDA graph.
// ------------------------------------------------------------------------------------------------------------
static constexpr int P2P_MAX_EMBD = 16384;
struct P2PInbox {
    float x[P2P_MAX_EMBD];
    unsigned long long seq_x;           // number of residual vectors delivered
    unsigned long long seq_tok;         // number of tokens delivered
    int32_t token;
    int32_t pad_[3];
};
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long * p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __noinline__ void p2p_timeout(const char * what, unsigned long long have, unsigned long long want) {
    printf("booster_b200: peer hand-off timeout waiting for %s: have %llu want %llu\n", what, have, want);
    __trap();
}
__device__ __forceinline__ void p2p_spin(const unsigned long long * seq, unsigned long long want, const char * what) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned spins = 0;
    while (ld_acquire_sys_u64(seq) < want) {
        if ((++spins & 4095u) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 30000000000ull) p2p_timeout(what, ld_acquire_sys_u64(seq), want);     // 30 s: a peer died
        }
    }
}
__global__ void k_p2p_send_x(const float * __restrict__ x, int n, P2PInbox * peer, unsigned long long * counts) {
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) reinterpret_cast<float4 *>(peer->x)[i] = reinterpret_cast<const float4 *>(x)[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { const unsigned long long s = ++counts[0]; st_release_sys_u64(&peer->seq_x, s); }
}
__global__ void k_p2p_recv_x(float * __restrict__ x, int n, const P2PInbox * mine, unsigned long long * counts) {
    __shared__ unsigned long long want;
    if (threadIdx.x == 0) { want = ++counts[1]; p2p_spin(&mine->seq_x, want, "the residual stream"); }
    __syncthreads();
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) reinterpret_cast<float4 *>(x)[i] = reinterpret_cast<const float4 *>(mine->x)[i];
}
__global__ void k_p2p_send_token(const DecodeState * st, P2PInbox * peer, unsigned long long * counts) {
    if (threadIdx.x != 0) return;
    peer->token = st->token;
    __threadfence_system();
    const unsigned long long s = ++counts[2];
    st_release_sys_u64(&peer->seq_tok, s);
}
__global__ void k_p2p_recv_token(DecodeState * st, const P2PInbox * mine, unsigned long long * counts) {
    if (threadIdx.x != 0) return;
    const unsigned long long want = ++counts[3];
    p2p_spin(&mine->seq_tok, want, "the sampled token");
    st->token = *reinterpret_cast<const volatile int32_t *>(&mine->token);
}

// this rank's inbox as a 64-byte CUDA IPC handle (allocates it on first use)
extern "C" int b200_p2p_handle(b200_ctx * c, uint8_t handle[64]) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        if (c->m->n_embd > P2P_MAX_EMBD || c->m->n_embd % 4) throw std::runtime_error("n_embd not supported by the peer hand-off");
        CU(cudaSetDevice(c->m->device));
        if (!c->inbox) {
            CU(cudaMalloc(&c->inbox, sizeof(P2PInbox)));
            CU(cudaMemset(c->inbox, 0, sizeof(P2PInbox)));
            CU(cudaMalloc(&c->p2p_counts, 4 * sizeof(unsigned long long)));
            CU(cudaMemset(c->p2p_counts, 0, 4 * sizeof(unsigned long long)));
        }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, c->inbox));
        std::memcpy(handle, &h, 64);
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
// map the inboxes this rank writes to: rank + 1's (next_handle; ignored on the last rank) and rank 0's (first_handle; used by
// the last rank only). After this call b200_pipeline_* hands off over NVLink stores instead of NCCL.
extern "C" int b200_p2p_connect(b200_ctx * c, int rank, int world, const uint8_t next_handle[64], const uint8_t first_handle[64]) {
    try {
        require_gpu();
        if (!c || !c->inbox) throw std::runtime_error("b200_p2p_handle must be called first");
        CU(cudaSetDevice(c->m->device));
        c->rank = rank; c->world = world;
        auto open = [&](const uint8_t * hb) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, hb, 64);
            void * p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            return (P2PInbox *) p;
        };
        if (rank + 1 < world) c->next_inbox = open(next_handle);
        if (rank == world - 1 && world > 1) c->first_inbox = open(first_handle);
        c->p2p = true;
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
// back to ncclSend / ncclRecv for this context (every rank of the group must take the same decision: a rank that could not
// map a peer's inbox reports it through the host-side process group and all of them call this)
extern "C" void b200_p2p_disable(b200_ctx * c) { if (c) c->p2p = false; }

// one token through this rank's stage: [recv x] -> forward -> [send x] ; last stage: argmax -> token to stage 0
static void enqueue_stage_step(b200_ctx * c, bool greedy) {
    b200_model & m = *c->m;
    const bool first = c->rank == 0, last = c->rank == c->world - 1;
    // The inbox of the peer hand-off is ONE slot: it is safe only while at most one token is in flight, which the greedy loop
    // guarantees (stage 0 cannot start token i + 1 before the last stage has handed token i's arg-max back). Prompt tokens
    // (greedy == false) have no hand-back — stage 0 would overwrite the slot while a slower stage still has to read it — so they
    // go through ncclSend / ncclRecv, whose FIFO is the flow control.
    if (c->p2p && greedy) {
        if (!first) { k_p2p_recv_x<<<1, 256, 0, c->st>>>(c->x, m.n_embd, c->inbox, c->p2p_counts); c->launches++; }
        enqueue_forward(c);
        if (!last) { k_p2p_send_x<<<1, 256, 0, c->st>>>(c->x, m.n_embd, c->next_inbox, c->p2p_counts); c->launches++; }
        if (greedy) {
            if (last) {
                enqueue_argmax(c, 1);
                if (c->world > 1) { k_p2p_send_token<<<1, 32, 0, c->st>>>(c->d_state, c->first_inbox, c->p2p_counts); c->launches++; }
            } else {
                k_advance<<<1, 32, 0, c->st>>>(c->d_state);
                if (first) { k_p2p_recv_token<<<1, 32, 0, c->st>>>(c->d_state, c->inbox, c->p2p_counts); c->launches++; }
            }
            c->launches++;
        }
        CU(cudaGetLastError());
        return;
    }
    if (!first) NC(g_nccl.Recv(c->x, (size_t) m.n_embd, NCCL_FLOAT32, c->rank - 1, c->comm, c->st));
    enqueue_forward(c);
    if (!last) NC(g_nccl.Send(c->x, (size_t) m.n_embd, NCCL_FLOAT32, c->rank + 1, c->comm, c->st));
    if (greedy) {
        if (last) {
            enqueue_argmax(c, 1);
            if (c->world > 1) NC(g_nccl.Send(&c->d_state->token, 1, NCCL_INT32, 0, c->comm, c->st));
        } else {
            k_advance<<<1, 32, 0, c->st>>>(c->d_state);
            if (first) NC(g_nccl.Recv(&c->d_state->token, 1, NCCL_INT32, c->world - 1, c->comm, c->st));
        }
        c->launches++;
    }
}

extern "C" int b200_pipeline_generate_greedy(b200_ctx * c, int32_t first_token, int pos0, int n_steps, int32_t * out_tokens) {
    try {
        require_gpu();
        if (!c || n_steps <= 0) throw std::runtime_error("bad arguments");
        if (c->world == 1) return b200_generate_greedy(c, first_token, pos0, n_steps, out_tokens);
        if (!c->comm) throw std::runtime_error("b200_comm_init not called");
        b200_model & m = *c->m;
        // the same validation as b200_generate_greedy, identical on EVERY rank (all ranks see the same arguments), so that
        // no rank blocks in ncclRecv while another one returns an error
        if (pos0 < 0 || pos0 + n_steps > c->n_ctx) throw std::runtime_error("positions exceed n_ctx");
        if (n_steps > c->out_tokens_cap) throw std::runtime_error("n_steps too large");
        if (first_token < 0 || first_token >= m.n_vocab) throw std::runtime_error("token id out of range");
        CU(cudaSetDevice(m.device));
        *c->h_state = identity_state(c, first_token, pos0, 0);
        note_positions(c, pos0 + n_steps);
        CU(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(DecodeState), cudaMemcpyHostToDevice, c->st));
        // device time of the burst on THIS rank's stream (first stage step enqueued -> last one complete; a stage's
        // stream idles inside ncclRecv while the other stages work, so every rank's span covers the whole burst)
        if (!c->ev_t0) { CU(cudaEventCreate(&c->ev_t0)); CU(cudaEventCreate(&c->ev_t1)); }
        // One stage step = [ncclRecv x] -> this stage's kernels -> [ncclSend x] -> arg-max / advance (+ token hand-back),
        // captured ONCE into a CUDA graph (NCCL point-to-point calls are capturable) and replayed per token, like the
        // single-GPU loop. The first burst runs un-graphed so that NCCL sets up its peer connections outside a capture;
        // BOOSTER_B200_PIPE_GRAPH=0 keeps it that way. A failed capture falls back to plain launches, never to another path.
        static int pipe_graph = -1;
        if (pipe_graph < 0) { const char * e = getenv("BOOSTER_B200_PIPE_GRAPH"); pipe_graph = e ? atoi(e) : 1; }
        if (pipe_graph && c->pipe_calls >= 1 && !c->g_pipe && !c->pipe_graph_failed) {
            const int64_t l0 = c->launches;
            try {
                c->g_pipe = capture(c, [&]() { enqueue_stage_step(c, true); }, &c->n_pipe);
            } catch (const std::exception & e) {
                cudaGraph_t junk = nullptr;
                cudaStreamEndCapture(c->st, &junk);            // leave capture mode whatever state it is in
                if (junk) cudaGraphDestroy(junk);
                cudaGetLastError();
                c->launches = l0; c->g_pipe = nullptr; c->pipe_graph_failed = true;
                fprintf(stderr, "booster_b200: pipeline stage graph capture failed (%s); using plain launches\n", e.what());
            }
        }
        c->pipe_calls++;
        CU(cudaEventRecord(c->ev_t0, c->st));
        if (c->g_pipe) {
            for (int s = 0; s < n_steps; s++) CU(cudaGraphLaunch(c->g_pipe, c->st));
            c->launches += c->n_pipe * (int64_t) n_steps;
        } else {
            for (int s = 0; s < n_steps; s++) enqueue_stage_step(c, true);
        }
        CU(cudaEventRecord(c->ev_t1, c->st));
        // every rank gets the ids: last rank broadcasts point-to-point
        const bool last = c->rank == c->world - 1;
        NC(g_nccl.GroupStart());
        if (last) { for (int r = 0; r < c->world - 1; r++) NC(g_nccl.Send(c->d_out_tokens, (size_t) n_steps, NCCL_INT32, r, c->comm, c->st)); }
        else      NC(g_nccl.Recv(c->d_out_tokens, (size_t) n_steps, NCCL_INT32, c->world - 1, c->comm, c->st));
        NC(g_nccl.GroupEnd());
        if (out_tokens) CU(cudaMemcpyAsync(out_tokens, c->d_out_tokens, (size_t) n_steps * 4, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        CU(cudaEventElapsedTime(&c->last_device_ms, c->ev_t0, c->ev_t1));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" int b200_pipeline_decode(b200_ctx * c, const int32_t * tokens, int n, int pos0, float * logits_out) {
    try {
        require_gpu();
        if (!c || !tokens || n <= 0) throw std::runtime_error("bad arguments");
        if (c->world == 1) return b200_decode(c, tokens, n, pos0, logits_out);
        if (!c->comm) throw std::runtime_error("b200_comm_init not called");
        b200_model & m = *c->m;
        if (pos0 < 0 || pos0 + n > c->n_ctx) throw std::runtime_error("positions exceed n_ctx");
        for (int i = 0; i < n; i++) if (tokens[i] < 0 || tokens[i] >= m.n_vocab) throw std::runtime_error("token id out of range");
        CU(cudaSetDevice(m.device));
        const int round_q = n > 1 ? 1 : 0;
        for (int i = 0; i < n; i++) {
            *c->h_state = identity_state(c, tokens[i], pos0 + i, round_q);
            note_positions(c, pos0 + i + 1);
            CU(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(DecodeState), cudaMemcpyHostToDevice, c->st));
            enqueue_stage_step(c, false);
            CU(cudaStreamSynchronize(c->st));
        }
        if (c->rank == c->world - 1 && logits_out) CU(cudaMemcpy(logits_out, c->logits, (size_t) m.n_vocab * 4, cudaMemcpyDeviceToHost));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}


// ------------------------------------------------------------------------------------------------------------
// in-process layer split: stage chain inside one process (what the Go server drives through the bridge)
// ------------------------------------------------------------------------------------------------------------
extern "C" int b200_stage_forward(b200_ctx * c, int32_t token, int pos, int batch_gt1, b200_ctx * prev) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        b200_model & m = *c->m;
        if (pos < 0 || pos >= c->n_ctx) throw std::runtime_error("position exceeds n_ctx");
        if (m.has_embd() && (token < 0 || token >= m.n_vocab)) throw std::runtime_error("token id out of range");
        if (m.has_embd() != (prev == nullptr)) throw std::runtime_error("stage chain mismatch: only the first stage has no predecessor");
        CU(cudaSetDevice(m.device));
        // back edge of the hand-off: this stage's kernels overwrite c->x, which the NEXT stage may still be copying for
        // the previous token (prompt tokens are enqueued back to back without a host sync) — wait until it has been taken
        if (c->taken_pending) { CU(cudaStreamWaitEvent(c->st, c->ev_taken, 0)); c->taken_pending = false; }
        if (prev) {
            // hand-off of the residual stream l_out -> next device (cf. cpp/ggml/src/ggml-cuda.cu:2386-2407)
            if (!prev->ev_done) { CU(cudaSetDevice(prev->m->device)); CU(cudaEventCreateWithFlags(&prev->ev_done, cudaEventDisableTiming)); CU(cudaSetDevice(m.device)); }
            // (an event is recorded on a stream of ITS device: ev_taken belongs to the consumer's device)
            if (!prev->ev_taken) CU(cudaEventCreateWithFlags(&prev->ev_taken, cudaEventDisableTiming));
            CU(cudaSetDevice(prev->m->device));
            CU(cudaEventRecord(prev->ev_done, prev->st));
            CU(cudaSetDevice(m.device));
            CU(cudaStreamWaitEvent(c->st, prev->ev_done, 0));
            CU(cudaMemcpyPeerAsync(c->x, m.device, prev->x, prev->m->device, (size_t) m.n_embd * 4, c->st));
            CU(cudaEventRecord(prev->ev_taken, c->st));
            prev->taken_pending = true;
        }
        const BatchPlace place = place_batch(c, pos, 1);
        note_positions(c, pos + 1);
        const DecodeState hs = token_state(c, place, 0, token, pos, batch_gt1 ? 1 : 0);
        k_set_state<<<1, 1, 0, c->st>>>(c->d_state, hs);
        c->launches++;
        enqueue_forward(c);
        CU(cudaGetLastError());
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
// block until everything enqueued on this stage's stream has completed (the bridge brackets each prompt chunk with it,
// as the reference brackets each blocking llama_decode with its timer: cpp/bridge.cpp:549-560)
extern "C" int b200_stage_sync(b200_ctx * c) {
    try {
        require_gpu();
        if (!c) throw std::runtime_error("null context");
        CU(cudaSetDevice(c->m->device));
        CU(cudaStreamSynchronize(c->st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
extern "C" int b200_stage_logits(b200_ctx * c, float * logits_out) {
    try {
        require_gpu();
        if (!c || !c->m->has_head()) throw std::runtime_error("not the last stage");
        if (!logits_out) throw std::runtime_error("null output");
        CU(cudaSetDevice(c->m->device));
        CU(cudaMemcpyAsync(c->h_logits, c->logits, (size_t) c->m->n_vocab * 4, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        std::memcpy(logits_out, c->h_logits, (size_t) c->m->n_vocab * 4);
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
// the last stage's logits in the context's own pinned host buffer (valid until the next call on this context; the caller may
// modify them — the bridge's sampler does, as the reference's sampler modifies llama_get_logits' buffer): no second copy
extern "C" float * b200_stage_logits_view(b200_ctx * c) {
    try {
        require_gpu();
        if (!c || !c->m->has_head()) throw std::runtime_error("not the last stage");
        CU(cudaSetDevice(c->m->device));
        CU(cudaMemcpyAsync(c->h_logits, c->logits, (size_t) c->m->n_vocab * 4, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        return c->h_logits;
    } catch (const std::exception & e) { set_err(e.what()); return nullptr; }
}
// llama_decode of one token on a single-stage context + llama_get_logits, as ONE graph replay (state in, forward, logits out
// to the pinned buffer) and one synchronisation; returns the pinned buffer (see b200_stage_logits_view)
extern "C" float * b200_decode_view(b200_ctx * c, int32_t token, int pos) {
    if (b200_decode(c, &token, 1, pos, nullptr) != 0) return nullptr;
    return c->taps ? nullptr : c->h_logits;
}
extern "C" int b200_stage_argmax(b200_ctx * c, int32_t * token_out) {
    try {
        require_gpu();
        if (!c || !c->m->has_head()) throw std::runtime_error("not the last stage");
        if (!token_out) throw std::runtime_error("null output");
        CU(cudaSetDevice(c->m->device));
        enqueue_argmax(c, 0);
        CU(cudaMemcpyAsync(&c->h_state->token, c->d_out_tokens, 4, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        *token_out = c->h_state->token;
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

// ------------------------------------------------------------------------------------------------------------
// tokenizer entry points (host side): llama_tokenize / llama_token_to_piece / llama_token_is_eog
// ------------------------------------------------------------------------------------------------------------
struct b200_tokenizer { std::unique_ptr<b200::Tokenizer> t; };
extern "C" b200_tokenizer * b200_tokenizer_load(const char * gguf_path) {
    try {
        if (!gguf_path) throw std::runtime_error("null path");
        std::string err;
        auto t = b200::make_tokenizer(gguf_path, err);
        if (!t) throw std::runtime_error(err);
        auto * h = new b200_tokenizer();
        h->t = std::move(t);
        return h;
    } catch (const std::exception & e) { set_err(e.what()); return nullptr; }
}
extern "C" void b200_tokenizer_free(b200_tokenizer * h) { delete h; }
extern "C" int32_t b200_tokenizer_n_vocab(const b200_tokenizer * h) { return h ? h->t->n_vocab() : 0; }
extern "C" int32_t b200_tokenize(const b200_tokenizer * h, const char * text, int32_t text_len, int32_t * tokens, int32_t n_max,
                                 int add_special, int parse_special) {
    try {
        if (!h || !text || text_len < 0) throw std::runtime_error("bad arguments");
        std::vector<int32_t> out;
        if (!h->t->tokenize(std::string(text, (size_t) text_len), add_special != 0, parse_special != 0, out))
            throw std::runtime_error("text cannot be tokenized (malformed UTF-8, a byte without a token, or ids out of range)");
        if ((int64_t) out.size() > (int64_t) n_max) return -(int32_t) out.size();      // llama_tokenize_impl: cpp/src/llama-vocab.cpp:1497-1516
        for (size_t i = 0; i < out.size(); i++) tokens[i] = out[i];
        return (int32_t) out.size();
    } catch (const std::exception & e) { set_err(e.what()); return INT32_MIN; }
}
extern "C" int32_t b200_token_to_piece(const b200_tokenizer * h, int32_t token, char * buf, int32_t length, int special) {
    if (!h) return 0;
    const std::string p = h->t->piece(token, special != 0);
    if ((int64_t) p.size() > (int64_t) length) return -(int32_t) p.size();
    std::memcpy(buf, p.data(), p.size());
    return (int32_t) p.size();
}
extern "C" int b200_token_is_eog(const b200_tokenizer * h, int32_t token) { return h && h->t->is_eog(token) ? 1 : 0; }
extern "C" int32_t b200_token_nl(const b200_tokenizer * h) { return h ? h->t->linefeed() : -1; }

// ------------------------------------------------------------------------------------------------------------
// operator-level entry points (tests): host in, host out, SAME kernels
// ------------------------------------------------------------------------------------------------------------
namespace {
struct DBuf {
    void * p = nullptr;
    explicit DBuf(size_t n) { CU(cudaMalloc(&p, n ? n : 1)); }
    ~DBuf() { cudaFree(p); }
    template <typename T> T * as() { return (T *) p; }
};
// a stream / a tiled matrix that an operator call owns: released on every exit, the error paths included
struct OpStream {
    cudaStream_t st = nullptr;
    OpStream() { CU(cudaStreamCreate(&st)); }
    ~OpStream() { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } }
};
struct OwnedMat {
    DevMat d;
    explicit OwnedMat(DevMat m) : d(m) {}
    ~OwnedMat() { cudaFree(d.alloc); }
};
}  // namespace

static int op_quantize(const float * x, int64_t k, void * out, int q80) {
    try {
        require_gpu();
        if (k <= 0 || k % 256 != 0) throw std::runtime_error("k must be a positive multiple of 256");
        const size_t ob = q80 ? (size_t) (k / 32) * 34 : (size_t) (k / 256) * 292;
        DBuf dx((size_t) k * 4), dout(ob);
        CU(cudaMemcpy(dx.p, x, (size_t) k * 4, cudaMemcpyHostToDevice));
        const size_t smem = act_smem_bytes((int) k, q80);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_quantize_export, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        k_quantize_export<<<1, 384, smem>>>(dx.as<float>(), (int) k, q80, dout.as<uint8_t>());
        CU(cudaGetLastError());
        CU(cudaMemcpy(out, dout.p, ob, cudaMemcpyDeviceToHost));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
extern "C" int b200_op_quantize_q8_K(const float * x, int64_t k, void * out) { return op_quantize(x, k, out, 0); }
extern "C" int b200_op_quantize_q8_0(const float * x, int64_t k, void * out) { return op_quantize(x, k, out, 1); }

extern "C" int b200_op_dequantize_row(int type, const void * w, int64_t k, float * y) {
    try {
        require_gpu();
        const size_t rb = (size_t) ggml_row_bytes((uint32_t) type, (uint64_t) k);
        if (rb == 0) throw std::runtime_error("unsupported type");
        DBuf dw(rb), dy((size_t) k * 4);
        CU(cudaMemcpy(dw.p, w, rb, cudaMemcpyHostToDevice));
        k_embed<<<(unsigned) ((k + 255) / 256), 256>>>(type, dw.as<uint8_t>(), rb, (int) k, nullptr, 0, dy.as<float>());
        CU(cudaGetLastError());
        CU(cudaMemcpy(y, dy.p, (size_t) k * 4, cudaMemcpyDeviceToHost));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" int b200_op_mul_mat_vec(int type, const void * w, int64_t n_rows, int64_t k, const float * x, float * y) {
    try {
        require_gpu();
        OpStream os;
        cudaStream_t st = os.st;
        // the operator accepts any row count: pad with all-zero blocks (d = 0 -> 0.0) up to the work-unit height
        const int64_t rows_pad = (n_rows + 31) / 32 * 32;
        std::vector<uint8_t> padded;
        if (rows_pad != n_rows) {
            const size_t rb = ggml_row_bytes(type, k);
            padded.assign((size_t) rows_pad * rb, 0);
            memcpy(padded.data(), w, (size_t) n_rows * rb);
            w = padded.data();
        }
        OwnedMat om(upload_matrix(type, w, rows_pad, k, st));
        const DevMat & d = om.d;
        DBuf dx((size_t) k * 4), dy((size_t) rows_pad * 4);
        CU(cudaMemcpyAsync(dx.p, x, (size_t) k * 4, cudaMemcpyHostToDevice, st));
        b200_ctx tmp;   // only st / sm_count / launches are used by launch_matvec
        tmp.st = st;
        int dev = 0; CU(cudaGetDevice(&dev));
        cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, dev));
        tmp.sm_count = prop.multiProcessorCount;
        tmp.device = dev;
        MatvecArgs a{};
        a.seg[0] = d.m; a.n_seg = 1; a.n_units = d.m.n_units; a.k = (int) k;
        a.x = dx.as<float>(); a.norm_w = nullptr; a.act_q8_0 = type == T_Q8_0; a.out = dy.as<float>();
        launch_matvec(&tmp, a, EPI_STORE);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(y, dy.p, (size_t) n_rows * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

// the batched mat-mul of a prompt batch as an operator: y[T][n_rows] = W x[t] for T <= 512 tokens, through k_quant_batch(_mma) +
// k_mma_batch / k_matmul_batch exactly as prefill_batch() launches them (tests: worst-case magnitudes, ragged T)
extern "C" int b200_op_mul_mat(int type, const void * w, int64_t n_rows, int64_t k, const float * x, int64_t T, float * y) {
    try {
        require_gpu();
        if (T <= 0 || T > PB_MAX_T) throw std::runtime_error("b200_op_mul_mat: 1..512 tokens");
        OpStream os;
        cudaStream_t st = os.st;
        const int64_t rows_pad = (n_rows + 63) / 64 * 64;
        std::vector<uint8_t> padded;
        if (rows_pad != n_rows) {
            const size_t rb = ggml_row_bytes(type, k);
            padded.assign((size_t) rows_pad * rb, 0);
            memcpy(padded.data(), w, (size_t) n_rows * rb);
            w = padded.data();
        }
        OwnedMat om(upload_matrix(type, w, rows_pad, k, st));
        const DevMat & d = om.d;
        const size_t rec_bytes = std::max((size_t) ((T + PB_CHUNK - 1) / PB_CHUNK) * (k / 256) * PB_CHUNK * pb_record_bytes(0, 1),
                                          std::max((size_t) ((T + MB_NT - 1) / MB_NT) * (k / 256) * MB_REC_BYTES, (size_t) ((T + UM_NT - 1) / UM_NT) * (k / 256) * UM_REC_BYTES));
        DBuf dx((size_t) T * k * 4), dy((size_t) T * rows_pad * 4), drec(rec_bytes);
        CU(cudaMemsetAsync(drec.p, 0, rec_bytes, st));
        CU(cudaMemcpyAsync(dx.p, x, (size_t) T * k * 4, cudaMemcpyHostToDevice, st));
        b200_ctx tmp;
        tmp.st = st;
        int dev = 0; CU(cudaGetDevice(&dev));
        tmp.device = dev;
        tmp.pb.rec = drec.as<uint8_t>();
        MatvecArgs a{};
        a.seg[0] = d.m; a.n_seg = 1; a.n_units = d.m.n_units; a.k = (int) k; a.act_q8_0 = type == T_Q8_0;
        const int was = type == T_Q6_K;
        pb_quant(&tmp, dx.as<float>(), (int) k, (int) T, nullptr, a.act_q8_0, was, pb_mma_mode(a));
        pb_matmul(&tmp, a, EPI_STORE, (int) T, 0, dy.as<float>(), (int) rows_pad, nullptr, was);
        CU(cudaGetLastError());
        std::vector<float> hy((size_t) T * rows_pad);
        CU(cudaMemcpyAsync(hy.data(), dy.p, hy.size() * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int64_t t = 0; t < T; t++) memcpy(y + t * n_rows, hy.data() + t * rows_pad, (size_t) n_rows * 4);
        tmp.pb.rec = nullptr;
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" int b200_op_rms_norm(const float * x, const float * w, int64_t k, float eps, float * y) {
    try {
        require_gpu();
        DBuf dx((size_t) k * 4), dw((size_t) k * 4), dy((size_t) k * 4);
        CU(cudaMemcpy(dx.p, x, (size_t) k * 4, cudaMemcpyHostToDevice));
        if (w) CU(cudaMemcpy(dw.p, w, (size_t) k * 4, cudaMemcpyHostToDevice));
        k_rms_norm<<<1, 256>>>(dx.as<float>(), w ? dw.as<float>() : nullptr, (int) k, eps, dy.as<float>());
        CU(cudaGetLastError());
        CU(cudaMemcpy(y, dy.p, (size_t) k * 4, cudaMemcpyDeviceToHost));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" int b200_op_rope(float * x, int n_heads, int head_dim, int pos, float freq_base, float freq_scale,
                            const float * freq_factors) {
    try {
        require_gpu();
        b200_model m;
        m.head_dim = head_dim; m.rope_freq_base = freq_base; m.rope_freq_scale = freq_scale; m.n_ctx_orig = 4096;
        if (freq_factors) m.rope_freq_factors.assign(freq_factors, freq_factors + head_dim / 2);
        std::vector<float2> tab;
        build_rope_table(m, pos + 1, tab);
        DBuf dt((size_t) (head_dim / 2) * sizeof(float2)), dx((size_t) n_heads * head_dim * 4);
        CU(cudaMemcpy(dt.p, tab.data() + (size_t) pos * (head_dim / 2), (size_t) (head_dim / 2) * sizeof(float2), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dx.p, x, (size_t) n_heads * head_dim * 4, cudaMemcpyHostToDevice));
        const int n = n_heads * head_dim / 2;
        k_rope<<<(n + 127) / 128, 128>>>(dx.as<float>(), n_heads, head_dim, dt.as<float2>());
        CU(cudaGetLastError());
        CU(cudaMemcpy(x, dx.p, (size_t) n_heads * head_dim * 4, cudaMemcpyDeviceToHost));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}

extern "C" int b200_op_attention(const float * q, const uint16_t * k_cache, const uint16_t * v_cache, int n_kv,
                                 int n_head, int n_head_kv, int head_dim, float scale, int round_q, float * out) {
    try {
        require_gpu();
        if (n_kv <= 0) throw std::runtime_error("n_kv must be positive");
        if (!q || !k_cache || !v_cache || !out || n_head <= 0 || n_head_kv <= 0 || n_head % n_head_kv) throw std::runtime_error("bad arguments");
        const int kvd = n_head_kv * head_dim, qd = n_head * head_dim;
        OpStream os;
        cudaStream_t st = os.st;
        int dev = 0; CU(cudaGetDevice(&dev));
        cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, dev));
        b200_ctx tmp;
        tmp.st = st; tmp.sm_count = prop.multiProcessorCount;
        const int n_pad = (n_kv + 31) / 32 * 32;
        DBuf dq((size_t) qd * 4), dk((size_t) n_kv * kvd * 2), dv((size_t) n_kv * kvd * 2), dout((size_t) qd * 4);
        DBuf dS((size_t) n_head * n_pad * 4), dT((size_t) n_head_kv * 4);
        CU(cudaMemsetAsync(dT.p, 0, (size_t) n_head_kv * 4, st));
        CU(cudaMemcpyAsync(dq.p, q, (size_t) qd * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dk.p, k_cache, (size_t) n_kv * kvd * 2, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dv.p, v_cache, (size_t) n_kv * kvd * 2, cudaMemcpyHostToDevice, st));
        AttnArgs a{};
        a.q = dq.as<float>(); a.k_cache = dk.as<__half>(); a.v_cache = dv.as<__half>();
        a.S = dS.as<float>(); a.s_stride = n_pad; a.out = dout.as<float>();
        a.n_head = n_head; a.n_head_kv = n_head_kv; a.head_dim = head_dim; a.kv_dim = kvd;
        a.scale = scale; a.st = nullptr; a.n_kv_override = n_kv; a.round_q_override = round_q; a.cell_pos = nullptr;
        launch_attention(&tmp, a, n_pad);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, dout.p, (size_t) qd * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception & e) { return set_err(e.what()); }
}
