// gguf.cpp — see gguf.hpp. Host-only, no CUDA.
#include "gguf.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace b200 {

namespace {

struct cursor {
    const uint8_t * p;
    const uint8_t * end;
    bool ok = true;
    template <typename T> T rd() {
        T v{};
        if (!ok || sizeof(T) > (size_t) (end - p)) { ok = false; return v; }
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        const uint64_t n = rd<uint64_t>();
        if (!ok || n > (uint64_t) (end - p)) { ok = false; return {}; }     // (no pointer arithmetic on an untrusted 64-bit length)
        std::string s((const char *) p, (size_t) n);
        p += n;
        return s;
    }
};

bool read_scalar(cursor & c, uint32_t t, gguf_value & v) {
    switch (t) {
        case GV_U8:   v.u = c.rd<uint8_t>();            v.f = (double) v.u; break;
        case GV_I8:   { int8_t  x = c.rd<int8_t>();   v.u = (uint64_t)(int64_t) x; v.f = x; } break;
        case GV_U16:  v.u = c.rd<uint16_t>();           v.f = (double) v.u; break;
        case GV_I16:  { int16_t x = c.rd<int16_t>();  v.u = (uint64_t)(int64_t) x; v.f = x; } break;
        case GV_U32:  v.u = c.rd<uint32_t>();           v.f = (double) v.u; break;
        case GV_I32:  { int32_t x = c.rd<int32_t>();  v.u = (uint64_t)(int64_t) x; v.f = x; } break;
        case GV_F32:  { float   x = c.rd<float>();    v.f = x; v.u = (uint64_t) x; } break;
        case GV_BOOL: v.u = c.rd<uint8_t>() != 0;       v.f = (double) v.u; break;
        case GV_U64:  v.u = c.rd<uint64_t>();           v.f = (double) v.u; break;
        case GV_I64:  { int64_t x = c.rd<int64_t>();  v.u = (uint64_t) x; v.f = (double) x; } break;
        case GV_F64:  { double  x = c.rd<double>();   v.f = x; v.u = (uint64_t) x; } break;
        case GV_STR:  v.s = c.str(); break;
        default: return false;
    }
    return c.ok;
}

}  // namespace

uint64_t ggml_row_bytes(uint32_t type, uint64_t k) {
    switch (type) {
        case 0:  return k * 4;                 // F32
        case 1:  return k * 2;                 // F16
        case 30: return k * 2;                 // BF16
        case 8:  return (k / 32) * 34;         // Q8_0   (cpp/ggml/src/ggml-common.h:186-190)
        case 12: return (k / 256) * 144;       // Q4_K   (:267-277)
        case 13: return (k / 256) * 176;       // Q5_K   (:284-295)
        case 14: return (k / 256) * 210;       // Q6_K   (:302-307)
        default: return 0;
    }
}

uint64_t ggml_block_elems(uint32_t type) {
    switch (type) {
        case 0: case 1: case 30: return 1;
        case 8: return 32;
        case 12: case 13: case 14: return 256;
        default: return 0;
    }
}

gguf_file::~gguf_file() {
    if (base) munmap((void *) base, size);
    if (fd >= 0) close(fd);
}

// a damaged file is an error string, never an exception or a crash (the callers sit behind extern "C" entry points)
std::string gguf_file::open(const std::string & path) {
    try {
        return open_impl(path);
    } catch (const std::exception & e) {
        return std::string("corrupt GGUF (") + e.what() + "): " + path;
    }
}

std::string gguf_file::open_impl(const std::string & path) {
    fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return "cannot open " + path;
    struct stat st;
    if (fstat(fd, &st) != 0) return "fstat failed: " + path;
    size = (uint64_t) st.st_size;
    void * m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) return "mmap failed: " + path;
    base = (const uint8_t *) m;

    cursor c{base, base + size};
    if (size < 24 || std::memcmp(base, "GGUF", 4) != 0) return "not a GGUF file: " + path;
    c.p += 4;
    version = c.rd<uint32_t>();
    if (version < 2 || version > 3) return "unsupported GGUF version " + std::to_string(version);
    const uint64_t n_tensors = c.rd<uint64_t>();
    const uint64_t n_kv      = c.rd<uint64_t>();
    // every kv pair / tensor record / array element occupies at least one byte of the file: counts beyond the
    // remaining bytes are corrupt (and must never reach reserve())
    if (!c.ok || n_kv > size || n_tensors > size) return "corrupt GGUF header (counts exceed the file size)";

    for (uint64_t i = 0; i < n_kv && c.ok; i++) {
        std::string key = c.str();
        gguf_value v;
        v.type = c.rd<uint32_t>();
        if (v.type == GV_ARR) {
            v.arr_type = c.rd<uint32_t>();
            v.arr_n    = c.rd<uint64_t>();
            if (!c.ok || v.arr_n > (uint64_t) (c.end - c.p)) return "corrupt GGUF array length in key " + key;
            if (v.arr_type == GV_STR) {
                v.arr_s.reserve((size_t) std::min<uint64_t>(v.arr_n, 1u << 20));   // (grows on demand: arr_n is untrusted)
                for (uint64_t j = 0; j < v.arr_n && c.ok; j++) v.arr_s.push_back(c.str());
            } else {
                v.arr_f.reserve((size_t) std::min<uint64_t>(v.arr_n, 1u << 20));
                for (uint64_t j = 0; j < v.arr_n && c.ok; j++) {
                    gguf_value e;
                    if (!read_scalar(c, v.arr_type, e)) return "bad array element type in key " + key;
                    v.arr_f.push_back(v.arr_type == GV_F32 || v.arr_type == GV_F64 ? e.f : (double)(int64_t) e.u);
                }
            }
        } else if (!read_scalar(c, v.type, v)) {
            return "bad value type for key " + key;
        }
        kv[key] = std::move(v);
    }
    if (!c.ok) return "truncated GGUF metadata";

    for (uint64_t i = 0; i < n_tensors && c.ok; i++) {
        gguf_tensor t;
        t.name   = c.str();
        t.n_dims = c.rd<uint32_t>();
        if (t.n_dims > 4) return "tensor " + t.name + ": n_dims > 4";
        for (uint32_t d = 0; d < t.n_dims; d++) t.ne[d] = c.rd<uint64_t>();
        t.type   = c.rd<uint32_t>();
        t.offset = c.rd<uint64_t>();
        if (!c.ok) break;
        tensor_order.push_back(t.name);
        tensors[t.name] = t;
    }
    if (!c.ok) return "truncated GGUF tensor table";

    alignment = get_u("general.alignment", 32);
    if (alignment == 0 || (alignment & (alignment - 1)) != 0 || alignment > (1u << 20))
        return "bad general.alignment " + std::to_string(alignment) + " (must be a power of two)";
    const uint64_t meta = (uint64_t) (c.p - base);
    data_off = (meta + alignment - 1) / alignment * alignment;
    if (data_off > size) return "truncated GGUF (no data section)";
    const uint64_t data_size = size - data_off;

    for (auto & it : tensors) {
        gguf_tensor & t = it.second;
        const uint64_t blk = ggml_block_elems(t.type);
        if (blk == 0) { t.nbytes = 0; t.data = nullptr; continue; }      // a type this path never reads: no data pointer
        if (t.ne[0] % blk != 0) return "tensor " + t.name + ": row length is not a multiple of its block size";
        // checked arithmetic: ne[] and offset are untrusted 64-bit values
        uint64_t nb = 0;
        if (__builtin_mul_overflow(t.ne[0] / blk, ggml_row_bytes(t.type, blk), &nb)) return "tensor " + t.name + ": size overflow";
        for (int d = 1; d < 4; d++)
            if (__builtin_mul_overflow(nb, t.ne[d], &nb)) return "tensor " + t.name + ": size overflow";
        t.nbytes = nb;
        if (t.offset > data_size || nb > data_size - t.offset) return "tensor " + t.name + " exceeds file size";
        t.data = base + data_off + t.offset;
    }
    return {};
}

uint64_t gguf_file::get_u(const std::string & k, uint64_t def) const {
    auto it = kv.find(k);
    if (it == kv.end() || it->second.type == GV_STR || it->second.type == GV_ARR) return def;
    if (it->second.type == GV_F32 || it->second.type == GV_F64) return (uint64_t) it->second.f;
    return it->second.u;
}
double gguf_file::get_f(const std::string & k, double def) const {
    auto it = kv.find(k);
    if (it == kv.end() || it->second.type == GV_STR || it->second.type == GV_ARR) return def;
    return it->second.f;
}
std::string gguf_file::get_s(const std::string & k, const std::string & def) const {
    auto it = kv.find(k);
    if (it == kv.end() || it->second.type != GV_STR) return def;
    return it->second.s;
}
const gguf_tensor * gguf_file::find(const std::string & name) const {
    auto it = tensors.find(name);
    return it == tensors.end() ? nullptr : &it->second;
}

}  // namespace b200
