// gguf.cpp — see gguf.hpp. Host-only, no CUDA.
#include "gguf.hpp"

#include <cstring>
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
        if (p + sizeof(T) > end) { ok = false; return v; }
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        const uint64_t n = rd<uint64_t>();
        if (!ok || p + n > end) { ok = false; return {}; }
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

gguf_file::~gguf_file() {
    if (base) munmap((void *) base, size);
    if (fd >= 0) close(fd);
}

std::string gguf_file::open(const std::string & path) {
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

    for (uint64_t i = 0; i < n_kv && c.ok; i++) {
        std::string key = c.str();
        gguf_value v;
        v.type = c.rd<uint32_t>();
        if (v.type == GV_ARR) {
            v.arr_type = c.rd<uint32_t>();
            v.arr_n    = c.rd<uint64_t>();
            if (v.arr_type == GV_STR) {
                v.arr_s.reserve((size_t) v.arr_n);
                for (uint64_t j = 0; j < v.arr_n && c.ok; j++) v.arr_s.push_back(c.str());
            } else {
                v.arr_f.reserve((size_t) v.arr_n);
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
        tensor_order.push_back(t.name);
        tensors[t.name] = t;
    }
    if (!c.ok) return "truncated GGUF tensor table";

    alignment = get_u("general.alignment", 32);
    const uint64_t meta = (uint64_t) (c.p - base);
    data_off = (meta + alignment - 1) / alignment * alignment;

    for (auto & it : tensors) {
        gguf_tensor & t = it.second;
        const uint64_t rb = ggml_row_bytes(t.type, t.ne[0]);
        t.nbytes = rb * t.ne[1] * t.ne[2] * t.ne[3];
        if (data_off + t.offset + t.nbytes > size && rb != 0) return "tensor " + t.name + " exceeds file size";
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
