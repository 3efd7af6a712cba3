// gguf.hpp — minimal read-only GGUF v2/v3 container parser (mmap), written from the format definition:
//   header  : magic "GGUF", version u32, n_tensors u64, n_kv u64      (cpp/ggml/src/ggml.c:20767-20773)
//   kv pair : key string (u64 len + bytes), value type u32, value     (cpp/ggml/include/ggml.h:2257-2272)
//   tensor  : name, n_dims u32, ne u64[n_dims], type u32, offset u64  (cpp/ggml/src/ggml.c:20775-20788)
//   data    : aligned to general.alignment (default 32)                (cpp/ggml/src/ggml.c:21100-21104)
// Replaces gguf_init_from_file (cpp/ggml/src/ggml.c:20896-) for the keys/tensors the LLaMA path reads.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace b200 {

enum gguf_vtype : uint32_t {
    GV_U8 = 0, GV_I8, GV_U16, GV_I16, GV_U32, GV_I32, GV_F32, GV_BOOL, GV_STR, GV_ARR, GV_U64, GV_I64, GV_F64
};

struct gguf_value {
    uint32_t type = 0;
    uint64_t u = 0;          // any integer / bool
    double   f = 0;          // any float
    std::string s;           // string
    // arrays
    uint32_t arr_type = 0;
    uint64_t arr_n = 0;
    std::vector<std::string> arr_s;   // string arrays (tokenizer.ggml.tokens, merges)
    std::vector<double>      arr_f;   // numeric arrays as double
};

struct gguf_tensor {
    std::string name;
    uint32_t n_dims = 0;
    uint64_t ne[4] = {1, 1, 1, 1};
    uint32_t type = 0;
    uint64_t offset = 0;      // relative to data section
    const uint8_t * data = nullptr;
    uint64_t nbytes = 0;
};

struct gguf_file {
    int fd = -1;
    const uint8_t * base = nullptr;
    uint64_t size = 0;
    uint32_t version = 0;
    uint64_t alignment = 32;
    uint64_t data_off = 0;
    std::map<std::string, gguf_value>  kv;
    std::map<std::string, gguf_tensor> tensors;
    std::vector<std::string> tensor_order;

    ~gguf_file();
    // returns empty string on success, else an error message
    std::string open(const std::string & path);
    std::string open_impl(const std::string & path);

    bool has(const std::string & k) const { return kv.count(k) > 0; }
    uint64_t    get_u(const std::string & k, uint64_t def) const;
    double      get_f(const std::string & k, double def) const;
    std::string get_s(const std::string & k, const std::string & def) const;
    const gguf_tensor * find(const std::string & name) const;
};

// bytes of one row of `k` elements in ggml block layout; 0 if the type is not one we handle
uint64_t ggml_row_bytes(uint32_t type, uint64_t k);
// elements per block of a type this path handles (1 for F32/F16/BF16); 0 for any other type
uint64_t ggml_block_elems(uint32_t type);

}  // namespace b200
