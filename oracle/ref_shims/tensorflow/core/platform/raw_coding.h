// Shim (test infrastructure): little-endian fixed32 decode used by uniq_hashtable.h's hash.
#pragma once
#include <cstdint>
#include <cstring>
namespace tensorflow { namespace core {
inline uint32_t DecodeFixed32(const char* ptr) { uint32_t r; std::memcpy(&r, ptr, sizeof(r)); return r; }
inline uint64_t DecodeFixed64(const char* ptr) { uint64_t r; std::memcpy(&r, ptr, sizeof(r)); return r; }
} }
