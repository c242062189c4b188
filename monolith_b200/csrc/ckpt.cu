// ckpt.cu — checkpoint files in the reference's on-disk format (host-only code).
//
// What MonolithMultiHashTableSave writes (RT/ops/multi_hash_table_save_restore_ops.cc:186-250):
//   <basename>-%05d-of-%05d       TFRecord stream, Snappy block compression (RecordWriterOptions
//                                 SNAPPY_COMPRESSION, :201-202): for every table, in table order, one record
//                                 per live entry = serialized EntryDump (embedding_hash_table.proto:45-50):
//                                   1 sfixed64 id | 2 repeated float num | 3 OptimizerDump opt | 4 int64 ts
//                                 OptimizerDump = repeated SingleOptimizerDump, one per segment, in segment
//                                 order (optimizer_combination.cc:73-83); oneof adagrad=1 {norm} sgd=2 {}
//                                 ftrl=3 {zero=1, norm=2} adam=7 {m, v, beta1_power, beta2_power}
//                                 (optimizer.proto:28-30,56-57,69-72,130-135; proto2: repeated floats unpacked)
//   <basename>.meta-%05d-of-%05d  TFRecord stream, uncompressed: one MultiHashTableMetadata {table_name,
//                                 num_entries} per table (embedding_hash_table.proto:139-142); restore reads
//                                 num_entries records for each table name (:349-388).
// TFRecord framing (tensorflow/core/lib/io/record_writer.cc): u64 length | u32 masked crc32c(length) | data |
// u32 masked crc32c(data), little endian, mask = rotr(crc, 15) + 0xa282ead8.
// Snappy block container (tensorflow/core/lib/io/snappy/snappy_outputbuffer.cc, TF 2.4 as pinned by the
// reference's WORKSPACE): per flushed 256 KiB input buffer ONE u32 big-endian COMPRESSED length, then the raw
// Snappy block (SnappyOutputBuffer::Deflate writes only that length; SnappyInputBuffer reads it and takes the
// uncompressed size from the Snappy preamble).  TensorFlow is not available in this image, so the container
// layout is restated from the TF source and NOT pinned by a TF-written fixture ("parity unpinned" for the
// container — round 1 wrote an extra uncompressed-length word, corrected in round 2; record framing,
// crc32c, Snappy codec and the protobuf wire bytes are pinned in tests/test_checkpoint_cpu.py against
// known-answer vectors, pyarrow's Snappy and the protobuf runtime).
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "engine.h"

namespace mono {
namespace {

// ---- crc32c (Castagnoli), slicing-by-8 ------------------------------------------------------------
struct Crc32cTable {
  uint32_t t[8][256];
  Crc32cTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
const Crc32cTable& crc_table() {
  static const Crc32cTable tab;
  return tab;
}
uint32_t crc32c(const void* data, size_t n) {
  const auto& T = crc_table().t;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo, hi;
    std::memcpy(&lo, p, 4);
    std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T[7][lo & 0xff] ^ T[6][(lo >> 8) & 0xff] ^ T[5][(lo >> 16) & 0xff] ^ T[4][lo >> 24] ^
        T[3][hi & 0xff] ^ T[2][(hi >> 8) & 0xff] ^ T[1][(hi >> 16) & 0xff] ^ T[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ T[0][(c ^ *p++) & 0xff];
  return c ^ 0xFFFFFFFFu;
}
uint32_t masked_crc(const void* data, size_t n) {
  const uint32_t c = crc32c(data, n);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

// ---- raw Snappy codec --------------------------------------------------------------------------------
void put_varint(std::string* out, uint64_t v) {
  while (v >= 0x80) {
    out->push_back((char)(v | 0x80));
    v >>= 7;
  }
  out->push_back((char)v);
}
void emit_literal(std::string* out, const uint8_t* p, size_t n) {
  while (n > 0) {
    const size_t len = n;  // one element holds up to 2^32 bytes
    const size_t l1 = len - 1;
    if (l1 < 60) {
      out->push_back((char)(l1 << 2));
    } else if (l1 < (1u << 8)) {
      out->push_back((char)(60 << 2));
      out->push_back((char)l1);
    } else if (l1 < (1u << 16)) {
      out->push_back((char)(61 << 2));
      out->push_back((char)(l1 & 0xff));
      out->push_back((char)(l1 >> 8));
    } else if (l1 < (1u << 24)) {
      out->push_back((char)(62 << 2));
      out->push_back((char)(l1 & 0xff));
      out->push_back((char)((l1 >> 8) & 0xff));
      out->push_back((char)(l1 >> 16));
    } else {
      out->push_back((char)(63 << 2));
      for (int k = 0; k < 4; ++k) out->push_back((char)((l1 >> (8 * k)) & 0xff));
    }
    out->append(reinterpret_cast<const char*>(p), len);
    p += len;
    n -= len;
  }
}
void emit_copy(std::string* out, size_t offset, size_t len) {  // offset < 65536
  while (len > 0) {
    size_t l = len > 64 ? 64 : len;
    if (len > 64 && len - 64 < 4) l = 60;  // never leave a tail shorter than a legal copy
    if (l >= 4 && l <= 11 && offset < 2048) {
      out->push_back((char)(1 | ((l - 4) << 2) | ((offset >> 8) << 5)));
      out->push_back((char)(offset & 0xff));
    } else {
      out->push_back((char)(2 | ((l - 1) << 2)));
      out->push_back((char)(offset & 0xff));
      out->push_back((char)(offset >> 8));
    }
    len -= l;
  }
}
// greedy 4-byte hash matcher, offsets below 64 KiB (always a valid Snappy stream)
void snappy_compress(const uint8_t* in, size_t n, std::string* out) {
  put_varint(out, n);
  if (n == 0) return;
  constexpr int kHashBits = 14;
  std::vector<int64_t> table((size_t)1 << kHashBits, -1);
  auto hash = [&](const uint8_t* p) {
    uint32_t v;
    std::memcpy(&v, p, 4);
    return (v * 0x1e35a7bdu) >> (32 - kHashBits);
  };
  size_t lit = 0, i = 0;
  while (i + 4 <= n) {
    const uint32_t h = hash(in + i);
    const int64_t cand = table[h];
    table[h] = (int64_t)i;
    if (cand >= 0 && i - (size_t)cand < 65536 && std::memcmp(in + cand, in + i, 4) == 0) {
      size_t len = 4;
      while (i + len < n && in[cand + len] == in[i + len]) ++len;
      if (i > lit) emit_literal(out, in + lit, i - lit);
      emit_copy(out, i - (size_t)cand, len);
      i += len;
      lit = i;
    } else {
      ++i;
    }
  }
  if (n > lit) emit_literal(out, in + lit, n - lit);
}
bool snappy_uncompress(const uint8_t* in, size_t n, std::string* out) {
  size_t p = 0;
  uint64_t ulen = 0;
  int shift = 0;
  while (true) {
    if (p >= n || shift > 35) return false;
    const uint8_t b = in[p++];
    ulen |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  // a copy element expands at most 3 bytes -> 64: a declared length beyond that is corrupt (and must not be
  // allowed to drive the allocation below)
  if (ulen > (uint64_t)n * 32 + 64) return false;
  const size_t base = out->size();
  out->reserve(base + ulen);
  while (p < n) {
    const uint8_t tag = in[p++];
    size_t len, offset;
    switch (tag & 3) {
      case 0: {
        len = (tag >> 2) + 1;
        if (len > 60) {
          const size_t nb = len - 60;
          if (p + nb > n) return false;
          len = 0;
          for (size_t k = 0; k < nb; ++k) len |= (size_t)in[p + k] << (8 * k);
          len += 1;
          p += nb;
        }
        if (p + len > n) return false;
        out->append(reinterpret_cast<const char*>(in + p), len);
        p += len;
        continue;
      }
      case 1:
        if (p + 1 > n) return false;
        len = ((tag >> 2) & 7) + 4;
        offset = ((size_t)(tag >> 5) << 8) | in[p];
        p += 1;
        break;
      case 2:
        if (p + 2 > n) return false;
        len = (tag >> 2) + 1;
        offset = in[p] | ((size_t)in[p + 1] << 8);
        p += 2;
        break;
      default:
        if (p + 4 > n) return false;
        len = (tag >> 2) + 1;
        offset = in[p] | ((size_t)in[p + 1] << 8) | ((size_t)in[p + 2] << 16) | ((size_t)in[p + 3] << 24);
        p += 4;
        break;
    }
    const size_t cur = out->size() - base;
    if (offset == 0 || offset > cur) return false;
    for (size_t k = 0; k < len; ++k) out->push_back((*out)[out->size() - offset]);  // may overlap
  }
  return out->size() - base == ulen;
}

// ---- protobuf wire helpers ---------------------------------------------------------------------------
void put_tag_len(std::string* out, uint32_t field, const std::string& body) {
  put_varint(out, (field << 3) | 2);
  put_varint(out, body.size());
  out->append(body);
}
void put_floats_unpacked(std::string* out, uint32_t field, const float* v, int n) {
  for (int i = 0; i < n; ++i) {
    put_varint(out, (field << 3) | 5);
    out->append(reinterpret_cast<const char*>(v + i), 4);
  }
}
void put_float(std::string* out, uint32_t field, float v) { put_floats_unpacked(out, field, &v, 1); }

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (true) {
      if (p >= end || shift > 63) {
        ok = false;
        return 0;
      }
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
  }
  bool bytes(size_t n, const uint8_t** out) {
    if ((size_t)(end - p) < n) {
      ok = false;
      return false;
    }
    *out = p;
    p += n;
    return true;
  }
  void skip(uint32_t wt) {
    const uint8_t* d;
    switch (wt) {
      case 0: varint(); break;
      case 1: bytes(8, &d); break;
      case 2: {
        const uint64_t n = varint();
        if (ok) bytes(n, &d);
        break;
      }
      case 5: bytes(4, &d); break;
      default: ok = false;
    }
  }
};
// repeated float (unpacked wire type 5 or packed wire type 2) appended to dst up to cap values
void read_floats(Reader& r, uint32_t wt, float* dst, int cap, int* n) {
  const uint8_t* d;
  if (wt == 5) {
    if (r.bytes(4, &d) && *n < cap) std::memcpy(dst + (*n)++, d, 4);
  } else if (wt == 2) {
    const uint64_t len = r.varint();
    if (r.ok && r.bytes(len, &d))
      for (uint64_t k = 0; k + 4 <= len && *n < cap; k += 4) std::memcpy(dst + (*n)++, d + k, 4);
  } else {
    r.ok = false;
  }
}

int state_floats_of(const mono_segment_cfg& s) {
  switch (s.opt_type) {
    case MONO_OPT_ADAGRAD: return s.dim;
    case MONO_OPT_FTRL: return 2 * s.dim;
    case MONO_OPT_ADAM: return 2 * s.dim + 2;
    case MONO_OPT_MOMENTUM: case MONO_OPT_RMSPROP: case MONO_OPT_RMSPROPV2: return s.dim;
    case MONO_OPT_ADADELTA: return 2 * s.dim;
    case MONO_OPT_AMSGRAD: return 3 * s.dim + 2;
    case MONO_OPT_GROUP_ADAGRAD: return 1;
    default: return 0;
  }
}

// one SingleOptimizerDump for segment s; st = the segment's state block in the engine's order
// (Adagrad: norm; FTRL: norm | zero; Adam: m | v | beta1_power | beta2_power — the reference's ctx order,
// ftrl_optimizer.cc:78-88, adam_optimizer.cc:86-101)
std::string encode_single_opt(const mono_segment_cfg& s, const float* st) {
  std::string body, msg;
  const int D = s.dim;
  switch (s.opt_type) {
    case MONO_OPT_ADAGRAD:
      put_floats_unpacked(&body, 1, st, D);
      put_tag_len(&msg, 1, body);
      break;
    case MONO_OPT_FTRL:
      put_floats_unpacked(&body, 1, st + D, D);  // zero = 1
      put_floats_unpacked(&body, 2, st, D);      // norm = 2
      put_tag_len(&msg, 3, body);
      break;
    case MONO_OPT_ADAM:
      put_floats_unpacked(&body, 1, st, D);
      put_floats_unpacked(&body, 2, st + D, D);
      put_float(&body, 3, st[2 * D]);
      put_float(&body, 4, st[2 * D + 1]);
      put_tag_len(&msg, 7, body);
      break;
    case MONO_OPT_MOMENTUM:  // MomentumOptimizerDump { n = 1 } = field 9 (optimizer.proto:232-252)
      put_floats_unpacked(&body, 1, st, D);
      put_tag_len(&msg, 9, body);
      break;
    case MONO_OPT_RMSPROP:   // RmspropOptimizerDump { n = 1 } = field 11
      put_floats_unpacked(&body, 1, st, D);
      put_tag_len(&msg, 11, body);
      break;
    case MONO_OPT_RMSPROPV2:  // RmspropV2OptimizerDump { n = 1 } = field 12
      put_floats_unpacked(&body, 1, st, D);
      put_tag_len(&msg, 12, body);
      break;
    case MONO_OPT_ADADELTA:  // AdadeltaOptimizerDump { accum = 1, accum_update = 2 } = field 6
      put_floats_unpacked(&body, 1, st, D);
      put_floats_unpacked(&body, 2, st + D, D);
      put_tag_len(&msg, 6, body);
      break;
    case MONO_OPT_AMSGRAD:  // AmsgradOptimizerDump { m = 1, v = 2, vhat = 3, beta1_power = 4, beta2_power = 5 } = field 8
      put_floats_unpacked(&body, 1, st, D);
      put_floats_unpacked(&body, 2, st + D, D);
      put_floats_unpacked(&body, 3, st + 2 * D, D);
      put_float(&body, 4, st[3 * D]);
      put_float(&body, 5, st[3 * D + 1]);
      put_tag_len(&msg, 8, body);
      break;
    case MONO_OPT_GROUP_ADAGRAD:  // GroupAdaGradOptimizerDump { grad_square_sum = 1 } = field 15
      put_float(&body, 1, st[0]);
      put_tag_len(&msg, 15, body);
      break;
    default:  // SGD: empty message, still present (sgd_optimizer.cc:51-55)
      put_tag_len(&msg, 2, body);
      break;
  }
  return msg;
}

bool decode_single_opt(const uint8_t* p, size_t n, const mono_segment_cfg& s, float* st) {
  Reader r{p, p + n};
  const int D = s.dim;
  while (r.ok && r.p < r.end) {
    const uint64_t tag = r.varint();
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    const bool mine = (field == 1 && s.opt_type == MONO_OPT_ADAGRAD) || (field == 3 && s.opt_type == MONO_OPT_FTRL) ||
                      (field == 7 && s.opt_type == MONO_OPT_ADAM) || (field == 9 && s.opt_type == MONO_OPT_MOMENTUM) ||
                      (field == 11 && s.opt_type == MONO_OPT_RMSPROP) || (field == 12 && s.opt_type == MONO_OPT_RMSPROPV2) ||
                      (field == 6 && s.opt_type == MONO_OPT_ADADELTA) || (field == 8 && s.opt_type == MONO_OPT_AMSGRAD) ||
                      (field == 15 && s.opt_type == MONO_OPT_GROUP_ADAGRAD);
    if (!mine || wt != 2) {
      r.skip(wt);
      continue;
    }
    const uint64_t len = r.varint();
    const uint8_t* d;
    if (!r.ok || !r.bytes(len, &d)) return false;
    Reader b{d, d + len};
    int n1 = 0, n2 = 0, n3 = 0;
    const bool one_vec = s.opt_type == MONO_OPT_MOMENTUM || s.opt_type == MONO_OPT_RMSPROP || s.opt_type == MONO_OPT_RMSPROPV2;
    while (b.ok && b.p < b.end) {
      const uint64_t t2 = b.varint();
      const uint32_t f2 = (uint32_t)(t2 >> 3), w2 = (uint32_t)(t2 & 7);
      if (one_vec && f2 == 1) read_floats(b, w2, st, D, &n1);
      else if (s.opt_type == MONO_OPT_ADADELTA && f2 == 1) read_floats(b, w2, st, D, &n1);
      else if (s.opt_type == MONO_OPT_ADADELTA && f2 == 2) read_floats(b, w2, st + D, D, &n2);
      else if (s.opt_type == MONO_OPT_AMSGRAD && f2 == 1) read_floats(b, w2, st, D, &n1);
      else if (s.opt_type == MONO_OPT_AMSGRAD && f2 == 2) read_floats(b, w2, st + D, D, &n2);
      else if (s.opt_type == MONO_OPT_AMSGRAD && f2 == 3) read_floats(b, w2, st + 2 * D, D, &n3);
      else if (s.opt_type == MONO_OPT_AMSGRAD && (f2 == 4 || f2 == 5) && w2 == 5) {
        const uint8_t* q;
        if (b.bytes(4, &q)) std::memcpy(st + 3 * D + (f2 - 4), q, 4);
      } else if (s.opt_type == MONO_OPT_GROUP_ADAGRAD && f2 == 1 && w2 == 5) {
        const uint8_t* q;
        if (b.bytes(4, &q)) std::memcpy(st, q, 4);
      } else if (s.opt_type == MONO_OPT_ADAGRAD && f2 == 1) read_floats(b, w2, st, D, &n1);
      else if (s.opt_type == MONO_OPT_FTRL && f2 == 1) read_floats(b, w2, st + D, D, &n1);
      else if (s.opt_type == MONO_OPT_FTRL && f2 == 2) read_floats(b, w2, st, D, &n2);
      else if (s.opt_type == MONO_OPT_ADAM && f2 == 1) read_floats(b, w2, st, D, &n1);
      else if (s.opt_type == MONO_OPT_ADAM && f2 == 2) read_floats(b, w2, st + D, D, &n2);
      else if (s.opt_type == MONO_OPT_ADAM && (f2 == 3 || f2 == 4) && w2 == 5) {
        const uint8_t* q;
        if (b.bytes(4, &q)) std::memcpy(st + 2 * D + (f2 - 3), q, 4);
      } else b.skip(w2);
    }
    if (!b.ok) return false;
  }
  return r.ok;
}

struct File {
  FILE* f = nullptr;
  ~File() {
    if (f) std::fclose(f);
  }
};

}  // namespace
}  // namespace mono

using namespace mono;

// ---- writer ---------------------------------------------------------------------------------------------
struct mono_ckpt_writer {
  std::string data_path, meta_path, data_tmp, meta_tmp;
  File data, meta;
  bool snappy = true;
  std::string block;  // uncompressed TFRecord bytes waiting for the next Snappy block
  std::vector<mono_segment_cfg> segs;
  std::string table_name;
  int dim = 0, state = 0;
  uint64_t table_entries = 0, total_entries = 0;
  bool in_table = false;
};
static constexpr size_t kSnappyInputBuffer = 256 * 1024;  // TF RecordWriter's compression input buffer

static void write_all(FILE* f, const void* p, size_t n) {
  if (n && std::fwrite(p, 1, n, f) != n) throw std::runtime_error("checkpoint: short write");
}
static void frame_record(std::string* out, const std::string& rec) {
  const uint64_t len = rec.size();
  char hdr[12];
  std::memcpy(hdr, &len, 8);
  const uint32_t c1 = masked_crc(hdr, 8);
  std::memcpy(hdr + 8, &c1, 4);
  out->append(hdr, 12);
  out->append(rec);
  const uint32_t c2 = masked_crc(rec.data(), rec.size());
  out->append(reinterpret_cast<const char*>(&c2), 4);
}
static void be32(char* p, uint32_t v) {
  p[0] = (char)(v >> 24); p[1] = (char)(v >> 16); p[2] = (char)(v >> 8); p[3] = (char)v;
}
static void flush_block(mono_ckpt_writer* w, bool all) {
  while (w->block.size() >= kSnappyInputBuffer || (all && !w->block.empty())) {
    const size_t n = std::min(w->block.size(), kSnappyInputBuffer);
    if (w->snappy) {
      std::string comp;
      snappy_compress(reinterpret_cast<const uint8_t*>(w->block.data()), n, &comp);
      char h[4];  // TF's SnappyOutputBuffer::Deflate: 4-byte big-endian COMPRESSED length, then the raw block
      be32(h, (uint32_t)comp.size());
      write_all(w->data.f, h, 4);
      write_all(w->data.f, comp.data(), comp.size());
    } else {
      write_all(w->data.f, w->block.data(), n);
    }
    w->block.erase(0, n);
  }
}

// ---- reader ---------------------------------------------------------------------------------------------
struct mono_ckpt_reader {
  File data, meta;
  bool snappy = true;
  std::string buf;   // decompressed, not yet consumed
  size_t pos = 0;
  std::vector<mono_segment_cfg> segs;
  int dim = 0, state = 0;
  uint64_t remaining = 0;  // records left in the current table
};
static bool fill(mono_ckpt_reader* r, size_t need) {  // make buf[pos, pos+need) available
  while (r->buf.size() - r->pos < need) {
    if (r->pos > (1u << 20)) {
      r->buf.erase(0, r->pos);
      r->pos = 0;
    }
    if (r->snappy) {
      unsigned char h[4];
      const size_t got = std::fread(h, 1, 4, r->data.f);
      if (got == 0) return false;
      if (got != 4) throw std::runtime_error("checkpoint: truncated Snappy block header");
      // TF's SnappyInputBuffer::ReadCompressedBlockLength: one big-endian length; the uncompressed size comes
      // from the Snappy preamble (Snappy_GetUncompressedLength), which snappy_uncompress validates
      const uint32_t clen = ((uint32_t)h[0] << 24) | (h[1] << 16) | (h[2] << 8) | h[3];
      if (clen > (64u << 20)) throw std::runtime_error("checkpoint: implausible Snappy block length");
      std::string comp(clen, '\0');
      if (clen && std::fread(&comp[0], 1, clen, r->data.f) != clen)
        throw std::runtime_error("checkpoint: truncated Snappy block");
      if (!snappy_uncompress(reinterpret_cast<const uint8_t*>(comp.data()), clen, &r->buf))
        throw std::runtime_error("checkpoint: corrupt Snappy block");
    } else {
      char tmp[1 << 16];
      const size_t got = std::fread(tmp, 1, sizeof(tmp), r->data.f);
      if (got == 0) return false;
      r->buf.append(tmp, got);
    }
  }
  return true;
}
static bool read_record_stream(mono_ckpt_reader* r, std::string* rec) {
  if (!fill(r, 12)) return false;
  uint64_t len;
  uint32_t c1;
  std::memcpy(&len, r->buf.data() + r->pos, 8);
  std::memcpy(&c1, r->buf.data() + r->pos + 8, 4);
  if (masked_crc(r->buf.data() + r->pos, 8) != c1) throw std::runtime_error("checkpoint: corrupt record length");
  if (!fill(r, 12 + len + 4)) throw std::runtime_error("checkpoint: truncated record");
  rec->assign(r->buf.data() + r->pos + 12, len);
  uint32_t c2;
  std::memcpy(&c2, r->buf.data() + r->pos + 12 + len, 4);
  if (masked_crc(rec->data(), rec->size()) != c2) throw std::runtime_error("checkpoint: corrupt record data");
  r->pos += 12 + len + 4;
  return true;
}
static bool read_record_file(FILE* f, std::string* rec) {  // uncompressed TFRecord straight from a file
  char hdr[12];
  const size_t got = std::fread(hdr, 1, 12, f);
  if (got == 0) return false;
  if (got != 12) throw std::runtime_error("checkpoint: truncated metadata record");
  uint64_t len;
  uint32_t c1;
  std::memcpy(&len, hdr, 8);
  std::memcpy(&c1, hdr + 8, 4);
  if (masked_crc(hdr, 8) != c1 || len > (1u << 20)) throw std::runtime_error("checkpoint: corrupt metadata record");
  rec->assign(len, '\0');
  uint32_t c2;
  if ((len && std::fread(&(*rec)[0], 1, len, f) != len) || std::fread(&c2, 1, 4, f) != 4 ||
      masked_crc(rec->data(), rec->size()) != c2)
    throw std::runtime_error("checkpoint: corrupt metadata record");
  return true;
}

namespace {
thread_local std::string g_ckpt_error;
template <class F>
int ck_guard(F f) {
  try {
    f();
    return MONO_OK;
  } catch (const ArgError& e) {
    g_ckpt_error = e.what();
    return MONO_ERR_INVALID_ARGUMENT;
  } catch (const std::exception& e) {
    g_ckpt_error = e.what();
    return MONO_ERR_INTERNAL;
  }
}
void set_segs(std::vector<mono_segment_cfg>* dst, int* dim, int* state, const mono_segment_cfg* segs, int n) {
  if (!segs || n <= 0) throw ArgError("checkpoint: table needs at least one segment");
  dst->assign(segs, segs + n);
  *dim = *state = 0;
  for (int i = 0; i < n; ++i) {
    if (segs[i].dim <= 0) throw ArgError("checkpoint: bad segment dim");
    *dim += segs[i].dim;
    *state += state_floats_of(segs[i]);
  }
}
}  // namespace

extern "C" {

const char* mono_ckpt_last_error(void) { return g_ckpt_error.c_str(); }

uint32_t mono_ckpt_crc32c(const void* data, int64_t n) { return crc32c(data, (size_t)n); }
uint32_t mono_ckpt_masked_crc32c(const void* data, int64_t n) { return masked_crc(data, (size_t)n); }

int64_t mono_ckpt_snappy_compress(const void* in, int64_t n, void* out, int64_t out_cap) {
  try {
    std::string s;
    snappy_compress(static_cast<const uint8_t*>(in), (size_t)n, &s);
    if ((int64_t)s.size() > out_cap) return -(int64_t)s.size();
    std::memcpy(out, s.data(), s.size());
    return (int64_t)s.size();
  } catch (const std::exception&) {
    return INT64_MIN;
  }
}
int64_t mono_ckpt_snappy_uncompress(const void* in, int64_t n, void* out, int64_t out_cap) {
  try {
    std::string s;
    if (!snappy_uncompress(static_cast<const uint8_t*>(in), (size_t)n, &s)) return -1;
    if ((int64_t)s.size() > out_cap) return -2;
    std::memcpy(out, s.data(), s.size());
    return (int64_t)s.size();
  } catch (const std::exception&) {
    return -3;
  }
}

// serialized EntryDump of one entry; row = [num(dim) | state | found | ts] (mono_mtable_export layout)
int64_t mono_ckpt_encode_entry(const mono_segment_cfg* segs, int32_t nsegs, int64_t id, const float* row,
                               void* out, int64_t out_cap) {
  std::string rec;
  int rc = ck_guard([&] {
    std::vector<mono_segment_cfg> sv;
    int dim, state;
    set_segs(&sv, &dim, &state, segs, nsegs);
    rec.push_back((char)0x09);  // 1: sfixed64
    rec.append(reinterpret_cast<const char*>(&id), 8);
    put_floats_unpacked(&rec, 2, row, dim);
    std::string opt;
    const float* st = row + dim;
    for (const auto& s : sv) {
      // MovingAverage::Save returns an empty OptimizerDump (moving_average_optimizer.cc:54-57): nothing is appended
      if (s.opt_type != MONO_OPT_MOVING_AVERAGE) put_tag_len(&opt, 1, encode_single_opt(s, st));
      st += state_floats_of(s);
    }
    put_tag_len(&rec, 3, opt);
    uint32_t ts;
    std::memcpy(&ts, row + dim + state + 1, 4);
    put_varint(&rec, (4u << 3) | 0);
    put_varint(&rec, (uint64_t)ts);
  });
  if (rc != MONO_OK) return -1;
  if ((int64_t)rec.size() > out_cap) return -(int64_t)rec.size();
  std::memcpy(out, rec.data(), rec.size());
  return (int64_t)rec.size();
}

// inverse; row_out gets [num | state | found = 1 | ts]; missing optional fields keep zeros
int mono_ckpt_decode_entry(const mono_segment_cfg* segs, int32_t nsegs, const void* rec, int64_t n,
                           int64_t* id_out, float* row_out) {
  return ck_guard([&] {
    std::vector<mono_segment_cfg> sv;
    int dim, state;
    set_segs(&sv, &dim, &state, segs, nsegs);
    std::memset(row_out, 0, sizeof(float) * (size_t)(dim + state + 2));
    *id_out = 0;
    Reader r{static_cast<const uint8_t*>(rec), static_cast<const uint8_t*>(rec) + n};
    int n_num = 0;
    uint32_t ts = 0;
    while (r.ok && r.p < r.end) {
      const uint64_t tag = r.varint();
      const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
      const uint8_t* d;
      if (field == 1 && wt == 1) {
        if (r.bytes(8, &d)) std::memcpy(id_out, d, 8);
      } else if (field == 2) {
        read_floats(r, wt, row_out, dim, &n_num);
      } else if (field == 3 && wt == 2) {
        const uint64_t len = r.varint();
        if (!r.ok || !r.bytes(len, &d)) break;
        Reader o{d, d + len};
        size_t si = 0;
        float* st = row_out + dim;
        while (o.ok && o.p < o.end) {
          const uint64_t t2 = o.varint();
          if ((t2 >> 3) == 1 && (t2 & 7) == 2) {
            const uint64_t l2 = o.varint();
            const uint8_t* q;
            if (!o.ok || !o.bytes(l2, &q)) break;
            while (si < sv.size() && sv[si].opt_type == MONO_OPT_MOVING_AVERAGE) ++si;  // contributes no dump, no state
            if (si < sv.size()) {
              if (!decode_single_opt(q, l2, sv[si], st)) throw std::runtime_error("checkpoint: bad OptimizerDump");
              st += state_floats_of(sv[si]);
              ++si;
            }
          } else {
            o.skip((uint32_t)(t2 & 7));
          }
        }
        if (!o.ok) throw std::runtime_error("checkpoint: bad OptimizerDump");
      } else if (field == 4 && wt == 0) {
        ts = (uint32_t)r.varint();
      } else {
        r.skip(wt);
      }
    }
    if (!r.ok) throw std::runtime_error("checkpoint: bad EntryDump");
    const uint32_t one = 1;
    std::memcpy(row_out + dim + state, &one, 4);
    std::memcpy(row_out + dim + state + 1, &ts, 4);
  });
}

int mono_ckpt_writer_open(const char* data_path, const char* meta_path, int32_t snappy, mono_ckpt_writer** out) {
  return ck_guard([&] {
    if (!data_path || !meta_path || !out) throw ArgError("checkpoint writer: null argument");
    auto w = std::make_unique<mono_ckpt_writer>();
    w->data_path = data_path;
    w->meta_path = meta_path;
    w->snappy = snappy != 0;
    // written under a temporary name and renamed on close, like the reference (:193-196, :244-247)
    char suffix[64];
    std::snprintf(suffix, sizeof(suffix), "-tmp-%llx", (unsigned long long)(uintptr_t)w.get());
    w->data_tmp = w->data_path + suffix;
    w->meta_tmp = w->meta_path + suffix;
    w->data.f = std::fopen(w->data_tmp.c_str(), "wb");
    w->meta.f = std::fopen(w->meta_tmp.c_str(), "wb");
    if (!w->data.f || !w->meta.f) throw std::runtime_error("checkpoint: cannot create " + w->data_tmp);
    *out = w.release();
  });
}

int mono_ckpt_writer_begin_table(mono_ckpt_writer* w, const char* name, const mono_segment_cfg* segs,
                                 int32_t nsegs) {
  return ck_guard([&] {
    if (!w || !name) throw ArgError("checkpoint writer: null argument");
    if (w->in_table) throw ArgError("checkpoint writer: previous table not ended");
    set_segs(&w->segs, &w->dim, &w->state, segs, nsegs);
    w->table_name = name;
    w->table_entries = 0;
    w->in_table = true;
  });
}

// rows: n x (dim + state + 2) floats in mono_mtable_export's layout.  Entries whose row has expired are
// dropped like the reference's save (:214-221): max_update_ts - ts >= expire_days(slot) * 86400, slot =
// (id >> 48) & 0x7fff (reader_util.h:36-38).  expire_days_by_slot: 32768 entries, or NULL = keep everything.
int mono_ckpt_writer_add(mono_ckpt_writer* w, const int64_t* ids, const float* rows, int64_t n,
                         int64_t max_update_ts, const int64_t* expire_days_by_slot, int64_t* n_written) {
  return ck_guard([&] {
    if (!w || !w->in_table || (n > 0 && (!ids || !rows))) throw ArgError("checkpoint writer: bad add");
    const int width = w->dim + w->state + 2;
    std::string rec, opt;
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) {
      const float* row = rows + (size_t)i * width;
      uint32_t ts;
      std::memcpy(&ts, row + w->dim + w->state + 1, 4);
      if (expire_days_by_slot) {
        const int64_t slot = (ids[i] >> 48) & 0x7fff;
        if (max_update_ts - (int64_t)ts >= expire_days_by_slot[slot] * 24 * 3600) continue;
      }
      rec.clear();
      opt.clear();
      rec.push_back((char)0x09);
      rec.append(reinterpret_cast<const char*>(ids + i), 8);
      put_floats_unpacked(&rec, 2, row, w->dim);
      const float* st = row + w->dim;
      for (const auto& s : w->segs) {
        if (s.opt_type != MONO_OPT_MOVING_AVERAGE) put_tag_len(&opt, 1, encode_single_opt(s, st));
        st += state_floats_of(s);
      }
      put_tag_len(&rec, 3, opt);
      put_varint(&rec, (4u << 3) | 0);
      put_varint(&rec, (uint64_t)ts);
      frame_record(&w->block, rec);
      if (w->block.size() >= kSnappyInputBuffer) flush_block(w, false);
      ++kept;
    }
    w->table_entries += (uint64_t)kept;
    if (n_written) *n_written = kept;
  });
}

int mono_ckpt_writer_end_table(mono_ckpt_writer* w) {
  return ck_guard([&] {
    if (!w || !w->in_table) throw ArgError("checkpoint writer: no open table");
    std::string meta, framed;
    put_tag_len(&meta, 1, w->table_name);
    put_varint(&meta, (2u << 3) | 0);
    put_varint(&meta, w->table_entries);
    frame_record(&framed, meta);
    write_all(w->meta.f, framed.data(), framed.size());
    w->total_entries += w->table_entries;
    w->in_table = false;
  });
}

int mono_ckpt_writer_close(mono_ckpt_writer* w, int32_t commit) {
  if (!w) return MONO_OK;
  int rc = ck_guard([&] {
    if (commit) {
      if (w->in_table) throw ArgError("checkpoint writer: table not ended");
      flush_block(w, true);
    }
  });
  if (w->data.f && std::fclose(w->data.f) != 0 && rc == MONO_OK) rc = MONO_ERR_INTERNAL;
  w->data.f = nullptr;
  if (w->meta.f && std::fclose(w->meta.f) != 0 && rc == MONO_OK) rc = MONO_ERR_INTERNAL;
  w->meta.f = nullptr;
  if (commit && rc == MONO_OK) {
    if (std::rename(w->data_tmp.c_str(), w->data_path.c_str()) != 0 ||
        std::rename(w->meta_tmp.c_str(), w->meta_path.c_str()) != 0) {
      g_ckpt_error = "checkpoint: rename failed";
      rc = MONO_ERR_INTERNAL;
    }
  }
  if (!commit || rc != MONO_OK) {  // nothing half-written is left behind
    std::remove(w->data_tmp.c_str());
    std::remove(w->meta_tmp.c_str());
  }
  delete w;
  return rc;
}

int mono_ckpt_reader_open(const char* data_path, const char* meta_path, int32_t snappy, mono_ckpt_reader** out) {
  return ck_guard([&] {
    if (!data_path || !meta_path || !out) throw ArgError("checkpoint reader: null argument");
    auto r = std::make_unique<mono_ckpt_reader>();
    r->snappy = snappy != 0;
    r->data.f = std::fopen(data_path, "rb");
    r->meta.f = std::fopen(meta_path, "rb");
    if (!r->data.f || !r->meta.f) throw std::runtime_error(std::string("checkpoint: cannot open ") + data_path);
    *out = r.release();
  });
}

// next table of the shard: name into name_out (NUL-terminated, truncated to cap), its entry count; *has = 0
// at the end of the metadata stream.  Unread entries of the previous table are skipped.
int mono_ckpt_reader_next_table(mono_ckpt_reader* r, char* name_out, int32_t cap, int64_t* num_entries,
                                int32_t* has) {
  return ck_guard([&] {
    if (!r || !name_out || cap <= 0 || !num_entries || !has) throw ArgError("checkpoint reader: null argument");
    std::string rec;
    while (r->remaining > 0) {
      if (!read_record_stream(r, &rec)) throw std::runtime_error("checkpoint: data ended before metadata count");
      --r->remaining;
    }
    if (!read_record_file(r->meta.f, &rec)) {
      *has = 0;
      return;
    }
    Reader p{reinterpret_cast<const uint8_t*>(rec.data()), reinterpret_cast<const uint8_t*>(rec.data()) + rec.size()};
    std::string name;
    uint64_t cnt = 0;
    while (p.ok && p.p < p.end) {
      const uint64_t tag = p.varint();
      if ((tag >> 3) == 1 && (tag & 7) == 2) {
        const uint64_t len = p.varint();
        const uint8_t* d;
        if (p.ok && p.bytes(len, &d)) name.assign(reinterpret_cast<const char*>(d), len);
      } else if ((tag >> 3) == 2 && (tag & 7) == 0) {
        cnt = p.varint();
      } else {
        p.skip((uint32_t)(tag & 7));
      }
    }
    if (!p.ok) throw std::runtime_error("checkpoint: bad MultiHashTableMetadata");
    std::snprintf(name_out, (size_t)cap, "%s", name.c_str());
    *num_entries = (int64_t)cnt;
    r->remaining = cnt;
    r->segs.clear();
    *has = 1;
  });
}

// up to max_n entries of the current table, decoded against `segs` into ids_out / rows_out
// (n x (dim + state + 2), mono_mtable_restore_rows layout); *n_read < max_n only at the end of the table
int mono_ckpt_reader_read(mono_ckpt_reader* r, const mono_segment_cfg* segs, int32_t nsegs, int64_t* ids_out,
                          float* rows_out, int64_t max_n, int64_t* n_read) {
  return ck_guard([&] {
    if (!r || !ids_out || !rows_out || !n_read) throw ArgError("checkpoint reader: null argument");
    set_segs(&r->segs, &r->dim, &r->state, segs, nsegs);
    const int width = r->dim + r->state + 2;
    std::string rec;
    int64_t k = 0;
    while (k < max_n && r->remaining > 0) {
      if (!read_record_stream(r, &rec)) throw std::runtime_error("checkpoint: data ended before metadata count");
      --r->remaining;
      if (mono_ckpt_decode_entry(segs, nsegs, rec.data(), (int64_t)rec.size(), ids_out + k,
                                 rows_out + (size_t)k * width) != MONO_OK)
        throw std::runtime_error("checkpoint: Parse entry failed");
      ++k;
    }
    *n_read = k;
  });
}

int mono_ckpt_reader_close(mono_ckpt_reader* r) {
  delete r;
  return MONO_OK;
}

}  // extern "C"
