// container.cpp -- wire format for many CABAC streams (host side, no GPU work).
//
// The reference knows one file per stream (SimpleCABACMex.cpp:195 opens the file for writing,
// :288 for reading) and ships the context initialisation separately as uint8 side information
// in a .mat file (ISS/ISS.m:197-201, cabacEncode.m:30).  A batch of N streams would be N files
// plus N side-info records; the container puts them in one buffer without touching the bytes
// of any stream: stream s is payload[byte_off[s] .. byte_off[s+1]) and those bytes are exactly
// the file the reference encoder would have written, so cabac_container_stream() hands out a
// range that the reference's decodeStart ... decodeFinish reads unchanged.
//
// Layout (little-endian, every section 8-byte aligned):
//   header (128 B)  magic "ISSCABAC", version, n_streams, n_ctx, flags, sym_width, symcfg,
//                   payload_bytes, section offsets, total_bytes, CRC-32 of tables / payload / header
//   byte_off        u64[n_streams + 1]
//   unit_off        u64[n_streams + 1]   symbols (or ops) per stream as exclusive offsets (optional)
//   ctx_init        u8[n_ctx] or u8[n_streams][n_ctx]  (state bytes, or the uint8 p(0) side info)
//   payload         u8[payload_bytes]
#include <stdint.h>
#include <string.h>

#include "../../include/isscabac.h"
#include "internal.h"

using isscabac_internal::set_error;

namespace {

constexpr uint32_t kVersion = 1;
constexpr uint32_t kHeaderBytes = 128;
constexpr char kMagic[8] = {'I', 'S', 'S', 'C', 'A', 'B', 'A', 'C'};
enum { F_PER_STREAM = 1u, F_HAS_CFG = 2u, F_HAS_UNITS = 4u, F_CTX_IS_PROB = 8u };

// CRC-32 (IEEE 802.3, reflected, as zlib), slicing-by-8
struct CrcTab {
  uint32_t t[8][256];
  CrcTab() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xffu];
  }
};
const CrcTab& crc_tab() {
  static const CrcTab T;
  return T;
}
uint32_t crc32_update(uint32_t crc, const uint8_t* p, uint64_t n) {
  const CrcTab& T = crc_tab();
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) { crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 0xffu]; --n; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= crc;
    crc = T.t[7][v & 0xff] ^ T.t[6][(v >> 8) & 0xff] ^ T.t[5][(v >> 16) & 0xff] ^ T.t[4][(v >> 24) & 0xff] ^
          T.t[3][(v >> 32) & 0xff] ^ T.t[2][(v >> 40) & 0xff] ^ T.t[1][(v >> 48) & 0xff] ^ T.t[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 0xffu];
  return ~crc;
}

inline uint64_t align8(uint64_t x) { return (x + 7u) & ~7ull; }
inline void put32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
inline void put64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }
inline uint32_t get32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t get64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct Sections {
  uint64_t off_boff, off_units, off_ctx, off_payload, total, ctx_bytes;
};
int layout(const isscabac_container_view* v, Sections& S) {
  if (!v) { set_error("container: null view"); return ISSCABAC_ERR_INVALID; }
  if (v->n_ctx > ISSCABAC_MAX_CTX) { set_error("container: n_ctx %u > %u", v->n_ctx, ISSCABAC_MAX_CTX); return ISSCABAC_ERR_INVALID; }
  const uint64_t tab = ((uint64_t)v->n_streams + 1) * 8;
  S.ctx_bytes = (uint64_t)v->n_ctx * (v->per_stream_init ? v->n_streams : 1);
  S.off_boff = kHeaderBytes;
  S.off_units = S.off_boff + tab;
  S.off_ctx = S.off_units + (v->unit_off ? tab : 0);
  S.off_payload = align8(S.off_ctx + S.ctx_bytes);
  S.total = align8(S.off_payload + v->payload_bytes);
  return ISSCABAC_OK;
}

}  // namespace

extern "C" {

uint32_t cabac_crc32(const uint8_t* p, uint64_t n) { return crc32_update(0, p, n); }

uint64_t cabac_container_size(const isscabac_container_view* v) {
  Sections S;
  return layout(v, S) == ISSCABAC_OK ? S.total : 0;
}

int cabac_container_write(const isscabac_container_view* v, uint8_t* out, uint64_t cap, uint64_t* written) {
  Sections S;
  int rc = layout(v, S);
  if (rc) return rc;
  if (!out || !v->byte_off || (S.ctx_bytes && !v->ctx_init) || (v->payload_bytes && !v->payload)) {
    set_error("cabac_container_write: null pointer");
    return ISSCABAC_ERR_INVALID;
  }
  if (cap < S.total) { set_error("container needs %llu bytes, capacity %llu", (unsigned long long)S.total, (unsigned long long)cap); return ISSCABAC_ERR_OVERFLOW; }
  const uint32_t n = v->n_streams;
  if (v->byte_off[0] != 0 || v->byte_off[n] != v->payload_bytes) { set_error("container: byte_off must run from 0 to payload_bytes"); return ISSCABAC_ERR_INVALID; }
  for (uint32_t s = 0; s < n; ++s) {
    if (v->byte_off[s + 1] < v->byte_off[s] || (v->unit_off && v->unit_off[s + 1] < v->unit_off[s])) {
      set_error("container: offsets not monotone at stream %u", s);
      return ISSCABAC_ERR_INVALID;
    }
  }
  memset(out, 0, kHeaderBytes);
  const uint64_t tab = ((uint64_t)n + 1) * 8;
  memcpy(out + S.off_boff, v->byte_off, tab);
  if (v->unit_off) memcpy(out + S.off_units, v->unit_off, tab);
  if (S.ctx_bytes) memcpy(out + S.off_ctx, v->ctx_init, S.ctx_bytes);
  memset(out + S.off_ctx + S.ctx_bytes, 0, S.off_payload - (S.off_ctx + S.ctx_bytes));
  if (v->payload_bytes) memcpy(out + S.off_payload, v->payload, v->payload_bytes);
  memset(out + S.off_payload + v->payload_bytes, 0, S.total - (S.off_payload + v->payload_bytes));

  memcpy(out, kMagic, 8);
  put32(out + 8, kVersion);
  put32(out + 12, kHeaderBytes);
  put32(out + 16, n);
  put32(out + 20, v->n_ctx);
  put32(out + 24, (v->per_stream_init ? F_PER_STREAM : 0u) | (v->has_cfg ? F_HAS_CFG : 0u) |
                      (v->unit_off ? F_HAS_UNITS : 0u) | (v->ctx_is_prob ? F_CTX_IS_PROB : 0u));
  put32(out + 28, (uint32_t)v->sym_width);
  if (v->has_cfg) {
    put32(out + 32, (uint32_t)v->cfg.profile); put32(out + 36, (uint32_t)v->cfg.method); put32(out + 40, v->cfg.Nq);
    put32(out + 44, (uint32_t)v->cfg.Nlbp); put32(out + 48, v->cfg.types); put32(out + 52, v->cfg.rows);
  }
  put64(out + 56, v->payload_bytes);
  put64(out + 64, S.off_boff);
  put64(out + 72, v->unit_off ? S.off_units : 0);
  put64(out + 80, S.off_ctx);
  put64(out + 88, S.off_payload);
  put64(out + 96, S.total);
  put32(out + 104, cabac_crc32(out + kHeaderBytes, S.off_payload - kHeaderBytes));
  put32(out + 108, cabac_crc32(out + S.off_payload, v->payload_bytes));
  put32(out + 112, cabac_crc32(out, 112));
  if (written) *written = S.total;
  return ISSCABAC_OK;
}

int cabac_container_parse(const uint8_t* buf, uint64_t n, int verify_payload_crc, isscabac_container_view* v) {
  if (!buf || !v) { set_error("cabac_container_parse: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (reinterpret_cast<uintptr_t>(buf) & 7u) { set_error("container buffer must be 8-byte aligned"); return ISSCABAC_ERR_INVALID; }
  if (n < kHeaderBytes || memcmp(buf, kMagic, 8) != 0) { set_error("not an ISSCABAC container"); return ISSCABAC_ERR_CORRUPT; }
  if (get32(buf + 8) != kVersion || get32(buf + 12) != kHeaderBytes) { set_error("container version %u not supported", get32(buf + 8)); return ISSCABAC_ERR_UNSUPPORTED; }
  if (get32(buf + 112) != cabac_crc32(buf, 112)) { set_error("container header checksum mismatch"); return ISSCABAC_ERR_CORRUPT; }
  memset(v, 0, sizeof *v);
  v->n_streams = get32(buf + 16);
  v->n_ctx = get32(buf + 20);
  const uint32_t flags = get32(buf + 24);
  v->per_stream_init = (flags & F_PER_STREAM) ? 1 : 0;
  v->has_cfg = (flags & F_HAS_CFG) ? 1 : 0;
  v->ctx_is_prob = (flags & F_CTX_IS_PROB) ? 1 : 0;
  v->sym_width = (int32_t)get32(buf + 28);
  v->cfg.profile = (int32_t)get32(buf + 32); v->cfg.method = (int32_t)get32(buf + 36); v->cfg.Nq = get32(buf + 40);
  v->cfg.Nlbp = (int32_t)get32(buf + 44); v->cfg.types = get32(buf + 48); v->cfg.rows = get32(buf + 52);
  v->payload_bytes = get64(buf + 56);
  // The header checksum is no defence against a crafted header (anyone can recompute it), so every section is
  // bounded against the buffer BEFORE anything is read or any pointer is derived: tables first (their size follows
  // from the counts alone and cannot overflow 64 bits: n_streams < 2^32, n_ctx <= 999), then the payload against
  // what is left.  Only then is the layout recomputed and compared with the stored section table.
  if (v->n_ctx > ISSCABAC_MAX_CTX) { set_error("container: n_ctx %u > %u", v->n_ctx, ISSCABAC_MAX_CTX); return ISSCABAC_ERR_CORRUPT; }
  {
    const uint64_t tab = ((uint64_t)v->n_streams + 1) * 8;
    const uint64_t ctx_bytes = (uint64_t)v->n_ctx * (v->per_stream_init ? v->n_streams : 1);
    const uint64_t off_payload = align8(kHeaderBytes + tab * ((flags & F_HAS_UNITS) ? 2 : 1) + ctx_bytes);
    if (off_payload > n || v->payload_bytes > n - off_payload || align8(v->payload_bytes) > n - off_payload) {
      set_error("container truncated or section sizes exceed the buffer (%llu bytes)", (unsigned long long)n);
      return ISSCABAC_ERR_CORRUPT;
    }
  }
  isscabac_container_view probe = *v;
  static const uint64_t dummy = 0;
  probe.unit_off = (flags & F_HAS_UNITS) ? &dummy : nullptr;
  Sections S;
  int rc = layout(&probe, S);
  if (rc) return ISSCABAC_ERR_CORRUPT;
  if (get64(buf + 64) != S.off_boff || get64(buf + 72) != ((flags & F_HAS_UNITS) ? S.off_units : 0) ||
      get64(buf + 80) != S.off_ctx || get64(buf + 88) != S.off_payload || get64(buf + 96) != S.total) {
    set_error("container section table inconsistent");
    return ISSCABAC_ERR_CORRUPT;
  }
  if (S.total > n) { set_error("container truncated: %llu of %llu bytes", (unsigned long long)n, (unsigned long long)S.total); return ISSCABAC_ERR_CORRUPT; }
  if (get32(buf + 104) != cabac_crc32(buf + kHeaderBytes, S.off_payload - kHeaderBytes)) { set_error("container table checksum mismatch"); return ISSCABAC_ERR_CORRUPT; }
  if (verify_payload_crc && get32(buf + 108) != cabac_crc32(buf + S.off_payload, v->payload_bytes)) { set_error("container payload checksum mismatch"); return ISSCABAC_ERR_CORRUPT; }
  v->byte_off = reinterpret_cast<const uint64_t*>(buf + S.off_boff);
  v->unit_off = (flags & F_HAS_UNITS) ? reinterpret_cast<const uint64_t*>(buf + S.off_units) : nullptr;
  v->ctx_init = S.ctx_bytes ? buf + S.off_ctx : nullptr;
  v->payload = buf + S.off_payload;
  const uint32_t ns = v->n_streams;
  if (v->byte_off[0] != 0 || v->byte_off[ns] != v->payload_bytes) { set_error("container offset table does not span the payload"); return ISSCABAC_ERR_CORRUPT; }
  if (v->unit_off && v->unit_off[0] != 0) { set_error("container unit table does not start at 0"); return ISSCABAC_ERR_CORRUPT; }
  for (uint32_t s = 0; s < ns; ++s)
    if (v->byte_off[s + 1] < v->byte_off[s] || (v->unit_off && v->unit_off[s + 1] < v->unit_off[s])) {
      set_error("container offsets not monotone at stream %u", s);
      return ISSCABAC_ERR_CORRUPT;
    }
  return ISSCABAC_OK;
}

int cabac_container_stream(const isscabac_container_view* v, uint32_t s, const uint8_t** bytes, uint64_t* n_bytes,
                           const uint8_t** ctx_init, uint64_t* n_units) {
  if (!v || !v->byte_off || s >= v->n_streams) { set_error("cabac_container_stream: stream %u out of range", s); return ISSCABAC_ERR_INVALID; }
  if (bytes) *bytes = v->payload + v->byte_off[s];
  if (n_bytes) *n_bytes = v->byte_off[s + 1] - v->byte_off[s];
  if (ctx_init) *ctx_init = v->ctx_init ? v->ctx_init + (v->per_stream_init ? (uint64_t)s * v->n_ctx : 0) : nullptr;
  if (n_units) *n_units = v->unit_off ? v->unit_off[s + 1] - v->unit_off[s] : 0;
  return ISSCABAC_OK;
}

}  // extern "C"
