// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) for the checkpoint files (host only): TensorFlow's tensor-bundle format
// stores a masked CRC-32C of every tensor and of every index block (tensorflow/core/lib/hash/crc32c.h; the reference writes such
// checkpoints through tf.train.Saver, models/base_model.py:62, :401-417).  SSE4.2 crc32 instruction when the CPU has it,
// slicing-by-8 tables otherwise.
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "../../include/pamrec_b200.h"

namespace {

struct Tables {
  uint32_t t[8][256];
  Tables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};

uint32_t crc_tables(uint32_t crc, const unsigned char* p, size_t n) {
  static const Tables T;
  while (n >= 8) {
    uint64_t w;
    std::memcpy(&w, p, 8);
    w ^= crc;                                                   // little-endian host (x86-64 / aarch64)
    crc = T.t[7][w & 0xff] ^ T.t[6][(w >> 8) & 0xff] ^ T.t[5][(w >> 16) & 0xff] ^ T.t[4][(w >> 24) & 0xff] ^
          T.t[3][(w >> 32) & 0xff] ^ T.t[2][(w >> 40) & 0xff] ^ T.t[1][(w >> 48) & 0xff] ^ T.t[0][(w >> 56) & 0xff];
    p += 8; n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 0xff];
  return crc;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t crc_hw(uint32_t crc, const unsigned char* p, size_t n) {
  uint64_t c = crc;
  while (n >= 8) {
    uint64_t w;
    std::memcpy(&w, p, 8);
    c = __builtin_ia32_crc32di(c, w);
    p += 8; n -= 8;
  }
  uint32_t c32 = (uint32_t)c;
  while (n--) c32 = __builtin_ia32_crc32qi(c32, *p++);
  return c32;
}
#endif

}  // namespace

extern "C" uint32_t pamrec_crc32c(uint32_t crc, const void* data, size_t n) {
  const unsigned char* p = (const unsigned char*)data;
  crc = ~crc;
#if defined(__x86_64__)
  if (__builtin_cpu_supports("sse4.2")) return ~crc_hw(crc, p, n);
#endif
  return ~crc_tables(crc, p, n);
}

extern "C" uint32_t pamrec_crc32c_portable(uint32_t crc, const void* data, size_t n) {
  return ~crc_tables(~crc, (const unsigned char*)data, n);
}
