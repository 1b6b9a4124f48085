// checksum.cu -- host-side CRC-32C (Castagnoli) for the TensorFlow checkpoint ("tensor bundle")
// importer/exporter (unflow_b200/e2eflow/core/tf_checkpoint.py).  The bundle format protects every
// index block and every tensor's bytes with a masked CRC-32C; the released FlowNetC checkpoints are
// ~157 MB per network, which a Python byte loop cannot check in reasonable time.  Slicing-by-8,
// table built once; no device code.
#include <cstddef>
#include <cstdint>
#include <mutex>

#include "common.cuh"

namespace {
uint32_t g_tab[8][256];
std::once_flag g_once;

void build_tables() {
  const uint32_t poly = 0x82f63b78u;  // reflected 0x1EDC6F41
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ poly : (c >> 1);
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xffu];
}
}  // namespace

extern "C" unsigned int unflow_crc32c(const void *data, size_t n, unsigned int crc) {
  std::call_once(g_once, build_tables);
  const unsigned char *p = static_cast<const unsigned char *>(data);
  uint32_t c = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) {
    c = g_tab[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
    --n;
  }
  while (n >= 8) {
    uint64_t w = *reinterpret_cast<const uint64_t *>(p) ^ c;  // little-endian hosts only (x86-64 / aarch64)
    c = g_tab[7][w & 0xff] ^ g_tab[6][(w >> 8) & 0xff] ^ g_tab[5][(w >> 16) & 0xff] ^
        g_tab[4][(w >> 24) & 0xff] ^ g_tab[3][(w >> 32) & 0xff] ^ g_tab[2][(w >> 40) & 0xff] ^
        g_tab[1][(w >> 48) & 0xff] ^ g_tab[0][(w >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_tab[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
  return ~c;
}
