// Mel codec kernels (placeholder until the STFT / Griffin-Lim kernels land in this file).
#include "../../include/b200ad.h"
namespace b200ad { int set_err(const char* fmt, ...); }
extern "C" size_t b200ad_mel_scratch_bytes(const b200ad_mel_config*, int) { return 0; }
extern "C" int b200ad_mel_encode(const b200ad_mel_config*, const float*, uint8_t*, int, void*, size_t, void*) {
  return b200ad::set_err("mel_encode: not built");
}
extern "C" int b200ad_mel_decode(const b200ad_mel_config*, const uint8_t*, float*, int, uint64_t, void*, size_t, void*) {
  return b200ad::set_err("mel_decode: not built");
}
