// kdf_inflate.h — whole-buffer DEFLATE decoder + CRC-32 used by the BGZF reader (kdf_host.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace kdf {
// Inflate one complete raw DEFLATE stream of `in_len` bytes into exactly `out_len` bytes.
// False on any malformed stream, on a stream that ends early or late, or when the inflated
// size is not `out_len`.  Never reads outside [in, in + in_len) nor writes outside
// [out, out + out_len).
bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
// CRC-32 of the gzip trailer (PCLMULQDQ folding where the CPU has it, zlib's otherwise).
uint32_t crc32_of(const uint8_t* buf, size_t len);
}  // namespace kdf
