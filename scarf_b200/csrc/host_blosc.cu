// Host-side decoder of the chunk codec Scarf stores are written with: numcodecs.Blosc(cname='lz4', clevel=5,
// shuffle=BITSHUFFLE) for count matrices and shuffle=SHUFFLE for the metadata columns (scarf/writers.py:79-89;
// numcodecs / c-blosc are not vendored in the reference).  Published frame format of c-blosc 1.x, restated:
//   header  u8 version, u8 versionlz, u8 flags, u8 typesize, u32 nbytes, u32 blocksize, u32 cbytes (little endian)
//   flags   0x01 byte shuffle, 0x02 stored (memcpy), 0x04 bit shuffle, 0x10 blocks are not split, bits 5-7 codec
//   body    int32 bstarts[nblocks], then per block `nsplits` streams of [int32 csize][LZ4 block | raw bytes]
// Only the LZ4 codec (format 1) and stored frames are decoded; everything else is an argument error.  Pure function
// of its inputs: callers decode chunks from several host threads at once (ctypes drops the GIL).
#include <string.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

constexpr uint8_t kByteShuffle = 0x01, kStored = 0x02, kBitShuffle = 0x04, kDontSplit = 0x10;

struct Header {
  uint8_t version, versionlz, flags, typesize;
  uint32_t nbytes, blocksize, cbytes;
};

inline uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

bool read_header(const uint8_t* f, int64_t n, Header* h) {
  if (n < 16) return false;
  h->version = f[0], h->versionlz = f[1], h->flags = f[2], h->typesize = f[3];
  h->nbytes = rd32(f + 4), h->blocksize = rd32(f + 8), h->cbytes = rd32(f + 12);
  return true;
}

// LZ4 block format: sequences of [token][literal length ext][literals][offset u16][match length ext]; the last
// sequence ends after its literals.  Returns the number of bytes produced or -1 on malformed / overrunning input.
int64_t lz4_block(const uint8_t* ip, int64_t n_in, uint8_t* dst, int64_t n_out) {
  const uint8_t* const iend = ip + n_in;
  uint8_t* op = dst;
  uint8_t* const oend = dst + n_out;
  while (ip < iend) {
    const uint32_t token = *ip++;
    int64_t lit = token >> 4;
    if (lit == 15) {
      uint8_t b;
      do {
        if (ip >= iend) return -1;
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if (lit > iend - ip || lit > oend - op) return -1;
    memcpy(op, ip, (size_t)lit);
    ip += lit, op += lit;
    if (ip >= iend) break;
    if (iend - ip < 2) return -1;
    const int64_t offset = ip[0] | (ip[1] << 8);
    ip += 2;
    if (offset == 0 || offset > op - dst) return -1;
    int64_t len = token & 15;
    if (len == 15) {
      uint8_t b;
      do {
        if (ip >= iend) return -1;
        b = *ip++;
        len += b;
      } while (b == 255);
    }
    len += 4;
    if (len > oend - op) return -1;
    const uint8_t* m = op - offset;
    if (offset >= len) {
      memcpy(op, m, (size_t)len);
    } else if (offset == 1) {
      memset(op, m[0], (size_t)len);  // runs of one byte: most of a bit-shuffled count matrix
    } else {
      // overlapping match = periodic extension of the last `offset` bytes: one period, then doubling copies
      memcpy(op, m, (size_t)offset);
      for (int64_t filled = offset; filled < len;) {
        const int64_t n = std::min(filled, len - filled);
        memcpy(op + filled, op, (size_t)n);
        filled += n;
      }
    }
    op += len;
  }
  return op - dst;
}

// src holds byte j of every element contiguously (typesize planes of n elements); trailing bytes are verbatim
void unshuffle_bytes(const uint8_t* src, uint8_t* dst, int64_t bsize, int ts) {
  const int64_t n = bsize / ts;
  for (int j = 0; j < ts; ++j) {
    const uint8_t* plane = src + (int64_t)j * n;
    for (int64_t i = 0; i < n; ++i) dst[i * ts + j] = plane[i];
  }
  memcpy(dst + n * ts, src + n * ts, (size_t)(bsize - n * ts));
}

inline uint64_t transpose8x8(uint64_t x) {  // byte r, bit c  <->  byte c, bit r
  uint64_t t;
  t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull, x = x ^ t ^ (t << 7);
  t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull, x = x ^ t ^ (t << 14);
  t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull, x = x ^ t ^ (t << 28);
  return x;
}

// bitshuffle layout over the first nelem - nelem % 8 elements: bit row (b * 8 + t) holds bit t of byte b of every
// element, element e at bit e % 8 of byte e / 8 of the row; the remaining bytes are verbatim
void unshuffle_bits(const uint8_t* src, uint8_t* dst, int64_t bsize, int ts) {
  const int64_t nelem = bsize / ts, n8 = nelem - nelem % 8, row = n8 / 8;
  memset(dst, 0, (size_t)(n8 * ts));
  for (int b = 0; b < ts; ++b) {
    const uint8_t* rows = src + (int64_t)b * 8 * row;  // the eight bit rows of byte b
    int64_t k = 0;
    // 64 elements at a time: one 8-byte word of each bit row.  Count matrices are sparse and their values small, so
    // most words (all of the high bytes, most high bits) are zero and the elements keep their zero bytes.
    for (; k + 8 <= row; k += 8) {
      uint64_t r[8], any = 0;
      for (int t = 0; t < 8; ++t) {
        memcpy(&r[t], rows + t * row + k, 8);
        any |= r[t];
      }
      if (!any) continue;
      for (int j = 0; j < 8; ++j) {
        uint64_t x = 0;
        for (int t = 0; t < 8; ++t) x |= ((r[t] >> (8 * j)) & 0xFFull) << (8 * t);
        if (!x) continue;
        x = transpose8x8(x);
        uint8_t* o = dst + (k + j) * 8 * ts + b;
        for (int e = 0; e < 8; ++e) o[(int64_t)e * ts] = (uint8_t)(x >> (8 * e));
      }
    }
    for (; k < row; ++k) {
      uint64_t x = 0;
      for (int t = 0; t < 8; ++t) x |= (uint64_t)rows[t * row + k] << (8 * t);
      x = transpose8x8(x);
      uint8_t* o = dst + k * 8 * ts + b;
      for (int e = 0; e < 8; ++e) o[(int64_t)e * ts] = (uint8_t)(x >> (8 * e));
    }
  }
  memcpy(dst + n8 * ts, src + n8 * ts, (size_t)(bsize - n8 * ts));
}

}  // namespace

extern "C" int32_t scf_host_blosc_info(const void* frame, int64_t frame_bytes, int64_t* nbytes, int32_t* typesize,
                                       int32_t* flags) {
  Header h;
  SCF_ARG(frame && read_header((const uint8_t*)frame, frame_bytes, &h), "frame shorter than the 16-byte header");
  if (nbytes) *nbytes = h.nbytes;
  if (typesize) *typesize = h.typesize;
  if (flags) *flags = h.flags;
  return 0;
}

extern "C" int32_t scf_host_blosc_decode(const void* frame, int64_t frame_bytes, void* dst, int64_t dst_bytes) {
  const uint8_t* f = (const uint8_t*)frame;
  Header h;
  SCF_ARG(f && read_header(f, frame_bytes, &h), "frame shorter than the 16-byte header");
  SCF_ARG(dst || h.nbytes == 0, "null destination");
  SCF_ARG((int64_t)h.nbytes == dst_bytes, "destination size differs from the frame's nbytes");
  SCF_ARG((int64_t)h.cbytes <= frame_bytes, "frame truncated (cbytes beyond the buffer)");
  if (h.nbytes == 0) return 0;
  if (h.flags & kStored) {
    SCF_ARG(16 + (int64_t)h.nbytes <= frame_bytes, "stored frame truncated");
    memcpy(dst, f + 16, h.nbytes);
    return 0;
  }
  SCF_ARG(((h.flags >> 5) & 7) == 1, "only the lz4 codec of Blosc is decoded (cname='lz4', scarf/writers.py:79-89)");
  SCF_ARG(h.blocksize > 0 && h.blocksize <= h.nbytes && h.typesize > 0, "bad header");
  const int64_t nblocks = ((int64_t)h.nbytes + h.blocksize - 1) / h.blocksize;
  SCF_ARG(16 + 4 * nblocks <= frame_bytes, "frame truncated (block table)");
  const int ts = h.typesize;
  const bool bitsh = (h.flags & kBitShuffle) != 0, bytesh = !bitsh && (h.flags & kByteShuffle) && ts > 1;
  std::vector<uint8_t> tmp((bitsh || bytesh) ? h.blocksize : 0);
  uint8_t* out = (uint8_t*)dst;
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t bsize = std::min<int64_t>(h.blocksize, (int64_t)h.nbytes - b * h.blocksize);
    const bool leftover = bsize != (int64_t)h.blocksize;
    const int nsplits = (!(h.flags & kDontSplit) && !leftover && ts <= 16 && bsize / ts >= 128) ? ts : 1;
    const int64_t neblock = bsize / nsplits;
    uint8_t* target = (bitsh || bytesh) ? tmp.data() : out + b * h.blocksize;
    int64_t pos = (int32_t)rd32(f + 16 + 4 * b);
    for (int s = 0; s < nsplits; ++s) {
      SCF_ARG(pos >= 0 && pos + 4 <= frame_bytes, "frame truncated (split header)");
      const int64_t csize = (int32_t)rd32(f + pos);
      pos += 4;
      SCF_ARG(csize >= 0 && pos + csize <= frame_bytes, "frame truncated (split body)");
      if (csize == neblock) {
        memcpy(target + s * neblock, f + pos, (size_t)neblock);
      } else {
        SCF_ARG(lz4_block(f + pos, csize, target + s * neblock, neblock) == neblock, "malformed LZ4 block");
      }
      pos += csize;
    }
    if (bitsh) unshuffle_bits(tmp.data(), out + b * h.blocksize, bsize, ts);
    else if (bytesh) unshuffle_bytes(tmp.data(), out + b * h.blocksize, bsize, ts);
  }
  return 0;
}
