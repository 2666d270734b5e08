// pgz.cu — parallel inflate of an ordinary (single-member) gzip file on the host threads, for `.fastq.gz` / `.fa.gz` input
// (the reference hands such files to bwa, which reads them through zlib on one thread; at 13 M reads/s on the GPU a
// single zlib stream (~0.2-0.3 GB/s of text, < 1 M reads/s) would be the whole run time).
//
// A deflate stream can only be decoded from its start: blocks begin at arbitrary bit offsets and matches reach 32 KiB back.  So:
//   1. the compressed bytes are cut into chunks; for every chunk a finder looks for the first bit offset at which a non-final
//      dynamic-Huffman block starts (header fields in range, complete code-length / literal / distance codes, and the block plus
//      the header after it decode cleanly);
//   2. chunks are decoded in parallel from their start to the next chunk's start into 16-bit symbols: a byte, or - where a match
//      reaches back past the chunk start - a marker holding the position in the unknown 32 KiB window;
//   3. windows are propagated chunk by chunk (32 KiB each, sequential), then all markers are replaced in parallel while the bytes
//      are written to the output and CRC-32'd.
// Every assumption is checked: a decode error, a chunk that does not end exactly where the next one starts, a marker in the first
// chunk, a second gzip member, or a CRC-32 / ISIZE mismatch against the gzip trailer makes pgz_inflate() return false, and the
// caller reads the file through zlib as before.  BGZF files (bgzip, BAM) have independent members and take bam.cu's block path.
// Checked against zlib's output in tests/test_gz_input.py.
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <memory>
#ifdef __SSE2__
#include <emmintrin.h>
#endif

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "gz.h"
#include "internal.cuh"

namespace catt {
namespace {

constexpr int FAST_BITS = 11;
constexpr uint16_t MARK = 0x8000;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// LSB-first bit reader over a buffer with >= 8 readable bytes of padding behind `nbits`
struct Bits {
  const uint8_t* base; uint64_t pos, nbits;
  inline uint64_t window() const { uint64_t w; memcpy(&w, base + (pos >> 3), 8); return w >> (pos & 7); }      // >= 57 valid bits
  inline uint32_t peek(int n) const { return (uint32_t)(window() & ((1ull << n) - 1)); }
  inline uint32_t take(int n) { const uint32_t v = peek(n); pos += n; return v; }
  inline bool over() const { return pos > nbits; }
};

struct Huff {
  uint16_t fast[1 << FAST_BITS];          // (symbol << 4) | length, 0 = longer than `fb` bits (or unused); the first 2^fb entries are used
  uint16_t count[16], symbol[288];
  int maxlen, fb;                         // fb: index width of `fast` for this table (small tables are cheaper to build per block)
};

// zlib's rule (inftrees.c): over-subscribed sets are invalid; an incomplete set is invalid for the code-length code (kind 0) and,
// for the literal/length and distance codes (kind 1), unless it is a single code of length 1 (or, distance only, no code at all)
bool build_huff(Huff& h, const uint8_t* lens, int n, int kind, int fb = FAST_BITS) {
  h.fb = fb;
  memset(h.count, 0, sizeof h.count);
  for (int i = 0; i < n; i++) h.count[lens[i]]++;
  h.count[0] = 0;
  int left = 1, total = 0;
  h.maxlen = 0;
  for (int l = 1; l <= 15; l++) {
    left = (left << 1) - h.count[l];
    if (left < 0) return false;                       // over-subscribed
    total += h.count[l];
    if (h.count[l]) h.maxlen = l;
  }
  if (left > 0 && (kind == 0 || h.maxlen > 1)) return false;     // incomplete
  uint16_t offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h.count[l];
  for (int i = 0; i < n; i++) if (lens[i]) h.symbol[offs[lens[i]]++] = (uint16_t)i;
  memset(h.fast, 0, sizeof(uint16_t) << fb);
  int code = 0, idx = 0;
  for (int l = 1; l <= std::min(15, fb); l++) {
    for (int k = 0; k < h.count[l]; k++, code++, idx++) {
      uint32_t rev = 0;
      for (int b = 0; b < l; b++) rev |= ((code >> b) & 1u) << (l - 1 - b);
      const uint16_t e = (uint16_t)((h.symbol[idx] << 4) | l);
      for (uint32_t x = rev; x < (1u << fb); x += 1u << l) h.fast[x] = e;
    }
    code <<= 1;
  }
  return true;
}

inline int decode_sym(const Huff& h, Bits& b) {
  const uint64_t w = b.window();
  const uint16_t e = h.fast[w & ((1u << h.fb) - 1)];
  if (e) { b.pos += e & 15; return e >> 4; }
  int code = 0, first = 0, index = 0;                 // canonical walk (codes longer than FAST_BITS are rare)
  for (int l = 1; l <= h.maxlen; l++) {
    code |= (int)((w >> (l - 1)) & 1);
    const int cnt = h.count[l];
    if (code - cnt < first) { b.pos += l; return h.symbol[index + (code - first)]; }
    index += cnt; first += cnt; first <<= 1; code <<= 1;
  }
  return -1;
}

// First-touch page faults are the cost that does not parallelise (and are slow in a VM): big buffers are 2 MiB aligned and ask for
// transparent huge pages, and the symbol buffers are reused from wave to wave.
void* big_alloc(size_t bytes) {
  void* p = nullptr;
  const size_t al = 2u << 20;
  bytes = (bytes + al - 1) / al * al;
  if (posix_memalign(&p, al, bytes) != 0) return nullptr;
  madvise(p, bytes, MADV_HUGEPAGE);
  return p;
}
void advise_huge(const void* p, size_t bytes) {
  const uintptr_t al = 2u << 20, a = ((uintptr_t)p + al - 1) / al * al, e = ((uintptr_t)p + bytes) / al * al;
  if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
}

// T = uint16_t: bytes and window markers (a chunk that starts in the middle of a stream); T = uint8_t: plain bytes (a stream decoded
// from its start, e.g. a BGZF member - a match that reaches back past the start is then an error)
template <class T>
struct OutBuf {
  T* p = nullptr; size_t n = 0, cap = 0;
  OutBuf() = default;
  OutBuf(const OutBuf&) = delete;
  OutBuf& operator=(const OutBuf&) = delete;
  OutBuf(OutBuf&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
  OutBuf& operator=(OutBuf&& o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; } return *this; }
  ~OutBuf() { free(p); }
  bool room(size_t extra) {
    if (n + extra <= cap) return true;
    const size_t want = std::max(cap * 2, n + extra + (4u << 20));
    T* q = (T*)big_alloc(want * sizeof(T));
    if (!q) return false;
    if (n) memcpy(q, p, n * sizeof(T));
    free(p);
    p = q; cap = want;
    return true;
  }
};

using SymBuf = OutBuf<uint16_t>;

// lit32 / dist32: the fast tables with everything a symbol needs folded in, so the hot loop does one lookup per code:
//   bits 0-3 code length, bits 4-7 number of extra bits, bits 8-9 kind (1 literal, 2 match length, 3 end of block; distance: 1),
//   bits 16-31 the literal, the length base or the distance base.  0 = code longer than the table index (or invalid): slow path.
constexpr int DIST_FAST_BITS = 10;
struct Tables { Huff lit, dist; uint32_t lit32[1 << FAST_BITS], dist32[1 << DIST_FAST_BITS]; };

void fold_tables(Tables& t) {
  for (uint32_t i = 0; i < (1u << FAST_BITS); i++) {
    const uint32_t e = t.lit.fast[i], sym = e >> 4, l = e & 15;
    uint32_t v = 0;
    if (e) {
      if (sym < 256) v = (sym << 16) | (1u << 8) | l;
      else if (sym == 256) v = (3u << 8) | l;
      else if (sym <= 285) v = ((uint32_t)LEN_BASE[sym - 257] << 16) | (2u << 8) | ((uint32_t)LEN_EXTRA[sym - 257] << 4) | l;
    }
    t.lit32[i] = v;
  }
  for (uint32_t i = 0; i < (1u << DIST_FAST_BITS); i++) {
    const uint32_t e = t.dist.fb == DIST_FAST_BITS ? t.dist.fast[i] : 0, sym = e >> 4, l = e & 15;
    t.dist32[i] = (e && sym <= 29) ? ((uint32_t)DIST_BASE[sym] << 16) | (1u << 8) | ((uint32_t)DIST_EXTRA[sym] << 4) | l : 0;
  }
}

// Reads a dynamic block header at b (BFINAL / BTYPE already consumed) and builds both tables.
bool read_dynamic_header(Bits& b, Tables& t) {
  if (b.pos + 14 + 12 > b.nbits) return false;
  const int hlit = (int)b.take(5) + 257, hdist = (int)b.take(5) + 1, hclen = (int)b.take(4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  uint8_t cl[19] = {0};
  for (int i = 0; i < hclen; i++) cl[CL_ORDER[i]] = (uint8_t)b.take(3);
  Huff& clh = t.dist;                                  // scratch: rebuilt below
  if (!build_huff(clh, cl, 19, 0, 7)) return false;
  uint8_t lens[286 + 30 + 138];
  int i = 0;
  while (i < hlit + hdist) {
    if (b.over()) return false;
    const int s = decode_sym(clh, b);
    if (s < 0) return false;
    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
    int rep, val = 0;
    if (s == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + (int)b.take(2); }
    else if (s == 17) rep = 3 + (int)b.take(3);
    else rep = 11 + (int)b.take(7);
    if (i + rep > hlit + hdist) return false;
    while (rep--) lens[i++] = (uint8_t)val;
  }
  if (lens[256] == 0) return false;                    // no end-of-block code
  if (!build_huff(t.lit, lens, hlit, 1, FAST_BITS)) return false;
  if (!build_huff(t.dist, lens + hlit, hdist, 1, DIST_FAST_BITS)) return false;
  fold_tables(t);
  return true;
}

// One deflate block at b, symbols appended to out (nullptr = validate only).  Returns 0 ok, 1 ok and it was the final block, < 0 error.
template <class T>
int decode_block(Bits& b, OutBuf<T>* out, size_t* produced, Tables& scratch) {
  if (b.pos + 3 > b.nbits) return -1;
  const int final = (int)b.take(1), type = (int)b.take(2);
  if (type == 3) return -1;
  if (type == 0) {
    b.pos = (b.pos + 7) & ~7ull;
    if (b.pos + 32 > b.nbits) return -1;
    const uint32_t len = b.take(16), nlen = b.take(16);
    if ((len ^ nlen) != 0xFFFF) return -1;
    if (b.pos + 8ull * len > b.nbits) return -1;
    if (out) {
      if (!out->room(len)) return -2;
      const uint8_t* s = b.base + (b.pos >> 3);
      for (uint32_t i = 0; i < len; i++) out->p[out->n++] = (T)s[i];
    }
    *produced += len;
    b.pos += 8ull * len;
    return final;
  }
  const Tables* t = &scratch;
  if (type == 2) { if (!read_dynamic_header(b, scratch)) return -1; }
  else {
    // fixed code: the distance code has 30 of 32 codes; build once, permissively
    static const Tables* fx = [] {
      Tables* q = new Tables();
      uint8_t l[288];
      for (int i = 0; i < 144; i++) l[i] = 8;
      for (int i = 144; i < 256; i++) l[i] = 9;
      for (int i = 256; i < 280; i++) l[i] = 7;
      for (int i = 280; i < 288; i++) l[i] = 8;
      build_huff(q->lit, l, 288, 1);
      uint8_t d[32];
      for (int i = 0; i < 32; i++) d[i] = 5;
      build_huff(q->dist, d, 32, 1, DIST_FAST_BITS);    // codes 30, 31 decode but are rejected below
      fold_tables(*q);
      return q;
    }();
    t = fx;
  }
  if (out) {
    // Hot loop with everything in locals.  One 8-byte load gives >= 57 valid bits: enough for three literals, or for a whole match
    // (length code <= 11 + 5 extra + distance code <= 10 + 13 extra = 39 bits) when its codes are in the fast tables.
    const uint8_t* const base = b.base;
    const uint64_t nbits = b.nbits;
    uint64_t pos = b.pos;
    const uint32_t* const lit32 = t->lit32;
    const uint32_t* const dist32 = t->dist32;
    constexpr uint64_t FM = (1u << FAST_BITS) - 1, DM = (1u << DIST_FAST_BITS) - 1;
    constexpr int W = 8 / (int)sizeof(T);                  // symbols per 8-byte word
    while (true) {
      if (pos > nbits) return -1;
      if (out->n + 600 > out->cap && !out->room(600)) return -2;
      T* const p = out->p;
      size_t n = out->n;
      const size_t stop = out->cap - 300;
      int len = 0, dist = 0, s = -3;                       // s: -3 = a match decoded in the fast path, -2 = re-check room / input, >= 0 = slow-path symbol
      while (true) {
        uint64_t w;
        memcpy(&w, base + (pos >> 3), 8);
        w >>= (pos & 7);
        uint32_t e = lit32[w & FM];
        if ((e & 0x300) == 0x100) {                        // literal run
          p[n++] = (T)(e >> 16); pos += e & 15; w >>= (e & 15);
          e = lit32[w & FM];
          if ((e & 0x300) == 0x100) {
            p[n++] = (T)(e >> 16); pos += e & 15; w >>= (e & 15);
            e = lit32[w & FM];
            if ((e & 0x300) == 0x100) { p[n++] = (T)(e >> 16); pos += e & 15; }
          }
          if (n < stop && pos <= nbits) continue;
          s = -2;
          break;
        }
        if ((e & 0x300) == 0x200) {                        // match: length, then distance, from the same window
          uint32_t l = e & 15, x = (e >> 4) & 15;
          w >>= l;
          len = (int)(e >> 16) + (int)(w & ((1u << x) - 1));
          w >>= x; pos += l + x;
          const uint32_t de = dist32[w & DM];
          if (de) {
            l = de & 15; x = (de >> 4) & 15;
            w >>= l;
            dist = (int)(de >> 16) + (int)(w & ((1u << x) - 1));
            pos += l + x;
          } else {                                         // long or invalid distance code
            b.pos = pos;
            const int ds = decode_sym(t->dist, b);
            if (ds < 0 || ds > 29) return -1;
            dist = DIST_BASE[ds] + (int)b.take(DIST_EXTRA[ds]);
            pos = b.pos;
          }
          break;
        }
        if ((e & 0x300) == 0x300) { pos += e & 15; s = 256; break; }
        b.pos = pos;                                       // long (or invalid) literal / length code
        s = decode_sym(t->lit, b);
        pos = b.pos;
        break;
      }
      out->n = n;
      if (s == -2) continue;
      if (s != -3) {
        if (s < 0) return -1;
        if (s < 256) { p[out->n++] = (T)s; continue; }
        if (s == 256) { b.pos = pos; *produced = out->n; return pos > nbits ? -1 : final; }
        if (s > 285) return -1;
        b.pos = pos;
        len = LEN_BASE[s - 257] + (int)b.take(LEN_EXTRA[s - 257]);
        const int ds = decode_sym(t->dist, b);
        if (ds < 0 || ds > 29) return -1;
        dist = DIST_BASE[ds] + (int)b.take(DIST_EXTRA[ds]);
        pos = b.pos;
      }
      int i = 0;
      if (sizeof(T) == 1) { if ((size_t)dist > n) return -1; }        // byte mode: nothing exists before the start
      else for (; i < len && (int64_t)n - dist < 0; i++, n++) p[n] = (T)(MARK | (uint16_t)(32768 + ((int64_t)n - dist)));   // before the chunk: [-32768, -1]
      if (i < len) {
        const T* src = p + (n - (size_t)dist);
        T* dstp = p + n;
        const int rest = len - i;
        T* const end = dstp + rest;
        if (dist >= 4 * W) {                               // 32 bytes per step: most matches are one step; may write up to 4W - 1 symbols past the match
          do { memcpy(dstp, src, 16); memcpy(dstp + 2 * W, src + 2 * W, 16); dstp += 4 * W; src += 4 * W; } while (dstp < end);
        } else if (dist >= W) {
          do { memcpy(dstp, src, 8); dstp += W; src += W; } while (dstp < end);
        } else for (int j = 0; j < rest; j++) dstp[j] = src[j];          // short-distance overlapping run
        n += (size_t)rest;
      }
      out->n = n;
    }
  }
  while (true) {                                           // validation only (the finder): no output
    if (b.over()) return -1;
    const int s = decode_sym(t->lit, b);
    if (s < 0) return -1;
    if (s < 256) { ++*produced; continue; }
    if (s == 256) return b.over() ? -1 : final;
    if (s > 285) return -1;
    const int len = LEN_BASE[s - 257] + (int)b.take(LEN_EXTRA[s - 257]);
    const int ds = decode_sym(t->dist, b);
    if (ds < 0 || ds > 29) return -1;
    b.pos += DIST_EXTRA[ds];
    *produced += (size_t)len;
  }
}

// First bit offset in [from, to) at which a non-final dynamic block starts whose header, body and following block header are valid.
uint64_t find_block(const uint8_t* base, uint64_t nbits, uint64_t from, uint64_t to) {
  Tables* t = new Tables();
  uint64_t found = UINT64_MAX;
  for (uint64_t p = from; p < to && p + 64 < nbits; p++) {
    Bits b{base, p, nbits};
    if (b.peek(3) != 4) continue;                      // BFINAL = 0, BTYPE = 2 (bits: 0, then 0 1)
    const uint32_t hdr = b.peek(17);
    if (((hdr >> 3) & 31) > 29 || ((hdr >> 8) & 31) > 29) continue;
    b.pos += 3;
    if (!read_dynamic_header(b, *t)) continue;
    // the block must decode to its end, and another plausible block must follow
    Bits body{base, p, nbits};
    size_t produced = 0;
    const int r = decode_block<uint16_t>(body, nullptr, &produced, *t);
    if (r != 0 || produced < 64) continue;
    if (body.pos + 3 > nbits) continue;
    Bits nxt = body;
    size_t produced2 = 0;
    const int r2 = decode_block<uint16_t>(nxt, nullptr, &produced2, *t);
    if (r2 < 0) continue;
    found = p;
    break;
  }
  delete t;
  return found;
}

bool parse_gzip_header(const uint8_t* in, size_t n, size_t* data_off, bool* bgzf) {
  if (n < 18 || in[0] != 0x1f || in[1] != 0x8b || in[2] != 8) return false;
  const int flg = in[3];
  size_t p = 10;
  *bgzf = false;
  if (flg & 4) {
    if (p + 2 > n) return false;
    const size_t xlen = in[p] | (in[p + 1] << 8);
    for (size_t x = p + 2; x + 4 <= p + 2 + xlen && x + 4 <= n;) {
      const size_t sl = in[x + 2] | (in[x + 3] << 8);
      if (in[x] == 'B' && in[x + 1] == 'C' && sl == 2) *bgzf = true;
      x += 4 + sl;
    }
    p += 2 + xlen;
  }
  if (flg & 8) { while (p < n && in[p]) p++; p++; }
  if (flg & 16) { while (p < n && in[p]) p++; p++; }
  if (flg & 2) p += 2;
  if (p >= n) return false;
  *data_off = p;
  return true;
}

}  // namespace

// ---- streaming interface: one wave of chunks per call ----------------------------------------------------------------
struct PgzStream::Impl {
  const uint8_t* base = nullptr; uint64_t nbits = 0;
  uint32_t want_crc = 0, want_isize = 0;
  size_t chunk_bytes = 0, nchunks = 0, next_chunk = 0;   // nominal chunks; next_chunk = first nominal chunk not yet searched
  uint64_t cur_start = 0;                                // start (bits) of the first undecoded chunk
  bool started = false, finished = false, failed = false;
  std::vector<uint8_t> window = std::vector<uint8_t>(32768, 0);
  size_t have_window = 0;
  uint32_t crc = 0; uint64_t total = 0;
  std::vector<SymBuf> sym;
  double t_find = 0, t_dec = 0, t_res = 0; size_t chunks_done = 0;
};

namespace {
// symbol buffers survive the stream: their pages are faulted in once per process, not once per file
std::mutex g_pool_mu;
std::vector<SymBuf> g_pool;
}  // namespace

PgzStream::PgzStream() : d(new Impl()) {}
PgzStream::~PgzStream() {
  if (!d->sym.empty()) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& sb : d->sym) if (sb.p && g_pool.size() < 64) { g_pool.emplace_back(); std::swap(g_pool.back(), sb); }
  }
  delete d;
}

bool PgzStream::open(const unsigned char* in, size_t n, size_t chunk_bytes) {
  size_t data_off = 0;
  bool bgzf = false;
  if (!parse_gzip_header(in, n, &data_off, &bgzf) || bgzf || n < data_off + 8 + 2) return false;
  d->base = in + data_off;
  d->nbits = 8ull * (n - 8 - data_off);
  d->want_crc = in[n - 8] | (in[n - 7] << 8) | (in[n - 6] << 16) | ((uint32_t)in[n - 5] << 24);
  d->want_isize = in[n - 4] | (in[n - 3] << 8) | (in[n - 2] << 16) | ((uint32_t)in[n - 1] << 24);
  d->chunk_bytes = std::max<size_t>(chunk_bytes, 64u << 10);
  d->nchunks = std::max<size_t>(1, (size_t)((d->nbits / 8 + d->chunk_bytes - 1) / d->chunk_bytes));
  d->next_chunk = 1; d->cur_start = 0;
  d->crc = (uint32_t)crc32(0L, Z_NULL, 0);
  return true;
}

uint32_t PgzStream::isize_field() const { return d->want_isize; }
uint64_t PgzStream::produced() const { return d->total; }

// Decodes the next wave of chunks and appends its bytes at dst (room for `cap` bytes).  1 = more to come, 0 = stream finished and
// verified against the trailer, -1 = failed (nothing written by this stream may be trusted), -2 = cap too small.
int PgzStream::next(char* dst, size_t cap, size_t* wrote) {
  Impl& S = *d;
  *wrote = 0;
  if (S.failed) return -1;
  if (S.finished) return 0;
  const int T = std::max(1, omp_get_max_threads());
  // a short first wave so the consumer can start early, then two chunks per thread (fewer, larger waves decode faster than a tail
  // of shrinking ones: tried, 379 -> 426 ms per 2 M reads)
  size_t want = S.started ? (size_t)T * 2 : (size_t)T;
  // the caller's room bounds the wave: text rarely exceeds 8x its compressed size (FASTQ: 3-6x); a wave that does anyway is reported (-2)
  want = std::max<size_t>(1, std::min(want, cap / (S.chunk_bytes * 8)));
  S.started = true;
  const bool trace = getenv("CATT_TRACE") != nullptr;
  double tt = omp_get_wtime();
  // starts of this wave's chunks: cur_start, then the first block found in each following nominal chunk
  std::vector<uint64_t> s{S.cur_start};
  {
    const size_t k0 = S.next_chunk, k1 = std::min(S.nchunks, k0 + want);
    std::vector<uint64_t> f(k1 > k0 ? k1 - k0 : 0, UINT64_MAX);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = (int64_t)k0; k < (int64_t)k1; k++)
      f[k - k0] = find_block(S.base, S.nbits, std::max<uint64_t>(8ull * k * S.chunk_bytes, S.cur_start + 1), 8ull * (k + 1) * S.chunk_bytes);
    for (uint64_t x : f) if (x != UINT64_MAX) s.push_back(x);
    S.next_chunk = k1;
    if (k1 == S.nchunks) s.push_back(S.nbits);           // the last wave runs to the end of the deflate data
  }
  const bool last_wave = S.next_chunk == S.nchunks;
  // chunks of this wave: [s[i], s[i+1]); when it is not the last wave the final entry of s is the next wave's first start
  const size_t m = s.size() - 1;
  S.t_find += omp_get_wtime() - tt; tt = omp_get_wtime();
  if (m == 0) return 1;                                  // no block start found in this stretch: the next wave continues from cur_start
  if (S.sym.size() < m) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    while (S.sym.size() < m) {
      S.sym.emplace_back();
      if (!g_pool.empty()) { std::swap(S.sym.back(), g_pool.back()); g_pool.pop_back(); }
    }
  }
  std::vector<char> good(m, 1), fin(m, 0);
  #pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < (int64_t)m; i++) {
    SymBuf& sb = S.sym[i];
    sb.n = 0;
    if (!sb.room(S.chunk_bytes * 5)) { good[i] = 0; continue; }
    Tables* t = new Tables();
    Bits b{S.base, s[i], S.nbits};
    size_t produced = 0;
    const uint64_t target = s[i + 1];
    const bool last_chunk = last_wave && (size_t)i + 1 == m;
    bool final_seen = false;
    while (b.pos < target) {
      const int r = decode_block(b, &sb, &produced, *t);
      if (r < 0) { good[i] = 0; break; }
      if (r == 1) { final_seen = true; break; }
    }
    delete t;
    if (!good[i]) continue;
    if (final_seen) {
      // only the last chunk may end the stream, exactly at the trailer (after byte alignment)
      if (!last_chunk || ((b.pos + 7) & ~7ull) != S.nbits) good[i] = 0;
      fin[i] = 1;
    } else if (b.pos != target || last_chunk) good[i] = 0;
  }
  bool ok = true, saw_final = false;
  for (size_t i = 0; i < m; i++) { if (!good[i]) ok = false; if (fin[i]) saw_final = true; }
  S.t_dec += omp_get_wtime() - tt; tt = omp_get_wtime();
  if (!ok) { S.failed = true; return -1; }
  // windows: sequential over the wave's chunks; win[i] = window BEFORE chunk i
  std::vector<std::vector<uint8_t>> win(m);
  std::vector<size_t> avail(m), off(m + 1, 0);
  for (size_t i = 0; i < m; i++) {
    win[i] = S.window; avail[i] = S.have_window;
    const SymBuf& sb = S.sym[i];
    off[i + 1] = off[i] + sb.n;
    std::vector<uint8_t> nw(32768);                      // new window = last 32 KiB of (window + resolved chunk)
    const size_t take = std::min<size_t>(sb.n, 32768);
    for (size_t j = 0; j < take; j++) {
      const uint16_t v = sb.p[sb.n - take + j];
      if (v & MARK) { if ((size_t)(32768 - (v & 0x7FFF)) > S.have_window) ok = false; nw[32768 - take + j] = S.window[v & 0x7FFF]; }
      else nw[32768 - take + j] = (uint8_t)v;
    }
    if (take < 32768) memcpy(nw.data(), S.window.data() + take, 32768 - take);
    S.window.swap(nw);
    S.have_window = std::min<size_t>(32768, S.have_window + sb.n);
  }
  if (!ok) { S.failed = true; return -1; }
  if (off[m] > cap) { S.failed = true; return -2; }
  std::vector<uint32_t> ccrc(m);
  bool bad_marker = false;
  #pragma omp parallel for schedule(dynamic, 1) reduction(|| : bad_marker)
  for (int64_t i = 0; i < (int64_t)m; i++) {
    const SymBuf& sb = S.sym[i];
    const uint8_t* wv = win[i].data();
    char* o = dst + off[i];
    const size_t av = avail[i];
    size_t j = 0;
    if (av < 32768) {                                      // the first 32 KiB of the stream: markers must stay inside what exists
      for (; j < sb.n; j++) {
        const uint16_t v = sb.p[j];
        if (v & MARK) { if ((size_t)(32768 - (v & 0x7FFF)) > av) bad_marker = true; o[j] = (char)wv[v & 0x7FFF]; }
        else o[j] = (char)v;
      }
    } else {
      // one table for both kinds of symbol (bytes map to themselves, markers to the window): no branch per symbol; groups of 16
      // symbols without a marker are packed with SSE2
      std::vector<uint8_t> lut(65536);
      for (int q = 0; q < 256; q++) lut[q] = (uint8_t)q;
      memcpy(lut.data() + MARK, wv, 32768);
#ifdef __SSE2__
      for (; j + 16 <= sb.n; j += 16) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(sb.p + j)), b2 = _mm_loadu_si128((const __m128i*)(sb.p + j + 8));
        if (_mm_movemask_epi8(_mm_or_si128(a, b2)) & 0xAAAA) {
          for (size_t q = j; q < j + 16; q++) o[q] = (char)lut[sb.p[q]];
        } else {
          _mm_storeu_si128((__m128i*)(o + j), _mm_packus_epi16(a, b2));
        }
      }
#endif
      for (; j < sb.n; j++) o[j] = (char)lut[sb.p[j]];
    }
    uint32_t c = (uint32_t)crc32(0L, Z_NULL, 0);          // zlib's crc32 takes uInt lengths: feed in slices
    for (size_t q = 0; q < sb.n; q += (1u << 30)) c = (uint32_t)crc32(c, (const Bytef*)o + q, (uInt)std::min<size_t>(sb.n - q, 1u << 30));
    ccrc[i] = c;
  }
  if (bad_marker) { S.failed = true; return -1; }
  for (size_t i = 0; i < m; i++) { S.crc = (uint32_t)crc32_combine(S.crc, ccrc[i], (z_off_t)S.sym[i].n); S.total += S.sym[i].n; }
  S.chunks_done += m;
  S.cur_start = s[m];
  *wrote = off[m];
  S.t_res += omp_get_wtime() - tt;
  if (!last_wave) return 1;
  if (trace) fprintf(stderr, "[catt trace] pgz: %zu chunks, find %.1f ms, decode %.1f ms, resolve + crc %.1f ms\n", S.chunks_done, S.t_find * 1e3,
                     S.t_dec * 1e3, S.t_res * 1e3);
  if (!saw_final || S.crc != S.want_crc || (uint32_t)(S.total & 0xFFFFFFFFull) != S.want_isize) { S.failed = true; return -1; }
  S.finished = true;
  return 0;
}

// true: `out` holds the inflated bytes (appended); false: not handled (not gzip, BGZF, multi-member, or any check failed).
// `in` must have 64 readable bytes behind in[n - 1] (the bit reader loads 8 bytes at a time and checks its position afterwards).
bool pgz_inflate(const unsigned char* in, size_t n, std::string* out, size_t chunk_bytes) {
  PgzStream st;
  if (!st.open(in, n, chunk_bytes)) return false;
  const size_t start_size = out->size();
  if (st.isize_field() >= n) {                          // exact for members below 4 GiB; only a hint otherwise
    out->reserve(start_size + st.isize_field());
    advise_huge(out->data() + start_size, st.isize_field());
  }
  std::string wave_buf;
  while (true) {
    // a wave's output size is not known before it is decoded: give it generous room, then shrink to what it wrote
    const size_t o0 = out->size();
    size_t room = std::max<size_t>(64u << 20, (size_t)omp_get_max_threads() * 2 * chunk_bytes * 8);
    int r; size_t wrote = 0;
    while (true) {
      out->resize(o0 + room);
      r = st.next(&(*out)[o0], room, &wrote);
      if (r != -2) break;
      out->resize(start_size);
      return false;                                      // (a wave larger than 8x its compressed size x2: leave it to zlib)
    }
    out->resize(o0 + wrote);
    if (r < 0) { out->resize(start_size); return false; }
    if (r == 0) return true;
  }
}

}  // namespace catt

namespace catt {

bool read_raw_file(const char* path, std::vector<unsigned char>* buf, size_t* size) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return false;
  struct stat stt;
  if (fstat(fd, &stt) != 0 || stt.st_size < 0) { close(fd); return false; }
  const size_t sz = (size_t)stt.st_size;
  buf->clear();
  buf->reserve(sz + 64);
  buf->resize(sz + 64);                                 // 64 zero bytes of slack for the bit readers (the rest is overwritten below)
  // parallel pread: the page-cache copy of a few hundred MB is worth spreading over the threads
  const size_t sub = 8u << 20;
  const int64_t ns = (int64_t)((sz + sub - 1) / sub);
  bool ok = true;
  #pragma omp parallel for schedule(dynamic, 1) reduction(&& : ok)
  for (int64_t k = 0; k < ns; k++) {
    size_t o = (size_t)k * sub;
    const size_t e = std::min(sz, o + sub);
    while (o < e) {
      const ssize_t r = pread(fd, buf->data() + o, e - o, (off_t)o);
      if (r <= 0) { ok = false; break; }
      o += (size_t)r;
    }
  }
  close(fd);
  *size = sz;
  return ok;
}

bool RawFile::open(const char* path) {
  close();
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) return false;
  struct stat stt;
  if (fstat(fd, &stt) != 0 || stt.st_size < 0) { ::close(fd); return false; }
  const size_t sz = (size_t)stt.st_size, page = 4096, tail = sz % page;
  if (sz > 0 && tail != 0 && tail + 64 <= page) {           // the slack behind the last byte lies inside the last mapped page (zeros)
    void* m = mmap(nullptr, sz, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (m != MAP_FAILED) {
      ::close(fd);
      madvise(m, sz, MADV_SEQUENTIAL);
      p_ = (const unsigned char*)m; n_ = sz; mapped_ = sz;
      return true;
    }
  }
  ::close(fd);
  if (!read_raw_file(path, &buf_, &n_)) return false;
  p_ = buf_.data();
  return true;
}

void RawFile::close() {
  if (mapped_) munmap(const_cast<unsigned char*>(p_), mapped_);
  p_ = nullptr; n_ = 0; mapped_ = 0;
  std::vector<unsigned char>().swap(buf_);
}

namespace {
inline uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint32_t le16(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
}  // namespace

// BGZF (SAMv1 4.1): a series of gzip members, each with a 'BC' extra subfield holding the member's total size - 1; the members
// inflate independently, on all host threads.
bool BgzfStream::open(const unsigned char* in, size_t n) {
  in_ = in; blocks_.clear(); next_ = 0; total_ = 0;
  size_t p = 0;
  while (p < n) {
    if (n - p < 18 || in[p] != 0x1f || in[p + 1] != 0x8b || in[p + 2] != 8 || !(in[p + 3] & 4)) return false;
    const size_t xlen = le16(&in[p + 10]);
    if (p + 12 + xlen > n) return false;
    size_t bsize = 0;
    for (size_t x = p + 12; x + 4 <= p + 12 + xlen;) {
      const size_t slen = le16(&in[x + 2]);
      if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2 && x + 6 <= p + 12 + xlen) bsize = le16(&in[x + 4]) + 1;
      x += 4 + slen;
    }
    if (!bsize || bsize < xlen + 20 || p + bsize > n) return false;
    Block b;
    b.cdata = p + 12 + xlen; b.clen = bsize - xlen - 20;
    b.crc = le32(&in[p + bsize - 8]); b.ulen = le32(&in[p + bsize - 4]);
    b.uoff = total_;
    if (b.ulen > 65536) return false;
    total_ += b.ulen;
    blocks_.push_back(b);
    p += bsize;
  }
  return !blocks_.empty();
}

int BgzfStream::next(char* dst, size_t cap, size_t* wrote) {
  *wrote = 0;
  if (next_ >= blocks_.size()) return 0;
  const size_t group = (size_t)std::max(1, omp_get_max_threads()) * 48;         // ~3 MiB of text per thread and call
  const size_t b0 = next_;
  size_t b1 = std::min(blocks_.size(), b0 + group);
  const uint64_t o0 = blocks_[b0].uoff;
  auto end_of = [&](size_t b) { return b < blocks_.size() ? blocks_[b].uoff : total_; };
  // the group follows the room the caller has (a small pinned window on a host with many threads): as many whole members as fit
  while (b1 > b0 + 1 && end_of(b1) - o0 > cap) b1 = b0 + (b1 - b0) / 2;
  const uint64_t o1 = end_of(b1);
  if (o1 - o0 > cap) return -2;                                                 // not even one 64 KiB member fits
  bool bad = false;
  const bool zlib_only = getenv("CATT_ZLIB_ONLY") != nullptr;
  #pragma omp parallel for schedule(dynamic, 4) reduction(|| : bad)
  for (int64_t i = (int64_t)b0; i < (int64_t)b1; i++) {
    const Block& b = blocks_[i];
    if (b.ulen == 0) continue;
    unsigned char* o = reinterpret_cast<unsigned char*>(dst) + (b.uoff - o0);
    // the member's deflate data through the byte-mode decoder above (about twice zlib's inflate on the GPU boxes' hosts); the
    // member's CRC-32 and size are the check, and zlib gets the member if anything is off
    static thread_local OutBuf<uint8_t> scratch;
    static thread_local std::unique_ptr<Tables> tables(new Tables());
    bool done = false;
    scratch.n = 0;
    if (!zlib_only && scratch.room(65536 + 1024)) {
      // the decoder may read a few bytes past the member's deflate data: the member's 8-byte trailer and the next header are there
      Bits bits{in_ + b.cdata, 0, 8ull * b.clen};
      size_t produced = 0;
      int r = 0;
      while (r == 0 && scratch.n <= 65536) r = decode_block<uint8_t>(bits, &scratch, &produced, *tables);
      if (r == 1 && scratch.n == b.ulen && crc32(crc32(0L, Z_NULL, 0), scratch.p, b.ulen) == b.crc) { memcpy(o, scratch.p, b.ulen); done = true; }
    }
    if (done) continue;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) { bad = true; continue; }
    zs.next_in = const_cast<unsigned char*>(&in_[b.cdata]); zs.avail_in = (uInt)b.clen;
    zs.next_out = o; zs.avail_out = b.ulen;
    const int r = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (r != Z_STREAM_END || zs.avail_out != 0 || crc32(crc32(0L, Z_NULL, 0), o, b.ulen) != b.crc) bad = true;
  }
  if (bad) return -1;
  next_ = b1;
  *wrote = (size_t)(o1 - o0);
  return b1 < blocks_.size() && o1 < total_ ? 1 : 0;       // trailing empty members (the EOF marker) hold nothing to check
}

// 1 = inflated and appended to out, 0 = not BGZF, < 0 = error (set_error called)
int bgzf_inflate(const unsigned char* in, size_t n, std::string* out, const char* path) {
  if (n < 18 || in[0] != 0x1f || in[1] != 0x8b || !(in[3] & 4)) return 0;
  BgzfStream st;
  if (!st.open(in, n)) {
    // distinguish "an ordinary gzip file" from "BGZF with a damaged block chain": the first member decides
    const size_t xlen = le16(&in[10]);
    bool first_is_bgzf = false;
    for (size_t x = 12; x + 4 <= 12 + xlen && x + 4 <= n;) {
      const size_t slen = le16(&in[x + 2]);
      if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2) first_is_bgzf = true;
      x += 4 + slen;
    }
    if (!first_is_bgzf) return 0;
    set_error("%s: corrupt or truncated BGZF block chain", path);
    return CATT_E_IO;
  }
  const size_t o0 = out->size();
  out->resize(o0 + st.total());
  size_t have = 0;
  while (true) {
    size_t wrote = 0;
    const int r = st.next(&(*out)[o0 + have], (size_t)st.total() - have, &wrote);
    have += wrote;
    if (r < 0) { out->resize(o0); set_error("%s: a BGZF block does not inflate / fails its CRC", path); return CATT_E_IO; }
    if (r == 0) break;
  }
  return 1;
}

// Whole file into memory: plain bytes, BGZF (parallel), single-member gzip (parallel, pgz_inflate) or anything else zlib reads.
bool slurp_file(const std::string& path, std::string& out, int* method) {
  int dummy;
  if (!method) method = &dummy;
  *method = 0;
  out.clear();
  std::vector<unsigned char> raw;
  size_t n = 0;
  if (!read_raw_file(path.c_str(), &raw, &n)) return false;
  if (n < 2 || raw[0] != 0x1f || raw[1] != 0x8b) { out.assign(reinterpret_cast<const char*>(raw.data()), n); return true; }
  if (!getenv("CATT_ZLIB_ONLY")) {
    const int r = bgzf_inflate(raw.data(), n, &out, path.c_str());
    if (r == 1) { *method = 1; return true; }
    out.clear();
    const char* cb = getenv("CATT_PGZ_CHUNK");                 // tests: small chunks on small files
    const size_t chunk = cb ? (size_t)atol(cb) : (1u << 20);
    if (r == 0 && (cb || n >= (1u << 20)) && pgz_inflate(raw.data(), n, &out, chunk)) { *method = 2; return true; }
    out.clear();
  }
  std::vector<unsigned char>().swap(raw);
  *method = 3;
  gzFile f = gzopen(path.c_str(), "rb");                 // multi-member, tiny or unusual gzip files: one zlib stream
  if (!f) return false;
  gzbuffer(f, 1u << 20);
  std::vector<char> buf(1u << 20);
  int got;
  while ((got = gzread(f, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), (size_t)got);
  const bool ok = got == 0;
  gzclose(f);
  return ok;
}

}  // namespace catt
