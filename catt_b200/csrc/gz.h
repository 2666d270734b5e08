// gz.h — compressed input on the host threads (pgz.cu): BGZF members and single-member gzip streams inflate in parallel.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace catt {
// the file's bytes plus 64 zero bytes of slack; *size = the file's size
bool read_raw_file(const char* path, std::vector<unsigned char>* buf, size_t* size);
// The same without the copy: the file mapped from the page cache (read-only); falls back to read_raw_file when the 64 bytes behind
// the last byte would not be mapped.  data()[0 .. size() + 64) is readable.
class RawFile {
 public:
  RawFile() = default;
  RawFile(const RawFile&) = delete;
  RawFile& operator=(const RawFile&) = delete;
  ~RawFile() { close(); }
  bool open(const char* path);
  void close();
  const unsigned char* data() const { return p_; }
  size_t size() const { return n_; }
 private:
  const unsigned char* p_ = nullptr;
  size_t n_ = 0, mapped_ = 0;
  std::vector<unsigned char> buf_;
};
// 1 = inflated and appended to out, 0 = not BGZF, < 0 = error (set_error called)
int bgzf_inflate(const unsigned char* in, size_t n, std::string* out, const char* path);
// single-member gzip, speculative parallel decode verified by the trailer's CRC-32; false = not handled (caller uses zlib).
// `in` needs 64 readable bytes behind in[n - 1].
bool pgz_inflate(const unsigned char* in, size_t n, std::string* out, size_t chunk_bytes);
// The same decoder wave by wave, so a consumer can start on the first bytes while the rest inflates.
class PgzStream {
 public:
  PgzStream();
  ~PgzStream();
  PgzStream(const PgzStream&) = delete;
  PgzStream& operator=(const PgzStream&) = delete;
  bool open(const unsigned char* in, size_t n, size_t chunk_bytes);       // false = not a single-member, non-BGZF gzip file
  uint32_t isize_field() const;                                            // ISIZE of the trailer (length mod 2^32)
  uint64_t produced() const;
  // next wave appended at dst (room for cap bytes): 1 = more to come, 0 = finished and verified (CRC-32, ISIZE),
  // -1 = failed: nothing this stream wrote may be trusted, -2 = cap too small
  int next(char* dst, size_t cap, size_t* wrote);
  struct Impl;
 private:
  Impl* d;
};
// BGZF members (bgzip, BAM) group by group; the total size is known from the block trailers before anything is inflated.
class BgzfStream {
 public:
  bool open(const unsigned char* in, size_t n);                            // false = not BGZF / damaged block chain
  uint64_t total() const { return total_; }
  int next(char* dst, size_t cap, size_t* wrote);                          // 1 more, 0 done, -1 a block fails to inflate / its CRC, -2 cap
 private:
  struct Block { size_t cdata, clen, uoff; uint32_t ulen, crc; };
  const unsigned char* in_ = nullptr;
  std::vector<Block> blocks_;
  size_t next_ = 0;
  uint64_t total_ = 0;
};
// A BGZF-compressed BAM as the FASTQ text `samtools fastq -N` would write (bam.cu), group by group: the file pipeline indexes and
// packs that text on the GPU like any FASTQ, so the conversion overlaps the alignment.  text_bound() >= the bytes it will write.
class BamTextStream {
 public:
  bool open(const unsigned char* in, size_t n);                            // false = not a BGZF file
  uint64_t text_bound() const { return bz_.total() / 3 * 4 + 4096; }       // 2L+ln+8 FASTQ bytes <= 4/3 of (1.5L+ln+37) BAM bytes
  int next(char* dst, size_t cap, size_t* wrote);                          // 1 more, 0 done, -1 malformed, -2 cap
 private:
  BgzfStream bz_;
  std::vector<unsigned char> buf_;                                         // inflated bytes not consumed yet (a partial record)
  size_t pending_ = 0;
  bool header_done_ = false, eof_ = false;
  double t_inf_ = 0, t_scan_ = 0, t_conv_ = 0;                             // CATT_TRACE laps
};
// plain / BGZF / gzip file into memory; *method = 0 plain bytes, 1 BGZF members in parallel, 2 parallel gzip, 3 one zlib stream
bool slurp_file(const std::string& path, std::string& out, int* method = nullptr);
}  // namespace catt
