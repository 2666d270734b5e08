// gz.h — compressed input on the host threads (pgz.cu): BGZF members and single-member gzip streams inflate in parallel.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace catt {
// the file's bytes plus 64 zero bytes of slack; *size = the file's size
bool read_raw_file(const char* path, std::vector<unsigned char>* buf, size_t* size);
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
// plain / BGZF / gzip file into memory; *method = 0 plain bytes, 1 BGZF members in parallel, 2 parallel gzip, 3 one zlib stream
bool slurp_file(const std::string& path, std::string& out, int* method = nullptr);
}  // namespace catt
