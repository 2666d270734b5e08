// bam.cu — `--bam` input (catt.jl:756-764).  The reference converts the file with `samtools fastq -N -@ t in.bam > tmp.fastq`
// and feeds that FASTQ to the same path as any single-end sample; this file does the conversion in memory, so the reads never
// become FASTQ text.  samtools is a third-party program that is absent from /root/reference and from this image ("samtools >= 1.8",
// README.md:141): what is restated here is its documented behaviour and the SAM/BAM specification (SAMv1 4.2, BGZF 4.1):
//   * records with the SECONDARY (0x100) or SUPPLEMENTARY (0x800) flag are dropped (`samtools fastq` defaults to -F 0x900);
//     everything else - unmapped reads included - is written, in file order, to the one output stream;
//   * `-N`: "/1" is appended to the name of a READ1-only record and "/2" to a READ2-only record (bwa strips the suffix again,
//     which is why mates collapse in the QNAME dedup);
//   * a record with the REVERSE flag (0x10) is written in its original orientation: sequence reverse-complemented (IUPAC-aware),
//     quality reversed;
//   * a missing quality string (0xff.. in BAM, `*` in SAM) becomes 'B' for every base;
//   * bases come out of the 4-bit alphabet "=ACMGRSVTWYHKDBN" (SAM text is folded through the same table, so case is lost).
// Parity unpinned: there is no samtools here to generate a golden; tests/test_bam_input.py writes BAM/SAM files with its own
// encoder and compares with the FASTQ path on the equivalent FASTQ.  Host code (zlib inflate of the BGZF blocks on all host
// threads, then a parallel decode into the read arenas); the hot path starts at run_all().
#include <ctype.h>
#include <omp.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "gz.h"
#include "host_types.h"
#include "internal.cuh"

namespace catt {
namespace {

const char NT16[] = "=ACMGRSVTWYHKDBN";
const unsigned char COMP16[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};

inline uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint32_t le16(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

struct Rec { size_t name, seq, qual; uint32_t lname, lseq; uint16_t flag; };       // offsets into the decoded bytes

inline int suffix_of(uint16_t flag) { const bool r1 = flag & 0x40, r2 = flag & 0x80; return r1 && !r2 ? 1 : (r2 && !r1 ? 2 : 0); }

// Second pass shared by BAM and SAM: sizes -> offsets -> parallel fill of the read arenas (appends to what hr already holds).
template <class SeqAt, class QualAt>
void emit_reads(const unsigned char* base, const std::vector<Rec>& recs, HostReads* hr, SeqAt seq_code, QualAt qual_char) {
  const size_t n = recs.size();
  if (hr->own_name_off.empty()) { hr->own_name_off.push_back(0); hr->own_seq_off.push_back(0); }
  const size_t n0 = hr->own_seq_off.size() - 1;
  hr->own_name_off.resize(n0 + n + 1); hr->own_seq_off.resize(n0 + n + 1);
  int64_t no = hr->own_name_off[n0], so = hr->own_seq_off[n0];
  for (size_t i = 0; i < n; i++) {
    no += recs[i].lname + (suffix_of(recs[i].flag) ? 2 : 0); so += recs[i].lseq;
    hr->own_name_off[n0 + i + 1] = no; hr->own_seq_off[n0 + i + 1] = so;
  }
  hr->own_names.resize((size_t)no); hr->own_seqs.resize((size_t)so); hr->own_quals.resize((size_t)so);
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++) {
    const Rec& r = recs[i];
    char* nm = &hr->own_names[(size_t)hr->own_name_off[n0 + i]];
    memcpy(nm, base + r.name, r.lname);
    if (const int sfx = suffix_of(r.flag)) { nm[r.lname] = '/'; nm[r.lname + 1] = (char)('0' + sfx); }
    char* s = &hr->own_seqs[(size_t)hr->own_seq_off[n0 + i]];
    char* q = &hr->own_quals[(size_t)hr->own_seq_off[n0 + i]];
    const bool rev = r.flag & 0x10;
    for (uint32_t k = 0; k < r.lseq; k++) {
      const uint32_t src = rev ? r.lseq - 1 - k : k;
      const unsigned code = seq_code(base, r, src);
      s[k] = NT16[rev ? COMP16[code] : code];
      q[k] = qual_char(base, r, src);
    }
  }
  hr->n = (int64_t)(n0 + n);
  hr->names = hr->own_names.data(); hr->name_off = hr->own_name_off.data();
  hr->seqs = hr->own_seqs.data(); hr->seq_off = hr->own_seq_off.data();
  hr->quals = hr->own_quals.data();
}

int parse_bam_bytes(const unsigned char* d, size_t N, const char* path, HostReads* hr) {
  if (N < 12 || memcmp(d, "BAM\1", 4) != 0) { set_error("%s: not a BAM file (bad magic)", path); return CATT_E_IO; }
  size_t p = 4;
  const size_t l_text = le32(&d[p]); p += 4;
  if (p + l_text + 4 > N) { set_error("%s: truncated BAM header", path); return CATT_E_IO; }
  p += l_text;
  const uint32_t n_ref = le32(&d[p]); p += 4;
  for (uint32_t i = 0; i < n_ref; i++) {
    if (p + 4 > N) { set_error("%s: truncated BAM reference list", path); return CATT_E_IO; }
    const size_t l_name = le32(&d[p]);
    p += 4 + l_name + 4;
    if (p > N) { set_error("%s: truncated BAM reference list", path); return CATT_E_IO; }
  }
  std::vector<Rec> recs;
  while (p < N) {
    if (p + 4 > N) { set_error("%s: truncated BAM record at byte %zu", path, p); return CATT_E_IO; }
    const size_t bs = le32(&d[p]);
    if (bs < 32 || p + 4 + bs > N) { set_error("%s: truncated BAM record at byte %zu", path, p); return CATT_E_IO; }
    const unsigned char* r = &d[p + 4];
    const uint32_t l_read_name = r[8], n_cigar = le16(r + 12), flag = le16(r + 14), l_seq = le32(r + 16);
    const size_t need = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + l_seq;
    if (need > bs || l_read_name == 0) { set_error("%s: inconsistent BAM record at byte %zu", path, p); return CATT_E_IO; }
    if (!(flag & 0x900)) {
      Rec x;
      x.name = p + 4 + 32; x.lname = l_read_name - 1;                      // stored NUL-terminated
      x.seq = x.name + l_read_name + 4 * (size_t)n_cigar; x.qual = x.seq + ((size_t)l_seq + 1) / 2;
      x.lseq = l_seq; x.flag = (uint16_t)flag;
      recs.push_back(x);
    }
    p += 4 + bs;
  }
  emit_reads(d, recs, hr,
             [](const unsigned char* b, const Rec& r, uint32_t i) -> unsigned { const unsigned char c = b[r.seq + (i >> 1)]; return (i & 1) ? (c & 15) : (c >> 4); },
             [](const unsigned char* b, const Rec& r, uint32_t i) -> char { return b[r.qual] == 0xff ? 'B' : (char)(b[r.qual + i] + 33); });
  return CATT_OK;
}

int parse_sam_bytes(const unsigned char* d, size_t N, const char* path, HostReads* hr) {
  unsigned char code_of[256];
  memset(code_of, 15, sizeof code_of);
  for (int i = 0; i < 16; i++) { code_of[(unsigned char)NT16[i]] = (unsigned char)i; code_of[(unsigned char)tolower(NT16[i])] = (unsigned char)i; }
  std::vector<Rec> recs;
  size_t p = 0;
  int64_t line_no = 0;
  while (p < N) {
    const unsigned char* nl = (const unsigned char*)memchr(&d[p], '\n', N - p);
    size_t e = nl ? (size_t)(nl - d) : N;
    const size_t next = e + 1;
    if (e > p && d[e - 1] == '\r') e--;
    line_no++;
    if (e > p && d[p] != '@') {
      size_t f[12]; int nf = 0;
      f[nf++] = p;
      for (size_t k = p; k < e && nf < 12; k++) if (d[k] == '\t') f[nf++] = k + 1;
      if (nf < 11) { set_error("%s: SAM line %lld has fewer than 11 fields", path, (long long)line_no); return CATT_E_IO; }
      const size_t qual_end = nf > 11 ? f[11] - 1 : e;
      long flag = 0;
      for (size_t k = f[1]; k + 1 < f[2]; k++) { if (d[k] < '0' || d[k] > '9') { flag = -1; break; } flag = flag * 10 + (d[k] - '0'); }
      if (flag < 0 || flag > 65535 || f[2] - f[1] < 2) { set_error("%s: SAM line %lld: bad FLAG", path, (long long)line_no); return CATT_E_IO; }
      if (!(flag & 0x900)) {
        Rec x;
        x.name = f[0]; x.lname = (uint32_t)(f[1] - 1 - f[0]);
        x.seq = f[9]; x.lseq = (uint32_t)(f[10] - 1 - f[9]); x.qual = f[10]; x.flag = (uint16_t)flag;
        if (x.lseq == 1 && d[x.seq] == '*') x.lseq = 0;
        const size_t lq = qual_end - f[10];
        const bool noq = lq == 1 && d[f[10]] == '*';
        if (!noq && lq != x.lseq) { set_error("%s: SAM line %lld: SEQ and QUAL differ in length", path, (long long)line_no); return CATT_E_IO; }
        if (noq) x.qual = (size_t)-1;
        if (x.lname == 0) { set_error("%s: SAM line %lld: empty QNAME", path, (long long)line_no); return CATT_E_IO; }
        recs.push_back(x);
      }
    }
    p = next;
  }
  emit_reads(d, recs, hr,
             [&](const unsigned char* b, const Rec& r, uint32_t i) -> unsigned { return code_of[b[r.seq + i]]; },
             [](const unsigned char* b, const Rec& r, uint32_t i) -> char { return r.qual == (size_t)-1 ? 'B' : (char)b[r.qual + i]; });
  return CATT_OK;
}

}  // namespace

}  // namespace catt (reopened below)

namespace catt {

bool BamTextStream::open(const unsigned char* in, size_t n) {
  buf_.clear(); pending_ = 0; header_done_ = false; eof_ = false; t_inf_ = t_scan_ = t_conv_ = 0;
  return bz_.open(in, n) && bz_.total() >= 12;
}

int BamTextStream::next(char* dst, size_t cap, size_t* wrote) {
  *wrote = 0;
  if (eof_) return 0;
  // inflate the next group of blocks behind the bytes left over from the previous call
  // ... but never more BAM bytes than the caller has room for as text: a record's FASTQ text is < 4/3 of its BAM bytes (2 x l_seq
  // + name + 8 against 1.5 x l_seq + name + 36), so 3/4 of `cap` is safe, and it is checked BEFORE any block is consumed
  const size_t group_bound = (size_t)std::max(1, omp_get_max_threads()) * 48 * 65536;
  const size_t room = cap / 4 * 3;
  if (room < pending_ + 65536 + 1024) return -2;
  const size_t want = std::min(group_bound, room - pending_);
  if (buf_.size() < pending_ + want) buf_.resize(pending_ + want);
  size_t got = 0;
  double tt = omp_get_wtime();
  const int r = bz_.next(reinterpret_cast<char*>(buf_.data()) + pending_, want, &got);
  t_inf_ += omp_get_wtime() - tt; tt = omp_get_wtime();
  if (r < 0) return -1;
  const bool last = r == 0;
  const size_t N = pending_ + got;
  const unsigned char* d = buf_.data();
  size_t p = 0;
  if (!header_done_) {
    if (N < 12) { if (last) return -1; pending_ = N; return 1; }
    if (memcmp(d, "BAM\1", 4) != 0) return -1;
    size_t q = 8 + (size_t)le32(d + 4);
    bool complete = q + 4 <= N;
    if (complete) {
      const uint32_t n_ref = le32(d + q);
      q += 4;
      for (uint32_t i = 0; i < n_ref && complete; i++) {
        if (q + 4 > N) { complete = false; break; }
        q += 4 + (size_t)le32(d + q) + 4;
        if (q > N) complete = false;
      }
    }
    if (!complete) { if (last) return -1; pending_ = N; return 1; }     // a header longer than one group: wait for more
    header_done_ = true;
    p = q;
  }
  // complete records of this stretch
  std::vector<Rec> recs;
  std::vector<size_t> off{0};
  recs.reserve(N / 200 + 16); off.reserve(N / 200 + 17);
  while (p + 4 <= N) {
    const size_t bs = le32(d + p);
    if (bs < 32) return -1;
    if (p + 4 + bs > N) break;
    // the chain of block_size fields is serial and every header is a cache miss (the bytes were just written by other cores):
    // records of one run have similar sizes, so the headers a few records ahead can be fetched early
    __builtin_prefetch(d + std::min(N - 1, p + 8 * (bs + 4)));
    __builtin_prefetch(d + std::min(N - 1, p + 8 * (bs + 4) + 64));
    const unsigned char* rr = d + p + 4;
    const uint32_t l_read_name = rr[8], n_cigar = le16(rr + 12), flag = le16(rr + 14), l_seq = le32(rr + 16);
    const size_t need = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + l_seq;
    if (need > bs || l_read_name == 0) return -1;
    if (!(flag & 0x900)) {
      Rec x;
      x.name = p + 4 + 32; x.lname = l_read_name - 1;
      x.seq = x.name + l_read_name + 4 * (size_t)n_cigar; x.qual = x.seq + ((size_t)l_seq + 1) / 2;
      x.lseq = l_seq; x.flag = (uint16_t)flag;
      recs.push_back(x);
      off.push_back(off.back() + 1 + x.lname + (suffix_of(x.flag) ? 2 : 0) + 1 + 2 * (size_t)l_seq + 4);
    }
    p += 4 + bs;
  }
  if (last && p != N) return -1;                           // a truncated record at the end of the file
  if (off.back() > cap) return -2;
  t_scan_ += omp_get_wtime() - tt; tt = omp_get_wtime();
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)recs.size(); i++) {
    const Rec& x = recs[i];
    char* o = dst + off[i];
    *o++ = '@';
    memcpy(o, d + x.name, x.lname); o += x.lname;
    if (const int sfx = suffix_of(x.flag)) { *o++ = '/'; *o++ = (char)('0' + sfx); }
    *o++ = '\n';
    char* s = o;
    char* q = o + x.lseq + 3;
    const bool rev = x.flag & 0x10;
    const bool noq = x.lseq && d[x.qual] == 0xff;
    for (uint32_t k = 0; k < x.lseq; k++) {
      const uint32_t src = rev ? x.lseq - 1 - k : k;
      const unsigned char c = d[x.seq + (src >> 1)];
      const unsigned code = (src & 1) ? (c & 15) : (c >> 4);
      s[k] = NT16[rev ? COMP16[code] : code];
      q[k] = noq ? 'B' : (char)(d[x.qual + src] + 33);
    }
    s[x.lseq] = '\n'; s[x.lseq + 1] = '+'; s[x.lseq + 2] = '\n';
    q[x.lseq] = '\n';
  }
  t_conv_ += omp_get_wtime() - tt;
  if (last && getenv("CATT_TRACE")) fprintf(stderr, "[catt trace] bam stream: inflate %.1f ms, record scan %.1f ms, text %.1f ms\n", t_inf_ * 1e3, t_scan_ * 1e3, t_conv_ * 1e3);
  *wrote = off.back();
  pending_ = N - p;
  if (pending_) memmove(buf_.data(), buf_.data() + p, pending_);
  if (last) { eof_ = true; return 0; }
  return 1;
}

// BAM (BGZF or uncompressed) or SAM text (plain or gzipped) -> the reads `samtools fastq -N` would write, appended to hr.
int parse_bam(const char* path, HostReads* hr) {
  std::vector<unsigned char> raw;
  std::string inflated;
  size_t n = 0;
  Trace tr;
  if (!read_raw_file(path, &raw, &n)) { set_error("cannot read %s", path); return CATT_E_IO; }
  tr.lap("bam: read file");
  const unsigned char* d = raw.data();
  if (n >= 2 && raw[0] == 0x1f && raw[1] == 0x8b) {
    const int r = bgzf_inflate(raw.data(), n, &inflated, path);
    if (r < 0) return r;
    if (r == 0 && !slurp_file(path, inflated)) { set_error("cannot read %s", path); return CATT_E_IO; }     // plain gzip (a gzipped SAM)
    std::vector<unsigned char>().swap(raw);
    d = reinterpret_cast<const unsigned char*>(inflated.data()); n = inflated.size();
    tr.lap("bam: inflate");
  }
  const int rc = n >= 4 && memcmp(d, "BAM\1", 4) == 0 ? parse_bam_bytes(d, n, path, hr) : parse_sam_bytes(d, n, path, hr);
  tr.lap("bam: decode records");
  return rc;
}

}  // namespace catt
