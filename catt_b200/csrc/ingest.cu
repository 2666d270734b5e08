// ingest.cu — FASTQ ingestion on the device: raw FASTQ text in HBM -> line index -> per-read QNAME / sequence /
// quality spans -> 2-bit packed reads.  Replaces the host parser + host packing for catt_run_file(s): what bwa's kseq reader
// (name = header up to the first whitespace, catt.jl:43) and FASTQ.Reader see, without the host touching a single record.
// The text stays resident: Stage B reads the QNAMEs and the quality strings straight from it.
//
// Only the strict 4-line FASTQ layout is handled here ('@' header, sequence, '+' line, quality of the same length, no blank
// lines; '\r\n' tolerated).  Anything else (FASTA, wrapped records, blank lines, malformed input) makes ingest_fastq return
// CATT_E_ARG with `*fallback = 1` and the caller takes the host parser, which owns the error messages.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cuda/std/functional>

#include "host_types.h"

namespace catt {
namespace {

struct IsNewline {
  const char* text;
  __host__ __device__ __forceinline__ bool operator()(const long long& i) const { return text[i] == '\n'; }
};

enum { IG_BAD_LAYOUT = 0, IG_TOO_LONG, IG_MAXLEN, IG_COUNT = 4 };

// newline count first, so the line index is allocated exactly (a hostile input could be all newlines)
__global__ void __launch_bounds__(256) count_nl_kernel(const char* __restrict__ text, long long t0, long long bytes, unsigned long long* __restrict__ out) {
  unsigned long long mine = 0;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < bytes; i += (long long)gridDim.x * blockDim.x * 16) {
    const long long e = min(i + 16, bytes);
    for (long long k = i; k < e; k++) mine += text[t0 + k] == '\n';
  }
  #pragma unroll
  for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}

// one thread per record: spans of the four lines, validation, read length and packed size
__global__ void record_kernel(const char* __restrict__ text, long long t0, const long long* __restrict__ nl, long long n,
                              long long* __restrict__ name_beg, long long* __restrict__ name_end, long long* __restrict__ seq_beg,
                              long long* __restrict__ qual_beg, int32_t* __restrict__ len, uint32_t* __restrict__ nwords,
                              unsigned long long* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { nwords[n] = 0; return; }
  const long long l0 = i ? nl[4 * i - 1] + 1 : t0, e0 = nl[4 * i];
  const long long l1 = e0 + 1, e1 = nl[4 * i + 1];
  const long long l2 = e1 + 1, e2 = nl[4 * i + 2];
  const long long l3 = e2 + 1, e3 = nl[4 * i + 3];
  auto trim = [&](long long b, long long e) { return (e > b && text[e - 1] == '\r') ? e - 1 : e; };
  const long long s_end = trim(l1, e1), q_end = trim(l3, e3), h_end = trim(l0, e0);
  bool bad = h_end <= l0 || text[l0] != '@' || trim(l2, e2) <= l2 || text[l2] != '+' || s_end <= l1 || (q_end - l3) != (s_end - l1);
  long long ne = l0 + 1;
  while (ne < h_end && text[ne] != ' ' && text[ne] != '\t') ne++;
  const long long L = s_end - l1;
  if (L > MAX_READ) { atomicAdd(&flags[IG_TOO_LONG], 1ull); bad = true; }
  if (bad) atomicAdd(&flags[IG_BAD_LAYOUT], 1ull);
  const int l = bad ? 0 : (int)L;
  name_beg[i] = l0 + 1; name_end[i] = ne; seq_beg[i] = l1; qual_beg[i] = l3;
  len[i] = l; nwords[i] = (uint32_t)read_words(l);
  if (l > 0) atomicMax(&flags[IG_MAXLEN], (unsigned long long)l);
}

// one warp per read: lane k builds base word k (16 bases) and, for k < mask words, N-mask word k (32 bases)
// One warp per read, 32 bases per step: every lane loads one character (one coalesced request per step), and three ballots give
// the N-mask word and the two bit planes of the codes; two lanes interleave the planes into the step's two 2-bit words.
__device__ __forceinline__ uint32_t spread16(uint32_t x) {        // bit i of the low half -> bit 2i
  x &= 0xffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__global__ void __launch_bounds__(256) pack_kernel(const char* __restrict__ text, const long long* __restrict__ seq_beg,
                                                   const int32_t* __restrict__ len, const uint32_t* __restrict__ off, long long n,
                                                   uint32_t* __restrict__ words) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (long long)gridDim.x * 8) {
    const int l = len[r];
    const char* __restrict__ s = text + seq_beg[r];
    uint32_t* __restrict__ b = words + off[r];
    const int bw = read_base_words(l);
    for (int i0 = 0; i0 < l; i0 += 32) {
      const int i = i0 + lane;
      const uint32_t c = i < l ? nt_code(s[i]) : 0u;
      const uint32_t p0 = __ballot_sync(0xffffffffu, c & 1u), p1 = __ballot_sync(0xffffffffu, c & 2u), pn = __ballot_sync(0xffffffffu, c > 3u);
      const int k = (i0 >> 4) + lane;                             // lanes 0 and 1: the two base words of this step
      if (lane < 2 && k < bw) b[k] = spread16(p0 >> (16 * lane)) | (spread16(p1 >> (16 * lane)) << 1);
      if (lane == 2) b[bw + (i0 >> 5)] = pn;
    }
  }
}

}  // namespace

// 2-bit packing of reads whose raw bases sit in HBM: read r = text[seq_beg[r], seq_beg[r] + len[r]) -> words + off[r]
int launch_pack_raw(const char* d_text, const int64_t* d_seq_beg, const int32_t* d_len, const uint32_t* d_off, int64_t n, uint32_t* d_words,
                    int sm_count, cudaStream_t st, int* launches) {
  if (n == 0) return CATT_OK;
  const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)sm_count * 32);
  pack_kernel<<<grid, 256, 0, st>>>(d_text, reinterpret_cast<const long long*>(d_seq_beg), d_len, d_off, n, d_words);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return CATT_OK;
}

// Parses the FASTQ text of one input file, resident at d_text[t0, t1) (ends with '\n'), appending its reads at index `base`.
// On success *n_out = records found; the ctx read buffers (rd_off is filled per file relative to `word_base`) are extended.
// Runs on `st`.  The per-read buffers it appends to may be in use by Stage A of the previous piece on the alignment streams: before
// any of them has to GROW (free + reallocate), those streams are drained.
int ingest_fastq(catt_ctx* c, const char* d_text, int64_t t0, int64_t t1, int64_t base, uint64_t word_base, int64_t* n_out,
                 uint64_t* words_out, int* max_len, int* fallback, int* launches, cudaStream_t st) {
  *fallback = 0; *n_out = 0; *words_out = 0;
  auto drain_if_growing = [&](std::initializer_list<std::pair<const DevBuf*, size_t>> need) -> int {
    for (const auto& x : need)
      if (x.second > x.first->cap) {
        CATT_CUDA(cudaStreamSynchronize(c->stream));
        CATT_CUDA(cudaStreamSynchronize(c->stream_j));
        break;
      }
    return CATT_OK;
  };
  const int64_t bytes = t1 - t0;
  if (bytes <= 0) return CATT_OK;
  int rc;
  // ---- newline positions
  unsigned long long* d_flags; long long* d_count;
  if ((rc = c->sb[SB_CNT].ensure(64 * sizeof(unsigned long long)))) return rc;
  d_flags = c->sb[SB_CNT].as<unsigned long long>() + 32;
  d_count = reinterpret_cast<long long*>(d_flags + IG_COUNT);
  CATT_CUDA(cudaMemsetAsync(d_flags, 0, (IG_COUNT + 2) * sizeof(unsigned long long), st));
  unsigned long long n_nl = 0;
  count_nl_kernel<<<c->sm_count * 8, 256, 0, st>>>(d_text, t0, bytes, d_flags + IG_COUNT + 1);
  CATT_CUDA(cudaGetLastError());
  CATT_CUDA(cudaMemcpyAsync(&n_nl, d_flags + IG_COUNT + 1, sizeof n_nl, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  if (n_nl == 0 || n_nl % 4 != 0) { *fallback = 1; set_error("not a 4-line FASTQ"); return CATT_E_ARG; }
  if ((rc = c->rd_nl.ensure((size_t)(n_nl + 1) * sizeof(long long)))) return rc;
  long long* d_nl = c->rd_nl.as<long long>();
  cub::CountingInputIterator<long long> first(t0);
  size_t need = 0;
  CATT_CUDA(cub::DeviceSelect::If(nullptr, need, first, d_nl, d_count, (long long)bytes, IsNewline{d_text}, st));
  if ((rc = c->sb[SB_CUB].ensure(std::max<size_t>(need, 256)))) return rc;
  need = c->sb[SB_CUB].cap;
  CATT_CUDA(cub::DeviceSelect::If(c->sb[SB_CUB].p, need, first, d_nl, d_count, (long long)bytes, IsNewline{d_text}, st));
  long long n_lines = 0;
  CATT_CUDA(cudaMemcpyAsync(&n_lines, d_count, sizeof n_lines, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  *launches += 4;
  if ((unsigned long long)n_lines != n_nl) { set_error("FASTQ line index mismatch"); return CATT_E_ARG; }
  const int64_t n = n_lines / 4;
  // ---- per-record spans
  const int64_t tot = base + n;
  if ((rc = drain_if_growing({{&c->rd_name_beg, (size_t)(tot + 1) * 8}, {&c->rd_name_end, (size_t)(tot + 1) * 8}, {&c->rd_seq_beg, (size_t)(tot + 1) * 8},
                              {&c->rd_qual_beg, (size_t)(tot + 1) * 8}, {&c->rd_len, (size_t)(tot + 1) * 4}, {&c->rd_off, (size_t)(tot + 2) * 4}}))) return rc;
  if ((rc = c->rd_name_beg.ensure_keep((size_t)(tot + 1) * 8, st)) || (rc = c->rd_name_end.ensure_keep((size_t)(tot + 1) * 8, st)) ||
      (rc = c->rd_seq_beg.ensure_keep((size_t)(tot + 1) * 8, st)) || (rc = c->rd_qual_beg.ensure_keep((size_t)(tot + 1) * 8, st)) ||
      (rc = c->rd_len.ensure_keep((size_t)(tot + 1) * 4, st)) || (rc = c->rd_off.ensure_keep((size_t)(tot + 2) * 4, st)) ||
      (rc = c->rd_nwords.ensure((size_t)(n + 1) * 4))) return rc;
  long long* name_beg = c->rd_name_beg.as<long long>() + base;
  long long* name_end = c->rd_name_end.as<long long>() + base;
  long long* seq_beg = c->rd_seq_beg.as<long long>() + base;
  long long* qual_beg = c->rd_qual_beg.as<long long>() + base;
  int32_t* len = c->rd_len.as<int32_t>() + base;
  uint32_t* off = c->rd_off.as<uint32_t>() + base;
  uint32_t* nwords = c->rd_nwords.as<uint32_t>();
  record_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(d_text, t0, d_nl, n, name_beg, name_end, seq_beg, qual_beg, len, nwords, d_flags);
  CATT_CUDA(cudaGetLastError());
  need = 0;                                                      // word offsets continue after the previous files' words
  CATT_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, need, nwords, off, cuda::std::plus<uint32_t>{}, (uint32_t)word_base, n + 1, st));
  if ((rc = c->sb[SB_CUB].ensure(std::max<size_t>(need, 256)))) return rc;
  need = c->sb[SB_CUB].cap;
  CATT_CUDA(cub::DeviceScan::ExclusiveScan(c->sb[SB_CUB].p, need, nwords, off, cuda::std::plus<uint32_t>{}, (uint32_t)word_base, n + 1, st));
  unsigned long long hflags[IG_COUNT];
  uint32_t end_words = 0;
  CATT_CUDA(cudaMemcpyAsync(hflags, d_flags, sizeof hflags, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(&end_words, off + n, sizeof end_words, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  const uint64_t total_words = (uint64_t)end_words - word_base;
  *launches += 3;
  if (hflags[IG_TOO_LONG]) { set_error("%llu reads are longer than the limit of %d nt", hflags[IG_TOO_LONG], MAX_READ); return CATT_E_LIMIT; }
  if (hflags[IG_BAD_LAYOUT]) { *fallback = 1; set_error("FASTQ records do not follow the strict 4-line layout"); return CATT_E_ARG; }
  if (end_words < word_base || (uint64_t)n * read_words(MAX_READ) + word_base > 0xFFFFFFF0ull) {
    set_error("batch too large: packed reads may exceed 2^32 words");
    return CATT_E_LIMIT;
  }
  // ---- pack
  if ((rc = drain_if_growing({{&c->rd_words, (size_t)(word_base + total_words + 4) * 4}}))) return rc;
  if ((rc = c->rd_words.ensure_keep((size_t)(word_base + total_words + 4) * 4, st))) return rc;
  const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)c->sm_count * 32);
  pack_kernel<<<grid, 256, 0, st>>>(d_text, seq_beg, len, off, n, c->rd_words.as<uint32_t>());
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  *n_out = n; *words_out = total_words; *max_len = std::max(*max_len, (int)hflags[IG_MAXLEN]);
  return CATT_OK;
}

}  // namespace catt
