// host_types.h — host-side objects behind the opaque handles of include/catt_b200.h (shared by api.cu / assemble.cu).
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <string>
#include <vector>

#include "comm.h"
#include "internal.cuh"

struct DeviceReads {
  uint32_t* words = nullptr; uint32_t* off = nullptr; int32_t* len = nullptr;
  int64_t n = 0; int max_len = 0;
  catt::ReadsDev view() const { return catt::ReadsDev{words, off, len, n, max_len}; }
  // reads [i0, i0+cnt): offsets stay global word offsets into `words`
  catt::ReadsDev view(int64_t i0, int64_t cnt) const { return catt::ReadsDev{words, off + i0, len + i0, cnt, max_len}; }
  void release() { cudaFree(words); cudaFree(off); cudaFree(len); words = nullptr; off = nullptr; len = nullptr; }
};

// grow-only device scratch: reused across calls instead of cudaMalloc/cudaFree per call
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return CATT_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return catt::cuda_fail(e, "scratch cudaMalloc", __FILE__, __LINE__);
    cap = want;
    return CATT_OK;
  }
  // grow, keeping the contents (appending the reads of the next input file)
  int ensure_keep(size_t bytes, cudaStream_t st) {
    if (bytes <= cap) return CATT_OK;
    void* old = p; const size_t old_cap = cap;
    p = nullptr; cap = 0;
    int rc = ensure(bytes);
    if (rc) { p = old; cap = old_cap; return rc; }
    if (old) {
      cudaError_t e = cudaMemcpyAsync(p, old, old_cap, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      cudaFree(old);
      if (e != cudaSuccess) return catt::cuda_fail(e, "scratch grow copy", __FILE__, __LINE__);
    }
    return CATT_OK;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
// grow-only pinned host staging
struct PinBuf {
  void* p = nullptr; size_t cap = 0;
  int ensure(size_t bytes, bool exact = false) {          // exact: no growth slack (pinning costs ~1 ms per MB, allocation + first touch)
    if (bytes <= cap) return CATT_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = exact ? bytes + 256 : bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) return catt::cuda_fail(e, "pinned cudaMallocHost", __FILE__, __LINE__);
    cap = want;
    return CATT_OK;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Stage B device scratch slots (assemble.cu)
enum {
  SB_NAMES = 0, SB_NAME_OFF, SB_SLOT_OF, SB_OWNER, SB_MINV, SB_MINJ, SB_KEYS_A, SB_SORT_V, SB_SORT_J, SB_FLAG, SB_KPOS_V,
  SB_KPOS_J, SB_ORD_IDX_V, SB_ORD_IDX_J, SB_ORD_FLAG_V, SB_ORD_FLAG_J, SB_OPOS_V, SB_OPOS_J, SB_PAIR_A, SB_PAIR_B, SB_PKEY_A, SB_PKEY_B, SB_ITEMS,
  SB_EOUT, SB_KFLAG, SB_KFLAG1, SB_KPOS2, SB_KPOS3, SB_PFLAG, SB_XPOS, SB_TAB_KEYS, SB_TAB_VALS, SB_PITEMS, SB_PARTIAL, SB_XOUT, SB_DESC, SB_SSZ, SB_QSZ, SB_SOFF, SB_QOFF, SB_OUT_SEQ, SB_OUT_QUAL, SB_OUT_PART, SB_CFLAG, SB_CPOS, SB_CMAP, SB_QRID, SB_QGATHER, SB_QENTRY, SB_CNT, SB_CUB,
  SB_COUNT
};
enum { PIN_PACK0 = 0, PIN_PACK1, PIN_QUAL0, PIN_QUAL1, PIN_OFF, PIN_LEN, PIN_CNT, PIN_TEXT, PIN_RING0, PIN_RING1, PIN_QRID, PIN_QGATHER, PIN_QENTRY, PIN_COUNT };
constexpr int SH_SLOTS = 48;  // scratch of the multi-GPU Stage B (assemble.cu: stage_b_sharded)
constexpr int EV_A = 6;     // per reference set: seed start, seed end, expand end, sw end, select end

// QNAMEs in HBM: name i = s[beg[i], end[i]).  Host-supplied names are contiguous (beg = off, end = off + 1); names parsed on the
// device point into the FASTQ text.
struct NamesDev { const char* s = nullptr; const int64_t* beg = nullptr; const int64_t* end = nullptr; };

// quality strings of the sample in HBM (uploaded under the Stage A kernels); quals == nullptr for FASTA input
struct DeviceQuals {
  const char* quals = nullptr; const int64_t* seq_off = nullptr;
  // only_cdr3 with host buffers: the quality strings stay on the host until Stage B knows the few reads that keep a CDR3; their
  // strings are then gathered and uploaded (host_quals + host_seq_off = the caller's buffers)
  const char* host_quals = nullptr; const int64_t* host_seq_off = nullptr;
};

namespace catt {
// per input file of the sample (paired-end = two files, catt.jl:738-745): first read, mapped V / J records
struct FileHits {
  std::vector<int64_t> start, v, j;
  void single(int64_t n, int64_t vh, int64_t jh) { start = {0, n}; v = {vh}; j = {jh}; }
};
}  // namespace catt

struct catt_ctx {
  catt::RefSet V, J;
  catt::FinderDev finder{};
  catt::FinderDev* d_finder = nullptr;
  int min_score = 12, kmer = CATT_KMER_AUTO, nthreads = 1, allow_v = 3, allow_j = 9, device = 0, sm_count = 148;
  int keep_records = 0;                   // download the per-read alignment records into every result (diagnostics)
  int only_cdr3 = 0;                      // return the lists as they stand after filter!(x -> x.cdr3 != "None") (catt.jl:581-582)
  int64_t chunk_reads = 1 << 20;          // Stage A batch size of the pipelined path
  cudaStream_t stream = nullptr;          // kernels
  cudaStream_t copy_stream = nullptr;     // H2D of the packed chunks, under the kernels of the previous chunk
  cudaStream_t stream_j = nullptr;        // Stage A of the J reference set, concurrent with the V set on `stream`
  cudaStream_t stream_ingest = nullptr;   // FASTQ indexing + packing of the next piece, under Stage A of the current one (file pipeline)
  cudaEvent_t ev_ingest = nullptr;        // the piece's packed reads are complete
  cudaEvent_t ev_inputs = nullptr;        // the chunk's packed reads are ready (fork point of the two Stage A streams)
  std::vector<std::string> d_names, d_seqs;
  char* d_dseq = nullptr; int32_t* d_doff = nullptr;
  catt::AlignBuffers ab[2];
  DevBuf rd_words, rd_off, rd_len;        // packed reads of the sample being run (catt_run_*, catt_assemble)
  int64_t last_n = 0, last_words = 0; int last_max_len = 0;   // what the last Stage A left in rd_* (catt_shard_counts)
  DevBuf rd_quals, rd_seqoff;
  DevBuf nb[4];                           // nbsorb scans (nbsorb.cu): raw sequences, query planes, target planes, aux             // its quality strings (as given) and int64 offsets, for the string kernel
  DevBuf rd_text, rd_nl, rd_name_beg, rd_name_end, rd_seq_beg, rd_qual_beg, rd_nwords;   // FASTQ text ingested on the device (ingest.cu)
  DevBuf recs_all[2];                     // one catt_aln per read and reference set, for the whole sample
  DevBuf sb[SB_COUNT];
  DevBuf sh[SH_SLOTS];
  DevBuf cl[16];                          // clonotype table scratch (clonotype.cu)
  DevBuf df_tied[2], df_mx[2], df_nc[2];  // Stage A split at the tie-break (run_file_rank): per read tied-candidate bitmap, best score, candidates
  int want_clonotypes = 0;
  // ---- one sample over several GPUs.  Either this context is rank 0 of an in-process group (catt_config.ngpu > 1: `peers` are the
  //      contexts of ranks 1.., `comms[r]` the communicator of rank r) or it is one rank of a multi-process group (catt_ctx_join).
  std::vector<catt_ctx*> peers;
  std::vector<catt::Comm*> comms;
  std::shared_ptr<catt::LocalGroup> lgroup;
  catt::Comm* comm = nullptr;             // multi-process: this process' communicator
  bool resident_quals = false;            // stage_a_stream: keep the quality strings on the device even with only_cdr3
  struct Shard {                          // catt_shard_upload / catt_shard_run: this rank's block, resident in the rd_* / name buffers
    bool valid = false;
    int64_t n_total = 0, lo = 0;
    std::vector<int64_t> gstart, ord0;
    catt::FileHits fh;
    DeviceReads dr; DeviceQuals dq; NamesDev nd;
  } shard;
  PinBuf pin[PIN_COUNT];
  cudaEvent_t ev[8]{};
  cudaEvent_t ev_a[2][EV_A]{};            // Stage A stage boundaries per reference set
  cudaEvent_t ev_h2d[2]{};                // sequence staging buffer c&1 has reached the device
  cudaEvent_t ev_q[2]{};                  // ... and the quality staging buffer
  std::vector<cudaEvent_t> ev_piece;      // file input: piece k of the FASTQ text has reached the device
  DevBuf rd_raw[2];                       // raw bases of the chunk in flight (packed on the device)
};

struct catt_batch {
  DevBuf quals, seqoff, names, name_off; bool have_quals = false;      // inputs of Stage B, resident after the first assemble
  DeviceReads dr;
  int64_t first_ordinal = 0;
  catt_info info{};
};

namespace catt {
// Host blocks for the result lists.  A result holds ~250 B per kept read; freshly mapped memory costs one page fault per
// 4 KiB on first touch, which dominated the materialisation step, so freed blocks are kept (bounded) and reused by the
// next result of the process, and new blocks are 2 MiB aligned with MADV_HUGEPAGE.
void* block_get(size_t bytes, size_t* cap);
void block_put(void* p, size_t cap);
}  // namespace catt

struct catt_result {
  catt_info info{};
  catt_read* part[2] = {nullptr, nullptr};
  int64_t npart[2] = {0, 0};
  std::vector<catt_aln> rec[2];
  std::vector<catt_clonotype> clono; std::vector<char> clono_seq, clono_qual; int64_t n_clono = 0;     // clonotype.cu
  char* seq_arena = nullptr; char* qual_arena = nullptr;     // every Myread string, NUL-terminated, back to back
  size_t cap_part[2] = {0, 0}, cap_seq = 0, cap_qual = 0;
  ~catt_result() {
    catt::block_put(part[0], cap_part[0]);      // part[1] points into the same block
    catt::block_put(seq_arena, cap_seq); catt::block_put(qual_arena, cap_qual);
  }
};

namespace catt {
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// CATT_TRACE=1: wall-clock trace of the host-side sections (stderr)
struct Trace {
  double t; bool on;
  Trace() : t(now_ms()), on(getenv("CATT_TRACE") != nullptr) {}
  void lap(const char* what) { if (on) { double n = now_ms(); fprintf(stderr, "[catt trace] %-28s %8.2f ms\n", what, n - t); t = n; } }
};
// Stage B (assemble.cu): everything after the per-read alignment records exist.  The records are read from
// c->recs_all[0/1] on the device; names/seqs/quals are host buffers (names are uploaded, strings are materialised on the host).
int stage_b(catt_ctx* c, int64_t n, const NamesDev& names, const ReadsDev& reads, const DeviceQuals& dq, const FileHits& fh, catt_result* R,
            int* launches);
// Stage B of one sample spread over the ranks of `comm` (assemble.cu).  This rank holds reads [lo, lo + n_loc) of the concatenated
// input files (records in c->recs_all, packed reads in c->rd_* with n_words words); the lists come out on rank `root`.
struct ShardIn {
  Comm* comm = nullptr;
  int64_t lo = 0, n_loc = 0, n_words = 0;
  std::vector<int64_t> gstart;            // global file boundaries (nfiles + 1)
  std::vector<int64_t> bounds;            // global block boundaries of the ranks (nranks + 1)
  int max_len = 0;                        // longest read of the whole sample
  int root = 0;
};
int order_only(catt_ctx* c, int64_t n, const NamesDev& nd, const int32_t* d_len, int64_t nm_v, int64_t nm_j, int32_t* h_items, int64_t* n_items, int64_t* counts);
int build_clonotypes(catt_ctx* c, int64_t E, const catt_read* d_part, const char* d_seq, const char* d_qual, const char* host_seq,
                     const char* host_qual, catt_result* R, int* launches);
int stage_b_sharded(catt_ctx* c, const ShardIn& in, const NamesDev& names, const DeviceQuals& dq, catt_result* R, int* launches);
// FASTQ text of one input file, resident at d_text[t0, t1) -> spans + packed reads appended at read index `base` (ingest.cu)
int ingest_fastq(catt_ctx* c, const char* d_text, int64_t t0, int64_t t1, int64_t base, uint64_t word_base, int64_t* n_out,
                 uint64_t* words_out, int* max_len, int* fallback, int* launches, cudaStream_t st);
int launch_pack_raw(const char* d_text, const int64_t* d_seq_beg, const int32_t* d_len, const uint32_t* d_off, int64_t n, uint32_t* d_words,
                    int sm_count, cudaStream_t st, int* launches);
// uploads host QNAMEs (contiguous, n+1 offsets) into the context's name buffers
int upload_names(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, NamesDev* out, catt_info* info, cudaStream_t st = nullptr);
}  // namespace catt
