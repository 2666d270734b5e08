// host_types.h — host-side objects behind the opaque handles of include/catt_b200.h (shared by api.cu / assemble.cu).
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <string>
#include <vector>

#include "internal.cuh"

struct DeviceReads {
  uint32_t* words = nullptr; uint32_t* off = nullptr; int32_t* len = nullptr;
  int64_t n = 0; int max_len = 0;
  catt::ReadsDev view() const { return catt::ReadsDev{words, off, len, n, max_len}; }
  void release() { cudaFree(words); cudaFree(off); cudaFree(len); words = nullptr; off = nullptr; len = nullptr; }
};

// grow-only device scratch: Stage B reuses these across calls instead of cudaMalloc/cudaFree per call
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
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
enum { SB_ITEMS = 0, SB_EOUT, SB_KEYS, SB_VALS, SB_PITEMS, SB_PARTIAL, SB_XOUT, SB_COUNT };

struct catt_ctx {
  catt::RefSet V, J;
  catt::FinderDev finder{};
  catt::FinderDev* d_finder = nullptr;
  int min_score = 12, kmer = CATT_KMER_AUTO, nthreads = 1, allow_v = 3, allow_j = 9, device = 0, sm_count = 148;
  cudaStream_t stream = nullptr;
  std::vector<std::string> d_names, d_seqs;
  char* d_dseq = nullptr; int32_t* d_doff = nullptr;
  catt::AlignBuffers ab[2];
  DevBuf sb[SB_COUNT];
  cudaEvent_t ev[8]{};
};

struct catt_batch {
  DeviceReads dr;
  int64_t first_ordinal = 0;
  catt_info info{};
};

struct catt_result {
  catt_info info{};
  std::vector<catt_read> part[2];
  std::vector<catt_aln> rec[2];
  std::vector<char> seq_arena, qual_arena;     // every Myread string, NUL-terminated, back to back
};

namespace catt {
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// CATT_TRACE=1: wall-clock trace of the host-side sections (stderr)
struct Trace {
  double t; bool on;
  Trace() : t(now_ms()), on(getenv("CATT_TRACE") != nullptr) {}
  void lap(const char* what) { if (on) { double n = now_ms(); fprintf(stderr, "[catt trace] %-28s %8.2f ms\n", what, n - t); t = n; } }
};
// Stage B (assemble.cu): everything after the per-read alignment records exist
int stage_b(catt_ctx* c, const HostReads& hr, const DeviceReads& dr, catt_result* R, int* launches);
}  // namespace catt
