// internal.cuh — shared declarations of libcatt_b200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "catt_b200.h"

namespace catt {

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define CATT_CUDA(call)                                                         \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) return ::catt::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// The opt-in to more than 48 KB of dynamic shared memory is a per-DEVICE attribute of a kernel: a process that drives several
// GPUs (catt_config.ngpu, or one context per device) must set it on each of them.  One word per call site remembers the devices
// of this process that already have it.
template <class K> inline int smem_optin(K kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  CATT_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return CATT_OK;
  CATT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.fetch_or(bit, std::memory_order_release);
  return CATT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// limits and packing
// ---------------------------------------------------------------------------------------------------------------
constexpr int MAX_READ = CATT_MAX_READ_LEN;     // rows of the DP, positions of the seeding arrays
constexpr int SEED_K = 10;                      // bwa mem -k 10
constexpr int MAX_CW = 24;                      // candidate bitmap words per read: 2*nrec <= 768
constexpr int MAX_REC = MAX_CW * 16;            // 384 records per reference set (IGHV has 376)
constexpr int MAX_REC_LEN = 320;                // longest record the SW kernel tiles (TRGV 311)
constexpr int MOTIF_MAX_ALT = 64;
constexpr int MOTIF_MAX_LEN = 8;

__host__ __device__ inline uint8_t nt_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

// Packed read layout in HBM (u32 words), per read at word offset off[i]:
//   [ceil(len/16) words of 2-bit bases, base b at bits 2*(b%16) of word b/16]
//   [ceil(len/32) words of N mask,      base b at bit  b%32     of word b/32]
__host__ __device__ inline int read_base_words(int len) { return (len + 15) >> 4; }
__host__ __device__ inline int read_mask_words(int len) { return (len + 31) >> 5; }
__host__ __device__ inline int read_words(int len) { return read_base_words(len) + read_mask_words(len); }
// words of the 2-bit packed 2-strand text (padded so 64-bit window reads and 16-byte bulk copies stay in bounds)
__host__ __device__ inline int t2_words(int L) { return (((2 * L + 15) / 16) + 2 + 3) & ~3; }

// ---------------------------------------------------------------------------------------------------------------
// reference set (V or J): host copy + device tables
// ---------------------------------------------------------------------------------------------------------------
struct RefDev {            // POD view handed to kernels
  const uint32_t* T2;      // 2-bit packed 2-strand text (forward + reverse complement), 2L bases
  const uint8_t* Tb;       // forward text, one byte per base (0..3): DP targets
  const uint8_t* fa;       // forward FASTA codes (0..4, N kept): real_score compares against these
  const uint32_t* bitmap;  // 2^20 bits: which 10-mers occur in the 2-strand text
  const uint32_t* kstart;  // 2^20+1 CSR offsets
  const uint32_t* kpos;    // text positions by 10-mer, grouped by the text base preceding the occurrence
  const unsigned long long* ksub;  // per 10-mer: cumulative sizes of the preceding-base groups 0..3 (4 x u16)
  const uint16_t* pos2rid; // forward position -> record
  const int32_t* rec_off;
  const int32_t* rec_len;
  const uint16_t* canon;   // candidate (rid*2+strand) -> candidate of the first byte-identical record (same strand)
  int32_t L, nrec, maxlen, CW;
  int32_t direct;          // 1: every candidate goes through the TRACK kernel once (score + end cell), no bulk pass (sw.cu: launch_sw)
};

struct RefSet {
  std::vector<std::string> names;
  std::vector<int32_t> off, len;
  std::vector<uint8_t> fa;   // L
  std::vector<uint8_t> T;    // 2L
  int L = 0, maxlen = 0;
  RefDev dev{};
  std::vector<void*> allocs;
  int load(const std::string& fasta);   // host side
  int upload();                         // device tables
  void release();
};

// ---------------------------------------------------------------------------------------------------------------
// motif tables (compiled regexes of reference.jl) and the per-context device config
// ---------------------------------------------------------------------------------------------------------------
struct MotifDev {
  int32_t len, nalt;
  // alt[k][r]: the set of alternatives (bit a) that accept residue r (letter-'A', 26 = '*') at position k; the motif
  // matches at p iff the AND over k of alt[k][aa[p+k]] is non-zero - len table reads instead of nalt x len tests
  unsigned long long alt[MOTIF_MAX_LEN][28];
};
enum { M_C = 0, M_F = 1, M_INNER_C = 2, M_INNER_F = 3, M_HARD_C = 4, M_HARD_F = 5, M_COUNT = 6 };
struct FinderDev {
  MotifDev m[M_COUNT];
  int32_t coffset, foffset;
};
int compile_motif(const char* re, MotifDev* out);   // host; returns 0 or CATT_E_ARG

// ---------------------------------------------------------------------------------------------------------------
// packed reads on the device
// ---------------------------------------------------------------------------------------------------------------
struct ReadsDev {
  const uint32_t* words;
  const uint32_t* off;     // word offset per read
  const int32_t* len;
  int64_t n;
  int32_t max_len;
};

struct HostReads {          // views into caller memory (or owned buffers for catt_run_file)
  int64_t n = 0;
  const char* names = nullptr; const int64_t* name_off = nullptr;
  const char* seqs = nullptr;  const int64_t* seq_off = nullptr;
  const char* quals = nullptr;
  std::string own_names, own_seqs, own_quals;
  std::vector<int64_t> own_name_off, own_seq_off;
};

struct PackedReads {
  std::vector<uint32_t> words;
  std::vector<uint32_t> off;
  std::vector<int32_t> len;
  int max_len = 0;
};
int pack_reads(int64_t n, const char* seqs, const int64_t* seq_off, PackedReads* out);
// layout pass (lengths, word offsets; parallel) and the packing of reads [i0, i1) into dst (dst[0] = word off[i0])
int pack_layout(int64_t n, const int64_t* seq_off, uint32_t* off, int32_t* len, int* max_len);
void pack_range(int64_t i0, int64_t i1, const char* seqs, const int64_t* seq_off, const uint32_t* off, uint32_t* dst);

// ---------------------------------------------------------------------------------------------------------------
// stage A device buffers and kernels
// ---------------------------------------------------------------------------------------------------------------
struct SwTask { uint32_t read; uint16_t a, b; };   // two same-strand candidates (cand = rid*2+strand; 0xFFFF = none)

struct AlignBuffers {       // per reference set, sized for one batch
  int64_t cap_reads = 0, cap_tasks = 0;
  uint32_t* candbits = nullptr;    // n * CW: (record, strand) pairs bwa-mem would extend
  uint32_t* canonbits = nullptr;   // n * CW: the same set with byte-identical records folded onto their first copy
  uint32_t* ntask = nullptr;       // n (+1): tasks per read, then exclusive offsets
  uint32_t* task_off = nullptr;    // n+1
  SwTask* tasks = nullptr;
  uint32_t* task_score = nullptr;  // per task: scoreA | scoreB << 16
  SwTask* wtasks = nullptr;        // n: the winning candidate of every mapped read (compacted), input of the TRACK pass
  uint32_t* wres = nullptr;        // 2 per winner: score<<20 | (1023-ei)<<10 | (1023-ej)
  uint32_t* worder = nullptr;      // n + MAX_REC + 2: winner indices grouped by record, bins padded to even sizes (sw.cu: dispatch_track2)
  uint32_t* whist = nullptr;       // MAX_REC counts, MAX_REC cursors, slot count
  uint32_t* gapped_list = nullptr; // read indices whose winner needs the traceback kernel
  unsigned long long* counters = nullptr;  // [0] cand, [1] cells, [2] n_gapped, [3] hits, [4] seed redo, [6] seed overflow,
                                           // [7] winners, [8] cells computed (canonical records), [9] task-buffer overflow
  void* cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
  uint32_t* seed_scratch = nullptr; size_t seed_scratch_runs = 0;
};

// seeding (seed.cu)
int launch_seed(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, int sm_count, cudaStream_t st, int* launches);
// task expansion + SW scoring + selection + traceback (sw.cu)
constexpr int N_COUNTERS = 16;
// counters 10..13: work cursors of the persistent kernels (seeding, its overflow pass, bulk DP, TRACK pass), zeroed with the rest
constexpr int CTR_WORK_SEED = 10, CTR_WORK_REDO = 11, CTR_WORK_SW = 12, CTR_WORK_TRACK = 13;
int ensure_task_buffers(AlignBuffers& ab, int64_t want);
int launch_expand(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, cudaStream_t st, int* launches);
int launch_sw(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, int sm_count, cudaStream_t st, int* launches);
int launch_select(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, catt_aln* recs, int64_t first_ordinal, int min_score,
                  int sm_count, cudaStream_t st, int* launches);
int launch_select_pre(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, uint32_t* tied, int16_t* mxs, int16_t* ncands, cudaStream_t st, int* launches);
int launch_select_post(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, uint32_t* tied, int16_t* mxs, int16_t* ncands, catt_aln* recs,
                       int64_t first_ordinal, int min_score, int sm_count, cudaStream_t st, int* launches);
int launch_sw_pairs(const RefDev& ref, const ReadsDev& reads, const int32_t* d_rid, const int32_t* d_strand, int32_t* d_out3,
                    int sm_count, cudaStream_t st);
int alu_peak(int sm_count, cudaStream_t st, double* ops_per_s, double* ms);

// ---------------------------------------------------------------------------------------------------------------
// stage B kernels (extract.cu, pool.cu)
// ---------------------------------------------------------------------------------------------------------------
struct ExtractItem {       // one SAM record (or V/J pair) to turn into a Myread
  uint32_t read;           // index into the packed reads (the V record's read for kind 2)
  uint32_t read2;          // kind 2: the J record's read (SEQ equality, catt.jl:298)
  int16_t kind;            // 0 V part, 1 J part, 2 both-mapped
  int16_t strand;
  int16_t qb, m_end, rb, rid;     // of the V record (kind 0,2) or the J record (kind 1)
  int16_t j_m_end, j_rid, j_strand, pad_;   // kind 2: GetEndCigar / RNAME / strand of the J record
};
struct ExtractOut {
  int16_t status;          // 0 dropped, 1 kept (able and passes the length / N filters); 2 = dropped for SEQ mismatch
  int16_t seq_off, seq_len;       // slice of the SAM-oriented read
  int16_t cdr3_off, cdr3_len;     // relative to the slice; -1 = "None"
  int16_t able;
  int32_t pad_;
};
int launch_extract(const RefDev& vref, const RefDev& jref, const ReadsDev& reads, const FinderDev* d_finder,
                   const ExtractItem* d_items, ExtractOut* d_out, int64_t n, int kmer, int allow_v, int allow_j, int sm_count,
                   cudaStream_t st, int* launches);
int launch_real_score(const RefDev& ref, const ReadsDev& reads, const int32_t* d_rid, const int32_t* d_strand, const int32_t* d_qb, const int32_t* d_m_end,
                      const int32_t* d_rb, int allowance, uint8_t* d_out, int sm_count, cudaStream_t st);
int launch_finder_raw(const ReadsDev& reads, const FinderDev* d_finder, int array_version, int32_t* d_out3, int sm_count, cudaStream_t st);

struct PoolItem {          // a Vpart/Jpart read as the k-mer table sees it
  uint32_t read;
  int16_t strand, seq_off, seq_len;
  int16_t side;            // 0 = V list (last k+10 nt, extend right), 1 = J list (first k+10 nt, extend left)
  uint32_t order;          // position in [Vpart; Jpart]: later writers win
  int32_t partial;         // 1 = cdr3 == "None": run the extension + finder!(array)
};
constexpr int EXT_MAX = 152;                 // extension stops at 150 nt (catt.jl:349,389)
constexpr int EXT_WORDS = (EXT_MAX + 15) / 16;
struct ExtendOut {         // 48 bytes
  int16_t n_ext;           // bases added
  int16_t cdr3_off, cdr3_len;   // in the extended sequence; -1 = still "None"
  int16_t able;
  uint32_t ext2[EXT_WORDS];     // added bases, 2 bits each, in emission order
};
struct OutDesc {           // everything the host needs to materialise one Myread (32 bytes)
  uint32_t read;
  int32_t xi;              // index into the ExtendOut array, -1 = not a partial read
  int16_t strand, kind, seq_off, seq_len, cdr3_off, cdr3_len, v_id, j_id, q_start1, j_pad;
  int32_t pad_;
};
// open addressing, `nsub` sub-tables of `sub_slots` slots back to back (one GPU: nsub == 1); see pool.cu for the addressing
// slot = {key, ~value}: one 16-byte word, so a lookup touches one 32-byte sector and an insert's CAS + update hit the same one.
// The value is stored inverted (an all-ones memset is the empty table: key = EMPTY, value = 0) and raised with atomicMin.
struct KmerTable { ulonglong2* slots; uint64_t sub_slots; uint32_t nsub; };
struct PoolEntry { unsigned long long key, val; };    // one (k-mer, order << 8 | position) pair on its way to the GPU that owns the key
constexpr int KMER_MAX = 31;                            // one 2-bit packed u64 per k-mer, ~0 = empty slot
int launch_pool(const ReadsDev& reads, const PoolItem* d_items, int64_t n, int kmer, KmerTable tab, int sm_count, cudaStream_t st, int* launches);
int launch_pool_emit(const ReadsDev& reads, const PoolItem* d_items, int64_t n, int kmer, KmerTable tab, int pass, unsigned long long* d_cursor,
                     PoolEntry* d_out, int sm_count, cudaStream_t st, int* launches);
int launch_pool_insert(const PoolEntry* d_in, int64_t n, KmerTable tab, int sm_count, cudaStream_t st, int* launches);
int launch_extend(const ReadsDev& reads, const FinderDev* d_finder, const PoolItem* d_items, const uint32_t* d_partial_idx,
                  int64_t n_partial, int kmer, KmerTable tab, ExtendOut* d_out, uint32_t* d_rescued, int sm_count, cudaStream_t st, int* launches);
int launch_best_d(const char* d_cdr3, const int32_t* d_off, int64_t n, const char* d_dseq, const int32_t* d_doff, int nd,
                  int32_t* d_out, cudaStream_t st);

}  // namespace catt
