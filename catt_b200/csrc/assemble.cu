// assemble.cu — Stage B: from per-read alignment records to Vpart / Jpart (catt.jl:186-339 after the SAM files exist,
// plus the pool / extension / finder!(array) part of catt(), catt.jl:544-577).
//
// Everything the reference does between the stages with samtools / awk / Julia containers is restated on the device, on
// the alignment records Stage A left in HBM (ctx->recs_all), so no per-read data crosses PCIe until the final lists:
//   * `samtools sort -t AS | view -F 2308 | tac | awk '!a[$1]++'` (catt.jl:45-48): one 61-bit radix key
//     (AS, tid, pos, strand, input index) sorted descending == descending (AS, tid, pos, strand) with later input first
//     on complete ties; the awk dedup keeps the record with the LARGEST key of its trimmed QNAME (QNAMEs live in a device
//     hash set, compared byte-exactly on every hash hit; name_kernel / kept_kernel below)                  [SURVEY A7]
//   * get_kmer_fromsam / get_err_fromsam (catt.jl:153-184): device sums, formatted on the host as samtools stats
//     prints them                                                                                          [SURVEY A8]
//   * share_name, the NR % nthreads chunk order (closed form), the name-sorted both-mapped pairing (LSD radix sort over
//     8-byte name digits) (catt.jl:221-319)
//   * assignV/J, real_score, finder!, k-mer table, extension: extract.cu / pool.cu
// The Myread strings are written on the device too (materialise_kernel) and come back as three bulk copies; the host only
// formats two numbers (mean length, error rate) the way `samtools stats` prints them.
#include <omp.h>
#include <stdio.h>
#include <string.h>

#include <sys/mman.h>

#include <algorithm>
#include <mutex>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "host_types.h"

namespace catt {
namespace {

constexpr uint32_t INF = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xffffffffu;
// device counters (SB_CNT, unsigned long long each)
enum { C_NMAP_V = 0, C_NMAP_J, C_F_V, C_F_J, C_NV, C_NJ, C_NP, C_MAXNAME, C_SUM_LEN, C_SUM_NM, C_SUM_MI, C_MISMATCH, C_M, C_VPART,
       C_FULL, C_NPARTIAL, C_RESCUED, C_E, C_EV, C_COUNT = 24 };


__device__ __forceinline__ int trimmed_len(const char* s, int l) {            // bwa trim_readno
  if (l > 2 && s[l - 2] == '/' && s[l - 1] >= '0' && s[l - 1] <= '9') l -= 2;
  return l;
}

// ---- QNAME classes and the dedup, without a sorted rank ----------------------------------------------------------------------------
// `samtools sort -t AS | view -F 2308 | tac | awk '!a[$1]++'` keeps, per trimmed QNAME, the record that comes first in the final
// order.  The sort key (AS, tid, pos, strand, input index) is unique per record, so "first in descending order" is simply the class's
// LARGEST key: every mapped read raises its class's V and J maxima when it enters the QNAME set (one 32-byte slot per class:
// owner, max V key, max J key - one sector), a second pass in input order compares each record with its class (kept = its key is
// the maximum; shared = the class has a V and a J record) and only then are the keys sorted - to place the kept records, not to
// decide them.  Non-kept records sort as 0; the `shared` flag rides in bit 62, above the 61 sorted bits.
struct QClass { unsigned long long maxv, maxj; uint32_t owner; uint32_t pad_[3]; };      // owner = read index + 1 of the first inserter
constexpr unsigned long long KEY_SHARED = 1ull << 62;

__device__ __forceinline__ unsigned long long sam_key(const catt_aln& a, uint32_t i) {
  return ((unsigned long long)(uint16_t)a.as << 51) | ((unsigned long long)(uint32_t)a.rid << 42) | ((unsigned long long)(uint16_t)a.rb << 33) |
         ((unsigned long long)(a.strand & 1) << 32) | (unsigned long long)i;
}

__global__ void name_kernel(int64_t n, NamesDev nd, const catt_aln* __restrict__ rv, const catt_aln* __restrict__ rj, QClass* __restrict__ cls,
                            uint32_t mask, uint32_t* __restrict__ slot_of) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const catt_aln av = rv[i], aj = rj[i];
  if (!(av.mapped || aj.mapped)) { slot_of[i] = INF; return; }
  const char* s = nd.s + nd.beg[i];
  const int l = trimmed_len(s, (int)(nd.end[i] - nd.beg[i]));
  uint64_t h = 0xcbf29ce484222325ull;
  for (int k = 0; k < l; k++) { h ^= (unsigned char)s[k]; h *= 0x100000001b3ull; }
  h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32;
  uint32_t slot = (uint32_t)h & mask;
  while (true) {
    const uint32_t o = atomicCAS(&cls[slot].owner, 0u, (uint32_t)i + 1u);
    if (o == 0u || o == (uint32_t)i + 1u) break;
    const char* t = nd.s + nd.beg[o - 1];
    const int lt = trimmed_len(t, (int)(nd.end[o - 1] - nd.beg[o - 1]));
    bool same = lt == l;
    for (int k = 0; same && k < l; k++) same = s[k] == t[k];
    if (same) break;
    slot = (slot + 1) & mask;
  }
  slot_of[i] = slot;
  if (av.mapped) atomicMax(&cls[slot].maxv, sam_key(av, (uint32_t)i));
  if (aj.mapped) atomicMax(&cls[slot].maxj, sam_key(aj, (uint32_t)i));
}

// input order: the sort keys of the mapped records (0 for a record that lost its QNAME to a better one), kept counts, and the sums
// samtools stats takes over the final V list
__global__ void kept_kernel(int64_t n, const catt_aln* __restrict__ rv, const catt_aln* __restrict__ rj, const int32_t* __restrict__ len,
                            const uint32_t* __restrict__ slot_of, const QClass* __restrict__ cls, unsigned long long* __restrict__ keys_v,
                            unsigned long long* __restrict__ keys_j, unsigned long long* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool mv = false, mj = false;
  unsigned long long kv = 0, kj = 0, sl = 0, sn = 0, sm = 0;
  if (i < n) {
    const catt_aln av = rv[i], aj = rj[i];
    mv = av.mapped; mj = aj.mapped;
    if (mv || mj) {
      const QClass c = cls[slot_of[i]];
      const unsigned long long sh = (c.maxv != 0 && c.maxj != 0) ? KEY_SHARED : 0ull;
      if (mv) { const unsigned long long k = sam_key(av, (uint32_t)i); if (k == c.maxv) { kv = k | sh; sl = (unsigned long long)len[i]; sn = (unsigned long long)av.nm; sm = (unsigned long long)av.mi; } }
      if (mj) { const unsigned long long k = sam_key(aj, (uint32_t)i); if (k == c.maxj) kj = k | sh; }
    }
  }
  const unsigned bv = __ballot_sync(FULL, mv), bj = __ballot_sync(FULL, mj);
  const unsigned fv = __ballot_sync(FULL, kv != 0), fj = __ballot_sync(FULL, kj != 0);
  if (bv) {
    unsigned long long base = 0;
    const int leader = __ffs(bv) - 1;
    if (lane == leader) { base = atomicAdd(&cnt[C_NMAP_V], (unsigned long long)__popc(bv)); if (fv) atomicAdd(&cnt[C_F_V], (unsigned long long)__popc(fv)); }
    base = __shfl_sync(FULL, base, leader);
    if (mv) keys_v[base + __popc(bv & ((1u << lane) - 1))] = kv;
  }
  if (bj) {
    unsigned long long base = 0;
    const int leader = __ffs(bj) - 1;
    if (lane == leader) { base = atomicAdd(&cnt[C_NMAP_J], (unsigned long long)__popc(bj)); if (fj) atomicAdd(&cnt[C_F_J], (unsigned long long)__popc(fj)); }
    base = __shfl_sync(FULL, base, leader);
    if (mj) keys_j[base + __popc(bj & ((1u << lane) - 1))] = kj;
  }
  if (fv) {
    #pragma unroll
    for (int d = 16; d; d >>= 1) { sl += __shfl_xor_sync(FULL, sl, d); sn += __shfl_xor_sync(FULL, sn, d); sm += __shfl_xor_sync(FULL, sm, d); }
    if (lane == 0) { atomicAdd(&cnt[C_SUM_LEN], sl); atomicAdd(&cnt[C_SUM_NM], sn); atomicAdd(&cnt[C_SUM_MI], sm); }
  }
}

// position of the nr-th (1-based) line of the final SAM after the `NR % T` split + sequential read-back (catt.jl:228,256-276)
__device__ __forceinline__ uint32_t chunk_pos(uint32_t nr, uint32_t F, uint32_t T) {
  if (T <= 1) return nr - 1;
  const uint32_t p = nr % T;
  uint32_t start = 0;
  if (p > 0) {
    // file 0 holds NR = T, 2T, ...; file q >= 1 holds q, q+T, ...
    start = F / T;
    for (uint32_t q = 1; q < p; q++) start += F >= q ? (F - q) / T + 1 : 0;
  }
  const uint32_t first = p == 0 ? T : p;
  return start + (nr - first) / T;
}

// kept records in the reference's processing order (rank r among the kept = position in the descending sort; the zero keys of the
// dropped records sort behind them); records of shared QNAMEs leave the part lists and (J side) seed the pairs
__global__ void order_kernel(int64_t nm, const unsigned long long* __restrict__ sorted, const unsigned long long* __restrict__ n_kept, int T, int which,
                             NamesDev nd, uint32_t* __restrict__ ord_idx, uint32_t* __restrict__ ord_flag, uint32_t* __restrict__ pair_j,
                             unsigned long long* __restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long key = r < nm ? sorted[r] : 0ull;
  const bool live = key != 0;
  bool pair = false;
  uint32_t idx = 0;
  if (live) {
    const uint32_t F = (uint32_t)*n_kept;
    idx = (uint32_t)key;
    const uint32_t y = chunk_pos((uint32_t)r + 1, F, (uint32_t)T);
    const bool shared = (key & KEY_SHARED) != 0;
    ord_idx[y] = idx;
    ord_flag[y] = !shared;
    pair = which == 1 && shared;
  }
  // the J records of shared QNAMEs seed the both-mapped pairs: one atomic per warp (a counter every thread hits serialises)
  const unsigned bal = __ballot_sync(FULL, pair);
  if (!bal) return;
  const int lane = threadIdx.x & 31, leader = __ffs(bal) - 1;
  unsigned long long base = 0;
  unsigned nl = 0;
  if (pair) {
    const char* s = nd.s + nd.beg[idx];
    nl = (unsigned)trimmed_len(s, (int)(nd.end[idx] - nd.beg[idx]));
  }
  nl = __reduce_max_sync(FULL, nl);
  if (lane == leader) { base = atomicAdd(&cnt[C_NP], (unsigned long long)__popc(bal)); atomicMax(&cnt[C_MAXNAME], (unsigned long long)nl); }
  base = __shfl_sync(FULL, base, leader);
  if (pair) pair_j[base + __popc(bal & ((1u << lane) - 1))] = idx;
}

__global__ void collect_kernel(int64_t nm_v, int64_t nm_j, const uint32_t* opos_v, const uint32_t* opos_j, unsigned long long* cnt) {
  cnt[C_NV] = nm_v ? opos_v[nm_v] : 0; cnt[C_NJ] = nm_j ? opos_j[nm_j] : 0;
}

// 8-byte big-endian digit `d` of the trimmed QNAME, zero padded: LSD passes over these digits give Julia's String order
__global__ void pair_key_kernel(int64_t np, const uint32_t* __restrict__ pair_j, NamesDev nd, int d, unsigned long long* __restrict__ keys) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const uint32_t idx = pair_j[p];
  const char* s = nd.s + nd.beg[idx];
  const int l = trimmed_len(s, (int)(nd.end[idx] - nd.beg[idx]));
  unsigned long long k = 0;
  for (int b = 0; b < 8; b++) { const int pos = d * 8 + b; k = (k << 8) | (pos < l ? (unsigned char)s[pos] : 0u); }
  keys[p] = k;
}

__device__ __forceinline__ ExtractItem make_item(const catt_aln& a, uint32_t read, int kind) {
  ExtractItem it;
  it.read = read; it.read2 = read; it.kind = (int16_t)kind; it.strand = a.strand; it.qb = a.qb; it.m_end = a.m_end; it.rb = a.rb;
  it.rid = (int16_t)a.rid; it.j_m_end = 0; it.j_rid = 0; it.j_strand = 0; it.pad_ = 0;
  return it;
}

__global__ void part_item_kernel(int64_t F, const uint32_t* __restrict__ ord_idx, const uint32_t* __restrict__ ord_flag,
                                 const uint32_t* __restrict__ opos, const catt_aln* __restrict__ recs, int kind, uint32_t lo,
                                 ExtractItem* __restrict__ items) {
  const int64_t y = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= F || !ord_flag[y]) return;
  const uint32_t idx = ord_idx[y];
  items[opos[y]] = make_item(recs[idx], lo + idx, kind);
}

__global__ void pair_item_kernel(int64_t np, const uint32_t* __restrict__ pair_j, const uint32_t* __restrict__ slot_of, const QClass* __restrict__ cls,
                                 const catt_aln* __restrict__ rv, const catt_aln* __restrict__ rj, uint32_t lo, ExtractItem* __restrict__ items) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const uint32_t ji = pair_j[p];
  const uint32_t vi = (uint32_t)cls[slot_of[ji]].maxv;          // the kept V record of the QNAME: its key ends in the read index
  ExtractItem it = make_item(rv[vi], lo + vi, 2);
  const catt_aln b = rj[ji];
  it.read2 = lo + ji; it.j_m_end = b.m_end; it.j_rid = (int16_t)b.rid; it.j_strand = b.strand;
  items[p] = it;
}

// kept items by list: kind 0 / 2 -> Vpart (flag0), kind 1 -> Jpart (flag1)
__global__ void keep_kernel(int64_t n_items, const ExtractItem* __restrict__ items, const ExtractOut* __restrict__ eout,
                            uint32_t* __restrict__ kflag0, uint32_t* __restrict__ kflag1, uint32_t* __restrict__ pflag,
                            unsigned long long* __restrict__ cnt) {
  const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (it > n_items) return;
  pflag[it] = 0;
  if (it == n_items) { kflag0[it] = 0; kflag1[it] = 0; return; }
  const int st = eout[it].status, kind = items[it].kind;
  kflag0[it] = st == 1 && kind != 1;
  kflag1[it] = st == 1 && kind == 1;
  if (st == 2) atomicAdd(&cnt[C_MISMATCH], 1ull);
  if (st == 1 && kind == 2) atomicAdd(&cnt[C_FULL], 1ull);
}

// kept items -> PoolItem / OutDesc at their rank m (= position in [Vpart; Jpart])
__global__ void pitem_kernel(int64_t n_items, const ExtractItem* __restrict__ items, const ExtractOut* __restrict__ eout,
                             const uint32_t* __restrict__ kflag0, const uint32_t* __restrict__ kflag1, const uint32_t* __restrict__ kpos0,
                             const uint32_t* __restrict__ kpos1, const int32_t* __restrict__ jlen, PoolItem* __restrict__ pitems,
                             OutDesc* __restrict__ desc, uint32_t* __restrict__ pflag, unsigned long long* __restrict__ cnt) {
  const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t vpart = kpos0[n_items];
  if (it == 0) { cnt[C_VPART] = vpart; cnt[C_M] = vpart + kpos1[n_items]; }
  if (it >= n_items || !(kflag0[it] | kflag1[it])) return;
  const ExtractItem x = items[it];
  const ExtractOut e = eout[it];
  const int side = x.kind == 1;
  const uint32_t m = side ? vpart + kpos1[it] : kpos0[it];
  PoolItem pi;
  pi.read = x.read; pi.strand = x.strand; pi.seq_off = e.seq_off; pi.seq_len = e.seq_len; pi.side = (int16_t)side;
  pi.order = m; pi.partial = e.cdr3_off < 0;
  pitems[m] = pi;
  pflag[m] = pi.partial;
  OutDesc d;
  d.read = x.read; d.xi = -1; d.strand = x.strand; d.kind = x.kind; d.seq_off = e.seq_off; d.seq_len = e.seq_len;
  d.cdr3_off = e.cdr3_off; d.cdr3_len = e.cdr3_len;
  d.v_id = x.kind == 1 ? -1 : x.rid; d.j_id = x.kind == 0 ? -1 : (x.kind == 1 ? x.rid : x.j_rid);
  d.q_start1 = (int16_t)(x.qb + 1); d.j_pad = x.kind == 1 ? (int16_t)(jlen[x.rid] - (x.rb + 1)) : 0; d.pad_ = 0;
  desc[m] = d;
}

__global__ void partial_kernel(int64_t n_items, const uint32_t* __restrict__ pflag, const uint32_t* __restrict__ xpos,
                               uint32_t* __restrict__ partial_idx, OutDesc* __restrict__ desc, unsigned long long* __restrict__ cnt) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m == 0) cnt[C_NPARTIAL] = xpos[n_items];
  if (m >= n_items || !pflag[m]) return;
  partial_idx[xpos[m]] = (uint32_t)m;
  desc[m].xi = (int32_t)xpos[m];
}

// ---- Myread strings on the device ------------------------------------------------------------------------------------------
// What catt.jl:104,125-126,302-304 (assign*), Jtool.jl:364-365,382-383 (finder! slices qual) and catt.jl:378-379,420-421
// (extension pads qual with 'G') leave in Myread.seq / Myread.qual, written straight into the result arenas.
struct StrSrc {
  const OutDesc* desc; const ExtendOut* xout;     // xout == nullptr: no extension pass ran (every partial read keeps n_ext = 0)
  const char* quals; const int64_t* seq_off;      // quals == nullptr: FASTA input, 'I' everywhere
  int64_t M, vpart;                               // all Vpart/Jpart reads; the first vpart are Vpart
  const uint32_t* map; int64_t E;                 // output entry e is read map[e] (map == nullptr: every read, E == M)
  const int64_t* q_entry_off;                     // != nullptr: quality strings were gathered per output entry (quals + q_entry_off[e])
  // one sample over several GPUs: this GPU materialises ITS reads at their positions in the global lists - read m of `desc` is
  // read gm[m] of [Vpart; Jpart], its output entry is epos[gm[m]] and the string offsets are indexed by gm[m] (nullptr: one GPU)
  const uint32_t* gm = nullptr; const uint32_t* epos = nullptr;
  int64_t ordinal_base = 0;                       // global input index of this GPU's first read
};
__device__ __forceinline__ int64_t str_read(const StrSrc& S, int64_t e) { return S.map ? (int64_t)S.map[e] : e; }

// only_cdr3: which reads survive `filter!(x -> x.cdr3 != "None", ...)` (catt.jl:581-582)
__global__ void cdr3_flag_kernel(const OutDesc* __restrict__ desc, const ExtendOut* __restrict__ xout, int64_t M, uint32_t* __restrict__ flag) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m > M) return;
  if (m == M) { flag[m] = 0; return; }
  const OutDesc p = desc[m];
  flag[m] = p.xi >= 0 ? (xout != nullptr && xout[p.xi].cdr3_off >= 0) : (p.cdr3_off >= 0);
}
__global__ void entry_read_kernel(const OutDesc* __restrict__ desc, const uint32_t* __restrict__ map, int64_t E, uint32_t* __restrict__ rid) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) rid[e] = desc[map[e]].read;
}
__global__ void cdr3_map_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, int64_t M, int64_t vpart,
                                uint32_t* __restrict__ map, unsigned long long* __restrict__ cnt) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m == 0) { cnt[C_E] = pos[M]; cnt[C_EV] = pos[vpart]; }
  if (m < M && flag[m]) map[pos[m]] = (uint32_t)m;
}
__device__ __forceinline__ void str_dims(const StrSrc& S, int64_t m, int* n_ext, int* c_off, int* c_len, int* slen, int* qlen, int* base_q) {
  const OutDesc p = S.desc[m];
  *n_ext = 0; *c_off = p.cdr3_off; *c_len = p.cdr3_len;
  if (p.xi >= 0) {
    if (S.xout) { const ExtendOut& x = S.xout[p.xi]; *n_ext = x.n_ext; *c_off = x.cdr3_off; *c_len = x.cdr3_len; }
    else { *c_off = -1; *c_len = 0; }
  }
  *base_q = p.kind == 1 ? (p.q_start1 - 1 + max(0, (int)p.j_pad)) : p.seq_len;
  *slen = p.seq_len + *n_ext;
  *qlen = *c_off >= 0 ? *c_len : *base_q + *n_ext;
}

__global__ void strsize_kernel(StrSrc S, unsigned long long* __restrict__ ssz, unsigned long long* __restrict__ qsz) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > S.E) return;
  if (e == S.E) { ssz[e] = 0; qsz[e] = 0; return; }
  int n_ext, c_off, c_len, slen, qlen, base_q;
  str_dims(S, str_read(S, e), &n_ext, &c_off, &c_len, &slen, &qlen, &base_q);
  ssz[e] = (unsigned long long)slen + 1; qsz[e] = (unsigned long long)qlen + 1;
}

constexpr int MW = 8;   // warps per CTA
__global__ void __launch_bounds__(MW * 32) materialise_kernel(StrSrc S, ReadsDev reads, const unsigned long long* __restrict__ soff,
                                                              const unsigned long long* __restrict__ qoff, char* __restrict__ seq_arena,
                                                              char* __restrict__ qual_arena, catt_read* __restrict__ part,
                                                              const char* host_seq, const char* host_qual) {
  const int lane = threadIdx.x & 31;
  const char NT[5] = {'A', 'C', 'G', 'T', 'N'};
  for (int64_t e_ = (int64_t)blockIdx.x * MW + (threadIdx.x >> 5); e_ < S.E; e_ += (int64_t)gridDim.x * MW) {
    const int64_t m = str_read(S, e_);
    const OutDesc p = S.desc[m];
    int n_ext, c_off, c_len, slen, qlen, base_q;
    str_dims(S, m, &n_ext, &c_off, &c_len, &slen, &qlen, &base_q);
    const ExtendOut* x = (p.xi >= 0 && S.xout) ? &S.xout[p.xi] : nullptr;
    int64_t oe = e_, oi = e_;                                      // output entry, index of its string offsets
    if (S.gm) { oi = (int64_t)S.gm[m]; oe = (int64_t)S.epos[oi]; }
    const int side = p.kind == 1;                                  // J-side reads form Jpart, everything else Vpart
    const int lead = side == 1 ? n_ext : 0;                       // a left extension is prepended
    const int len = reads.len[p.read];
    const uint32_t* __restrict__ w = reads.words + reads.off[p.read];
    const uint32_t* __restrict__ nm = w + read_base_words(len);
    char* s = seq_arena + soff[oi];
    for (int i = lane; i <= slen; i += 32) {
      char ch = 0;
      if (i < slen) {
        const int b = i - lead;
        if (b >= 0 && b < p.seq_len) {
          const int pos0 = p.seq_off + b, pp = p.strand ? len - 1 - pos0 : pos0;
          uint32_t code = (w[pp >> 4] >> ((pp & 15) * 2)) & 3u;
          if (p.strand) code = 3u - code;
          if ((nm[pp >> 5] >> (pp & 31)) & 1u) code = 4u;
          ch = NT[code];
        } else {
          const int e = b < 0 ? n_ext - 1 - i : b - p.seq_len;   // reverse(ext) * seq  |  seq * ext
          ch = NT[(x->ext2[e >> 4] >> (2 * (e & 15))) & 3u];
        }
      }
      s[i] = ch;
    }
    // quality: 'G' x lead | base part | 'G' x (n_ext - lead), sliced to the CDR3 range when one was found
    const int qfull = base_q + n_ext;
    const int real_lo = lead, real_hi = lead + (p.kind == 1 ? min(base_q, p.q_start1 - 1) : base_q);
    const int src0 = p.kind == 1 ? 0 : p.seq_off;
    const int a0 = c_off >= 0 ? c_off : 0;
    const char* rq = S.quals ? (S.q_entry_off ? S.quals + S.q_entry_off[e_] : S.quals + S.seq_off[p.read]) : nullptr;
    char* q = qual_arena + qoff[oi];
    for (int i = lane; i <= qlen; i += 32) {
      char ch = 0;
      if (i < qlen) {
        const int fi = a0 + i;
        ch = 'G';
        if (fi >= real_lo && fi < real_hi) {
          const int j = src0 + (fi - real_lo);
          ch = rq ? (p.strand ? rq[len - 1 - j] : rq[j]) : 'I';
        }
      }
      q[i] = ch;
    }
    if (lane == 0) {
      catt_read rd;
      rd.seq = host_seq + soff[oi]; rd.qual = host_qual + qoff[oi]; rd.seq_len = slen; rd.qual_len = qlen;
      rd.v_id = p.v_id; rd.j_id = p.j_id; rd.cdr3_off = c_off; rd.cdr3_len = c_off >= 0 ? c_len : 0;
      rd.ordinal = (int32_t)(S.ordinal_base + p.read); rd.able = 1;
      rd.flags = (uint8_t)((p.kind == 0 ? 1 : p.kind == 1 ? 2 : 4) | ((c_off >= 0 && c_off + c_len > qfull) ? 8 : 0));   // 8: reference BoundsError
      rd.pad_[0] = 0; rd.pad_[1] = 0;
      part[oe] = rd;
    }
  }
}

inline unsigned nblk(int64_t n, int b = 256) { return (unsigned)std::max<int64_t>(1, (n + b - 1) / b); }

template <class T> int ensure(catt_ctx* c, int slot, int64_t count, T** out) {
  int rc = c->sb[slot].ensure((size_t)std::max<int64_t>(count, 1) * sizeof(T));
  *out = c->sb[slot].as<T>();
  return rc;
}

int ensure_cub(catt_ctx* c, size_t need) { return c->sb[SB_CUB].ensure(std::max<size_t>(need, 256)); }

int scan_u32(catt_ctx* c, const uint32_t* in, uint32_t* out, int64_t count, cudaStream_t st) {
  size_t need = 0;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, count, st));
  int rc = ensure_cub(c, need);
  if (rc) return rc;
  need = c->sb[SB_CUB].cap;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(c->sb[SB_CUB].p, need, in, out, count, st));
  return CATT_OK;
}

int scan_u64(catt_ctx* c, const unsigned long long* in, unsigned long long* out, int64_t count, cudaStream_t st) {
  size_t need = 0;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, count, st));
  int rc = ensure_cub(c, need);
  if (rc) return rc;
  need = c->sb[SB_CUB].cap;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(c->sb[SB_CUB].p, need, in, out, count, st));
  return CATT_OK;
}

}  // namespace

// ---- reusable host blocks (see host_types.h) -------------------------------------------------------------------------------
namespace {
struct BlockCache {
  std::mutex mu;
  struct Ent { void* p; size_t cap; };
  std::vector<Ent> free_list;
  size_t held = 0;
  ~BlockCache() { for (auto& e : free_list) free(e.p); }     // process exit: the CUDA context may be gone, no unregister
};
BlockCache g_blocks;
constexpr size_t BLOCK_SMALL = 1 << 20;            // below this plain malloc/free is fine
constexpr size_t BLOCK_KEEP_MAX = (size_t)6 << 30;  // never sit on more than this much freed memory
constexpr size_t BLOCK_PIN_MAX = (size_t)1 << 30;   // larger blocks are not pinned (pinning costs more than it saves per call)
}  // namespace

void* block_get(size_t bytes, size_t* cap) {
  if (bytes < BLOCK_SMALL) { *cap = bytes; return malloc(std::max<size_t>(bytes, 1)); }
  size_t pin_max = BLOCK_PIN_MAX;
  if (const char* e = getenv("CATT_PIN_MAX")) pin_max = (size_t)atoll(e);      // tests force the ring path on small results
  if (bytes > pin_max) {                                 // too large to pin per call: filled through the pinned ring (d2h_ring)
    const size_t want = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    void* p = nullptr;
    if (posix_memalign(&p, 2u << 20, want) != 0) return nullptr;
    madvise(p, want, MADV_HUGEPAGE);
    *cap = want | 1;                                     // odd capacity = "not registered, not cached"
    return p;
  }
  {
    std::lock_guard<std::mutex> lk(g_blocks.mu);
    size_t best = SIZE_MAX;
    for (size_t i = 0; i < g_blocks.free_list.size(); i++) {
      const size_t c = g_blocks.free_list[i].cap;
      if (c >= bytes && c <= bytes * 2 + (64u << 20) && (best == SIZE_MAX || c < g_blocks.free_list[best].cap)) best = i;
    }
    if (best != SIZE_MAX) {
      const BlockCache::Ent e = g_blocks.free_list[best];
      g_blocks.free_list.erase(g_blocks.free_list.begin() + best);
      g_blocks.held -= e.cap;
      *cap = e.cap;
      return e.p;
    }
  }
  const size_t want = (bytes + bytes / 8 + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
  void* p = nullptr;
  if (posix_memalign(&p, 2u << 20, want) != 0) return nullptr;
  madvise(p, want, MADV_HUGEPAGE);
  memset(p, 0, want);                                    // touch once, then pin: the arenas arrive as bulk device->host copies
  if (cudaHostRegister(p, want, cudaHostRegisterPortable) != cudaSuccess) cudaGetLastError();   // unpinned still works, slower
  *cap = want;
  return p;
}

void block_put(void* p, size_t cap) {
  if (!p) return;
  if (cap < BLOCK_SMALL || (cap & 1)) { free(p); return; }
  std::lock_guard<std::mutex> lk(g_blocks.mu);
  if (g_blocks.held + cap > BLOCK_KEEP_MAX || g_blocks.free_list.size() >= 16) {
    if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
    free(p);
    return;
  }
  g_blocks.free_list.push_back({p, cap});
  g_blocks.held += cap;
}

// device -> host copy of one result arena.  Pinned (cached) blocks take one bulk copy; larger, unpinned blocks are filled through
// two pinned 64 MiB ring buffers: the copy engine fills one while the host threads empty the other into the destination.
static int d2h_arena(catt_ctx* c, void* dst, size_t cap, const void* d_src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return CATT_OK;
  if (!(cap & 1) || bytes < BLOCK_SMALL) { CATT_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st)); return CATT_OK; }
  const size_t piece = (size_t)64 << 20;
  int rc;
  if ((rc = c->pin[PIN_RING0].ensure(piece)) || (rc = c->pin[PIN_RING1].ensure(piece))) return rc;
  const size_t np = (bytes + piece - 1) / piece;
  auto issue = [&](size_t i) -> int {
    const size_t o = i * piece, len = std::min(piece, bytes - o);
    CATT_CUDA(cudaMemcpyAsync(c->pin[PIN_RING0 + (i & 1)].p, (const char*)d_src + o, len, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaEventRecord(c->ev_q[i & 1], st));
    return CATT_OK;
  };
  auto drain = [&](size_t i) -> int {
    const size_t o = i * piece, len = std::min(piece, bytes - o);
    CATT_CUDA(cudaEventSynchronize(c->ev_q[i & 1]));
    const char* src = c->pin[PIN_RING0 + (i & 1)].as<char>();
    const int64_t sub = 1 << 20, ns = (int64_t)((len + sub - 1) / sub);
    #pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < ns; k++) memcpy((char*)dst + o + k * sub, src + k * sub, (size_t)std::min<int64_t>(sub, (int64_t)len - k * sub));
    return CATT_OK;
  };
  if ((rc = issue(0))) return rc;
  for (size_t i = 1; i < np; i++) { if ((rc = issue(i)) || (rc = drain(i - 1))) return rc; }
  return drain(np - 1);
}

int upload_names(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, NamesDev* out, catt_info* info, cudaStream_t stream) {
  cudaStream_t st = stream ? stream : c->stream;
  *out = NamesDev{};
  if (n == 0) return CATT_OK;
  char* d_names; int64_t* d_off;
  int rc;
  // name_off may be a view into a larger sample (one rank's block of a GPU group): only this block's bytes are uploaded, the
  // device copy is addressed with the caller's absolute offsets
  const int64_t base = name_off[0], name_bytes = name_off[n] - base;
  if ((rc = ensure(c, SB_NAMES, name_bytes + 16, &d_names)) || (rc = ensure(c, SB_NAME_OFF, n + 1, &d_off))) return rc;
  CATT_CUDA(cudaMemcpyAsync(d_names, names + base, (size_t)name_bytes, cudaMemcpyHostToDevice, st));
  CATT_CUDA(cudaMemcpyAsync(d_off, name_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (info) info->bytes_h2d += name_bytes + (n + 1) * 8;
  out->s = d_names - base; out->beg = d_off; out->end = d_off + 1;
  return CATT_OK;
}

// ---- per-file ordering: names -> SAM order -> QNAME dedup -> share_name -> work items --------------------------------------
// (read_alignrs, catt.jl:210-326, up to the point where records become Myreads)
struct OrderScratch {
  uint32_t* slot_of; QClass* cls; uint64_t cap;
  unsigned long long *keys[2], *sorted[2];
  uint32_t *ord_idx[2], *ord_flag[2], *opos[2], *pair_a, *pair_b;
};
struct OrderOut { int64_t F[2], n_v, n_p, n_j; };

static int order_alloc(catt_ctx* c, int64_t max_reads, int64_t max_v, int64_t max_j, OrderScratch* S) {
  int rc;
  uint64_t cap = 1024;
  while (cap < (uint64_t)std::min<int64_t>(max_reads, max_v + max_j) * 2 + 2) cap <<= 1;      // one class per mapped read at most
  if (cap > 0x80000000ull) { set_error("too many mapped reads for the QNAME table"); return CATT_E_LIMIT; }
  S->cap = cap;
  if ((rc = ensure(c, SB_SLOT_OF, max_reads, &S->slot_of)) || (rc = ensure(c, SB_OWNER, (int64_t)cap, &S->cls))) return rc;
  if ((rc = ensure(c, SB_KEYS_A, max_v + 1, &S->keys[0])) || (rc = ensure(c, SB_MINV, max_j + 1, &S->keys[1])) || (rc = ensure(c, SB_SORT_V, max_v + 1, &S->sorted[0])) ||
      (rc = ensure(c, SB_SORT_J, max_j + 1, &S->sorted[1])) || (rc = ensure(c, SB_ORD_IDX_V, max_v + 1, &S->ord_idx[0])) ||
      (rc = ensure(c, SB_ORD_IDX_J, max_j + 1, &S->ord_idx[1])) || (rc = ensure(c, SB_ORD_FLAG_V, max_v + 1, &S->ord_flag[0])) ||
      (rc = ensure(c, SB_ORD_FLAG_J, max_j + 1, &S->ord_flag[1])) ||
      (rc = ensure(c, SB_OPOS_V, max_v + 1, &S->opos[0])) || (rc = ensure(c, SB_OPOS_J, max_j + 1, &S->opos[1])) ||
      (rc = ensure(c, SB_PAIR_A, max_j + 1, &S->pair_a)) || (rc = ensure(c, SB_PAIR_B, max_j + 1, &S->pair_b))) return rc;
  return CATT_OK;
}

// One input file: nf reads (records rec[0/1], names nd, lengths len, all indexed 0..nf), nm[] of them mapped per reference set.
// Writes the file's work items [V part | both-mapped | J part] in the reference's processing order to `seg` (read indices
// = read_base + index), updates k (args["kmer"], catt.jl:156-172) and the_ERR / the log counters in `info`.
static int order_file(catt_ctx* c, const OrderScratch& S, int64_t nf, const NamesDev& nd, const catt_aln* const rec[2], const int32_t* len,
                      const int64_t nm[2], uint32_t read_base, int* kmer, catt_info& info, unsigned long long* cnt, unsigned long long* hcnt,
                      ExtractItem* seg, int64_t seg_cap, OrderOut* out, int* launches) {
  cudaStream_t st = c->stream;
  int rc;
  int& k = *kmer;
  CATT_CUDA(cudaMemsetAsync(cnt, 0, C_COUNT * sizeof(unsigned long long), st));
  if (nm[0] + nm[1] > 0) {
    CATT_CUDA(cudaMemsetAsync(S.cls, 0, S.cap * sizeof(QClass), st));
    name_kernel<<<nblk(nf), 256, 0, st>>>(nf, nd, rec[0], rec[1], S.cls, (uint32_t)(S.cap - 1), S.slot_of);
    kept_kernel<<<nblk(nf), 256, 0, st>>>(nf, rec[0], rec[1], len, S.slot_of, S.cls, S.keys[0], S.keys[1], cnt);
    CATT_CUDA(cudaGetLastError());
    *launches += 2;
  }
  // per reference set: the final SAM order of the kept records, then their places in the reference's processing order
  for (int which = 0; which < 2; which++) {
    if (nm[which] == 0) continue;
    size_t need = 0;
    CATT_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, need, S.keys[which], S.sorted[which], nm[which], 0, 61, st));
    if ((rc = ensure_cub(c, need))) return rc;
    need = c->sb[SB_CUB].cap;
    CATT_CUDA(cub::DeviceRadixSort::SortKeysDescending(c->sb[SB_CUB].p, need, S.keys[which], S.sorted[which], nm[which], 0, 61, st));
    CATT_CUDA(cudaMemsetAsync(S.ord_flag[which], 0, (size_t)(nm[which] + 1) * sizeof(uint32_t), st));
    order_kernel<<<nblk(nm[which]), 256, 0, st>>>(nm[which], S.sorted[which], cnt + C_F_V + which, c->nthreads, which, nd, S.ord_idx[which], S.ord_flag[which],
                                                  S.pair_a, cnt);
    CATT_CUDA(cudaGetLastError());
    if ((rc = scan_u32(c, S.ord_flag[which], S.opos[which], nm[which] + 1, st))) return rc;
    *launches += 12;     // radix sort (histogram + 7 onesweep passes), order, scan
  }
  collect_kernel<<<1, 1, 0, st>>>(nm[0], nm[1], S.opos[0], S.opos[1], cnt);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  CATT_CUDA(cudaMemcpyAsync(hcnt, cnt, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  if ((int64_t)hcnt[C_NMAP_V] != nm[0] || (int64_t)hcnt[C_NMAP_J] != nm[1]) {
    set_error("stage B: mapped-record count mismatch (V %llu vs %lld, J %llu vs %lld)", hcnt[C_NMAP_V], (long long)nm[0], hcnt[C_NMAP_J], (long long)nm[1]);
    return CATT_E_ARG;
  }
  const int64_t F[2] = {(int64_t)hcnt[C_F_V], (int64_t)hcnt[C_F_J]};
  const int64_t n_v = (int64_t)hcnt[C_NV], n_j = (int64_t)hcnt[C_NJ], n_p = (int64_t)hcnt[C_NP];
  info.v_final += F[0]; info.j_final += F[1]; info.share += n_p;
  // get_kmer_fromsam / get_err_fromsam on this file's final V list, as samtools stats prints them
  {
    const uint64_t tot = hcnt[C_SUM_LEN], snm = hcnt[C_SUM_NM], smi = hcnt[C_SUM_MI];
    const float avg = F[0] == 0 ? 0.f : (float)(tot / (uint64_t)F[0]);       // samtools: integer division, printed %.0f
    char buf[64];
    snprintf(buf, sizeof buf, "%.0f", avg);
    const double avg_r = atof(buf);
    info.avg_len = avg_r;
    if (k == CATT_KMER_AUTO) {
      if (avg_r >= 50) k = 15;
      if (avg_r >= 75) k = 19;
      if (avg_r >= 100) k = 23;
      if (avg_r >= 150) k = 25;
    }
    const float er = smi ? (float)snm / (float)smi : 0.f;
    snprintf(buf, sizeof buf, "%e", (double)er);
    info.the_err = atof(buf);
    if (!(info.the_err == info.the_err)) info.the_err = 0.0;
  }
  // both-mapped pairs in QNAME order: LSD radix over 8-byte digits of the trimmed name
  uint32_t* pairs = S.pair_a;
  if (n_p > 1) {
    unsigned long long *pk_a, *pk_b;
    if ((rc = ensure(c, SB_PKEY_A, n_p, &pk_a)) || (rc = ensure(c, SB_PKEY_B, n_p, &pk_b))) return rc;
    const int digits = std::max(1, (int)((hcnt[C_MAXNAME] + 7) / 8));
    uint32_t* other = S.pair_b;
    for (int d = digits - 1; d >= 0; d--) {
      pair_key_kernel<<<nblk(n_p), 256, 0, st>>>(n_p, pairs, nd, d, pk_a);
      CATT_CUDA(cudaGetLastError());
      // the last digit of the longest name holds fewer than 8 characters: the zero padding below them needs no passes
      const int lo_bit = 64 - 8 * (int)std::min<unsigned long long>(8, std::max<unsigned long long>(hcnt[C_MAXNAME], 8ull * (unsigned)d + 1) - 8ull * (unsigned)d);
      size_t need = 0;
      CATT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, pk_a, pk_b, pairs, other, n_p, lo_bit, 64, st));
      if ((rc = ensure_cub(c, need))) return rc;
      need = c->sb[SB_CUB].cap;
      CATT_CUDA(cub::DeviceRadixSort::SortPairs(c->sb[SB_CUB].p, need, pk_a, pk_b, pairs, other, n_p, lo_bit, 64, st));
      std::swap(pairs, other);
      *launches += 10;
    }
  }
  // work items in the reference's processing order: [V part | both-mapped | J part] of this file
  const int64_t nseg = n_v + n_p + n_j;
  if (nseg > seg_cap) { set_error("stage B: item count exceeds the mapped-record bound"); return CATT_E_ARG; }
  if (F[0] > 0) {
    part_item_kernel<<<nblk(F[0]), 256, 0, st>>>(F[0], S.ord_idx[0], S.ord_flag[0], S.opos[0], rec[0], 0, read_base, seg);
    CATT_CUDA(cudaGetLastError());
    *launches += 1;
  }
  if (n_p > 0) {
    pair_item_kernel<<<nblk(n_p), 256, 0, st>>>(n_p, pairs, S.slot_of, S.cls, rec[0], rec[1], read_base, seg + n_v);
    CATT_CUDA(cudaGetLastError());
    *launches += 1;
  }
  if (F[1] > 0) {
    part_item_kernel<<<nblk(F[1]), 256, 0, st>>>(F[1], S.ord_idx[1], S.ord_flag[1], S.opos[1], rec[1], 1, read_base, seg + n_v + n_p);
    CATT_CUDA(cudaGetLastError());
    *launches += 1;
  }
  out->F[0] = F[0]; out->F[1] = F[1]; out->n_v = n_v; out->n_p = n_p; out->n_j = n_j;
  return CATT_OK;
}

// The ordering stage alone (parity hook): n reads of ONE file with their records already in c->recs_all and QNAMEs `nd`; the work
// items come back as (kind, read, read2) triples in processing order, counts = {final V, final J, V-only, pairs, J-only}.
int order_only(catt_ctx* c, int64_t n, const NamesDev& nd, const int32_t* d_len, int64_t nm_v, int64_t nm_j, int32_t* h_items, int64_t* n_items, int64_t* counts) {
  int rc, launches = 0;
  unsigned long long* cnt;
  if ((rc = ensure(c, SB_CNT, C_COUNT, &cnt))) return rc;
  unsigned long long* hcnt = c->pin[PIN_CNT].as<unsigned long long>() + 2 * N_COUNTERS;
  OrderScratch OS;
  if ((rc = order_alloc(c, n, nm_v, nm_j, &OS))) return rc;
  ExtractItem* items;
  if ((rc = ensure(c, SB_ITEMS, nm_v + nm_j, &items))) return rc;
  const catt_aln* rec[2] = {c->recs_all[0].as<catt_aln>(), c->recs_all[1].as<catt_aln>()};
  const int64_t nm[2] = {nm_v, nm_j};
  catt_info info{};
  int k = c->kmer;
  OrderOut oo;
  if ((rc = order_file(c, OS, n, nd, rec, d_len, nm, 0u, &k, info, cnt, hcnt, items, nm_v + nm_j, &oo, &launches))) return rc;
  const int64_t t = oo.n_v + oo.n_p + oo.n_j;
  std::vector<ExtractItem> h((size_t)std::max<int64_t>(t, 1));
  if (t) CATT_CUDA(cudaMemcpyAsync(h.data(), items, (size_t)t * sizeof(ExtractItem), cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < t; i++) { h_items[3 * i] = h[i].kind; h_items[3 * i + 1] = (int32_t)h[i].read; h_items[3 * i + 2] = (int32_t)h[i].read2; }
  *n_items = t;
  counts[0] = oo.F[0]; counts[1] = oo.F[1]; counts[2] = oo.n_v; counts[3] = oo.n_p; counts[4] = oo.n_j;
  return CATT_OK;
}

// only_cdr3 with host buffers: gather the quality strings of the E surviving reads (read indices d_rid, local to h_quals) on the
// host and upload them; S.quals / S.q_entry_off then index them per output entry
static int gather_host_quals(catt_ctx* c, const DeviceQuals& dq, const uint32_t* d_rid, int64_t E, StrSrc* S, catt_info& info, cudaStream_t st) {
  int rc;
  char* d_gather; int64_t* d_entry;
  if ((rc = c->pin[PIN_QRID].ensure((size_t)E * 4)) || (rc = c->pin[PIN_QENTRY].ensure((size_t)(E + 1) * 8))) return rc;
  uint32_t* h_rid = c->pin[PIN_QRID].as<uint32_t>();
  int64_t* h_entry = c->pin[PIN_QENTRY].as<int64_t>();
  CATT_CUDA(cudaMemcpyAsync(h_rid, d_rid, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  h_entry[0] = 0;
  for (int64_t e = 0; e < E; e++) h_entry[e + 1] = h_entry[e] + (dq.host_seq_off[h_rid[e] + 1] - dq.host_seq_off[h_rid[e]]);
  const int64_t gbytes = h_entry[E];
  if ((rc = c->pin[PIN_QGATHER].ensure((size_t)gbytes + 16)) || (rc = ensure(c, SB_QGATHER, gbytes + 16, &d_gather)) ||
      (rc = ensure(c, SB_QENTRY, E + 1, &d_entry))) return rc;
  char* h_gather = c->pin[PIN_QGATHER].as<char>();
  #pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; e++)
    memcpy(h_gather + h_entry[e], dq.host_quals + dq.host_seq_off[h_rid[e]], (size_t)(h_entry[e + 1] - h_entry[e]));
  CATT_CUDA(cudaMemcpyAsync(d_gather, h_gather, (size_t)gbytes, cudaMemcpyHostToDevice, st));
  CATT_CUDA(cudaMemcpyAsync(d_entry, h_entry, (size_t)(E + 1) * 8, cudaMemcpyHostToDevice, st));
  info.bytes_h2d += gbytes + (E + 1) * 8; info.bytes_d2h += E * 4;
  S->quals = d_gather; S->q_entry_off = d_entry;
  return CATT_OK;
}

int stage_b(catt_ctx* c, int64_t n, const NamesDev& names, const ReadsDev& reads, const DeviceQuals& dq, const FileHits& fh, catt_result* R,
            int* launches) {
  catt_info& info = R->info;
  info.kmer = c->kmer;
  if (n == 0) return CATT_OK;
  cudaStream_t st = c->stream;
  Trace tr;
  int rc;
  unsigned long long* cnt; unsigned long long* hcnt;
  if ((rc = ensure(c, SB_CNT, C_COUNT, &cnt))) return rc;
  hcnt = c->pin[PIN_CNT].as<unsigned long long>() + 2 * N_COUNTERS;      // the first 2*N_COUNTERS entries belong to Stage A
  const int nfiles = (int)fh.start.size() - 1;
  int64_t tot_hits = 0, max_v = 0, max_j = 0, max_reads = 0;
  for (int f = 0; f < nfiles; f++) {
    tot_hits += fh.v[f] + fh.j[f];
    max_v = std::max(max_v, fh.v[f]); max_j = std::max(max_j, fh.j[f]); max_reads = std::max(max_reads, fh.start[f + 1] - fh.start[f]);
  }
  // ---- per-file QNAME classes; work items of all files accumulate in `items` (n_items <= mapped records)
  OrderScratch OS;
  if ((rc = order_alloc(c, max_reads, max_v, max_j, &OS))) return rc;
  ExtractItem* items; ExtractOut* eout; uint32_t *kflag0, *kflag1, *kpos0, *kpos1, *pflag, *xpos;
  if ((rc = ensure(c, SB_ITEMS, tot_hits, &items)) || (rc = ensure(c, SB_EOUT, tot_hits, &eout)) || (rc = ensure(c, SB_KFLAG, tot_hits + 1, &kflag0)) ||
      (rc = ensure(c, SB_KFLAG1, tot_hits + 1, &kflag1)) || (rc = ensure(c, SB_KPOS2, tot_hits + 1, &kpos0)) || (rc = ensure(c, SB_KPOS3, tot_hits + 1, &kpos1)) ||
      (rc = ensure(c, SB_PFLAG, tot_hits + 1, &pflag)) || (rc = ensure(c, SB_XPOS, tot_hits + 1, &xpos))) return rc;
  int64_t n_items = 0;
  int k = c->kmer;
  float ms_order = 0, ms_extract = 0;

  // ---- read_alignrs, once per input file (catt.jl:210-326): the SAM order, the QNAME dedup, share_name and the both-mapped
  //      pairing are per file; args["kmer"] is fixed by the first file that reaches a mean length of 50, the_ERR by the last
  for (int f = 0; f < nfiles; f++) {
    const int64_t lo = fh.start[f], nf = fh.start[f + 1] - lo;
    const int64_t nm[2] = {fh.v[f], fh.j[f]};
    const catt_aln* rec[2] = {c->recs_all[0].as<catt_aln>() + lo, c->recs_all[1].as<catt_aln>() + lo};
    const NamesDev nd{names.s, names.beg + lo, names.end + lo};
    CATT_CUDA(cudaEventRecord(c->ev[5], st));
    OrderOut oo;
    if ((rc = order_file(c, OS, nf, nd, rec, reads.len + lo, nm, (uint32_t)lo, &k, info, cnt, hcnt, items + n_items, tot_hits - n_items, &oo, launches))) return rc;
    const int64_t nseg = oo.n_v + oo.n_p + oo.n_j;
    CATT_CUDA(cudaEventRecord(c->ev[6], st));
    // K5: assignV/J, real_score, clipping, finder! with the k that is current for this file (catt.jl:279,301)
    if ((rc = launch_extract(c->V.dev, c->J.dev, reads, c->d_finder, items + n_items, eout + n_items, nseg, k, c->allow_v, c->allow_j, c->sm_count, st, launches))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[7], st));
    CATT_CUDA(cudaStreamSynchronize(st));
    { float a = 0, b = 0; cudaEventElapsedTime(&a, c->ev[5], c->ev[6]); cudaEventElapsedTime(&b, c->ev[6], c->ev[7]); ms_order += a; ms_extract += b; }
    n_items += nseg;
  }
  info.kmer = k;
  tr.lap("B: names + order + dedup + extract");

  // ---- Vpart = kept V part reads and kept both-mapped reads, file by file; Jpart = kept J part reads
  PoolItem* pitems; OutDesc* desc; uint32_t* partial_idx;
  if ((rc = ensure(c, SB_PITEMS, n_items, &pitems)) || (rc = ensure(c, SB_DESC, n_items, &desc)) || (rc = ensure(c, SB_PARTIAL, n_items, &partial_idx))) return rc;
  CATT_CUDA(cudaMemsetAsync(cnt, 0, C_COUNT * sizeof(unsigned long long), st));
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  keep_kernel<<<nblk(n_items + 1), 256, 0, st>>>(n_items, items, eout, kflag0, kflag1, pflag, cnt);
  CATT_CUDA(cudaGetLastError());
  if ((rc = scan_u32(c, kflag0, kpos0, n_items + 1, st)) || (rc = scan_u32(c, kflag1, kpos1, n_items + 1, st))) return rc;
  pitem_kernel<<<nblk(std::max<int64_t>(n_items, 1)), 256, 0, st>>>(n_items, items, eout, kflag0, kflag1, kpos0, kpos1, c->J.dev.rec_len, pitems, desc, pflag, cnt);
  CATT_CUDA(cudaGetLastError());
  if ((rc = scan_u32(c, pflag, xpos, n_items + 1, st))) return rc;
  partial_kernel<<<nblk(std::max<int64_t>(n_items, 1)), 256, 0, st>>>(n_items, pflag, xpos, partial_idx, desc, cnt);
  CATT_CUDA(cudaGetLastError());
  *launches += 9;
  CATT_CUDA(cudaEventRecord(c->ev[6], st));
  CATT_CUDA(cudaMemcpyAsync(hcnt, cnt, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  { float a = 0; cudaEventElapsedTime(&a, c->ev[5], c->ev[6]); info.ms_order += ms_order + a; info.ms_extract += ms_extract; }
  tr.lap("B: lists");
  const int64_t M = (int64_t)hcnt[C_M], n_partial = (int64_t)hcnt[C_NPARTIAL];
  info.vpart = (int64_t)hcnt[C_VPART]; info.jpart = M - info.vpart; info.full_pushed = (int64_t)hcnt[C_FULL];
  info.pair_seq_mismatch = (int64_t)hcnt[C_MISMATCH];
  info.left = n_partial; info.direct = M - n_partial;

  // ---- K6/K7: k-mer table + extension of the partial reads
  ExtendOut* xout = nullptr;
  const bool run_pool = M > 0 && k <= KMER_MAX && n_partial > 0;
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  if (run_pool) {
    const uint64_t slots = std::max<uint64_t>(1024, (uint64_t)M * 11ull * 2ull);
    KmerTable tab{};
    if ((rc = ensure(c, SB_TAB_KEYS, (int64_t)slots, &tab.slots)) || (rc = ensure(c, SB_XOUT, n_partial, &xout))) return rc;
    tab.sub_slots = slots; tab.nsub = 1;
    CATT_CUDA(cudaMemsetAsync(tab.slots, 0xff, slots * sizeof(ulonglong2), st));
    CATT_CUDA(cudaEventRecord(c->ev[0], st));
    if ((rc = launch_pool(reads, pitems, M, k, tab, c->sm_count, st, launches))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[1], st));
    if ((rc = launch_extend(reads, c->d_finder, pitems, partial_idx, n_partial, k, tab, xout, reinterpret_cast<uint32_t*>(cnt + C_RESCUED), c->sm_count, st, launches))) return rc;
  }
  CATT_CUDA(cudaEventRecord(c->ev[6], st));
  // ---- the Myread strings: sizes -> offsets (scans) -> one warp per read writes both strings and the catt_read entry; the
  //      arenas go to the host as three bulk copies into pinned result blocks
  StrSrc S{desc, xout, dq.quals, dq.seq_off, M, info.vpart, nullptr, M, nullptr};
  int64_t E = M, EV = info.vpart;
  if (c->only_cdr3 && M > 0) {                      // the lists as they stand after catt.jl:581-582
    uint32_t *cflag, *cpos, *cmap;
    if ((rc = ensure(c, SB_CFLAG, M + 1, &cflag)) || (rc = ensure(c, SB_CPOS, M + 1, &cpos)) || (rc = ensure(c, SB_CMAP, M, &cmap))) return rc;
    cdr3_flag_kernel<<<nblk(M + 1), 256, 0, st>>>(desc, xout, M, cflag);
    CATT_CUDA(cudaGetLastError());
    if ((rc = scan_u32(c, cflag, cpos, M + 1, st))) return rc;
    cdr3_map_kernel<<<nblk(M), 256, 0, st>>>(cflag, cpos, M, info.vpart, cmap, cnt);
    CATT_CUDA(cudaGetLastError());
    *launches += 4;
    unsigned long long ev2[2];
    CATT_CUDA(cudaMemcpyAsync(ev2, cnt + C_E, sizeof ev2, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    E = (int64_t)ev2[0]; EV = (int64_t)ev2[1];
    S.map = cmap; S.E = E;
    if (dq.host_quals && E > 0) {                   // gather the quality strings of the E surviving reads on the host, upload them
      uint32_t* d_rid;
      if ((rc = ensure(c, SB_QRID, E, &d_rid))) return rc;
      entry_read_kernel<<<nblk(E), 256, 0, st>>>(desc, cmap, E, d_rid);
      CATT_CUDA(cudaGetLastError());
      *launches += 1;
      if ((rc = gather_host_quals(c, dq, d_rid, E, &S, info, st))) return rc;
    }
  }
  unsigned long long *ssz, *qsz, *soff, *qoff;
  if ((rc = ensure(c, SB_SSZ, E + 1, &ssz)) || (rc = ensure(c, SB_QSZ, E + 1, &qsz)) || (rc = ensure(c, SB_SOFF, E + 1, &soff)) ||
      (rc = ensure(c, SB_QOFF, E + 1, &qoff))) return rc;
  strsize_kernel<<<nblk(E + 1), 256, 0, st>>>(S, ssz, qsz);
  CATT_CUDA(cudaGetLastError());
  if ((rc = scan_u64(c, ssz, soff, E + 1, st)) || (rc = scan_u64(c, qsz, qoff, E + 1, st))) return rc;
  *launches += 5;
  CATT_CUDA(cudaMemcpyAsync(hcnt, cnt, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(hcnt + C_COUNT, soff + E, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(hcnt + C_COUNT + 1, qoff + E, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  { float a = 0; cudaEventElapsedTime(&a, c->ev[5], c->ev[6]); info.ms_pool += a; }
  if (run_pool) { float a = 0; cudaEventElapsedTime(&a, c->ev[0], c->ev[1]); info.ms_pool_kernel += a; }
  info.rescued += (int64_t)(uint32_t)hcnt[C_RESCUED];
  tr.lap("B: pool + extend + string sizes");
  const double t2 = now_ms();
  const size_t seq_bytes = (size_t)hcnt[C_COUNT] + 1, qual_bytes = (size_t)hcnt[C_COUNT + 1] + 1, part_bytes = (size_t)std::max<int64_t>(E, 1) * sizeof(catt_read);
  R->seq_arena = (char*)block_get(seq_bytes, &R->cap_seq);
  R->qual_arena = (char*)block_get(qual_bytes, &R->cap_qual);
  R->part[0] = (catt_read*)block_get(part_bytes, &R->cap_part[0]);
  if (!R->seq_arena || !R->qual_arena || !R->part[0]) { set_error("out of host memory for the result"); return CATT_E_LIMIT; }
  R->npart[0] = EV; R->npart[1] = E - EV;
  R->part[1] = R->part[0] + EV;                     // one block: Vpart entries, then Jpart entries
  info.ms_host += now_ms() - t2;
  tr.lap("B: result blocks");
  if (E > 0) {
    char *d_seq, *d_qual; catt_read* d_part;
    if ((rc = ensure(c, SB_OUT_SEQ, (int64_t)seq_bytes, &d_seq)) || (rc = ensure(c, SB_OUT_QUAL, (int64_t)qual_bytes, &d_qual)) ||
        (rc = ensure(c, SB_OUT_PART, E, &d_part))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[5], st));
    const int grid = (int)std::min<int64_t>((E + MW - 1) / MW, (int64_t)c->sm_count * 16);
    materialise_kernel<<<grid, MW * 32, 0, st>>>(S, reads, soff, qoff, d_seq, d_qual, d_part, R->seq_arena, R->qual_arena);
    CATT_CUDA(cudaGetLastError());
    *launches += 1;
    CATT_CUDA(cudaEventRecord(c->ev[6], st));
    if ((rc = d2h_arena(c, R->seq_arena, R->cap_seq, d_seq, seq_bytes - 1, st)) || (rc = d2h_arena(c, R->qual_arena, R->cap_qual, d_qual, qual_bytes - 1, st)) ||
        (rc = d2h_arena(c, R->part[0], R->cap_part[0], d_part, (size_t)E * sizeof(catt_read), st))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[7], st));
    CATT_CUDA(cudaStreamSynchronize(st));
    info.bytes_d2h += (int64_t)(seq_bytes + qual_bytes - 2) + E * (int64_t)sizeof(catt_read) + (2 + nfiles) * C_COUNT * 8;
    { float a = 0, b2 = 0; cudaEventElapsedTime(&a, c->ev[5], c->ev[6]); cudaEventElapsedTime(&b2, c->ev[6], c->ev[7]); info.ms_strings += a; info.ms_d2h += b2; }
    if (c->want_clonotypes && (rc = build_clonotypes(c, E, d_part, d_seq, d_qual, R->seq_arena, R->qual_arena, R, launches))) return rc;
  }
  tr.lap("B: strings + D2H");
  return CATT_OK;
}


// =================================================================================================================================
// One sample over N GPUs.  Every rank owns a contiguous block of the reads, has run Stage A on it (records in
// c->recs_all, packed reads in c->rd_*) and now takes part in ONE global Stage B:
//   1. the records + QNAMEs of the MAPPED reads are all-gathered (64 B + name per mapped read) and every rank derives the same
//      global SAM order / QNAME dedup / share_name / pairing with the single-GPU code above (catt.jl:45-48, 221-237, 288-297):
//      ~0.7 ms per million reads, cheaper than any exchange that would shard it;
//   2. each work item is processed by the rank that owns its read: assignV/J, real_score, finder! (extract_kernel) run sharded;
//      a both-mapped pair whose two records sit on different ranks (mates far apart in the file) gets the other read's packed
//      words through one sparse all-reduce;
//   3. one byte per item (kept / list / partial) is all-reduced, so every rank knows every read's position in [Vpart; Jpart];
//   4. the k-mer table is global and last-writer-wins (catt.jl:426-446, 481-497): every rank buckets its (k-mer, order << 8 | pos)
//      pairs by the owner of the key (hash % N), one all-to-all moves them, the owners insert with atomicMax, the sub-tables are
//      all-gathered, and the extension (catt.jl:341-424, 552-577) runs sharded against the replicated table;
//   5. every rank writes the strings of ITS reads at their global offsets into zeroed arenas; one sum-reduce to the root merges
//      them (each byte has exactly one writer).
// =================================================================================================================================
namespace {

struct GRec { catt_aln v, j; uint32_t gidx; int32_t len; int32_t name_len; int32_t pad_; };    // one mapped read on the wire (80 B)
struct RankBounds { long long b[17]; int n; };
__device__ __forceinline__ int owner_of(const RankBounds& rb, long long g) {
  int r = 0;
  while (r + 1 < rb.n && g >= rb.b[r + 1]) r++;
  return r;
}
enum { SH_FLAG = 0, SH_POS, SH_NLEN, SH_NPOS, SH_GREC_LOC, SH_NAMES_LOC, SH_GREC, SH_NAMES, SH_RECV, SH_RECJ, SH_GIDX, SH_LEN, SH_NLEN64, SH_NBEG,
       SH_SEG, SH_OWN, SH_CROSS, SH_OWNPOS, SH_CROSSPOS, SH_ITEMS_LOC, SH_EOUT_LOC, SH_GIT, SH_XBUF, SH_GFL, SH_LKEEP, SH_LKPOS, SH_LPFLAG, SH_LXPOS,
       SH_GM, SH_CURSOR, SH_SEND, SH_RECVBUF, SH_SZW, SH_EFLAG, SH_EPOS, SH_LMAP, SH_LEFLAG, SH_LEPOS, SH_ALLREC, SH_COUNT_ };

template <class T> int ensure_sh(catt_ctx* c, int slot, int64_t count, T** out) {
  static_assert(SH_COUNT_ <= SH_SLOTS, "raise SH_SLOTS in host_types.h");
  int rc = c->sh[slot].ensure((size_t)std::max<int64_t>(count, 1) * sizeof(T));
  *out = c->sh[slot].as<T>();
  return rc;
}

// local reads [a, b) of the current file: which are mapped, and how long are their QNAMEs
__global__ void mapped_flag_kernel(int64_t a, int64_t cnt_, const catt_aln* __restrict__ rv, const catt_aln* __restrict__ rj, NamesDev nd,
                                   uint32_t* __restrict__ flag, uint32_t* __restrict__ nlen) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > cnt_) return;
  if (t == cnt_) { flag[t] = 0; nlen[t] = 0; return; }
  const int64_t i = a + t;
  const bool m = rv[i].mapped || rj[i].mapped;
  flag[t] = m;
  nlen[t] = m ? (uint32_t)(nd.end[i] - nd.beg[i]) : 0u;
}
__global__ void grec_pack_kernel(int64_t a, int64_t cnt_, const catt_aln* __restrict__ rv, const catt_aln* __restrict__ rj, NamesDev nd,
                                 const int32_t* __restrict__ len, uint32_t gidx0, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                 const uint32_t* __restrict__ npos, GRec* __restrict__ out, char* __restrict__ names_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt_ || !flag[t]) return;
  const int64_t i = a + t;
  GRec g;
  g.v = rv[i]; g.j = rj[i]; g.gidx = gidx0 + (uint32_t)t; g.len = len[i]; g.name_len = (int32_t)(nd.end[i] - nd.beg[i]); g.pad_ = 0;
  out[pos[t]] = g;
  const char* s = nd.s + nd.beg[i];
  char* d = names_out + npos[t];
  for (int k = 0; k < g.name_len; k++) d[k] = s[k];
}
__global__ void grec_unpack_kernel(int64_t n, const GRec* __restrict__ in, catt_aln* __restrict__ rv, catt_aln* __restrict__ rj, uint32_t* __restrict__ gidx,
                                   int32_t* __restrict__ len, unsigned long long* __restrict__ nlen, unsigned long long* __restrict__ cnt2) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool mv = false, mj = false;
  if (t < n) {
    const GRec g = in[t];
    rv[t] = g.v; rj[t] = g.j; gidx[t] = g.gidx; len[t] = g.len; nlen[t] = (unsigned long long)g.name_len;
    mv = g.v.mapped != 0; mj = g.j.mapped != 0;
  } else if (t == n) {
    nlen[t] = 0;
  }
  const unsigned bv = __ballot_sync(FULL, mv), bj = __ballot_sync(FULL, mj);
  if ((threadIdx.x & 31) == 0) { if (bv) atomicAdd(&cnt2[0], (unsigned long long)__popc(bv)); if (bj) atomicAdd(&cnt2[1], (unsigned long long)__popc(bj)); }
}

// which items of the (globally identical) item list does this rank process, and which both-mapped pairs straddle two ranks
__global__ void own_kernel(int64_t n_seg, const ExtractItem* __restrict__ seg, const uint32_t* __restrict__ gidx, long long gstart, RankBounds rb, int me,
                           uint32_t* __restrict__ own, uint32_t* __restrict__ cross) {
  const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (it > n_seg) return;
  if (it == n_seg) { own[it] = 0; cross[it] = 0; return; }
  const ExtractItem x = seg[it];
  const int o1 = owner_of(rb, gstart + gidx[x.read]);
  const int o2 = x.kind == 2 ? owner_of(rb, gstart + gidx[x.read2]) : o1;
  own[it] = o1 == me;
  cross[it] = o1 != o2;
}
// this rank's items, in list order, with LOCAL read indices; the other read of a straddling pair is extra read `cross position`
__global__ void own_compact_kernel(int64_t n_seg, const ExtractItem* __restrict__ seg, const uint32_t* __restrict__ gidx, long long gstart, long long lo,
                                   const uint32_t* __restrict__ own, const uint32_t* __restrict__ cross, const uint32_t* __restrict__ ownpos,
                                   const uint32_t* __restrict__ crosspos, uint32_t extra_base, uint32_t item_base, ExtractItem* __restrict__ out,
                                   uint32_t* __restrict__ git) {
  const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= n_seg || !own[it]) return;
  ExtractItem x = seg[it];
  const uint32_t r2 = x.read2;
  x.read = (uint32_t)(gstart + gidx[x.read] - lo);
  x.read2 = x.kind != 2 ? x.read : (cross[it] ? extra_base + crosspos[it] : (uint32_t)(gstart + gidx[r2] - lo));
  out[ownpos[it]] = x;
  git[ownpos[it]] = item_base + (uint32_t)it;
}
// the J-side read of every straddling pair, written by the rank that owns it into slot `cross position`: [len, packed words]
__global__ void xfill_kernel(int64_t n_seg, const ExtractItem* __restrict__ seg, const uint32_t* __restrict__ gidx, long long gstart, long long lo,
                             RankBounds rb, int me, const uint32_t* __restrict__ cross, const uint32_t* __restrict__ crosspos, ReadsDev reads, int W,
                             uint32_t* __restrict__ xbuf) {
  const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= n_seg || !cross[it]) return;
  const long long g2 = gstart + gidx[seg[it].read2];
  if (owner_of(rb, g2) != me) return;
  const int64_t r = g2 - lo;
  const int len = reads.len[r], nw = read_words(len);
  uint32_t* d = xbuf + (size_t)crosspos[it] * (size_t)(W + 1);
  d[0] = (uint32_t)len;
  const uint32_t* w = reads.words + reads.off[r];
  for (int k = 0; k < nw; k++) d[1 + k] = w[k];
}
// ... appended behind the local reads as extra reads (fixed stride W words)
__global__ void xappend_kernel(int64_t nx, const uint32_t* __restrict__ xbuf, int W, int64_t read0, uint32_t word0, uint32_t* __restrict__ words,
                               uint32_t* __restrict__ off, int32_t* __restrict__ len) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nx) return;
  const uint32_t* s = xbuf + (size_t)x * (size_t)(W + 1);
  const uint32_t o = word0 + (uint32_t)x * (uint32_t)W;
  off[read0 + x] = o; len[read0 + x] = (int32_t)s[0];
  for (int k = 0; k < W; k++) words[o + k] = s[1 + k];
}

// one byte per global item, written by the item's rank: 1 kept Vpart read, 2 kept Jpart read, 4 partial (cdr3 == "None"),
// 8 dropped for SEQ mismatch (catt.jl:298), 16 kept both-mapped read
__global__ void gflag_fill_kernel(int64_t nl, const ExtractItem* __restrict__ items, const ExtractOut* __restrict__ eout, const uint32_t* __restrict__ git,
                                  uint8_t* __restrict__ gfl) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nl) return;
  const int st = eout[k].status, kind = items[k].kind;
  const bool kept = st == 1;
  gfl[git[k]] = (uint8_t)((kept && kind != 1 ? 1 : 0) | (kept && kind == 1 ? 2 : 0) | (kept && eout[k].cdr3_off < 0 ? 4 : 0) | (st == 2 ? 8 : 0) |
                          (kept && kind == 2 ? 16 : 0));
}
__global__ void gflag_expand_kernel(int64_t n, const uint8_t* __restrict__ gfl, uint32_t* __restrict__ kflag0, uint32_t* __restrict__ kflag1,
                                    unsigned long long* __restrict__ cnt) {
  const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t b = 0;
  if (it < n) b = gfl[it];
  if (it <= n) { kflag0[it] = b & 1; kflag1[it] = (b >> 1) & 1; }
  const unsigned bp = __ballot_sync(FULL, b & 4), bm = __ballot_sync(FULL, b & 8), bf = __ballot_sync(FULL, b & 16);
  if ((threadIdx.x & 31) == 0) {
    if (bp) atomicAdd(&cnt[C_NPARTIAL], (unsigned long long)__popc(bp));
    if (bm) atomicAdd(&cnt[C_MISMATCH], (unsigned long long)__popc(bm));
    if (bf) atomicAdd(&cnt[C_FULL], (unsigned long long)__popc(bf));
  }
}
__global__ void lkeep_kernel(int64_t nl, const ExtractOut* __restrict__ eout, uint32_t* __restrict__ lkeep) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= nl) lkeep[k] = k < nl && eout[k].status == 1;
}
// this rank's kept items -> PoolItem / OutDesc at their LOCAL rank j, carrying their GLOBAL position m in [Vpart; Jpart]
__global__ void pitem_shard_kernel(int64_t nl, const ExtractItem* __restrict__ items, const ExtractOut* __restrict__ eout, const uint32_t* __restrict__ git,
                                   const uint32_t* __restrict__ lkeep, const uint32_t* __restrict__ lkpos, const uint32_t* __restrict__ kpos0,
                                   const uint32_t* __restrict__ kpos1, uint32_t vpart, const int32_t* __restrict__ jlen, PoolItem* __restrict__ pitems,
                                   OutDesc* __restrict__ desc, uint32_t* __restrict__ gm, uint32_t* __restrict__ lpflag) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k == nl) lpflag[lkpos[nl]] = 0;                           // sentinel behind the last kept item
  if (k >= nl || !lkeep[k]) return;
  const ExtractItem x = items[k];
  const ExtractOut e = eout[k];
  const int side = x.kind == 1;
  const uint32_t g = git[k];
  const uint32_t m = side ? vpart + kpos1[g] : kpos0[g];
  const uint32_t j = lkpos[k];
  PoolItem pi;
  pi.read = x.read; pi.strand = x.strand; pi.seq_off = e.seq_off; pi.seq_len = e.seq_len; pi.side = (int16_t)side;
  pi.order = m; pi.partial = e.cdr3_off < 0;
  pitems[j] = pi;
  lpflag[j] = pi.partial;
  gm[j] = m;
  OutDesc d;
  d.read = x.read; d.xi = -1; d.strand = x.strand; d.kind = x.kind; d.seq_off = e.seq_off; d.seq_len = e.seq_len;
  d.cdr3_off = e.cdr3_off; d.cdr3_len = e.cdr3_len;
  d.v_id = x.kind == 1 ? -1 : x.rid; d.j_id = x.kind == 0 ? -1 : (x.kind == 1 ? x.rid : x.j_rid);
  d.q_start1 = (int16_t)(x.qb + 1); d.j_pad = x.kind == 1 ? (int16_t)(jlen[x.rid] - (x.rb + 1)) : 0; d.pad_ = 0;
  desc[j] = d;
}
__global__ void lpartial_kernel(int64_t ml, const uint32_t* __restrict__ lpflag, const uint32_t* __restrict__ lxpos, uint32_t* __restrict__ partial_idx,
                                OutDesc* __restrict__ desc) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ml || !lpflag[j]) return;
  partial_idx[lxpos[j]] = (uint32_t)j;
  desc[j].xi = (int32_t)lxpos[j];
}
// one word per global list position, written by the read's rank: bit 31 = the read is returned, bits 12.. = qual bytes, 0..11 = seq bytes
__global__ void szw_fill_kernel(StrSrc S, int only_cdr3, uint32_t* __restrict__ szw, uint32_t* __restrict__ leflag) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > S.M) return;
  if (j == S.M) { leflag[j] = 0; return; }
  int n_ext, c_off, c_len, slen, qlen, base_q;
  str_dims(S, j, &n_ext, &c_off, &c_len, &slen, &qlen, &base_q);
  const bool present = !only_cdr3 || c_off >= 0;
  leflag[j] = present;
  if (present) szw[S.gm[j]] = 0x80000000u | ((uint32_t)(qlen + 1) << 12) | (uint32_t)(slen + 1);
}
__global__ void szw_expand_kernel(int64_t M, const uint32_t* __restrict__ szw, uint32_t* __restrict__ eflag, unsigned long long* __restrict__ ssz,
                                  unsigned long long* __restrict__ qsz) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m > M) return;
  const uint32_t w = m < M ? szw[m] : 0u;
  const bool p = (w >> 31) != 0;
  eflag[m] = p; ssz[m] = p ? (w & 0xfffu) : 0u; qsz[m] = p ? ((w >> 12) & 0xfffu) : 0u;
}
__global__ void lmap_kernel(int64_t ml, const uint32_t* __restrict__ leflag, const uint32_t* __restrict__ lepos, const OutDesc* __restrict__ desc,
                            uint32_t* __restrict__ lmap, uint32_t* __restrict__ rid) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ml || !leflag[j]) return;
  lmap[lepos[j]] = (uint32_t)j;
  rid[lepos[j]] = desc[j].read;
}

}  // namespace

int stage_b_sharded(catt_ctx* c, const ShardIn& in, const NamesDev& names, const DeviceQuals& dq, catt_result* R, int* launches) {
  Comm* comm = in.comm;
  const int N = comm->nranks, me = comm->rank;
  catt_info& info = R->info;
  info.kmer = c->kmer;
  cudaStream_t st = c->stream;
  Trace tr;
  int rc;
  if (N > 16) { set_error("GPU group of %d ranks (limit 16)", N); return CATT_E_ARG; }
  if (in.bounds.back() == 0) return CATT_OK;                      // an empty sample (every rank sees that)
  unsigned long long* cnt; unsigned long long* hcnt;
  if ((rc = ensure(c, SB_CNT, C_COUNT, &cnt))) return rc;
  hcnt = c->pin[PIN_CNT].as<unsigned long long>() + 2 * N_COUNTERS;
  const int nfiles = (int)in.gstart.size() - 1;
  const int64_t lo = in.lo, n_loc = in.n_loc;
  RankBounds rb{};
  rb.n = N;
  for (int r = 0; r <= N; r++) rb.b[r] = in.bounds[r];
  const catt_aln* lrv = c->recs_all[0].as<catt_aln>();
  const catt_aln* lrj = c->recs_all[1].as<catt_aln>();
  const int W = read_words(in.max_len);
  float ms_order = 0, ms_extract = 0;
  // per collective: device time between two events on the stream (read after the next sync) and the bytes this rank moved
  int64_t bytes_mark = comm->bytes_moved;
  auto comm_lap = [&](int which, cudaEvent_t a, cudaEvent_t b) {
    float x = 0;
    cudaEventElapsedTime(&x, a, b);
    info.ms_coll[which] += x; info.ms_comm += x;
  };
  auto comm_bytes = [&](int which) { info.bytes_coll[which] += comm->bytes_moved - bytes_mark; info.bytes_comm += comm->bytes_moved - bytes_mark; bytes_mark = comm->bytes_moved; };

  // local items of all files, in list order
  ExtractItem* items_loc = nullptr; ExtractOut* eout_loc = nullptr; uint32_t* git = nullptr;
  int64_t nl = 0, nl_cap = 0;                     // local items so far / capacity
  int64_t n_items_g = 0;                          // global items so far
  int64_t n_extra = 0;                            // extra reads appended behind the local ones
  uint64_t extra_words = 0;
  int k = c->kmer;

  for (int f = 0; f < nfiles; f++) {
    const int64_t gs = in.gstart[f], ge = in.gstart[f + 1];
    const int64_t a = std::min(std::max(lo, gs), lo + n_loc) - lo, b = std::max(std::min(lo + n_loc, ge), lo) - lo;     // local reads [a, b) of this file
    const int64_t cn = std::max<int64_t>(b - a, 0);
    // ---- 1. mapped reads of this block -> wire format
    uint32_t *flag, *pos, *nlen, *npos;
    if ((rc = ensure_sh(c, SH_FLAG, cn + 1, &flag)) || (rc = ensure_sh(c, SH_POS, cn + 1, &pos)) || (rc = ensure_sh(c, SH_NLEN, cn + 1, &nlen)) ||
        (rc = ensure_sh(c, SH_NPOS, cn + 1, &npos))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[5], st));
    mapped_flag_kernel<<<nblk(cn + 1), 256, 0, st>>>(a, cn, lrv, lrj, names, flag, nlen);
    CATT_CUDA(cudaGetLastError());
    if ((rc = scan_u32(c, flag, pos, cn + 1, st)) || (rc = scan_u32(c, nlen, npos, cn + 1, st))) return rc;
    uint32_t h2[2];
    CATT_CUDA(cudaMemcpyAsync(&h2[0], pos + cn, 4, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaMemcpyAsync(&h2[1], npos + cn, 4, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    const int64_t nmap_loc = h2[0], nb_loc = h2[1];
    GRec* grec_loc; char* names_loc;
    if ((rc = ensure_sh(c, SH_GREC_LOC, nmap_loc, &grec_loc)) || (rc = ensure_sh(c, SH_NAMES_LOC, nb_loc + 16, &names_loc))) return rc;
    if (cn > 0) {
      grec_pack_kernel<<<nblk(cn), 256, 0, st>>>(a, cn, lrv, lrj, names, c->rd_len.as<int32_t>(), (uint32_t)(lo + a - gs), flag, pos, npos, grec_loc, names_loc);
      CATT_CUDA(cudaGetLastError());
    }
    *launches += 6;
    // ---- 2. all-gather them
    std::vector<long long> mine{(long long)nmap_loc, (long long)nb_loc}, all((size_t)2 * N);
    if ((rc = comm->host_allgather(mine.data(), all.data(), 2 * sizeof(long long), st))) return rc;
    std::vector<size_t> cnt_r(N), dsp_r(N), cnt_n(N), dsp_n(N);
    int64_t nmap = 0, nbytes = 0;
    for (int r = 0; r < N; r++) {
      cnt_r[r] = (size_t)all[2 * r] * sizeof(GRec); dsp_r[r] = (size_t)nmap * sizeof(GRec); nmap += all[2 * r];
      cnt_n[r] = (size_t)all[2 * r + 1]; dsp_n[r] = (size_t)nbytes; nbytes += all[2 * r + 1];
    }
    if (nmap >= 0xFFFFFFF0ll) { set_error("too many mapped reads in one file for the sharded path"); return CATT_E_LIMIT; }
    GRec* grec; char* gnames;
    if ((rc = ensure_sh(c, SH_GREC, nmap, &grec)) || (rc = ensure_sh(c, SH_NAMES, nbytes + 16, &gnames))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[2], st));
    if ((rc = comm->allgatherv(grec_loc, grec, cnt_r.data(), dsp_r.data(), st)) || (rc = comm->allgatherv(names_loc, gnames, cnt_n.data(), dsp_n.data(), st))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[3], st));
    comm_bytes(CATT_COLL_RECORDS);
    catt_aln *grv, *grj; uint32_t* gidx; int32_t* glen; unsigned long long *gnlen, *gnbeg;
    if ((rc = ensure_sh(c, SH_RECV, nmap, &grv)) || (rc = ensure_sh(c, SH_RECJ, nmap, &grj)) || (rc = ensure_sh(c, SH_GIDX, nmap, &gidx)) ||
        (rc = ensure_sh(c, SH_LEN, nmap, &glen)) || (rc = ensure_sh(c, SH_NLEN64, nmap + 1, &gnlen)) || (rc = ensure_sh(c, SH_NBEG, nmap + 1, &gnbeg))) return rc;
    CATT_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), st));
    grec_unpack_kernel<<<nblk(nmap + 1), 256, 0, st>>>(nmap, grec, grv, grj, gidx, glen, gnlen, cnt);
    CATT_CUDA(cudaGetLastError());
    if ((rc = scan_u64(c, gnlen, gnbeg, nmap + 1, st))) return rc;
    unsigned long long hm[2];
    CATT_CUDA(cudaMemcpyAsync(hm, cnt, sizeof hm, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    comm_lap(CATT_COLL_RECORDS, c->ev[2], c->ev[3]);
    *launches += 3;
    const int64_t nm[2] = {(int64_t)hm[0], (int64_t)hm[1]};
    // ---- 3. the global order, identically on every rank (indices = positions in the gathered arrays: input order is kept)
    OrderScratch OS;
    if ((rc = order_alloc(c, nmap, nm[0], nm[1], &OS))) return rc;
    ExtractItem* seg;
    if ((rc = ensure_sh(c, SH_SEG, nm[0] + nm[1], &seg))) return rc;
    const NamesDev gnd{gnames, reinterpret_cast<const int64_t*>(gnbeg), reinterpret_cast<const int64_t*>(gnbeg) + 1};
    const catt_aln* rec[2] = {grv, grj};
    OrderOut oo;
    if ((rc = order_file(c, OS, nmap, gnd, rec, glen, nm, 0u, &k, info, cnt, hcnt, seg, nm[0] + nm[1], &oo, launches))) return rc;
    const int64_t nseg = oo.n_v + oo.n_p + oo.n_j;
    if (n_items_g + nseg >= 0xFFFFFFF0ll) { set_error("too many work items for the sharded path"); return CATT_E_LIMIT; }
    // ---- 4. this rank's items; pairs that straddle two ranks
    uint32_t *own, *cross, *ownpos, *crosspos;
    if ((rc = ensure_sh(c, SH_OWN, nseg + 1, &own)) || (rc = ensure_sh(c, SH_CROSS, nseg + 1, &cross)) || (rc = ensure_sh(c, SH_OWNPOS, nseg + 1, &ownpos)) ||
        (rc = ensure_sh(c, SH_CROSSPOS, nseg + 1, &crosspos))) return rc;
    own_kernel<<<nblk(nseg + 1), 256, 0, st>>>(nseg, seg, gidx, gs, rb, me, own, cross);
    CATT_CUDA(cudaGetLastError());
    if ((rc = scan_u32(c, own, ownpos, nseg + 1, st)) || (rc = scan_u32(c, cross, crosspos, nseg + 1, st))) return rc;
    CATT_CUDA(cudaMemcpyAsync(&h2[0], ownpos + nseg, 4, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaMemcpyAsync(&h2[1], crosspos + nseg, 4, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    const int64_t n_own = h2[0], n_cross = h2[1];
    *launches += 5;
    if (nl + n_own > nl_cap) {
      nl_cap = (nl + n_own) + (nl + n_own) / 4 + 1024;
      if ((rc = c->sh[SH_ITEMS_LOC].ensure_keep((size_t)nl_cap * sizeof(ExtractItem), st)) || (rc = c->sh[SH_EOUT_LOC].ensure_keep((size_t)nl_cap * sizeof(ExtractOut), st)) ||
          (rc = c->sh[SH_GIT].ensure_keep((size_t)nl_cap * sizeof(uint32_t), st))) return rc;
    }
    items_loc = c->sh[SH_ITEMS_LOC].as<ExtractItem>(); eout_loc = c->sh[SH_EOUT_LOC].as<ExtractOut>(); git = c->sh[SH_GIT].as<uint32_t>();
    if (nseg > 0) {
      own_compact_kernel<<<nblk(nseg), 256, 0, st>>>(nseg, seg, gidx, gs, lo, own, cross, ownpos, crosspos, (uint32_t)(n_loc + n_extra), (uint32_t)n_items_g,
                                                     items_loc + nl, git + nl);
      CATT_CUDA(cudaGetLastError());
      *launches += 1;
    }
    if (n_cross > 0) {                             // the other read of every straddling pair: sparse fill + all-reduce, then appended locally
      uint32_t* xbuf;
      const size_t xw = (size_t)n_cross * (size_t)(W + 1);
      if ((rc = ensure_sh(c, SH_XBUF, (int64_t)xw, &xbuf))) return rc;
      CATT_CUDA(cudaMemsetAsync(xbuf, 0, xw * 4, st));
      ReadsDev lr{c->rd_words.as<uint32_t>(), c->rd_off.as<uint32_t>(), c->rd_len.as<int32_t>(), n_loc, in.max_len};
      xfill_kernel<<<nblk(nseg), 256, 0, st>>>(nseg, seg, gidx, gs, lo, rb, me, cross, crosspos, lr, W, xbuf);
      CATT_CUDA(cudaGetLastError());
      CATT_CUDA(cudaEventRecord(c->ev[2], st));
      if ((rc = comm->allreduce_sum_u32(xbuf, xw, st))) return rc;
      CATT_CUDA(cudaEventRecord(c->ev[3], st));
      comm_bytes(CATT_COLL_MATES);
      const uint64_t word0 = (uint64_t)in.n_words + extra_words;
      if (word0 + (uint64_t)n_cross * W >= 0xFFFFFFF0ull) { set_error("packed reads exceed 2^32 words"); return CATT_E_LIMIT; }
      if ((rc = c->rd_words.ensure_keep((size_t)(word0 + (uint64_t)n_cross * W + 4) * 4, st)) ||
          (rc = c->rd_off.ensure_keep((size_t)(n_loc + n_extra + n_cross + 2) * 4, st)) || (rc = c->rd_len.ensure_keep((size_t)(n_loc + n_extra + n_cross + 1) * 4, st))) return rc;
      xappend_kernel<<<nblk(n_cross), 256, 0, st>>>(n_cross, xbuf, W, n_loc + n_extra, (uint32_t)word0, c->rd_words.as<uint32_t>(), c->rd_off.as<uint32_t>(),
                                                    c->rd_len.as<int32_t>());
      CATT_CUDA(cudaGetLastError());
      CATT_CUDA(cudaStreamSynchronize(st));
      comm_lap(CATT_COLL_MATES, c->ev[2], c->ev[3]);
      n_extra += n_cross; extra_words += (uint64_t)n_cross * W;
      *launches += 2;
    }
    CATT_CUDA(cudaEventRecord(c->ev[6], st));
    // ---- 5. assignV/J, real_score, finder! on this rank's items
    const ReadsDev reads{c->rd_words.as<uint32_t>(), c->rd_off.as<uint32_t>(), c->rd_len.as<int32_t>(), n_loc + n_extra, in.max_len};
    if ((rc = launch_extract(c->V.dev, c->J.dev, reads, c->d_finder, items_loc + nl, eout_loc + nl, n_own, k, c->allow_v, c->allow_j, c->sm_count, st, launches))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[7], st));
    CATT_CUDA(cudaStreamSynchronize(st));
    { float x = 0, y = 0; cudaEventElapsedTime(&x, c->ev[5], c->ev[6]); cudaEventElapsedTime(&y, c->ev[6], c->ev[7]); ms_order += x; ms_extract += y; }
    nl += n_own; n_items_g += nseg;
  }
  info.kmer = k;
  tr.lap("B/N: gather + order + extract");
  const ReadsDev reads{c->rd_words.as<uint32_t>(), c->rd_off.as<uint32_t>(), c->rd_len.as<int32_t>(), n_loc + n_extra, in.max_len};

  // ---- 6. every read's position in [Vpart; Jpart]: one byte per item, all-reduced
  uint8_t* gfl; uint32_t *kflag0, *kflag1, *kpos0, *kpos1;
  const size_t gfl_words = ((size_t)n_items_g + 4) / 4 + 1;
  if ((rc = ensure_sh(c, SH_GFL, (int64_t)gfl_words * 4, &gfl)) || (rc = ensure(c, SB_KFLAG, n_items_g + 1, &kflag0)) || (rc = ensure(c, SB_KFLAG1, n_items_g + 1, &kflag1)) ||
      (rc = ensure(c, SB_KPOS2, n_items_g + 1, &kpos0)) || (rc = ensure(c, SB_KPOS3, n_items_g + 1, &kpos1))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  CATT_CUDA(cudaMemsetAsync(gfl, 0, gfl_words * 4, st));
  CATT_CUDA(cudaMemsetAsync(cnt, 0, C_COUNT * sizeof(unsigned long long), st));
  if (nl > 0) { gflag_fill_kernel<<<nblk(nl), 256, 0, st>>>(nl, items_loc, eout_loc, git, gfl); CATT_CUDA(cudaGetLastError()); }
  CATT_CUDA(cudaEventRecord(c->ev[2], st));
  if ((rc = comm->allreduce_sum_u32(reinterpret_cast<uint32_t*>(gfl), gfl_words, st))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[3], st));
  comm_bytes(CATT_COLL_FLAGS);
  gflag_expand_kernel<<<nblk(n_items_g + 1), 256, 0, st>>>(n_items_g, gfl, kflag0, kflag1, cnt);
  CATT_CUDA(cudaGetLastError());
  if ((rc = scan_u32(c, kflag0, kpos0, n_items_g + 1, st)) || (rc = scan_u32(c, kflag1, kpos1, n_items_g + 1, st))) return rc;
  uint32_t hv[2];
  CATT_CUDA(cudaMemcpyAsync(&hv[0], kpos0 + n_items_g, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(&hv[1], kpos1 + n_items_g, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(hcnt, cnt, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  comm_lap(CATT_COLL_FLAGS, c->ev[2], c->ev[3]);
  const int64_t vpart = hv[0], M = (int64_t)hv[0] + hv[1], n_partial_g = (int64_t)hcnt[C_NPARTIAL];
  info.vpart = vpart; info.jpart = M - vpart; info.full_pushed = (int64_t)hcnt[C_FULL]; info.pair_seq_mismatch = (int64_t)hcnt[C_MISMATCH];
  info.left = n_partial_g; info.direct = M - n_partial_g;
  // this rank's kept reads
  uint32_t *lkeep, *lkpos, *lpflag, *lxpos, *gm, *partial_idx;
  PoolItem* pitems; OutDesc* desc;
  if ((rc = ensure_sh(c, SH_LKEEP, nl + 1, &lkeep)) || (rc = ensure_sh(c, SH_LKPOS, nl + 1, &lkpos)) || (rc = ensure_sh(c, SH_LPFLAG, nl + 1, &lpflag)) ||
      (rc = ensure_sh(c, SH_LXPOS, nl + 1, &lxpos)) || (rc = ensure_sh(c, SH_GM, nl + 1, &gm)) || (rc = ensure(c, SB_PITEMS, nl, &pitems)) ||
      (rc = ensure(c, SB_DESC, nl, &desc)) || (rc = ensure(c, SB_PARTIAL, nl, &partial_idx))) return rc;
  lkeep_kernel<<<nblk(nl + 1), 256, 0, st>>>(nl, eout_loc, lkeep);
  CATT_CUDA(cudaGetLastError());
  if ((rc = scan_u32(c, lkeep, lkpos, nl + 1, st))) return rc;
  pitem_shard_kernel<<<nblk(nl + 1), 256, 0, st>>>(nl, items_loc, eout_loc, git, lkeep, lkpos, kpos0, kpos1, (uint32_t)vpart, c->J.dev.rec_len, pitems, desc, gm, lpflag);
  CATT_CUDA(cudaGetLastError());
  uint32_t hml = 0;
  CATT_CUDA(cudaMemcpyAsync(&hml, lkpos + nl, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  const int64_t ml = hml;                          // this rank's share of [Vpart; Jpart]
  if ((rc = scan_u32(c, lpflag, lxpos, ml + 1, st))) return rc;
  if (ml > 0) { lpartial_kernel<<<nblk(ml), 256, 0, st>>>(ml, lpflag, lxpos, partial_idx, desc); CATT_CUDA(cudaGetLastError()); }
  uint32_t hpl = 0;
  CATT_CUDA(cudaMemcpyAsync(&hpl, lxpos + ml, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaEventRecord(c->ev[6], st));
  CATT_CUDA(cudaStreamSynchronize(st));
  const int64_t n_partial = hpl;                   // ... and of the partial reads
  { float x = 0; cudaEventElapsedTime(&x, c->ev[5], c->ev[6]); info.ms_order += ms_order + x; info.ms_extract += ms_extract; }
  *launches += 12;
  tr.lap("B/N: lists");

  // ---- 7. the global k-mer table, then the extension of this rank's partial reads
  ExtendOut* xout = nullptr;
  const bool run_pool = M > 0 && k <= KMER_MAX && n_partial_g > 0;
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  if (run_pool) {
    KmerTable tab{};
    tab.nsub = (uint32_t)N;
    // 1.5 slots per insert (the one-GPU table uses 2): the sub-tables cross NVLink once, and reads of one clonotype share most
    // of their k-mers, so the real load factor is below 2/3
    tab.sub_slots = std::max<uint64_t>(1024, ((uint64_t)M * 11ull * 3ull / 2ull + N - 1) / N);
    const uint64_t slots = tab.sub_slots * (uint64_t)N;
    unsigned long long* cursor; PoolEntry *sendb, *recvb;
    if ((rc = ensure(c, SB_TAB_KEYS, (int64_t)slots, &tab.slots)) ||
        (rc = ensure(c, SB_XOUT, n_partial, &xout)) || (rc = ensure_sh(c, SH_CURSOR, 2 * N, &cursor)) || (rc = ensure_sh(c, SH_SEND, ml * 11, &sendb))) return rc;
    CATT_CUDA(cudaMemsetAsync(tab.slots + (uint64_t)me * tab.sub_slots, 0xff, tab.sub_slots * sizeof(ulonglong2), st));
    CATT_CUDA(cudaMemsetAsync(cursor, 0, (size_t)2 * N * sizeof(unsigned long long), st));
    CATT_CUDA(cudaEventRecord(c->ev[0], st));
    if ((rc = launch_pool_emit(reads, pitems, ml, k, tab, 0, cursor, nullptr, c->sm_count, st, launches))) return rc;
    std::vector<unsigned long long> hc((size_t)N), allc((size_t)N * N), off((size_t)N);
    CATT_CUDA(cudaMemcpyAsync(hc.data(), cursor, (size_t)N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    unsigned long long run = 0;
    for (int r = 0; r < N; r++) { off[r] = run; run += hc[r]; }
    CATT_CUDA(cudaMemcpyAsync(cursor + N, off.data(), (size_t)N * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    if ((rc = launch_pool_emit(reads, pitems, ml, k, tab, 1, cursor + N, sendb, c->sm_count, st, launches))) return rc;
    if ((rc = comm->host_allgather(hc.data(), allc.data(), (size_t)N * sizeof(unsigned long long), st))) return rc;
    std::vector<size_t> sc(N), sd(N), rcn(N), rd(N);
    size_t rtot = 0;
    for (int r = 0; r < N; r++) {
      sc[r] = (size_t)hc[r] * sizeof(PoolEntry); sd[r] = (size_t)off[r] * sizeof(PoolEntry);
      rcn[r] = (size_t)allc[(size_t)r * N + me] * sizeof(PoolEntry); rd[r] = rtot; rtot += rcn[r];
    }
    if ((rc = ensure_sh(c, SH_RECVBUF, (int64_t)(rtot / sizeof(PoolEntry)), &recvb))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[2], st));
    if ((rc = comm->alltoallv(sendb, sc.data(), sd.data(), recvb, rcn.data(), rd.data(), st))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[3], st));
    comm_bytes(CATT_COLL_KMER_A2A);
    if ((rc = launch_pool_insert(recvb, (int64_t)(rtot / sizeof(PoolEntry)), tab, c->sm_count, st, launches))) return rc;
    std::vector<size_t> gc(N, (size_t)tab.sub_slots * sizeof(ulonglong2)), gd(N);
    for (int r = 0; r < N; r++) gd[r] = (size_t)r * tab.sub_slots * sizeof(ulonglong2);
    CATT_CUDA(cudaEventRecord(c->ev[4], st));
    if ((rc = comm->allgatherv(tab.slots + (uint64_t)me * tab.sub_slots, tab.slots, gc.data(), gd.data(), st))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[1], st));
    comm_bytes(CATT_COLL_KMER_TABLE);
    if ((rc = launch_extend(reads, c->d_finder, pitems, partial_idx, n_partial, k, tab, xout, reinterpret_cast<uint32_t*>(cnt + C_RESCUED), c->sm_count, st, launches))) return rc;
  }
  CATT_CUDA(cudaEventRecord(c->ev[6], st));

  // ---- 8. string sizes at global positions -> global offsets
  StrSrc S{desc, xout, dq.quals, dq.seq_off, ml, vpart, nullptr, ml, nullptr};
  S.gm = gm; S.ordinal_base = lo;
  uint32_t *szw, *eflag, *epos, *leflag, *lepos, *lmap, *lrid;
  unsigned long long *ssz, *qsz, *soff, *qoff;
  if ((rc = ensure_sh(c, SH_SZW, M + 1, &szw)) || (rc = ensure_sh(c, SH_EFLAG, M + 1, &eflag)) || (rc = ensure_sh(c, SH_EPOS, M + 1, &epos)) ||
      (rc = ensure_sh(c, SH_LEFLAG, ml + 1, &leflag)) || (rc = ensure_sh(c, SH_LEPOS, ml + 1, &lepos)) || (rc = ensure_sh(c, SH_LMAP, ml + 1, &lmap)) ||
      (rc = ensure(c, SB_QRID, ml + 1, &lrid)) ||
      (rc = ensure(c, SB_SSZ, M + 1, &ssz)) || (rc = ensure(c, SB_QSZ, M + 1, &qsz)) || (rc = ensure(c, SB_SOFF, M + 1, &soff)) || (rc = ensure(c, SB_QOFF, M + 1, &qoff))) return rc;
  CATT_CUDA(cudaMemsetAsync(szw, 0, (size_t)(M + 1) * 4, st));
  szw_fill_kernel<<<nblk(ml + 1), 256, 0, st>>>(S, c->only_cdr3, szw, leflag);
  CATT_CUDA(cudaGetLastError());
  CATT_CUDA(cudaEventRecord(c->ev[7], st));
  if ((rc = comm->allreduce_sum_u32(szw, (size_t)M + 1, st))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev_a[0][0], st));
  comm_bytes(CATT_COLL_SIZES);
  szw_expand_kernel<<<nblk(M + 1), 256, 0, st>>>(M, szw, eflag, ssz, qsz);
  CATT_CUDA(cudaGetLastError());
  if ((rc = scan_u32(c, eflag, epos, M + 1, st)) || (rc = scan_u64(c, ssz, soff, M + 1, st)) || (rc = scan_u64(c, qsz, qoff, M + 1, st)) ||
      (rc = scan_u32(c, leflag, lepos, ml + 1, st))) return rc;
  if (ml > 0) { lmap_kernel<<<nblk(ml), 256, 0, st>>>(ml, leflag, lepos, desc, lmap, lrid); CATT_CUDA(cudaGetLastError()); }
  uint32_t he[3]; unsigned long long hs[2];
  CATT_CUDA(cudaMemcpyAsync(&he[0], epos + M, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(&he[1], epos + vpart, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(&he[2], lepos + ml, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(&hs[0], soff + M, 8, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(&hs[1], qoff + M, 8, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(hcnt, cnt, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  *launches += 10;
  { float x = 0; cudaEventElapsedTime(&x, c->ev[5], c->ev[6]); info.ms_pool += x; }
  if (run_pool) {
    float x = 0; cudaEventElapsedTime(&x, c->ev[0], c->ev[1]); info.ms_pool_kernel += x;
    comm_lap(CATT_COLL_KMER_A2A, c->ev[2], c->ev[3]); comm_lap(CATT_COLL_KMER_TABLE, c->ev[4], c->ev[1]);
  }
  comm_lap(CATT_COLL_SIZES, c->ev[7], c->ev_a[0][0]);
  info.rescued += (int64_t)(uint32_t)hcnt[C_RESCUED];     // this rank's partial reads; summed over the ranks by the caller
  const int64_t E = he[0], EV = he[1], el = he[2];
  tr.lap("B/N: k-mer table + extension + sizes");
  S.map = lmap; S.E = el; S.epos = epos;
  if (dq.host_quals && el > 0 && (rc = gather_host_quals(c, dq, lrid, el, &S, info, st))) return rc;

  // ---- 9. strings of this rank's reads at their global offsets; sum-reduce to the root (one writer per byte)
  const size_t seq_bytes = (size_t)hs[0] + 1, qual_bytes = (size_t)hs[1] + 1;
  const size_t seq_w = (seq_bytes + 3) / 4, qual_w = (qual_bytes + 3) / 4, part_w = (size_t)E * sizeof(catt_read) / 4;
  char *d_seq, *d_qual; catt_read* d_part;
  if ((rc = ensure(c, SB_OUT_SEQ, (int64_t)seq_w * 4, &d_seq)) || (rc = ensure(c, SB_OUT_QUAL, (int64_t)qual_w * 4, &d_qual)) ||
      (rc = ensure(c, SB_OUT_PART, E, &d_part))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  CATT_CUDA(cudaMemsetAsync(d_seq, 0, seq_w * 4, st));
  CATT_CUDA(cudaMemsetAsync(d_qual, 0, qual_w * 4, st));
  CATT_CUDA(cudaMemsetAsync(d_part, 0, std::max<size_t>(part_w * 4, 4), st));
  const bool root = me == in.root;
  if (root) {
    const double t2 = now_ms();
    const size_t part_bytes = (size_t)std::max<int64_t>(E, 1) * sizeof(catt_read);
    R->seq_arena = (char*)block_get(seq_bytes, &R->cap_seq);
    R->qual_arena = (char*)block_get(qual_bytes, &R->cap_qual);
    R->part[0] = (catt_read*)block_get(part_bytes, &R->cap_part[0]);
    if (!R->seq_arena || !R->qual_arena || !R->part[0]) { set_error("out of host memory for the result"); return CATT_E_LIMIT; }
    R->npart[0] = EV; R->npart[1] = E - EV;
    R->part[1] = R->part[0] + EV;
    info.ms_host += now_ms() - t2;
  }
  // every rank writes the ROOT's host addresses into its catt_read entries: the root publishes its two arena bases
  unsigned long long bases[2] = {root ? (unsigned long long)(uintptr_t)R->seq_arena : 0ull, root ? (unsigned long long)(uintptr_t)R->qual_arena : 0ull};
  std::vector<unsigned long long> allb((size_t)2 * N);
  if ((rc = comm->host_allgather(bases, allb.data(), sizeof bases, st))) return rc;
  const char* host_seq = reinterpret_cast<const char*>((uintptr_t)allb[(size_t)2 * in.root]);
  const char* host_qual = reinterpret_cast<const char*>((uintptr_t)allb[(size_t)2 * in.root + 1]);
  if (el > 0) {
    const int grid = (int)std::min<int64_t>((el + MW - 1) / MW, (int64_t)c->sm_count * 16);
    materialise_kernel<<<grid, MW * 32, 0, st>>>(S, reads, soff, qoff, d_seq, d_qual, d_part, host_seq, host_qual);
    CATT_CUDA(cudaGetLastError());
    *launches += 1;
  }
  CATT_CUDA(cudaEventRecord(c->ev[6], st));
  if (E > 0) {
    if ((rc = comm->reduce_sum_u32(reinterpret_cast<uint32_t*>(d_seq), seq_w, in.root, st)) ||
        (rc = comm->reduce_sum_u32(reinterpret_cast<uint32_t*>(d_qual), qual_w, in.root, st)) ||
        (rc = comm->reduce_sum_u32(reinterpret_cast<uint32_t*>(d_part), part_w, in.root, st))) return rc;
  }
  CATT_CUDA(cudaEventRecord(c->ev[7], st));
  comm_bytes(CATT_COLL_RESULT);
  if (root && E > 0) {
    if ((rc = d2h_arena(c, R->seq_arena, R->cap_seq, d_seq, seq_bytes - 1, st)) || (rc = d2h_arena(c, R->qual_arena, R->cap_qual, d_qual, qual_bytes - 1, st)) ||
        (rc = d2h_arena(c, R->part[0], R->cap_part[0], d_part, (size_t)E * sizeof(catt_read), st))) return rc;
    info.bytes_d2h += (int64_t)(seq_bytes + qual_bytes - 2) + E * (int64_t)sizeof(catt_read);
  }
  CATT_CUDA(cudaEventRecord(c->ev[0], st));
  CATT_CUDA(cudaStreamSynchronize(st));
  if (root && E > 0 && c->want_clonotypes && (rc = build_clonotypes(c, E, d_part, d_seq, d_qual, R->seq_arena, R->qual_arena, R, launches))) return rc;
  { float x = 0, y = 0; cudaEventElapsedTime(&x, c->ev[5], c->ev[6]); cudaEventElapsedTime(&y, c->ev[7], c->ev[0]); info.ms_strings += x; info.ms_d2h += y; }
  comm_lap(CATT_COLL_RESULT, c->ev[6], c->ev[7]);
  tr.lap("B/N: strings + reduce + D2H");
  return CATT_OK;
}

}  // namespace catt
