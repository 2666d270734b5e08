// assemble.cu — Stage B: from per-read alignment records to Vpart / Jpart (catt.jl:186-339 after the SAM files exist,
// plus the pool / extension / finder!(array) part of catt(), catt.jl:544-577).
//
// Everything the reference does between the stages with samtools / awk / Julia containers is restated on the device, on
// the alignment records Stage A left in HBM (ctx->recs_all), so no per-read data crosses PCIe until the final lists:
//   * `samtools sort -t AS | view -F 2308 | tac | awk '!a[$1]++'` (catt.jl:45-48): one 61-bit radix key
//     (AS, tid, pos, strand, input index) sorted descending == descending (AS, tid, pos, strand) with later input first
//     on complete ties; the awk dedup is an atomicMin of the sorted rank per trimmed QNAME (QNAMEs live in a device
//     hash set, compared byte-exactly on every hash hit)                                                  [SURVEY A7]
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

// ---- QNAME set: slot_of[i] = equivalence class of read i's trimmed name (only for reads with a V or J hit) ------------
__global__ void name_kernel(int64_t n, NamesDev nd, const catt_aln* __restrict__ rv, const catt_aln* __restrict__ rj, int32_t* owner,
                            uint32_t mask, uint32_t* __restrict__ slot_of) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(rv[i].mapped || rj[i].mapped)) { slot_of[i] = INF; return; }
  const char* s = nd.s + nd.beg[i];
  const int l = trimmed_len(s, (int)(nd.end[i] - nd.beg[i]));
  uint64_t h = 0xcbf29ce484222325ull;
  for (int k = 0; k < l; k++) { h ^= (unsigned char)s[k]; h *= 0x100000001b3ull; }
  h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32;
  uint32_t slot = (uint32_t)h & mask;
  while (true) {
    int32_t o = atomicCAS(&owner[slot], -1, (int32_t)i);
    if (o == -1 || o == (int32_t)i) break;
    const char* t = nd.s + nd.beg[o];
    const int lt = trimmed_len(t, (int)(nd.end[o] - nd.beg[o]));
    bool same = lt == l;
    for (int k = 0; same && k < l; k++) same = s[k] == t[k];
    if (same) break;
    slot = (slot + 1) & mask;
  }
  slot_of[i] = slot;
}

// ---- sort keys of the mapped records --------------------------------------------------------------------------------
__global__ void key_kernel(int64_t n, const catt_aln* __restrict__ recs, unsigned long long* __restrict__ keys,
                           unsigned long long* __restrict__ counter) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool m = i < n && recs[i].mapped;
  const unsigned bal = __ballot_sync(FULL, m);
  if (!bal) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == __ffs(bal) - 1) base = atomicAdd(counter, (unsigned long long)__popc(bal));
  base = __shfl_sync(FULL, base, __ffs(bal) - 1);
  if (m) {
    const catt_aln a = recs[i];
    keys[base + __popc(bal & ((1u << lane) - 1))] = ((unsigned long long)(uint16_t)a.as << 51) | ((unsigned long long)(uint32_t)a.rid << 42) |
                                                    ((unsigned long long)(uint16_t)a.rb << 33) | ((unsigned long long)(a.strand & 1) << 32) |
                                                    (unsigned long long)(uint32_t)i;
  }
}

__global__ void minrank_kernel(int64_t nm, const unsigned long long* __restrict__ sorted, const uint32_t* __restrict__ slot_of,
                               uint32_t* __restrict__ minrank) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nm) atomicMin(&minrank[slot_of[(uint32_t)sorted[r]]], (uint32_t)r);
}

// kept = first record of its QNAME in sorted order; the V list also feeds samtools stats' sums
__global__ void flag_kernel(int64_t nm, const unsigned long long* __restrict__ sorted, const uint32_t* __restrict__ slot_of,
                            const uint32_t* __restrict__ minrank, const catt_aln* __restrict__ recs, const int32_t* __restrict__ len,
                            int stats, uint32_t* __restrict__ flag, unsigned long long* __restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long sl = 0, sn = 0, sm = 0;
  if (r < nm) {
    const uint32_t idx = (uint32_t)sorted[r];
    const bool kept = minrank[slot_of[idx]] == (uint32_t)r;
    flag[r] = kept;
    if (kept && stats) { sl = (unsigned long long)len[idx]; sn = (unsigned long long)recs[idx].nm; sm = (unsigned long long)recs[idx].mi; }
  } else if (r == nm) {
    flag[r] = 0;
  }
  if (stats) {
    #pragma unroll
    for (int d = 16; d; d >>= 1) { sl += __shfl_xor_sync(FULL, sl, d); sn += __shfl_xor_sync(FULL, sn, d); sm += __shfl_xor_sync(FULL, sm, d); }
    if ((threadIdx.x & 31) == 0 && sl) { atomicAdd(&cnt[C_SUM_LEN], sl); atomicAdd(&cnt[C_SUM_NM], sn); atomicAdd(&cnt[C_SUM_MI], sm); }
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

// kept records in the reference's processing order; records of shared QNAMEs leave the part lists and (J side) seed the pairs
__global__ void order_kernel(int64_t nm, const unsigned long long* __restrict__ sorted, const uint32_t* __restrict__ flag,
                             const uint32_t* __restrict__ kpos, int T, const uint32_t* __restrict__ slot_of,
                             const uint32_t* __restrict__ minv, const uint32_t* __restrict__ minj, int which, NamesDev nd,
                             uint32_t* __restrict__ ord_idx, uint32_t* __restrict__ ord_flag, uint32_t* __restrict__ pair_j,
                             unsigned long long* __restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nm || !flag[r]) return;
  const uint32_t F = kpos[nm];
  const uint32_t idx = (uint32_t)sorted[r];
  const uint32_t y = chunk_pos(kpos[r] + 1, F, (uint32_t)T);
  const uint32_t slot = slot_of[idx];
  const bool shared = minv[slot] != INF && minj[slot] != INF;
  ord_idx[y] = idx;
  ord_flag[y] = !shared;
  if (which == 1 && shared) {
    const unsigned long long p = atomicAdd(&cnt[C_NP], 1ull);
    pair_j[p] = idx;
    const char* s = nd.s + nd.beg[idx];
    atomicMax(&cnt[C_MAXNAME], (unsigned long long)trimmed_len(s, (int)(nd.end[idx] - nd.beg[idx])));
  }
}

__global__ void collect_kernel(const uint32_t* kpos_v, int64_t nm_v, const uint32_t* kpos_j, int64_t nm_j, const uint32_t* opos_v,
                               const uint32_t* opos_j, unsigned long long* cnt) {
  cnt[C_F_V] = nm_v ? kpos_v[nm_v] : 0; cnt[C_F_J] = nm_j ? kpos_j[nm_j] : 0;
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

__global__ void pair_item_kernel(int64_t np, const uint32_t* __restrict__ pair_j, const uint32_t* __restrict__ slot_of,
                                 const uint32_t* __restrict__ minv, const unsigned long long* __restrict__ sorted_v,
                                 const catt_aln* __restrict__ rv, const catt_aln* __restrict__ rj, uint32_t lo,
                                 ExtractItem* __restrict__ items) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const uint32_t ji = pair_j[p];
  const uint32_t vi = (uint32_t)sorted_v[minv[slot_of[ji]]];
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
    const int side = m < S.vpart ? 0 : 1;
    const int lead = side == 1 ? n_ext : 0;                       // a left extension is prepended
    const int len = reads.len[p.read];
    const uint32_t* __restrict__ w = reads.words + reads.off[p.read];
    const uint32_t* __restrict__ nm = w + read_base_words(len);
    char* s = seq_arena + soff[e_];
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
    char* q = qual_arena + qoff[e_];
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
      rd.seq = host_seq + soff[e_]; rd.qual = host_qual + qoff[e_]; rd.seq_len = slen; rd.qual_len = qlen;
      rd.v_id = p.v_id; rd.j_id = p.j_id; rd.cdr3_off = c_off; rd.cdr3_len = c_off >= 0 ? c_len : 0;
      rd.ordinal = (int32_t)p.read; rd.able = 1;
      rd.flags = (uint8_t)((p.kind == 0 ? 1 : p.kind == 1 ? 2 : 4) | ((c_off >= 0 && c_off + c_len > qfull) ? 8 : 0));   // 8: reference BoundsError
      rd.pad_[0] = 0; rd.pad_[1] = 0;
      part[e_] = rd;
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
  const int64_t name_bytes = name_off[n];
  if ((rc = ensure(c, SB_NAMES, name_bytes + 16, &d_names)) || (rc = ensure(c, SB_NAME_OFF, n + 1, &d_off))) return rc;
  CATT_CUDA(cudaMemcpyAsync(d_names, names, (size_t)name_bytes, cudaMemcpyHostToDevice, st));
  CATT_CUDA(cudaMemcpyAsync(d_off, name_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (info) info->bytes_h2d += name_bytes + (n + 1) * 8;
  out->s = d_names; out->beg = d_off; out->end = d_off + 1;
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
  uint32_t* slot_of; int32_t* owner; uint32_t* minr[2];
  uint64_t cap = 1024;
  while (cap < (uint64_t)(max_v + max_j) * 2 + 2) cap <<= 1;
  if (cap > 0x80000000ull) { set_error("too many mapped reads for the QNAME table"); return CATT_E_LIMIT; }
  if ((rc = ensure(c, SB_SLOT_OF, max_reads, &slot_of)) || (rc = ensure(c, SB_OWNER, (int64_t)cap, &owner)) ||
      (rc = ensure(c, SB_MINV, (int64_t)cap, &minr[0])) || (rc = ensure(c, SB_MINJ, (int64_t)cap, &minr[1]))) return rc;
  unsigned long long *keys_a, *sorted[2];
  uint32_t *flag, *kpos[2], *ord_idx[2], *ord_flag[2], *opos[2], *pair_a, *pair_b;
  const int64_t nmax = std::max(max_v, max_j);
  if ((rc = ensure(c, SB_KEYS_A, nmax, &keys_a)) || (rc = ensure(c, SB_SORT_V, max_v, &sorted[0])) ||
      (rc = ensure(c, SB_SORT_J, max_j, &sorted[1])) || (rc = ensure(c, SB_FLAG, nmax + 1, &flag)) || (rc = ensure(c, SB_KPOS_V, max_v + 1, &kpos[0])) ||
      (rc = ensure(c, SB_KPOS_J, max_j + 1, &kpos[1])) || (rc = ensure(c, SB_ORD_IDX_V, max_v + 1, &ord_idx[0])) ||
      (rc = ensure(c, SB_ORD_IDX_J, max_j + 1, &ord_idx[1])) || (rc = ensure(c, SB_ORD_FLAG_V, max_v + 1, &ord_flag[0])) ||
      (rc = ensure(c, SB_ORD_FLAG_J, max_j + 1, &ord_flag[1])) ||
      (rc = ensure(c, SB_OPOS_V, max_v + 1, &opos[0])) || (rc = ensure(c, SB_OPOS_J, max_j + 1, &opos[1])) ||
      (rc = ensure(c, SB_PAIR_A, max_j + 1, &pair_a)) || (rc = ensure(c, SB_PAIR_B, max_j + 1, &pair_b))) return rc;
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
    CATT_CUDA(cudaMemsetAsync(cnt, 0, C_COUNT * sizeof(unsigned long long), st));
    CATT_CUDA(cudaEventRecord(c->ev[5], st));
    if (nm[0] + nm[1] > 0) {
      CATT_CUDA(cudaMemsetAsync(owner, 0xff, cap * sizeof(int32_t), st));
      CATT_CUDA(cudaMemsetAsync(minr[0], 0xff, cap * sizeof(uint32_t), st));
      CATT_CUDA(cudaMemsetAsync(minr[1], 0xff, cap * sizeof(uint32_t), st));
      name_kernel<<<nblk(nf), 256, 0, st>>>(nf, nd, rec[0], rec[1], owner, (uint32_t)(cap - 1), slot_of);
      CATT_CUDA(cudaGetLastError());
      *launches += 1;
    }
    // per reference set: sorted final SAM order, QNAME dedup
    for (int which = 0; which < 2; which++) {
      if (nm[which] == 0) continue;
      key_kernel<<<nblk(nf), 256, 0, st>>>(nf, rec[which], keys_a, cnt + C_NMAP_V + which);
      CATT_CUDA(cudaGetLastError());
      size_t need = 0;
      CATT_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, need, keys_a, sorted[which], nm[which], 0, 61, st));
      if ((rc = ensure_cub(c, need))) return rc;
      need = c->sb[SB_CUB].cap;
      CATT_CUDA(cub::DeviceRadixSort::SortKeysDescending(c->sb[SB_CUB].p, need, keys_a, sorted[which], nm[which], 0, 61, st));
      minrank_kernel<<<nblk(nm[which]), 256, 0, st>>>(nm[which], sorted[which], slot_of, minr[which]);
      CATT_CUDA(cudaGetLastError());
      *launches += 10;     // key, radix sort (histogram + 7 onesweep passes), minrank
    }
    for (int which = 0; which < 2; which++) {
      if (nm[which] == 0) continue;
      flag_kernel<<<nblk(nm[which] + 1), 256, 0, st>>>(nm[which], sorted[which], slot_of, minr[which], rec[which], reads.len + lo, which == 0, flag, cnt);
      CATT_CUDA(cudaGetLastError());
      if ((rc = scan_u32(c, flag, kpos[which], nm[which] + 1, st))) return rc;
      CATT_CUDA(cudaMemsetAsync(ord_flag[which], 0, (size_t)(nm[which] + 1) * sizeof(uint32_t), st));
      order_kernel<<<nblk(nm[which]), 256, 0, st>>>(nm[which], sorted[which], flag, kpos[which], c->nthreads, slot_of, minr[0], minr[1], which, nd,
                                                    ord_idx[which], ord_flag[which], pair_a, cnt);
      CATT_CUDA(cudaGetLastError());
      if ((rc = scan_u32(c, ord_flag[which], opos[which], nm[which] + 1, st))) return rc;
      *launches += 6;
    }
    collect_kernel<<<1, 1, 0, st>>>(kpos[0], nm[0], kpos[1], nm[1], opos[0], opos[1], cnt);
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
    uint32_t* pairs = pair_a;
    if (n_p > 1) {
      unsigned long long *pk_a, *pk_b;
      if ((rc = ensure(c, SB_PKEY_A, n_p, &pk_a)) || (rc = ensure(c, SB_PKEY_B, n_p, &pk_b))) return rc;
      const int digits = std::max(1, (int)((hcnt[C_MAXNAME] + 7) / 8));
      uint32_t* other = pair_b;
      for (int d = digits - 1; d >= 0; d--) {
        pair_key_kernel<<<nblk(n_p), 256, 0, st>>>(n_p, pairs, nd, d, pk_a);
        CATT_CUDA(cudaGetLastError());
        size_t need = 0;
        CATT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, pk_a, pk_b, pairs, other, n_p, 0, 64, st));
        if ((rc = ensure_cub(c, need))) return rc;
        need = c->sb[SB_CUB].cap;
        CATT_CUDA(cub::DeviceRadixSort::SortPairs(c->sb[SB_CUB].p, need, pk_a, pk_b, pairs, other, n_p, 0, 64, st));
        std::swap(pairs, other);
        *launches += 10;
      }
    }
    // work items in the reference's processing order: [V part | both-mapped | J part] of this file
    const int64_t nseg = n_v + n_p + n_j;
    if (n_items + nseg > tot_hits) { set_error("stage B: item count exceeds the mapped-record bound"); return CATT_E_ARG; }
    ExtractItem* seg = items + n_items;
    if (F[0] > 0) {
      part_item_kernel<<<nblk(F[0]), 256, 0, st>>>(F[0], ord_idx[0], ord_flag[0], opos[0], rec[0], 0, (uint32_t)lo, seg);
      CATT_CUDA(cudaGetLastError());
      *launches += 1;
    }
    if (n_p > 0) {
      pair_item_kernel<<<nblk(n_p), 256, 0, st>>>(n_p, pairs, slot_of, minr[0], sorted[0], rec[0], rec[1], (uint32_t)lo, seg + n_v);
      CATT_CUDA(cudaGetLastError());
      *launches += 1;
    }
    if (F[1] > 0) {
      part_item_kernel<<<nblk(F[1]), 256, 0, st>>>(F[1], ord_idx[1], ord_flag[1], opos[1], rec[1], 1, (uint32_t)lo, seg + n_v + n_p);
      CATT_CUDA(cudaGetLastError());
      *launches += 1;
    }
    CATT_CUDA(cudaEventRecord(c->ev[6], st));
    // K5: assignV/J, real_score, clipping, finder! with the k that is current for this file (catt.jl:279,301)
    if ((rc = launch_extract(c->V.dev, c->J.dev, reads, c->d_finder, seg, eout + n_items, nseg, k, c->allow_v, c->allow_j, st, launches))) return rc;
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
  const bool run_pool = M > 0 && k <= 31 && n_partial > 0;
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  if (run_pool) {
    uint64_t slots = 1024;
    while (slots < (uint64_t)M * 11ull * 2ull) slots <<= 1;
    KmerTable tab{};
    if ((rc = ensure(c, SB_TAB_KEYS, (int64_t)slots, &tab.keys)) || (rc = ensure(c, SB_TAB_VALS, (int64_t)slots, &tab.vals)) ||
        (rc = ensure(c, SB_XOUT, n_partial, &xout))) return rc;
    tab.mask = slots - 1;
    CATT_CUDA(cudaMemsetAsync(tab.keys, 0xff, slots * sizeof(unsigned long long), st));
    CATT_CUDA(cudaMemsetAsync(tab.vals, 0, slots * sizeof(unsigned long long), st));
    CATT_CUDA(cudaEventRecord(c->ev[0], st));
    if ((rc = launch_pool(reads, pitems, M, k, tab, st, launches))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev[1], st));
    if ((rc = launch_extend(reads, c->d_finder, pitems, partial_idx, n_partial, k, tab, xout, reinterpret_cast<uint32_t*>(cnt + C_RESCUED), st, launches))) return rc;
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
      uint32_t* d_rid; char* d_gather; int64_t* d_entry;
      if ((rc = ensure(c, SB_QRID, E, &d_rid)) || (rc = c->pin[PIN_QRID].ensure((size_t)E * 4)) || (rc = c->pin[PIN_QENTRY].ensure((size_t)(E + 1) * 8))) return rc;
      entry_read_kernel<<<nblk(E), 256, 0, st>>>(desc, cmap, E, d_rid);
      CATT_CUDA(cudaGetLastError());
      *launches += 1;
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
      S.quals = d_gather; S.q_entry_off = d_entry;
    }
  }
  unsigned long long *ssz, *qsz, *soff, *qoff;
  if ((rc = ensure(c, SB_SSZ, E + 1, &ssz)) || (rc = ensure(c, SB_QSZ, E + 1, &qsz)) || (rc = ensure(c, SB_SOFF, E + 1, &soff)) ||
      (rc = ensure(c, SB_QOFF, E + 1, &qoff))) return rc;
  strsize_kernel<<<nblk(E + 1), 256, 0, st>>>(S, ssz, qsz);
  CATT_CUDA(cudaGetLastError());
  for (int which = 0; which < 2; which++) {
    size_t need = 0;
    CATT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, which ? qsz : ssz, which ? qoff : soff, E + 1, st));
    if ((rc = ensure_cub(c, need))) return rc;
    need = c->sb[SB_CUB].cap;
    CATT_CUDA(cub::DeviceScan::ExclusiveSum(c->sb[SB_CUB].p, need, which ? qsz : ssz, which ? qoff : soff, E + 1, st));
  }
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
  }
  tr.lap("B: strings + D2H");
  return CATT_OK;
}

}  // namespace catt
